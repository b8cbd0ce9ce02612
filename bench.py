#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 superpixel engine (see the contract in DESIGN.md "Measurement").

Workload (BASELINE.json metric "SLIC megapixels/sec (10 iters, 4K)"; configs[4]: 1024 synthetic 4K frames sharded at 1/2/4/8):
  a job of F = 1024 frames drawn from 16 synthetic seeds is dealt to the ranks by shard.frames_for_rank (frame i -> rank
  i mod N, no data-path collective; STRONG scaling: the job is fixed, a rank holds F/N frame buffers, already Lab, 8UC3,
  resident in HBM, every one its own buffer >> L2);
  a STEP runs, for every frame of the rank's share, createSuperpixelSLIC(frame, SLIC, 20, 10) [seeds + perturbation],
  iterate(10) and getLabels into a device buffer, through the C-ABI batch entry point.
  value  = megapixels/s over all GPUs with inputs resident (device timing, max over ranks);
  e2e    = same metric through the host-buffer C-ABI calls (pinned host frames in, labels out: H2D + D2H inside);
  roofline = dominant kernel (slic assignment) : 7 B/px algorithmic / measured launch time vs measured HBM peak;
  cpu_baseline = the oracle restatement on the box's host cores, bounded sample.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F]
"""
import argparse
import ctypes as C
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
PKG = "superpixels-segmentation-gui-opencv_b200"

W4K, H4K, REGION, RULER, ITERS = 3840, 2160, 20, 10.0, 10
METRIC = "slic_megapixels_per_sec_10iters_4k"
JOB_FRAMES, JOB_SEEDS = 1024, 16
WORKLOAD = ("SLIC region_size=20 ruler=10, 10 iterations, 3840x2160 synthetic frames (createSuperpixelSLIC + iterate + getLabels "
            "per frame); job = 1024 frames from 16 seeds (BASELINE.json configs[4])")


def job_config(frames):
    """identical in both arms (the driver compares the two lines' config)"""
    return {"workload": WORKLOAD, "frames": frames, "seeds": JOB_SEEDS, "l2": "every frame is its own 24.9 MB buffer: inputs larger than L2"}


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` summary (profiles/), or None"""
    import glob
    import re
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_slic_tile_v*.txt"))):
        try:
            txt = open(f).read()
            rd = re.search(r"dram__bytes_read.sum\s+([0-9.]+)\s+(\w+)", txt)
            wr = re.search(r"dram__bytes_write.sum\s+([0-9.]+)\s+(\w+)", txt)
            if not rd or not wr:
                continue
            unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            best = {"bytes": float(rd.group(1)) * unit[rd.group(2)] + float(wr.group(1)) * unit[wr.group(2)],
                    "source": os.path.relpath(f, ROOT)}
        except Exception:
            continue
    return best


def ncu_issue():
    """issue utilisation and executed warp-instructions of the dominant kernel from the newest committed ncu summary"""
    import glob
    import re
    out = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_slic_tile_v*.txt"))):
        txt = open(f).read()
        a = re.search(r"smsp__issue_active.avg.pct_of_peak_sustained_active\s+([0-9.]+)", txt)
        b = re.search(r"smsp__inst_executed.sum\s+([0-9.]+)", txt)
        c = re.search(r"smsp__thread_inst_executed_per_inst_executed.ratio\s+([0-9.]+)", txt)
        if a and b:
            out = {"issue_frac": float(a.group(1)) / 100.0, "warp_inst": float(b.group(1)),
                   "threads_per_inst": float(c.group(1)) if c else 32.0, "source": os.path.relpath(f, ROOT)}
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples nvidia-smi clocks / throttle reasons during the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 8 for i in range(4) if r[4 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_frames(synth, n_distinct):
    return [synth.synth(W4K, H4K, s)[0] for s in range(n_distinct)]


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path, timed on the host cores.  The reference's arithmetic
    lives in OpenCV-contrib ximgproc, which is neither vendored nor installable here, so this arm times the
    oracle restatement (kind "port") with every host thread; one step = one 4K frame (bounded sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    synth = importlib.import_module(PKG + ".synth")
    cores = os.cpu_count() or 1
    O.set_threads(cores)
    bgr = make_frames(synth, 1)[0]
    lab = O.bgr2lab8(bgr)

    def step():
        s = O.SuperpixelSLIC(lab, O.SLIC, REGION, RULER)
        s.iterate(ITERS)
        return s.getLabels()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    mps = W4K * H4K / 1e6 / dt
    out = {"impl": "reference", "metric": METRIC, "value": mps, "unit": "MP/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": job_config(args.frames),
           "cpu_baseline": {"value": mps, "unit": "MP/s", "cores": cores, "kind": "port",
                            "sample": "bounded sample of the job: %d steps x ONE 4K frame (seed 0) each, oracle restatement of ximgproc "
                                      "SLIC (NOT ximgproc itself: OpenCV-contrib is not installable here) with OpenMP row bands" % args.steps},
           "e2e": {"value": mps, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(out)


def other_configs(spx, synth, L, cpu=True):
    """BASELINE.json's remaining configs on rank 0 (not bench lines: device time of iterate() with the image resident,
    wall clock around the synchronous C-ABI calls, best of 3).  MP/s = pixels / iterate time.  With `cpu`, the CPU oracle
    (kind "port": the restatement, not ximgproc) runs the same call sequence ONCE on all host threads beside it."""
    out = {}
    O = None
    if cpu:
        from oracle import oracle as O
        O.set_threads(os.cpu_count() or 1)

    def best(f, n=3):
        t = 1e30
        for _ in range(n):
            t0 = time.perf_counter(); f(); t = min(t, time.perf_counter() - t0)
        return t

    def once(f):
        t0 = time.perf_counter(); r = f(); return time.perf_counter() - t0, r

    # ---- configs[0]: the whole COMPUTE click (mainwindow.cpp:2096-2111) on the reference's example image, host buffers in and out
    try:
        import cv2
        ex = cv2.imread(os.path.join(ROOT, "tests", "golden", "logo-absurdephoton-segmentation-image.png"), cv2.IMREAD_COLOR)
    except Exception:
        ex = None
    if ex is not None:
        def click():
            lab = spx.cvtColor(ex, spx.COLOR_BGR2Lab)
            s = spx.createSuperpixelSLIC(lab, spx.SLIC, 20, 10.0)
            s.iterate(ITERS); s.enforceLabelConnectivity(25)
            k = s.getNumberOfSuperpixels(); lab_ = s.getLabels(); m = s.getLabelContourMask(True)
            s.release()
            return k, lab_, m
        click()
        t = best(click)
        k, gl, gm = click()
        e = {"image": "example/logo-absurdephoton-segmentation-image.png %dx%d" % (ex.shape[1], ex.shape[0]),
             "calls": "cvtColor(BGR2Lab) + createSuperpixelSLIC(SLIC,20,10) + iterate(10) + enforceLabelConnectivity(25) + "
                      "getNumberOfSuperpixels + getLabels + getLabelContourMask, host buffers",
             "click_ms": t * 1e3, "MP_s": ex.shape[0] * ex.shape[1] / 1e6 / t, "superpixels": k}
        if O is not None:
            def cclick():
                lab = O.bgr2lab8(ex)
                s = O.SuperpixelSLIC(lab, O.SLIC, 20, 10.0)
                s.iterate(ITERS); s.enforceLabelConnectivity(25)
                return s.getNumberOfSuperpixels(), s.getLabels(), s.getLabelContourMask(True)
            cclick()
            ct, (ck, cl, cm) = once(cclick)
            e.update({"cpu_oracle_click_ms": ct * 1e3, "cpu_cores": os.cpu_count(), "cpu_kind": "port",
                      "identical_to_cpu_oracle": bool(ck == k and (cl == gl).all() and (cm == gm).all())})
        out["cfg1_example_S20"] = e

    def slic_family(name, alg, W, H, S, cpu_alg=None):
        bgr = synth.synth(W, H, 0)[0]
        lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
        s = spx.createSuperpixelSLIC(lab, alg, S, RULER)
        s.iterate(ITERS)

        def it():
            s.reset(lab); L.spx_slic_synchronize(s._h)
            t0 = time.perf_counter(); s.iterate(ITERS); L.spx_slic_synchronize(s._h)
            return time.perf_counter() - t0
        t = min(it() for _ in range(3))
        tm = best(lambda: s.getLabelContourMask(False))
        out[name] = {"iterate10_ms": t * 1e3, "MP_s": W * H / 1e6 / t, "frac_of_hbm_roofline_7B_px_iter": 70.0 * W * H / t / 1e9 / peaks()[0],
                     "contour_mask_ms_incl_d2h": tm * 1e3}
        if O is not None and cpu_alg is not None:
            c = O.SuperpixelSLIC(lab, cpu_alg, S, RULER)
            ct, _ = once(lambda: c.iterate(ITERS))
            out[name].update({"cpu_oracle_iterate10_ms": ct * 1e3, "cpu_MP_s": W * H / 1e6 / ct, "cpu_cores": os.cpu_count(), "cpu_kind": "port",
                              "labels_identical_to_cpu_oracle": bool((c.getLabels() == s.getLabels()).all())})
        s.release()

    slic_family("SLIC_1080p_S20", spx.SLIC, 1920, 1080, 20)          # north-star: 1080p / 4K / 8K at one GPU
    slic_family("SLIC_4K_S20", spx.SLIC, W4K, H4K, 20)
    slic_family("SLIC_8K_S20", spx.SLIC, 7680, 4320, 20)
    slic_family("cfg2_SLICO_4K_S25", spx.SLICO, W4K, H4K, 25, O.SLICO if O else None)
    slic_family("cfg3_MSLIC_8K_S30", spx.MSLIC, 7680, 4320, 30, O.MSLIC if O else None)
    bgr = synth.synth(7680, 4320, 0)[0]
    lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
    l = spx.createSuperpixelLSC(lab, 30, 0.075)
    t1, _ = once(lambda: (l.iterate(ITERS), l.getNumberOfSuperpixels()))
    t = best(lambda: (l.iterate(ITERS), l.getNumberOfSuperpixels()), 2)       # further iterations on the same object
    e = {"iterate10_ms": t * 1e3, "first_iterate10_ms": t1 * 1e3, "MP_s": 7680 * 4320 / 1e6 / t}
    l.release()
    tc = []
    for _ in range(2):                                     # the first call also grows the stream-ordered memory pool
        l = spx.createSuperpixelLSC(lab, 30, 0.075); l.iterate(ITERS); l.getNumberOfSuperpixels()
        tc.append(once(lambda: l.enforceLabelConnectivity(25))[0])
        if len(tc) < 2:
            l.release()
    e.update({"enforce_connectivity25_ms": tc[1] * 1e3, "enforce_connectivity25_first_call_ms": tc[0] * 1e3,
              "merge_order": "upstream raster order (default)", "superpixels": l.getNumberOfSuperpixels()})
    if O is not None:
        c = O.SuperpixelLSC(lab, 30, 0.075)
        ct, _ = once(lambda: c.iterate(ITERS))
        cc, _ = once(lambda: c.enforceLabelConnectivity(25))
        e.update({"cpu_oracle_iterate10_ms": ct * 1e3, "cpu_oracle_enforce_connectivity25_ms": cc * 1e3, "cpu_MP_s": 7680 * 4320 / 1e6 / ct,
                  "cpu_cores": os.cpu_count(), "cpu_kind": "port", "labels_identical_to_cpu_oracle": bool((c.getLabels() == l.getLabels()).all())})
    out["cfg3_LSC_8K_S30_ratio0.075"] = e
    l.release()
    bgr = synth.synth(1920, 1080, 0)[0]
    lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
    sd = spx.createSuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)
    sd.iterate(lab, ITERS)
    t = best(lambda: sd.iterate(lab, ITERS))
    out["cfg4_SEEDS_1080p_2000sp_4lv_5bins"] = {"iterate10_ms_incl_h2d": t * 1e3, "MP_s": 1920 * 1080 / 1e6 / t}
    if O is not None:
        c = O.SuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)       # upstream's sequential visiting order
        ct, _ = once(lambda: c.iterate(lab, ITERS))
        out["cfg4_SEEDS_1080p_2000sp_4lv_5bins"].update({"cpu_oracle_iterate10_ms": ct * 1e3, "cpu_MP_s": 1920 * 1080 / 1e6 / ct,
                                                          "cpu_cores": os.cpu_count(), "cpu_kind": "port (sequential order, one thread)"})
    sd.release()
    return out


def strip_leg(spx, dist, rank, world, local):
    """row-strip mode (one image over all ranks of the job) on a 16384 x 16384 frame: every rank builds only its
    band (+2 halo rows); rank 0 also runs the single-GPU object on the whole frame for the comparison.  All ranks call this."""
    import torch
    strips = importlib.import_module(PKG + ".strips")
    W = H = 16384

    def rows(a, b):
        """rows [a, b) of the frame: 64 x 64 blocks of random colour xor per-row noise, the same on every rank"""
        out = np.empty((b - a, W, 3), np.uint8)
        for y0 in range(a - a % 64, b, 64):
            base = np.random.default_rng(7000 + y0 // 64).integers(0, 256, (1, (W + 63) // 64, 3), dtype=np.uint8)
            blk = np.kron(base, np.ones((64, 64, 1), np.uint8))[:, :W] ^ np.random.default_rng(9000 + y0 // 64).integers(0, 8, (64, W, 3), dtype=np.uint8)
            lo, hi = max(a, y0), min(b, y0 + 64)
            out[lo - a:hi - a] = blk[lo - y0:hi - y0]
        return out

    res = {}
    labels = None
    for mode in ("peer", "neighbour"):                     # replicated state + all-to-all first (comparison), then the partitioned default
        for it in range(2):                                # the first object warms kernels, peer mappings and NCCL up
            s = strips.StripSLIC(rows, 20, 10.0, dist, local, width=W, height=H, exchange=mode)
            if it == 0:
                s.iterate(1); s.release(); continue
            torch.cuda.synchronize(); dist.barrier()
            t0 = time.perf_counter(); s.iterate(ITERS); torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter() - t0
            tt = torch.tensor([t], dtype=torch.float64, device=torch.device("cuda", local))
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            if mode == "peer":
                res.update({"all_to_all_iterate10_ms": float(tt.item()) * 1e3, "all_to_all_sent_bytes_per_rank_per_iteration": s.exchangedBytes() / ITERS})
            else:
                labels = s.getLabels()                     # assembled on rank 0
                res.update({"image": "16384x16384 synthetic (64x64 colour blocks xor noise), SLIC S=20 ruler=10, 10 iterations", "n_gpus": world,
                            "strip_iterate10_ms": float(tt.item()) * 1e3, "MP_s": W * H / 1e6 / float(tt.item()),
                            "sent_bytes_per_rank_per_iteration": s.exchangedBytes() / ITERS,
                            "exchange": "%s: NVLink peer stores + flags (library kernels)" % s.exchange,
                            "superpixels": s.getNumberOfSuperpixels()})
            s.release()
    if rank == 0:
        img = rows(0, H)
        g = spx.createSuperpixelSLIC(img, spx.SLIC, 20, 10.0, device=local)
        spx.lib().spx_slic_synchronize(g._h)
        t0 = time.perf_counter(); g.iterate(ITERS); spx.lib().spx_slic_synchronize(g._h); t1 = time.perf_counter() - t0
        res.update({"single_gpu_iterate10_ms": t1 * 1e3, "speedup_vs_one_gpu": t1 * 1e3 / res["strip_iterate10_ms"],
                    "labels_identical": bool((g.getLabels() == labels).all())})
        g.release()
    dist.barrier()
    return res


def bind_to_gpu_numa_node(index):
    """pin this process (and the pinned host buffers it allocates from now on: first touch) to the CPUs of the NUMA node the
    GPU hangs off (sysfs local_cpulist of its PCI function); returns the CPU count, or None when the topology is not exposed"""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = "/sys/bus/pci/devices/%s:%s/local_cpulist" % (dom[-4:].lower(), rest.lower())
        cpus = set()
        for part in open(path).read().strip().split(","):
            if not part:
                continue
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return len(cpus)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    spx = importlib.import_module(PKG)
    synth = importlib.import_module(PKG + ".synth")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or spx.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; this engine has no CPU fallback")
    L = spx.lib()
    numa_cpus = None if args.no_numa_bind else bind_to_gpu_numa_node(local)
    # measured on the 8-GPU box (profiles/r1_e2e_host_sweep_8gpu.log): 2..6 pipeline threads per rank, spinning or sleeping
    # waits, NUMA binding on or off all give the same end-to-end rate -- the box's host I/O path is the limit -- and at one GPU
    # sleeping waits cost 6 %, so the driver's default (spin) stays unless asked otherwise
    blocking = 1 if args.blocking_sync else 0
    L.spx_set_blocking_sync.argtypes = [C.c_int, C.c_int]
    assert L.spx_set_blocking_sync(local, blocking) == 0, L.spx_last_error()
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    shard = importlib.import_module(PKG + ".shard")
    F = args.frames
    mine = shard.frames_for_rank(F, rank, world)          # frame i -> rank i mod world; no data-path collective
    B = len(mine)
    npx = W4K * H4K

    # ---- inputs: the rank's share of the job, every frame in its own device buffer (B * 24.9 MB >> 126 MB L2), already Lab
    seeds_dev = {}
    for sd in sorted({i % JOB_SEEDS for i in mine}):
        t = torch.from_numpy(synth.synth(W4K, H4K, sd)[0]).to(dev)
        rc = L.spx_bgr2lab8_dev(t.data_ptr(), W4K * 3, t.data_ptr(), W4K * 3, W4K, H4K, None)
        assert rc == 0, L.spx_last_error()
        seeds_dev[sd] = t
    torch.cuda.synchronize()
    frames_dev = [seeds_dev[i % JOB_SEEDS].clone() for i in mine]
    frames_host = []
    n_host = min(B, args.e2e_frames)
    for i in range(n_host):
        frames_host.append(frames_dev[i].cpu().pin_memory())
    labels_dev = [torch.empty((H4K, W4K), dtype=torch.int32, device=dev) for _ in range(B)]
    fptr = [t.data_ptr() for t in frames_dev]
    lptr = [t.data_ptr() for t in labels_dev]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_batch(fp, lp, steps, warmup, sample_clocks):
        """`steps` passes over the given frames through the C-ABI batch call: (seconds = max over ranks, launches of all ranks, wall, clocks)"""
        n_l = [0]

        def step():
            n_l[0] += spx.slic_batch_dev(fp, W4K, H4K, 3, spx.SLIC, REGION, RULER, ITERS, 0, lp, None, False, args.streams)

        for _ in range(warmup):
            step()
        barrier()
        sampler = ClockSampler(local)
        if sample_clocks and rank == 0:
            sampler.start()
        n_l[0] = 0
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()          # blocks until the batch is done on the library's streams
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        ev_ms = e0.elapsed_time(e1)
        clk = sampler.stop() if sample_clocks and rank == 0 else None
        el = torch.tensor([max(ev_ms / 1e3, 0.0)], device=dev, dtype=torch.float64)
        nl = torch.tensor([float(n_l[0])], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(el, op=dist.ReduceOp.MAX)
            dist.all_reduce(nl, op=dist.ReduceOp.SUM)      # kernels launched by all ranks inside the timed region
        return float(el.item()), int(nl.item()), wall, clk

    elapsed, total_launches, wall, clocks = timed_batch(fptr, lptr, args.steps, args.warmup, True)
    value = F * args.steps * npx / 1e6 / elapsed
    # the round-1 line for comparison: fixed work per GPU (weak scaling), 32 frames per GPU and step
    weak = None
    if B >= 32 and not args.no_other_configs:
        w_el, _, _, _ = timed_batch(fptr[:32], lptr[:32], 10, 3, False)
        weak = {"value": world * 32 * 10 * npx / 1e6 / w_el, "unit": "MP/s", "frames_per_gpu_per_step": 32, "ms_per_step": w_el / 10 * 1e3,
                "scaling": "weak"}

    # ---- e2e: host buffers in, host labels out, through the reference-facing calls (H2D + D2H timed)
    n_thr = min(args.e2e_threads, n_host)
    host_labels = [torch.empty((H4K, W4K), dtype=torch.int32).pin_memory() for _ in range(n_thr)]
    objs = [None] * n_thr

    def e2e_worker(t, nrep):
        torch.cuda.set_device(local)
        for r in range(nrep):
            for i in range(t, n_host, n_thr):
                f = frames_host[i]
                if objs[t] is None:
                    h = C.c_void_p(None)
                    rc = L.spx_slic_create(f.data_ptr(), W4K * 3, W4K, H4K, 3, spx.SLIC, REGION, RULER, local, C.byref(h))
                    assert rc == 0, L.spx_last_error()
                    objs[t] = h
                else:
                    assert L.spx_slic_reset(objs[t], f.data_ptr(), W4K * 3) == 0, L.spx_last_error()
                assert L.spx_slic_iterate(objs[t], ITERS) == 0, L.spx_last_error()
                assert L.spx_slic_get_labels(objs[t], host_labels[t].data_ptr(), W4K * 4) == 0, L.spx_last_error()

    def e2e_pass(nrep):
        th = [threading.Thread(target=e2e_worker, args=(t, nrep)) for t in range(n_thr)]
        for x in th:
            x.start()
        for x in th:
            x.join()

    e2e_pass(1)
    barrier()
    t0 = time.perf_counter()
    e2e_pass(args.e2e_reps)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = world * n_host * args.e2e_reps * npx / 1e6 / float(e2e_t.item())
    e2e_ref_labels = host_labels[0].numpy().copy()
    for h in objs:
        if h is not None:
            L.spx_slic_destroy(h)

    # ---- roofline of the dominant kernel: time the steady-state iterate(10) of ONE frame on its own stream
    roof = None
    cpu = None
    if rank == 0:
        peak, peak_src = peaks()
        h = C.c_void_p(None)
        st = torch.cuda.Stream(device=dev)
        assert L.spx_slic_create_dev(fptr[0], W4K * 3, W4K, H4K, 3, spx.SLIC, REGION, RULER, C.c_void_p(st.cuda_stream), C.byref(h)) == 0
        times = []
        with torch.cuda.stream(st):
            for r in range(8):
                assert L.spx_slic_reset_dev(h, fptr[r % B], W4K * 3) == 0
                a0 = torch.cuda.Event(enable_timing=True); a1 = torch.cuda.Event(enable_timing=True)
                a0.record(st)
                assert L.spx_slic_iterate(h, ITERS) == 0
                a1.record(st)
                st.synchronize()
                if r >= 3:
                    times.append(a0.elapsed_time(a1))
        it_ms = float(np.median(times))
        # the dominant kernel alone: CUDA events around every assignment launch of iterate(10) run launch by launch on the
        # object's stream (from the seeds, 3 runs + 1 warm-up), L2 flushed with a 256 MB memset before every launch and,
        # for comparison, L2-warm; kernel_ms = mean over the 10 launches
        arr = (C.c_float * ITERS)(); arr_warm = (C.c_float * ITERS)()
        assert L.spx_slic_time_iterations(h, ITERS, 3, 1, arr) == 0, L.spx_last_error()
        assert L.spx_slic_time_iterations(h, ITERS, 3, 0, arr_warm) == 0, L.spx_last_error()
        kms = C.c_float(float(np.mean(list(arr)))); kms_warm = C.c_float(float(np.mean(list(arr_warm))))
        L.spx_slic_destroy(h)
        alg_bytes = 7.0 * npx                      # per assignment launch: 3 B Lab read + 4 B label write per pixel
        achieved = alg_bytes / (kms.value * 1e-3) / 1e9
        iss = ncu_issue() or {}
        roof = {"bound": "hbm", "kernel": "slic_assign (fused assignment + centre accumulation)",
                # the real limiter (instruction issue), from the committed ncu --set full capture of this kernel
                "issue_frac": iss.get("issue_frac"), "issue_source": iss.get("source"),
                "inst_per_px": (iss["warp_inst"] * iss["threads_per_inst"] / npx) if iss else None,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (ncu_traffic() or {}).get("bytes"), "traffic_source": (ncu_traffic() or {}).get("source"),
                "peak_source": peak_src, "kernel_ms": kms.value, "kernel_ms_l2_warm": kms_warm.value,
                "kernel_ms_per_launch": [round(float(v), 5) for v in arr],
                "iterate10_ms_one_frame": it_ms, "assign_share_of_iterate": ITERS * kms_warm.value / it_ms,
                "algorithmic_bytes_per_launch": alg_bytes,
                "batch_frac": (7.0 * ITERS * npx) * (value * 1e6 / npx) / 1e9 / peak}
        # ---- cpu baseline: oracle restatement, bounded sample (2 frames), all host threads
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        O.set_threads(cores)
        i_last = max(i for i in range(0, n_host, n_thr))          # the frame whose labels thread 0 left in host_labels[0]
        lab0 = frames_host[i_last].numpy()
        t0 = time.perf_counter()
        nb = 2
        for _ in range(nb):
            s = O.SuperpixelSLIC(lab0, O.SLIC, REGION, RULER); s.iterate(ITERS); ol = s.getLabels()
        cdt = (time.perf_counter() - t0) / nb
        cpu = {"value": npx / 1e6 / cdt, "unit": "MP/s", "cores": cores, "kind": "port",
               "sample": "%d x one 4K frame (create + iterate(10) + getLabels), oracle restatement, OpenMP %d threads" % (nb, cores),
               "label_agreement_vs_gpu": float((ol == e2e_ref_labels).mean())}

    other = None
    strip = None
    if world > 1 and not args.no_other_configs and not args.no_strip:
        strip = strip_leg(spx, dist, rank, world, local)                    # all ranks: one 16384^2 frame in row strips
    if rank == 0 and not args.no_other_configs:
        other = other_configs(spx, synth, L, cpu=not args.no_cpu_other)
        if strip:
            other["strip_16k"] = strip

    if rank == 0:
        if other is not None and weak is not None:
            other["weak_32_frames_per_gpu"] = weak
        out = {"metric": METRIC, "value": value, "unit": "MP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": elapsed / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": job_config(F),
               "run": {"frames_per_step": F, "frames_per_gpu_per_step": shard.shard_sizes(F, world), "sharding": "shard.frames_for_rank: frame i -> rank i mod N",
                       "streams": args.streams, "device_frame_bytes_per_gpu": B * npx * 3},
               "frames_per_sec": value * 1e6 / npx,
               "e2e": {"value": e2e_value, "unit": "MP/s", "h2d_bytes_per_step": n_host * npx * 3,
                       "d2h_bytes_per_step": n_host * npx * 4, "frames": n_host, "threads": n_thr,
                       "host": {"cores": os.cpu_count(), "numa_bound_cpus": numa_cpus, "blocking_sync": bool(blocking)}},
               "gpu_launches": total_launches, "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
               "wall_s_timed": wall, "other_configs": other}
        emit(out)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line of the contract goes to the real stdout; everything else a library prints (NCCL's version banner,
    warnings) was redirected to stderr at start-up"""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)                      # fd 1 -> stderr for the rest of the process (C libraries included)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=JOB_FRAMES, help="frames of the whole job per step (dealt to the ranks)")
    ap.add_argument("--streams", type=int, default=8, help="pooled engines (streams) per GPU: 2: 12.9, 4: 14.5, 8: 15.0, 16: 15.0 GP/s")
    ap.add_argument("--e2e-frames", type=int, default=16)
    ap.add_argument("--e2e-reps", type=int, default=4)
    ap.add_argument("--e2e-threads", type=int, default=4, help="pipeline threads per rank for the end-to-end leg")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the rank to the CPUs of its GPU's NUMA node")
    ap.add_argument("--blocking-sync", action="store_true", help="host threads sleep while waiting for the device (spx_set_blocking_sync)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the extra (non-headline) configs on rank 0")
    ap.add_argument("--no-cpu-other", action="store_true", help="other configs without their CPU-oracle timings")
    ap.add_argument("--no-strip", action="store_true", help="N > 1: skip the row-strip leg (one 16384^2 frame over all ranks)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
