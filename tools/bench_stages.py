#!/usr/bin/env python
"""Times every stage of the hot path on the GPU (through the host-buffer C ABI, wall clock around synchronous calls,
best of N) next to the CPU oracle on the same input, for the configurations of BASELINE.json.  Writes one JSON line
per stage.  usage: python tools/bench_stages.py [--quick] [--no-cpu]"""
import importlib, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
from oracle import oracle as O

quick = "--quick" in sys.argv
cpu = "--no-cpu" not in sys.argv
O.set_threads(os.cpu_count() or 1)
PEAK = 6530.3
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def best(f, n=5):
    t = 1e30
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); t = min(t, time.perf_counter() - t0)
    return t, r


def line(stage, cfg, px, bytes_per_px, t_gpu, t_cpu, extra=None):
    out = {"stage": stage, "config": cfg, "gpu_ms": t_gpu * 1e3, "gpu_mpx_s": px / 1e6 / t_gpu,
           "algorithmic_GBs": px * bytes_per_px / t_gpu / 1e9, "frac_of_hbm_peak": px * bytes_per_px / t_gpu / 1e9 / PEAK,
           "cpu_oracle_ms": None if t_cpu is None else t_cpu * 1e3, "cpu_threads": O.lib().spxo_get_threads(),
           "note": "wall clock around synchronous host-buffer C-ABI calls (includes H2D/D2H of that call)"}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)


def run():
    W, H = (1920, 1080) if quick else (3840, 2160)
    bgr, gt = synth.synth(W, H, 0)
    px = W * H
    # Lab conversion
    t, lab = best(lambda: spx.cvtColor(bgr, spx.COLOR_BGR2Lab))
    tc = best(lambda: O.bgr2lab8(bgr), 2)[0] if cpu else None
    line("cvtColor BGR2Lab", "%dx%d" % (W, H), px, 6, t, tc)
    # SLIC family: create + iterate(10) + sync
    for alg, name, S in ((spx.SLIC, "SLIC", 20), (spx.SLICO, "SLICO", 25), (spx.MSLIC, "MSLIC", 30)):
        def g():
            s = spx.createSuperpixelSLIC(lab, alg, S, 10.0); s.iterate(10); s.getNumberOfSuperpixels(); return s
        g()
        t, s = best(g, 3)
        s2 = spx.createSuperpixelSLIC(lab, alg, S, 10.0)
        t_it = best(lambda: (s2.reset(lab), spx.lib().spx_slic_synchronize(s2._h), 0)[2], 1)[0]
        def it_only():
            s2.reset(lab); spx.lib().spx_slic_synchronize(s2._h)
            t0 = time.perf_counter(); s2.iterate(10); spx.lib().spx_slic_synchronize(s2._h); return time.perf_counter() - t0
        t_iter = min(it_only() for _ in range(5))
        tc = None
        if cpu:
            def c():
                o = O.SuperpixelSLIC(lab, alg, S, 10.0); o.iterate(10); return o
            tc = best(c, 1)[0]
        line("%s create+iterate(10)" % name, "%dx%d S=%d" % (W, H, S), px, 70, t, tc, {"iterate10_only_ms": t_iter * 1e3,
             "iterate10_frac_of_hbm_peak": px * 70 / t_iter / 1e9 / PEAK})
        if alg == spx.SLIC:
            # connectivity + contour on the SLIC result
            L0 = s.getLabels(); k0 = s.getNumberOfSuperpixels()
            def conn():
                s.setLabels(L0, k0); t0 = time.perf_counter(); s.enforceLabelConnectivity(25); s.getNumberOfSuperpixels(); return time.perf_counter() - t0
            conn(); t = min(conn() for _ in range(5))
            tc = best(lambda: O.enforce_label_connectivity(L0, k0, 25), 1)[0] if cpu else None
            line("enforceLabelConnectivity(25)", "%dx%d, labels resident" % (W, H), px, 8, t, tc)
            L1 = s.getLabels()
            for thick in (False, True):
                t, _ = best(lambda: s.getLabelContourMask(thick), 3)
                tc = best(lambda: O.label_contour_mask(L1, thick), 1)[0] if cpu else None
                line("getLabelContourMask(thick_line=%s)" % thick, "%dx%d" % (W, H), px, 5, t, tc)
    # LSC at 8K (or 4K when quick)
    W2, H2 = (3840, 2160) if quick else (7680, 4320)
    bgr2, _ = synth.synth(W2, H2, 0)
    lab2 = spx.cvtColor(bgr2, spx.COLOR_BGR2Lab)
    def gl():
        l = spx.createSuperpixelLSC(lab2, 30, 0.075); l.iterate(10); l.getNumberOfSuperpixels(); return l
    gl(); t, l = best(gl, 2)
    tc = None
    if cpu:
        def cl():
            o = O.SuperpixelLSC(lab2, 30, 0.075); o.iterate(10); return o
        tc = best(cl, 1)[0]
    line("LSC create+iterate(10)", "%dx%d S=30 ratio=0.075" % (W2, H2), W2 * H2, 70, t, tc)
    Ll, kl = l.getLabels(), l.getNumberOfSuperpixels()
    l.enforceLabelConnectivity(25)                       # first call: the device memory pool grows (hundreds of MB)
    l.setLabels(Ll, kl)
    t0 = time.perf_counter(); l.enforceLabelConnectivity(25); k = l.getNumberOfSuperpixels(); t = time.perf_counter() - t0
    line("LSC enforceLabelConnectivity(25)", "%dx%d" % (W2, H2), W2 * H2, 8, t, None, {"K_after": k})
    if not quick:
        def gm():
            m = spx.createSuperpixelSLIC(lab2, spx.MSLIC, 30, 10.0); m.iterate(10); m.getNumberOfSuperpixels(); return m
        gm(); t, m = best(gm, 2)
        line("MSLIC create+iterate(10)", "%dx%d S=30" % (W2, H2), W2 * H2, 70, t, None)
    # SEEDS 1080p
    bgr3, _ = synth.synth(1920, 1080, 0)
    lab3 = spx.cvtColor(bgr3, spx.COLOR_BGR2Lab)
    sd = spx.createSuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)
    sd.iterate(lab3, 10)
    t, _ = best(lambda: sd.iterate(lab3, 10), 3)
    tc = None
    if cpu:
        so = O.SuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)
        O.set_threads(1)
        tc = best(lambda: so.iterate(lab3, 10), 1)[0]
    line("SEEDS iterate(img, 10)", "1920x1080 2000 superpixels 4 levels 5 bins", 1920 * 1080, 3 + 20 * 10, t, tc, {"cpu_threads": 1, "launches": sd.launchCount()})


run()
