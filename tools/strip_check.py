#!/usr/bin/env python
"""Row-strip SLIC over N GPUs against the single-GPU object on the same image (must be bit-identical), with timings.
launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/strip_check.py [W H]"""
import importlib, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
strips = importlib.import_module("superpixels-segmentation-gui-opencv_b200.strips")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 4096)
S, RULER, IT = 20, 10.0, 10
bgr, _ = synth.synth(W, H, 3)                               # every rank builds the same image and slices its band
lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab, device=local)
final_mode = {}
post_times = {}
def run(exchange):
    w = strips.StripSLIC(lab, S, RULER, dist, local, exchange=exchange)
    w.iterate(1)                                            # kernel / NCCL / peer-mapping warm-up
    w.release()
    s = strips.StripSLIC(lab, S, RULER, dist, local, exchange=exchange)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter(); s.iterate(IT); torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter() - t0
    out = [s.getLabels(), s.getCentres(), t, s.exchangedBytes(), s.getNumberOfSuperpixels()]
    final_mode[exchange] = s.exchange                        # "neighbour" falls back to "peer" when the drift check fires
    if exchange == "peer":                                  # the rest of the COMPUTE click on the strip object (gathered on rank 0's device)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter(); s.gatherLabels(); torch.cuda.synchronize(); dist.barrier(); tg = time.perf_counter() - t0
        t0 = time.perf_counter(); s._gathered = False; s.gatherLabels(); torch.cuda.synchronize(); dist.barrier(); tg2 = time.perf_counter() - t0
        t0 = time.perf_counter(); s.enforceLabelConnectivity(25); torch.cuda.synchronize(); dist.barrier(); tc = time.perf_counter() - t0
        post_times.update({"gather_first_ms": tg * 1e3, "gather_again_ms": tg2 * 1e3, "connectivity_first_call_ms": tc * 1e3})
        out += [s.getLabels(), s.getLabelContourMask(False), s.getNumberOfSuperpixels(), tc]
        t0 = time.perf_counter(); s.enforceLabelConnectivity(25); torch.cuda.synchronize(); dist.barrier()
        post_times["connectivity_second_call_ms"] = (time.perf_counter() - t0) * 1e3
    s.release()
    return out


os.environ["SPX_STRIP_SPLIT"] = "1"                         # A/B: the tile rows next to the cuts first, interior during the exchange
_, _, tb_split, _, _ = run("neighbour")
del os.environ["SPX_STRIP_SPLIT"]
Lb, Cb, tb, xbb, _ = run("neighbour")                       # partitioned cluster state, boundary sums to rank-1 / rank+1 only
Ls, Cs, t, xb, K, Lc, Mc, Kc, tc = run("peer")                              # replicated state, touched range to every peer
Ln, Cn, tn, xbn, _ = run("nccl")                            # NCCL all-reduce of the accumulators
if rank == 0:
    g = spx.createSuperpixelSLIC(lab, spx.SLIC, S, RULER, device=local)
    spx.lib().spx_slic_synchronize(g._h)
    t1 = time.perf_counter(); g.iterate(IT); spx.lib().spx_slic_synchronize(g._h); t1 = time.perf_counter() - t1
    Lg, Cg = g.getLabels(), g.getCentres()
    g.enforceLabelConnectivity(25)
    post_detail = {"K": bool(Kc == g.getNumberOfSuperpixels()), "labels": float((Lc == g.getLabels()).mean()), "mask": float((Mc == g.getLabelContourMask(False)).mean())}
    post_ok = bool(post_detail["K"] and post_detail["labels"] == 1.0 and post_detail["mask"] == 1.0)
    print(json.dumps({"strip_check": "%dx%d SLIC S=%d %d iterations" % (W, H, S, IT), "n_gpus": world, "bands": strips.row_bands(H, world),
                      "labels_identical": bool((Ls == Lg).all() and (Ln == Lg).all() and (Lb == Lg).all()),
                      "centres_identical": bool((Cs == Cg).all() and (Cn == Cg).all() and (Cb == Cg).all()),
                      "neighbour_mode_ran_as": final_mode.get("neighbour"), "strip_iterate_ms_neighbour": tb * 1e3, "strip_iterate_ms_neighbour_cut_rows_first": tb_split * 1e3, "sent_MB_per_rank_per_iteration_neighbour": xbb / IT / 1e6,
                      "connectivity_and_contour_identical": post_ok, "post_on_rank0": post_times, "post_detail": post_detail, "K_connected": Kc,
                      "strip_iterate_ms": t * 1e3, "strip_iterate_ms_nccl_allreduce": tn * 1e3, "single_gpu_iterate_ms": t1 * 1e3,
                      "exchange": "peer memory (CUDA IPC over NVLink): stores + flags by the library's kernels",
                      "sent_MB_per_rank_per_iteration": xb / IT / 1e6, "allreduce_payload_MB_per_iteration": xbn / IT / 1e6, "K": K}))
    assert (Ls == Lg).all() and (Cs == Cg).all() and (Ln == Lg).all() and (Cn == Cg).all() and post_ok
    assert (Lb == Lg).all() and (Cb == Cg).all()
dist.barrier()
dist.destroy_process_group()
