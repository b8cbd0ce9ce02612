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
s = strips.StripSLIC(lab, S, RULER, dist, local)
s.iterate(1)                                                # NCCL / kernel warm-up (one more iteration than the reference below)
s2 = strips.StripSLIC(lab, S, RULER, dist, local)
s = s2
torch.cuda.synchronize(); dist.barrier()
t0 = time.perf_counter(); s.iterate(IT); torch.cuda.synchronize(); dist.barrier(); t = time.perf_counter() - t0
Ls = s.getLabels()
Cs = s.getCentres()
if rank == 0:
    g = spx.createSuperpixelSLIC(lab, spx.SLIC, S, RULER, device=local)
    spx.lib().spx_slic_synchronize(g._h)
    t1 = time.perf_counter(); g.iterate(IT); spx.lib().spx_slic_synchronize(g._h); t1 = time.perf_counter() - t1
    Lg, Cg = g.getLabels(), g.getCentres()
    print(json.dumps({"strip_check": "%dx%d SLIC S=%d %d iterations" % (W, H, S, IT), "n_gpus": world, "bands": strips.row_bands(H, world),
                      "labels_identical": bool((Ls == Lg).all()), "centres_identical": bool((Cs == Cg).all()),
                      "strip_iterate_ms": t * 1e3, "single_gpu_iterate_ms": t1 * 1e3,
                      "exchanged_MB_per_iteration": s.exchanged_bytes / IT / 1e6, "K": s.getNumberOfSuperpixels()}))
    assert (Ls == Lg).all() and (Cs == Cg).all()
dist.barrier()
dist.destroy_process_group()
