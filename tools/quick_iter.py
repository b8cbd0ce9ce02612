#!/usr/bin/env python
"""single-frame iterate(10) wall/device time at 4K / 1080p (host-timed around a synchronising call), A/B of env switches"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
import numpy as np
S = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for (W, H) in ((3840, 2160), (1920, 1080)):
    img, _ = synth.synth(W, H, 0)
    lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
    for alg in (100, 101):
        L = spx.lib()
        s2 = spx.createSuperpixelSLIC(lab, alg, S, 10.0)
        ts = []
        for r in range(30):
            s2.reset(lab); L.spx_slic_synchronize(s2._h)
            t = time.perf_counter(); s2.iterate(10); L.spx_slic_synchronize(s2._h); ts.append(time.perf_counter() - t)
        lab_out = s2.getLabels()
        print(W, H, alg, "S", S, "iterate(10) min %.3f ms median %.3f ms" % (min(ts) * 1e3, sorted(ts)[len(ts) // 2] * 1e3), "hash", hash(lab_out.tobytes()) & 0xffffffff)
