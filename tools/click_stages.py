#!/usr/bin/env python
"""The COMPUTE click of the GUI (mainwindow.cpp:2096-2111) on the reference's example image, call by call, host buffers, a NEW
object per click as the GUI does: where the milliseconds of a click go.  usage: python tools/click_stages.py [W H]"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
if len(sys.argv) > 2:
    bgr = synth.synth(int(sys.argv[1]), int(sys.argv[2]), 0)[0]
else:
    import cv2
    bgr = cv2.imread(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "logo-absurdephoton-segmentation-image.png"))
names = ["cvtColor", "create", "iterate(10)+sync", "enforceLabelConnectivity", "getNumberOfSuperpixels", "getLabels", "getLabelContourMask", "release"]
best = [1e9] * len(names)
tot = 1e9
for rep in range(8):
    ts = []
    t0 = time.perf_counter(); lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); s = spx.createSuperpixelSLIC(lab, spx.SLIC, 20, 10.0); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); s.iterate(10); spx.lib().spx_slic_synchronize(s._h); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); s.enforceLabelConnectivity(25); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); k = s.getNumberOfSuperpixels(); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); L = s.getLabels(); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); m = s.getLabelContourMask(False); ts.append(time.perf_counter() - t0)
    t0 = time.perf_counter(); s.release(); ts.append(time.perf_counter() - t0)
    best = [min(a, b) for a, b in zip(best, ts)]
    tot = min(tot, sum(ts))
print("%dx%d click, best of 8: total %.3f ms" % (bgr.shape[1], bgr.shape[0], tot * 1e3))
for n, t in zip(names, best):
    print("  %-28s %.3f ms" % (n, t * 1e3))
