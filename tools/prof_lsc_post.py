#!/usr/bin/env python
"""LSC enforceLabelConnectivity timing at several sizes, visiting order 0 (default: fixed-point kernel) and 1 (hashed rounds).
usage: python tools/prof_lsc_post.py [check]   (check: compare with the CPU oracle at 1080p)"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
for (W, H, S) in [(1920, 1080, 30), (3840, 2160, 30), (7680, 4320, 30)]:
    img, _ = synth.synth(W, H, 0)
    lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
    for order in (0, 1):
        for rep in range(2):
            l = spx.createSuperpixelLSC(lab, S, 0.075); l.setMergeOrder(order); l.iterate(10); n = l.getNumberOfSuperpixels()
            t = time.time(); l.enforceLabelConnectivity(25); t2 = time.time() - t
        print("LSC %dx%d S %d order %d: connectivity %.2f ms, K %d -> %d" % (W, H, S, order, t2 * 1e3, n, l.getNumberOfSuperpixels()), flush=True)
