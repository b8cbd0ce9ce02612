#!/usr/bin/env python
"""SEEDS iterate(img, 10) at 1080p (BASELINE.json configs[3]) for ncu launch lists.  usage: python tools/prof_seeds.py"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
img, _ = synth.synth(1920, 1080, 0)
lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
s = spx.createSuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)
s.iterate(lab, 10)
t = time.time(); s.iterate(lab, 10); n = s.getNumberOfSuperpixels(); t1 = time.time() - t
print("SEEDS 1080p: iterate(10) %.2f ms, K %d, launches %d" % (t1 * 1e3, n, s.launchCount()))
