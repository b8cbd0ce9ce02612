#!/usr/bin/env python
"""SEEDS iterate(img, 10) at 1080p (BASELINE.json configs[3]) for ncu launch lists.  usage: python tools/prof_seeds.py"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
img, _ = synth.synth(1920, 1080, 0)
lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
s = spx.createSuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)
ts = []
for _ in range(4):                                   # 1st: direct launches, 2nd: captures the graph, 3rd on: replays it
    t = time.time(); s.iterate(lab, 10); n = s.getNumberOfSuperpixels(); ts.append((time.time() - t) * 1e3)
print("SEEDS 1080p: iterate(10) %.2f ms (first call, direct launches %.2f ms; second, graph capture %.2f ms), K %d, launches %d" % (
    min(ts[2:]), ts[0], ts[1], n, s.launchCount()))
