#!/usr/bin/env python
"""Determinism stress of the post stages on one very large label map: enforceLabelConnectivity and both contour masks are run
REPS times on the same input and must reproduce the first result bit for bit (a race shows up as a difference).
usage: python tools/stress_post.py [W H REPS]"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
W, H, REPS = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (16384, 16384, 6)
img, _ = synth.synth(W, H, 3)
lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
s = spx.createSuperpixelSLIC(lab, spx.SLIC, 20, 10.0)
s.iterate(10)
L0 = s.getLabels(); K0 = s.getNumberOfSuperpixels()
ref = None
bad = 0
for r in range(REPS):
    s.setLabels(L0, K0)
    s.enforceLabelConnectivity(25)
    cur = (s.getNumberOfSuperpixels(), s.getLabels(), s.getLabelContourMask(False), s.getLabelContourMask(True))
    if ref is None:
        ref = cur
    else:
        same = [cur[0] == ref[0]] + [bool((a == b).all()) for a, b in zip(cur[1:], ref[1:])]
        if not all(same):
            bad += 1
            print("rep", r, "DIFFERS: K/labels/thin/thick same =", same, "labels agree", float((cur[1] == ref[1]).mean()),
                  "thin agree", float((cur[2] == ref[2]).mean()), "thick agree", float((cur[3] == ref[3]).mean()))
print("%dx%d: %d repetitions, %d differ from the first; K %d -> %d" % (W, H, REPS, bad, K0, ref[0]))
sys.exit(1 if bad else 0)
