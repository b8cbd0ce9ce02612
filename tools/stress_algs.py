#!/usr/bin/env python
"""Run-to-run determinism of LSC (iterate + enforceLabelConnectivity + contour) and SEEDS at sizes beyond the parity tests:
every repetition must reproduce the first bit for bit (the algorithms use grid barriers, work lists and atomics; a race shows up
as a difference).  usage: python tools/stress_algs.py [W H REPS]"""
import hashlib, importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
W, H, REPS = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8192, 8192, 3)
img, _ = synth.synth(W, H, 5)
lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
hh = lambda a: hashlib.sha1(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]
bad = 0
ref = None
for r in range(REPS):
    l = spx.createSuperpixelLSC(lab, 20, 0.075)
    l.iterate(5); a = hh(l.getLabels())
    l.enforceLabelConnectivity(25)
    cur = (a, l.getNumberOfSuperpixels(), hh(l.getLabels()), hh(l.getLabelContourMask(False)), hh(l.getLabelContourMask(True)))
    del l
    if ref is None: ref = cur
    elif cur != ref: bad += 1; print("LSC rep", r, "differs:", [x == y for x, y in zip(cur, ref)])
print("LSC %dx%d: %d repetitions, %d differ; K after connectivity %d" % (W, H, REPS, bad, ref[1]))
ref = None
for r in range(REPS):
    s = spx.createSuperpixelSEEDS(W, H, 3, 20000, 5, 2, 5, True)
    s.iterate(lab, 6)
    cur = (hh(s.getLabels()), hh(s.getLabelContourMask(False)), hh(s.getLabelContourMask(True)))
    del s
    if ref is None: ref = cur
    elif cur != ref: bad += 1; print("SEEDS rep", r, "differs:", [x == y for x, y in zip(cur, ref)])
print("SEEDS %dx%d: %d repetitions" % (W, H, REPS), "ok" if not bad else "%d differences in total" % bad)
sys.exit(1 if bad else 0)
