#!/usr/bin/env python
"""getLabelContourMask on a very wide label map (default 16384 x 16384: the 1024-thread band kernel) against the persistent
sweep kernel (SPX_CONTOUR_SWEEPS, ends on a full verification sweep) and against the CPU oracle's sequential scan.
usage: python tools/big_contour_check.py [W H]"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
from oracle import oracle as O
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 16384)
img, _ = synth.synth(W, H, 3)
s = spx.createSuperpixelSLIC(spx.cvtColor(img, spx.COLOR_BGR2Lab), spx.SLIC, 20, 10.0)
s.iterate(4); s.enforceLabelConnectivity(25)
L = s.getLabels()
ok = True
for thick in (False, True):
    os.environ.pop("SPX_CONTOUR_SWEEPS", None)
    a = s.getLabelContourMask(thick)
    os.environ["SPX_CONTOUR_SWEEPS"] = "1"
    b = s.getLabelContourMask(thick)
    os.environ.pop("SPX_CONTOUR_SWEEPS", None)
    c = O.label_contour_mask(L, thick)
    same = (bool((a == c).all()), bool((b == c).all()))
    print("thick" if thick else "thin", "band kernel == oracle:", same[0], " sweep kernel == oracle:", same[1], " contour px", int((c != 0).sum()))
    ok &= all(same)
sys.exit(0 if ok else 1)
