#!/usr/bin/env python
"""One 4K frame through the C ABI (Lab -> SLIC S=20 ruler=10 x 10 iterations -> connectivity -> contour):
the command profiled with ncu (see profiles/README.md).  No torch needed."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
alg = int(sys.argv[1]) if len(sys.argv) > 1 else 100
S = int(sys.argv[2]) if len(sys.argv) > 2 else 20
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (3840, 2160)
img, _ = synth.synth(W, H, 0)
lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
s = spx.createSuperpixelSLIC(lab, alg, S, 10.0)
t = time.time(); s.iterate(10); n = s.getNumberOfSuperpixels(); t1 = time.time() - t
s.enforceLabelConnectivity(25)
m = s.getLabelContourMask(False)
print("alg", alg, "K", n, "->", s.getNumberOfSuperpixels(), "iterate(10) wall %.2f ms" % (t1 * 1e3), "contour px", int((m != 0).sum()), "launches", s.launchCount())
