#!/usr/bin/env python
"""One frame through the LSC C ABI (create -> iterate(10) -> enforceLabelConnectivity -> contour): the command profiled
with ncu for the LSC launch lists in profiles/.  usage: python tools/prof_lsc.py [S W H]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
S = int(sys.argv[1]) if len(sys.argv) > 1 else 30
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (7680, 4320)
img, _ = synth.synth(W, H, 0)
lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
t = time.time(); l = spx.createSuperpixelLSC(lab, S, 0.075); t0 = time.time() - t
t = time.time(); l.iterate(10); n = l.getNumberOfSuperpixels(); t1 = time.time() - t
t = time.time(); l.enforceLabelConnectivity(25); t2 = time.time() - t
m = l.getLabelContourMask(False)
print("LSC %dx%d S %d: create %.2f ms, iterate(10) %.2f ms, connectivity %.2f ms, K %d -> %d" % (W, H, S, t0 * 1e3, t1 * 1e3, t2 * 1e3, n, l.getNumberOfSuperpixels()))
