#!/usr/bin/env python
"""connectivity + both contour masks on a 4K SLIC result (for ncu launch lists)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
img, _ = synth.synth(3840, 2160, 0)
lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
s = spx.createSuperpixelSLIC(lab, spx.SLIC, 20, 10.0)
s.iterate(10)
s.enforceLabelConnectivity(25)
m0 = s.getLabelContourMask(False)
m1 = s.getLabelContourMask(True)
print(s.getNumberOfSuperpixels(), int((m0 != 0).sum()), int((m1 != 0).sum()), s.launchCount())
