#!/usr/bin/env python
"""The COMPUTE click for LSC and SEEDS (mainwindow.cpp:2114-2127), call by call, new objects, pageable host buffers.
usage: python tools/click_stages_other.py [W H]"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (3840, 2160)
bgr = synth.synth(W, H, 0)[0]
lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
def report(title, names, runs):
    best = [min(r[i] for r in runs) for i in range(len(names))]
    print("%s %dx%d, best of %d: total %.3f ms" % (title, W, H, len(runs), min(sum(r) for r in runs) * 1e3))
    for n, t in zip(names, best): print("  %-28s %.3f ms" % (n, t * 1e3))
def timed(calls):
    ts = []
    for f in calls:
        t0 = time.perf_counter(); f(); ts.append(time.perf_counter() - t0)
    return ts
runs = []
for rep in range(5):
    box = {}
    runs.append(timed([lambda: box.__setitem__("l", spx.createSuperpixelLSC(lab, 30, 0.075)), lambda: box["l"].iterate(10), lambda: box["l"].enforceLabelConnectivity(25),
                       lambda: box["l"].getNumberOfSuperpixels(), lambda: box["l"].getLabels(), lambda: box["l"].getLabelContourMask(False), lambda: box["l"].release()]))
report("LSC click", ["create", "iterate(10)", "enforceLabelConnectivity", "getNumberOfSuperpixels", "getLabels", "getLabelContourMask", "release"], runs)
runs = []
for rep in range(5):
    box = {}
    runs.append(timed([lambda: box.__setitem__("s", spx.createSuperpixelSEEDS(W, H, 3, 1000, 4, 2, 5, False)), lambda: box["s"].iterate(lab, 10),
                       lambda: box["s"].getNumberOfSuperpixels(), lambda: box["s"].getLabels(), lambda: box["s"].getLabelContourMask(False), lambda: box["s"].release()]))
report("SEEDS click", ["create", "iterate(img, 10)", "getNumberOfSuperpixels", "getLabels", "getLabelContourMask", "release"], runs)
