#!/usr/bin/env python
"""Times the dominant kernel (one SLIC assignment launch at 4K, S=20 ruler=10) and iterate(10) through the C ABI.
No torch needed.  usage: python tools/time_assign.py [alg S ruler W H]"""
import ctypes as C, importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
synth = importlib.import_module("superpixels-segmentation-gui-opencv_b200.synth")
alg = int(sys.argv[1]) if len(sys.argv) > 1 else 100
S = int(sys.argv[2]) if len(sys.argv) > 2 else 20
ruler = float(sys.argv[3]) if len(sys.argv) > 3 else 10.0
W, H = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (3840, 2160)
img, _ = synth.synth(W, H, 0)
lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
s = spx.createSuperpixelSLIC(lab, alg, S, ruler)
s.iterate(10)
L = spx.lib()
for flush in (1, 0):
    ms = C.c_float(0)
    assert L.spx_slic_time_assign(s._h, 20, flush, C.byref(ms)) == 0
    print("assign kernel: %.2f us (%s)" % (ms.value * 1e3, "L2 flushed" if flush else "L2 warm"))
for flush in (1, 0):
    arr = (C.c_float * 10)()
    assert L.spx_slic_time_iterations(s._h, 10, 5, flush, arr) == 0, L.spx_last_error()
    v = [x * 1e3 for x in arr]
    print("assign launches of iterate(10), us (%s): %s  mean %.2f" % ("L2 flushed" if flush else "L2 warm", " ".join("%.1f" % x for x in v), sum(v) / 10))
best = 1e9
for r in range(5):
    s.reset(lab)
    L.spx_slic_synchronize(s._h)
    t0 = time.perf_counter(); s.iterate(10); L.spx_slic_synchronize(s._h); best = min(best, time.perf_counter() - t0)
print("iterate(10): %.1f us wall (best of 5), %d px" % (best * 1e6, W * H))
