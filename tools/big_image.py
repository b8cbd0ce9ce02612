#!/usr/bin/env python
"""One very large frame through SLIC (tiled path): index arithmetic, workspace sizes and the post stages at 16384 x 16384
(268 Mpx); checks size-independent properties (label range, every seed index used at most once per component id, contour mask
only where labels differ).  usage: python tools/big_image.py [W H]"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (16384, 16384)
rng = np.random.default_rng(0)
base = rng.integers(0, 256, ((H + 63) // 64, (W + 63) // 64, 3), dtype=np.uint8)
img = np.kron(base, np.ones((64, 64, 1), np.uint8))[:H, :W]
img = np.ascontiguousarray(img) ^ rng.integers(0, 8, (H, W, 3), dtype=np.uint8)
t = time.time(); s = spx.createSuperpixelSLIC(img, spx.SLIC, 20, 10.0); k0 = s.getNumberOfSuperpixels(); t0 = time.time() - t
t = time.time(); s.iterate(10); spx.lib().spx_slic_synchronize(s._h); t1 = time.time() - t
L = s.getLabels()
assert L.min() >= 0 and L.max() < k0, (L.min(), L.max(), k0)
t = time.time(); s.enforceLabelConnectivity(25); t2 = time.time() - t
t = time.time(); s.enforceLabelConnectivity(25); t2b = time.time() - t
k1 = s.getNumberOfSuperpixels(); L2 = s.getLabels()
assert L2.min() >= 0 and L2.max() < k1, (L2.min(), L2.max(), k1)
m = s.getLabelContourMask(False)
d = np.zeros((H, W), bool); d[:, 1:] |= L2[:, 1:] != L2[:, :-1]; d[1:, :] |= L2[1:, :] != L2[:-1, :]; d[:, :-1] |= d[:, 1:].copy(); d[:-1, :] |= d[1:, :].copy()
dd = d.copy(); dd[1:, :] |= d[:-1, :]; dd[:-1, :] |= d[1:, :]; dd[:, 1:] |= d[:, :-1]; dd[:, :-1] |= d[:, 1:]
assert not (m.astype(bool) & ~dd).any(), "contour pixel away from any label change"
print("%dx%d: K %d -> %d, create %.1f ms, iterate(10) %.1f ms (%.0f MP/s), connectivity %.1f ms (again: %.1f ms), contour px %d" % (
    W, H, k0, k1, t0 * 1e3, t1 * 1e3, W * H / 1e6 / t1, t2 * 1e3, t2b * 1e3, int((m != 0).sum())))
