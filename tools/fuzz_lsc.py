#!/usr/bin/env python
"""Randomised parity stress of SuperpixelLSC (tiled and per-pixel assignment kernels, connectivity, contour) against the
oracle: random sizes, region sizes, ratios, iteration counts, noise / blocky / flat images.
usage: python tools/fuzz_lsc.py [cases] [seed]"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
from oracle import oracle as O
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
for i in range(n):
    w, h = int(rng.integers(16, 400)), int(rng.integers(16, 260))
    S = int(rng.choice([6, 10, 11, 12, 13, 14, 15, 16, 20, 23, 30, 32, 40, 47, 64]))
    ratio = float(rng.choice([0.02, 0.075, 0.2, 0.5]))
    it = int(rng.integers(1, 6))
    kind = int(rng.integers(0, 3))
    if kind == 0:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    elif kind == 1:
        base = rng.integers(0, 256, ((h + 15) // 16, (w + 15) // 16, 3), dtype=np.uint8)
        img = np.kron(base, np.ones((16, 16, 1), np.uint8))[:h, :w].copy()
        img = np.clip(img.astype(np.int64) + rng.integers(-5, 6, img.shape), 0, 255).astype(np.uint8)
    else:
        img = np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8); img[rng.random((h, w)) < 0.05] = rng.integers(0, 256, 3)
    name = "case %d: %dx%d S %d ratio %g it %d kind %d" % (i, w, h, S, ratio, it, kind)
    try:
        g = spx.createSuperpixelLSC(img, S, ratio)
    except spx.SpxError as e:
        if e.code == -5:
            continue
        print("FAIL", name, str(e)[:100]); bad += 1; continue
    try:
        o = O.SuperpixelLSC(img, S, ratio)
        assert g.getNumberOfSuperpixels() == o.getNumberOfSuperpixels(), "K differs"
        g.iterate(it); o.iterate(it)
        assert (g.getLabels() == o.getLabels()).all(), "labels differ (%d px)" % int((g.getLabels() != o.getLabels()).sum())
        assert (g.getCentres() == o.getCentres()).all(), "centres differ"
        g.enforceLabelConnectivity(25); o.enforceLabelConnectivity(25)
        assert g.getNumberOfSuperpixels() == o.getNumberOfSuperpixels(), "K after connectivity differs"
        assert (g.getLabels() == o.getLabels()).all(), "connectivity differs"
        assert (g.getLabelContourMask(False) == o.getLabelContourMask(False)).all(), "contour differs"
    except (AssertionError, spx.SpxError) as e:
        print("FAIL", name, str(e)[:100]); bad += 1
        if isinstance(e, spx.SpxError) and e.code == -217:
            break
print("%d cases, %d failed" % (n, bad))
sys.exit(1 if bad else 0)
