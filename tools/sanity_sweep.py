#!/usr/bin/env python
"""Degenerate and ragged sizes of all five algorithms against the oracle (labels, counts, both contour masks): every case
either agrees bit for bit or is rejected with SPX_ERR_BAD_ARG.  usage: python tools/sanity_sweep.py [--small]
tests/test_gpu_parity.py::test_sanity_sweep runs the --small set."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

SIZES = ((1, 1), (2, 3), (5, 1), (33, 31), (64, 64), (100, 37), (257, 129), (640, 481), (99, 65), (130, 70), (161, 33))
SMALL = ((2, 3), (33, 31), (64, 64), (100, 37), (99, 65), (161, 33))


def run(sizes=SIZES, verbose=False):
    spx = importlib.import_module("superpixels-segmentation-gui-opencv_b200")
    from oracle import oracle as O
    ok, rejected, failed = 0, 0, []

    def chk(name, f):
        nonlocal ok, rejected
        try:
            f(); ok += 1
            if verbose: print("ok  ", name, flush=True)
        except spx.SpxError as e:
            if e.code == -5:
                rejected += 1
                if verbose: print("rej ", name, "->", str(e)[:90], flush=True)
            else:
                failed.append((name, str(e)[:120]))
                if verbose: print("FAIL", name, "->", str(e)[:120], flush=True)
        except AssertionError as e:
            failed.append((name, str(e)[:120]))
            if verbose: print("FAIL", name, "->", str(e)[:120], flush=True)

    rng = np.random.default_rng(0)
    for (w, h) in sizes:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for alg in (100, 101, 102):
            for S in (1, 3, 14, 20, 40, 64):
                def f():
                    s = spx.createSuperpixelSLIC(img, alg, S, 10.0); s.iterate(3); s.enforceLabelConnectivity(25)
                    L = s.getLabels(); m = s.getLabelContourMask(False); m2 = s.getLabelContourMask(True)
                    o = O.SuperpixelSLIC(img, alg, S, 10.0); o.iterate(3); o.enforceLabelConnectivity(25)
                    assert o.getNumberOfSuperpixels() == s.getNumberOfSuperpixels(), "count differs"
                    assert (o.getLabels() == L).all(), "labels differ"
                    assert (o.getLabelContourMask(False) == m).all() and (o.getLabelContourMask(True) == m2).all(), "contour differs"
                chk("slic %dx%d alg %d S %d" % (w, h, alg, S), f)
        for S in (2, 8, 30):
            def f():
                l = spx.createSuperpixelLSC(img, S, 0.075); l.iterate(2); l.enforceLabelConnectivity(25); L = l.getLabels(); l.getLabelContourMask(True)
                o = O.SuperpixelLSC(img, S, 0.075); o.iterate(2); o.enforceLabelConnectivity(25)
                assert (o.getLabels() == L).all(), "labels differ"
            chk("lsc %dx%d S %d" % (w, h, S), f)
        for (num, lev) in ((4, 1), (16, 2), (100, 4), (2000, 6)):
            def f():
                s = spx.createSuperpixelSEEDS(w, h, 3, num, lev, 2, 5, True); s.iterate(img, 3); L = s.getLabels()
                s.getLabelContourMask(False); s.getLabelContourMask(True)
                o = O.SuperpixelSEEDS(w, h, 3, num, lev, 2, 5, True); o.setOrder(1); o.iterate(img, 3)
                assert (o.getLabels() == L).all(), "labels differ"
            chk("seeds %dx%d num %d lev %d" % (w, h, num, lev), f)
    return ok, rejected, failed


if __name__ == "__main__":
    ok, rejected, failed = run(SMALL if "--small" in sys.argv else SIZES, verbose=True)
    print("%d agree, %d rejected with SPX_ERR_BAD_ARG, %d FAILED" % (ok, rejected, len(failed)))
    sys.exit(1 if failed else 0)
