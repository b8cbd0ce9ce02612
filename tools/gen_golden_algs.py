#!/usr/bin/env python
"""Writes tests/golden/algs_hashes.json: SHA-256 of the outputs of the four algorithms the reference's fixtures do NOT pin
(SLICO, MSLIC, LSC, SEEDS -- SURVEY 8c: parity unpinned upstream) on a 640x480 crop of the reference's example image, computed
with the CPU oracle.  They pin OUR restatement: the oracle and the CUDA path must keep reproducing these bytes, so a change of
either shows up even where no upstream vector exists.
usage: python tools/gen_golden_algs.py [--check]"""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "algs_hashes.json")
CROP = (slice(300, 780), slice(420, 1060))          # rows, columns of the 1500x1500 example image


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _post(s, out, contour_thick):
    s.enforceLabelConnectivity(25)
    out["K_connected"] = int(s.getNumberOfSuperpixels())
    out["labels_connected"] = sha(s.getLabels().astype(np.int32))
    out["contour"] = sha(s.getLabelContourMask(contour_thick))


def compute(lab, slic, lsc, seeds):
    """slic(lab, alg, S, ruler), lsc(lab, S, ratio), seeds(w, h, c, n, levels, prior, bins, double_step) -> objects with the
    reference's call surface (oracle classes or the product's); SLICO = 101, MSLIC = 102"""
    lab = np.ascontiguousarray(lab[CROP])
    h, w = lab.shape[:2]
    out = {"crop": sha(lab)}
    for name, alg, S in (("SLICO_S25", 101, 25), ("MSLIC_S30", 102, 30)):
        s = slic(lab, alg, S, 10.0)
        o = {"K_seeds": int(s.getNumberOfSuperpixels())}
        s.iterate(10)
        o["K_10it"] = int(s.getNumberOfSuperpixels())
        o["labels_10it"] = sha(s.getLabels().astype(np.int32))
        _post(s, o, False)
        out[name] = o
    s = lsc(lab, 30, 0.075)
    o = {"K_seeds": int(s.getNumberOfSuperpixels())}
    s.iterate(10)
    o["labels_10it"] = sha(s.getLabels().astype(np.int32))
    o["centres_10it"] = sha(np.asarray(s.getCentres(), dtype=np.float32))
    _post(s, o, False)
    out["LSC_S30_ratio0.075"] = o
    s = seeds(w, h, 3, 400, 4, 2, 5, False)
    if hasattr(s, "setOrder"):
        s.setOrder(1)                                  # phase-parallel order (what the CUDA path runs)
    o = {"K": int(s.getNumberOfSuperpixels())}
    s.iterate(lab, 10)
    o["labels_10it"] = sha(s.getLabels().astype(np.int32))
    o["contour_thin"] = sha(s.getLabelContourMask(False))
    o["contour_thick"] = sha(s.getLabelContourMask(True))
    out["SEEDS_400sp_4lv_5bins"] = o
    return out


def main():
    import cv2
    from oracle import oracle as O
    O.set_threads(min(8, os.cpu_count() or 1))
    bgr = cv2.imread(os.path.join(ROOT, "tests", "golden", "logo-absurdephoton-segmentation-image.png"), cv2.IMREAD_COLOR)
    got = compute(O.bgr2lab8(bgr), lambda lab, alg, S, r: O.SuperpixelSLIC(lab, alg, S, r), O.SuperpixelLSC, O.SuperpixelSEEDS)
    if "--check" in sys.argv:
        want = json.load(open(OUT))
        assert want == got, "golden hashes changed"
        print("golden hashes reproduce")
    else:
        json.dump(got, open(OUT, "w"), indent=1, sort_keys=True)
        print("wrote", OUT)


if __name__ == "__main__":
    main()
