#!/usr/bin/env python
"""Writes tests/golden/cfg1_hashes.json: SHA-256 of every stage of BASELINE.json configs[0] (the reference's example image,
BGR2Lab -> SLIC S=20 ruler=10 x 10 iterations -> enforceLabelConnectivity(25) -> both contour masks) and of the S=16 ruler=30
fixture run, computed with the CPU oracle (which tests/test_oracle.py pins against the reference's shipped session files and
against cv2).  Regression anchors: the oracle and the CUDA path must keep reproducing these bytes.
usage: python tools/gen_golden.py [--check]"""
import hashlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden", "cfg1_hashes.json")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def stages(mod, lab, S, ruler, make):
    """mod: oracle module or product package (same call surface); make(lab, S, ruler) -> SuperpixelSLIC object"""
    out = {}
    s = make(lab, S, ruler)
    out["K_seeds"] = int(s.getNumberOfSuperpixels())
    s.iterate(10)
    out["labels_10it"] = sha(s.getLabels().astype(np.int32))
    out["centres_10it"] = sha(np.asarray(s.getCentres(), dtype=np.float32))
    s.enforceLabelConnectivity(25)
    out["K_connected"] = int(s.getNumberOfSuperpixels())
    out["labels_connected"] = sha(s.getLabels().astype(np.int32))
    out["contour_thin"] = sha(s.getLabelContourMask(False))
    out["contour_thick"] = sha(s.getLabelContourMask(True))
    return out


def compute(mod, bgr, lab_of, make):
    lab = lab_of(bgr)
    return {"lab": sha(lab), "cfg1_S20_ruler10": stages(mod, lab, 20, 10.0, make), "fixture_S16_ruler30": stages(mod, lab, 16, 30.0, make)}


def main():
    import cv2
    from oracle import oracle as O
    O.set_threads(min(8, os.cpu_count() or 1))
    bgr = cv2.imread(os.path.join(ROOT, "tests", "golden", "logo-absurdephoton-segmentation-image.png"), cv2.IMREAD_COLOR)
    got = compute(O, bgr, O.bgr2lab8, lambda lab, S, r: O.SuperpixelSLIC(lab, O.SLIC, S, r))
    if "--check" in sys.argv:
        want = json.load(open(OUT))
        assert want == got, "golden hashes changed"
        print("golden hashes reproduce")
    else:
        json.dump(got, open(OUT, "w"), indent=1, sort_keys=True)
        print("wrote", OUT)


if __name__ == "__main__":
    main()
