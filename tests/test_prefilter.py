"""Image pre-processing that feeds the path: oracle/prefilter.py against the cv2 4.13 calls of
on_button_apply_clicked (mainwindow.cpp:1979-2047), and csrc/prefilter.cu against the oracle."""
import numpy as np
import pytest


def _images(rng):
    yield rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    yield rng.integers(40, 200, (64, 200, 3)).astype(np.uint8)
    yield (rng.normal(120, 25, (120, 77, 3))).clip(3, 250).astype(np.uint8)
    yield np.full((9, 11, 3), 77, np.uint8)
    g = np.zeros((33, 47, 3), np.uint8); g[..., 0] = np.arange(47)[None, :] * 5; g[..., 1] = 200; g[..., 2] = np.arange(33)[:, None] * 7
    yield g


def _cv_balance(src, percent):
    import cv2
    half = np.float32(percent) / np.float32(200.0)
    chans = list(cv2.split(src))
    for i in range(3):
        flat = np.sort(chans[i].reshape(1, -1), axis=1)          # cv::sort(flat, flat, SORT_EVERY_ROW + SORT_ASCENDING)
        n = flat.shape[1]
        lo = int(flat[0, min(int(np.floor(np.float64(np.float32(n) * half))), n - 1)])
        hi = int(flat[0, min(int(np.ceil(np.float64(np.float32(n)) * (1.0 - np.float64(half)))), n - 1)])
        c = chans[i].copy(); c[c < lo] = lo; c[c > hi] = hi                                # setTo(lowval, ch < lowval) ...
        chans[i] = cv2.normalize(c, None, 0, 255, cv2.NORM_MINMAX)
    return cv2.merge(chans)


def test_oracle_stages_match_cv2():
    import cv2
    from oracle import prefilter as OP
    rng = np.random.default_rng(11)
    for a in _images(rng):
        assert (OP.gaussian_blur3(a) == cv2.GaussianBlur(a, (3, 3), 0, 0)).all()
        yc = cv2.cvtColor(a, cv2.COLOR_BGR2YCrCb)
        assert (OP.bgr2ycrcb(a) == yc).all() and (OP.ycrcb2bgr(yc) == cv2.cvtColor(yc, cv2.COLOR_YCrCb2BGR)).all()
        assert (OP.equalize_hist(yc[..., 0]) == cv2.equalizeHist(np.ascontiguousarray(yc[..., 0]))).all()
        yc2 = yc.copy(); yc2[..., 0] = cv2.equalizeHist(np.ascontiguousarray(yc[..., 0]))
        assert (OP.equalize_histogram(a) == cv2.cvtColor(yc2, cv2.COLOR_YCrCb2BGR)).all()
        assert (OP.normalize_minmax(a, 1.0, 255.0) == cv2.normalize(a, None, 1, 255, cv2.NORM_MINMAX, -1)).all()
        for pct in (1, 5, 20):
            assert (OP.simplest_color_balance(a, pct) == _cv_balance(a, pct)).all()


def test_oracle_normalize_all_ranges_match_cv2():
    """every (min, max) pair: the float32 convertTo rounding is the delicate part"""
    import cv2
    from oracle import prefilter as OP
    ramp = np.arange(256, dtype=np.uint8)
    for mn in range(0, 256, 7):
        for mx in range(mn, 256, 5):
            a = np.clip(ramp, mn, mx).reshape(16, 16, 1).repeat(3, axis=2)
            for lo, hi in ((1, 255), (0, 255)):
                assert (OP.normalize_minmax(a, float(lo), float(hi)) == cv2.normalize(a, None, lo, hi, cv2.NORM_MINMAX, -1)).all(), (mn, mx, lo, hi)


@pytest.mark.gpu
def test_preprocess_matches_oracle(spx):
    from oracle import prefilter as OP
    rng = np.random.default_rng(12)
    for a in list(_images(rng)) + [rng.integers(0, 256, (1, 1, 3), dtype=np.uint8), rng.integers(0, 256, (1, 40, 3), dtype=np.uint8),
                                   rng.integers(0, 256, (40, 1, 3), dtype=np.uint8), rng.integers(0, 256, (481, 640, 3), dtype=np.uint8)]:
        for kw in (dict(equalize=True), dict(normalize=True), dict(color_balance=1), dict(color_balance=10), dict(gaussian_blur=True),
                   dict(equalize=True, normalize=True, color_balance=2, gaussian_blur=True), dict()):
            assert np.array_equal(spx.preprocess(a, **kw), OP.preprocess(a, **kw)), (a.shape, kw)


@pytest.mark.gpu
def test_preprocess_errors(spx):
    with pytest.raises(spx.SpxError):
        spx.preprocess(np.zeros((4, 4, 3), np.uint8), color_balance=101)
