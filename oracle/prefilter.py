"""oracle/prefilter.py -- TEST INFRASTRUCTURE ONLY: numpy restatement of the reference's image pre-processing
(MainWindow::on_button_apply_clicked, mainwindow.cpp:1979-2047; mat-image-tools.cpp:184-229), the checker for csrc/prefilter.cu.
Pinned: every function is compared with the cv2 4.13 call the reference makes (tests/test_prefilter.py, CPU)."""
import numpy as np


def gaussian_blur3(a):
    """cv::GaussianBlur(Size(3,3), 0): 1-2-1 taps, fixed point, (sum + 8) >> 4, BORDER_REFLECT_101 (mainwindow.cpp:2031)"""
    h, w = a.shape[:2]
    p = np.pad(a.astype(np.int32), ((1, 1) if h > 1 else (0, 0), (1, 1) if w > 1 else (0, 0), (0, 0)), mode="reflect")
    if h == 1:
        p = np.concatenate([p, p, p], axis=0)
    if w == 1:
        p = np.concatenate([p, p, p], axis=1)
    hs = p[:, :-2] + 2 * p[:, 1:-1] + p[:, 2:]
    v = hs[:-2] + 2 * hs[1:-1] + hs[2:]
    return ((v + 8) >> 4).astype(np.uint8)


def bgr2ycrcb(a):
    b, g, r = [a[..., i].astype(np.int64) for i in range(3)]
    y = (b * 1868 + g * 9617 + r * 4899 + 8192) >> 14
    cr = ((r - y) * 11682 + 128 * 16384 + 8192) >> 14
    cb = ((b - y) * 9241 + 128 * 16384 + 8192) >> 14
    return np.stack([y, np.clip(cr, 0, 255), np.clip(cb, 0, 255)], -1).astype(np.uint8)


def ycrcb2bgr(a):
    y, cr, cb = [a[..., i].astype(np.int64) for i in range(3)]
    b = y + (((cb - 128) * 29049 + 8192) >> 14)
    g = y + (((cb - 128) * -5636 + (cr - 128) * -11698 + 8192) >> 14)
    r = y + (((cr - 128) * 22987 + 8192) >> 14)
    return np.clip(np.stack([b, g, r], -1), 0, 255).astype(np.uint8)


def equalize_hist(y):
    """cv::equalizeHist"""
    hist = np.bincount(y.ravel(), minlength=256)
    i0 = int(np.nonzero(hist)[0][0])
    if hist[i0] == y.size:
        return np.full_like(y, i0)
    scale = np.float32(255.0) / np.float32(y.size - hist[i0])
    lut = np.zeros(256, np.uint8)
    s = 0
    for i in range(i0 + 1, 256):
        s += int(hist[i])
        lut[i] = np.clip(np.rint(np.float32(s) * scale), 0, 255)
    return lut[y]


def equalize_histogram(a):
    """EqualizeHistogram (mat-image-tools.cpp:184-202)"""
    yc = bgr2ycrcb(a)
    yc[..., 0] = equalize_hist(yc[..., 0])
    return ycrcb2bgr(yc)


def normalize_minmax(a, lo, hi):
    """cv::normalize(a, a, lo, hi, NORM_MINMAX): one range for the whole array, then convertTo(scale, shift): float32
    fma(v, scale, shift) in cv2 4.13's dispatched build (verified over every (min, max) pair; the unfused form differs in 12 %
    of them); float64 holds the exact product + sum, one rounding to float32 = the fma"""
    mn, mx = float(a.min()), float(a.max())
    scale = (hi - lo) / (mx - mn) if mx - mn > 2.220446049250313e-16 else 0.0
    shift = lo - mn * scale
    v = (a.astype(np.float64) * np.float64(np.float32(scale)) + np.float64(np.float32(shift))).astype(np.float32)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def simplest_color_balance(a, percent):
    """SimplestColorBalance (mat-image-tools.cpp:204-229); the upper index is clamped to the last element (upstream reads one
    past the end for percent == 0)"""
    half = np.float32(percent) / np.float32(200.0)
    out = np.empty_like(a)
    for c in range(3):
        ch = a[..., c]
        flat = np.sort(ch.ravel())
        n = flat.size
        ilo = int(np.floor(np.float64(np.float32(n) * half)))
        ihi = int(np.ceil(np.float64(np.float32(n)) * (1.0 - np.float64(half))))
        lowval, highval = int(flat[min(ilo, n - 1)]), int(flat[min(ihi, n - 1)])
        out[..., c] = normalize_minmax(np.clip(ch, lowval, highval), 0.0, 255.0)
    return out


def preprocess(a, equalize=False, normalize=False, color_balance=None, gaussian_blur=False):
    """mainwindow.cpp:2028-2031, in the reference's order"""
    if equalize:
        a = equalize_histogram(a)
    if normalize:
        a = normalize_minmax(a, 1.0, 255.0)
    if color_balance is not None:
        a = simplest_color_balance(a, color_balance)
    if gaussian_blur:
        a = gaussian_blur3(a)
    return a
