"""Synthetic frames synth(W, H, seed) -- the input recipe fixed in SURVEY.md section 8(d):
Voronoi partition of W*H/96^2 random sites, one random BGR colour in [32,223]^3 per region, a linear
luminance ramp of +-16, 3x3 Gaussian blur, i.i.d. Gaussian noise sigma=4.  Returns (bgr uint8, gt int32).
The draw order of the RNG calls is part of the recipe."""
import numpy as np


def synth(W, H, seed=0):
    import cv2
    rng = np.random.default_rng(seed)
    M = max(4, W * H // (96 * 96))
    ys = rng.integers(0, H, M)
    xs = rng.integers(0, W, M)
    sites = np.ones((H, W), np.uint8)
    sites[ys, xs] = 0
    gt = cv2.distanceTransformWithLabels(sites, cv2.DIST_L2, 5, labelType=cv2.DIST_LABEL_PIXEL)[1] - 1
    cols = rng.integers(32, 224, (int(gt.max()) + 1, 3)).astype(np.float32)
    img = cols[gt]
    ramp = (np.linspace(-16, 16, W, dtype=np.float32)[None, :] * np.float32(rng.uniform(-1, 1))
            + np.linspace(-16, 16, H, dtype=np.float32)[:, None] * np.float32(rng.uniform(-1, 1)))
    img = img + ramp[:, :, None]
    img = cv2.GaussianBlur(img, (3, 3), 0)
    img = img + rng.normal(0, 4, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8), gt.astype(np.int32)


def boundary_map(labels):
    """pixel with a differing 4-neighbour label (both sides marked)"""
    b = np.zeros(labels.shape, bool)
    d = labels[:, 1:] != labels[:, :-1]
    b[:, 1:] |= d; b[:, :-1] |= d
    d = labels[1:, :] != labels[:-1, :]
    b[1:, :] |= d; b[:-1, :] |= d
    return b


def boundary_recall(sp, gt):
    """BR with 2-px Chebyshev tolerance (5x5 dilation), SURVEY.md 8(d)"""
    import cv2
    gb = boundary_map(gt)
    sb = cv2.dilate(boundary_map(sp).astype(np.uint8), np.ones((5, 5), np.uint8)) > 0
    return float((gb & sb).sum()) / max(1, int(gb.sum()))


def undersegmentation_error(sp, gt):
    """UE = (1/N) sum_G sum_{P cap G != 0} min(|P cap G|, |P \\ G|)  (Neubert-Protzel)"""
    sp = sp.ravel().astype(np.int64); gt = gt.ravel().astype(np.int64)
    ns = int(sp.max()) + 1
    pair = gt * ns + sp
    up, cnt = np.unique(pair, return_counts=True)
    psize = np.bincount(sp, minlength=ns)
    p = up % ns
    return float(np.minimum(cnt, psize[p] - cnt).sum()) / sp.size
