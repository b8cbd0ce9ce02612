"""Segmentation working set on the device: csrc/session.cu against oracle/session.py, and the
oracle against the cv2 calls the reference makes (mainwindow.cpp:1936-1969, 2130-2171)."""
import numpy as np
import pytest

from conftest import random_label_map


def _state(rng, w, h):
    labels = random_label_map(rng, h, w, cell=9, noise=0.02)
    contour = np.where(rng.random((h, w)) < 0.15, 255, 0).astype(np.uint8)
    image = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    return labels, contour, image


def _script(S, rng, w, h, labels, contour, image):
    """the same sequence of GUI actions on either implementation; returns everything observable"""
    out = []
    S.begin(labels, contour, (0, 0, 255))
    if hasattr(S, "upload"):
        S.upload(0, image)
    else:
        S.image = image.copy()
    out.append(S.maxLabel())
    for i in range(12):                                    # clicks: paint with two label ids, erase some
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        out.append(S.whichCell(x, y, i % 4 != 3, (10 + 20 * i, 255 - 9 * i, 37 * i % 256), 1 + i % 2))
    out.append(S.eqMask(4, int(labels[h // 2, w // 2])))
    out.append(S.eqMask(5, 1))
    S.setWhereEq(5, 2, 1, (0, 0, 0), -1, 0)                # delete label 2 from the mask (mainwindow.cpp:228-229)
    S.setWhereEq(5, 1, 1, (200, 100, 50), -1, 0)           # colour change of label 1 (1334-1335)
    S.setWhereEq(5, 2, 1, (200, 100, 50), 5, 1)            # join label 2 into label 1 (258-260)
    S.recolourGrid((255, 255, 0))
    return out


def _planes(S):
    if hasattr(S, "download"):
        return [S.download(p) for p in range(6)]
    return [S.image, S.mask, S.grid, S.selection, S.labels, S.labels_mask]


def test_oracle_add_weighted_matches_cv2():
    import cv2
    from oracle import session as OS
    a = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 256, 1); b = a.T.copy()
    for al, be in ((0.37, 0.63), (0.55, 0.45), (1.0, 0.37), (0.0, 0.45), (1.0, 1.0), (0.91, 0.13), (1, 0.99), (1 - 0.35, 0.35)):
        assert (OS.add_weighted8(a, al, b, be) == cv2.addWeighted(a, al, b, be, 0, dtype=-1)).all(), (al, be)


def test_oracle_dilate_gray_compare_match_cv2():
    import cv2
    from oracle import session as OS
    rng = np.random.default_rng(5)
    a = np.where(rng.random((61, 83, 3)) < 0.03, rng.integers(1, 256, (61, 83, 3)), 0).astype(np.uint8)
    for d in (1, 2, 3):
        el = cv2.getStructuringElement(cv2.MORPH_RECT, (2 * d + 1, 2 * d + 1), (-1, -1))
        assert (OS.dilate_rect(a, d) == cv2.dilate(a, el)).all()
    c = rng.integers(0, 256, (64, 64, 3), dtype=np.uint8); c[:8] = rng.integers(0, 3, (8, 64, 3))
    assert (OS.rgb2gray8(c) == cv2.cvtColor(c, cv2.COLOR_RGB2GRAY)).all()


def test_oracle_compose_matches_cv2_sequence():
    """ShowSegmentation restated call by call with cv2 (mainwindow.cpp:1936-1969)"""
    import cv2
    from oracle import session as OS
    rng = np.random.default_rng(9)
    w, h = 97, 71
    labels, contour, image = _state(rng, w, h)
    S = OS.Session(w, h)
    _script(S, rng, w, h, labels, contour, image)
    S.selection[rng.random((h, w)) < 0.01] = (255, 255, 255)
    vx, vy, vw, vh = 11, 7, 60, 50
    bi, bm, bg, dil = 35, 50, 80, 2
    crop = lambda a: a[vy:vy + vh, vx:vx + vw]
    disp = np.zeros((vh, vw, 3), np.uint8)
    disp = cv2.addWeighted(disp, 1 - bi / 100, np.ascontiguousarray(crop(S.image)), bi / 100, 0, dtype=-1)
    disp = cv2.addWeighted(disp, 1, np.ascontiguousarray(crop(S.mask)), bm / 100, 0, dtype=-1)
    disp = cv2.addWeighted(disp, 1, np.ascontiguousarray(crop(S.grid)), bg / 100, 0, dtype=-1)
    el = cv2.getStructuringElement(cv2.MORPH_RECT, (2 * dil + 1, 2 * dil + 1), (-1, -1))
    sel = cv2.dilate(np.ascontiguousarray(crop(S.selection)), el)
    disp = np.where(sel != 0, sel, disp)                    # Mat::copyTo with a 3-channel mask is element-wise
    hm = cv2.inRange(S.mask, (0, 0, 0), (0, 0, 0))
    holes = np.zeros((h, w, 3), np.uint8); holes[hm != 0] = (255, 255, 255)
    disp = cv2.addWeighted(disp, 1, np.ascontiguousarray(crop(holes)), 1, 0, dtype=-1)
    got, n = S.compose((vx, vy, vw, vh), bi, bm, bg, dil, True)
    assert (got == disp).all() and n == cv2.countNonZero(hm)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h", [(97, 71), (640, 481), (33, 5)])
def test_session_matches_oracle(spx, w, h):
    from oracle import session as OS
    rng = np.random.default_rng(w * 1000 + h)
    labels, contour, image = _state(rng, w, h)
    G, R = spx.Session(w, h), OS.Session(w, h)
    og = _script(G, np.random.default_rng(1), w, h, labels, contour, image)
    orr = _script(R, np.random.default_rng(1), w, h, labels, contour, image)
    for a, b in zip(og, orr):
        assert np.array_equal(a, b)
    sel = np.zeros((h, w, 3), np.uint8); sel[rng.random((h, w)) < 0.01] = (255, 255, 255)
    G.upload(3, sel); R.selection = sel.copy()
    for a, b in zip(_planes(G), _planes(R)):
        assert np.array_equal(a, b)
    bg, cg = G.cellIndex(); br, cr = R.cellIndex()
    assert np.array_equal(cg, cr) and np.array_equal(bg[cr > 0], br[cr > 0])
    vp = (w // 7, h // 5, w - w // 3, h - h // 4)
    for kw in (dict(image=35, mask=50, grid=80, selection=2, holes=True), dict(image=100), dict(mask=33, grid=100), dict(image=60, selection=0),
               dict(image=1, mask=99, grid=7, selection=3, holes=True)):
        dg, ng = G.compose(vp, **kw); dr, nr = R.compose(vp, **kw)
        assert np.array_equal(dg, dr) and ng == nr
    # paint pure white over a few cells, then turn the white area into a new cell (mainwindow.cpp:313-331)
    for i in range(3):
        x, y = int(rng.integers(0, w)), int(rng.integers(0, h))
        G.whichCell(x, y, True, (255, 255, 255), 3); R.whichCell(x, y, True, (255, 255, 255), 3)
    assert G.newCellFromWhite((1, 2, 3), 4) == R.newCellFromWhite((1, 2, 3), 4)
    for a, b in zip(_planes(G), _planes(R)):
        assert np.array_equal(a, b)
    assert G.maxLabel() == R.maxLabel()
    G.release()


@pytest.mark.gpu
def test_session_errors(spx):
    s = spx.Session(16, 8)
    with pytest.raises(spx.SpxError):
        s.whichCell(3, 3, True, (1, 2, 3), 1)               # no labels yet
    s.begin(np.zeros((8, 16), np.int32), np.zeros((8, 16), np.uint8), (0, 0, 255))
    with pytest.raises(spx.SpxError):
        s.whichCell(16, 0, False)
    with pytest.raises(spx.SpxError):
        s.compose((10, 0, 10, 4), image=50)
    s.release()
