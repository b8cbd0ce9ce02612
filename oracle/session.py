"""oracle/session.py -- TEST INFRASTRUCTURE ONLY (see oracle/oracle.py): numpy restatement of the passes the reference GUI runs
over its segmentation working set, the checker for csrc/session.cu.

Pinned: every function below is checked against the cv2 4.13 calls the reference makes (tests/test_session.py, CPU).
State = the Mats of mainwindow.h:176-185: labels, labels_mask (int32), image, mask, grid, selection (uint8 x3, BGR order)."""
import numpy as np


def add_weighted8(d, alpha, s, beta):
    """cv::addWeighted(d, alpha, s, beta, 0) on CV_8U (mainwindow.cpp:1941-1951): float32 fma(d, alpha, fl(s*beta)),
    cvRound (half to even), saturate.  float64 holds the exact product + sum, one rounding to float32 = the fma."""
    a, b = np.float32(alpha), np.float32(beta)
    inner = (s.astype(np.float32) * b).astype(np.float32)
    v = (d.astype(np.float64) * np.float64(a) + inner.astype(np.float64)).astype(np.float32)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def dilate_rect(a, d):
    """DilatePixels (mat-image-tools.cpp:235-248): cv::dilate with a (2d+1)^2 rectangle; outside pixels are ignored"""
    if d <= 0:
        return a
    h, w = a.shape[:2]
    out = a.copy()
    for dy in range(-d, d + 1):
        for dx in range(-d, d + 1):
            ys0, ys1 = max(0, dy), min(h, h + dy)
            xs0, xs1 = max(0, dx), min(w, w + dx)
            out[ys0 - dy:ys1 - dy, xs0 - dx:xs1 - dx] = np.maximum(out[ys0 - dy:ys1 - dy, xs0 - dx:xs1 - dx], a[ys0:ys1, xs0:xs1])
    return out


def rgb2gray8(a):
    """cv::cvtColor(COLOR_RGB2GRAY) on CV_8UC3: (R*9798 + G*19235 + B*3735 + 2^14) >> 15, channel 0 = R"""
    c = a.astype(np.int64)
    return ((c[..., 0] * 9798 + c[..., 1] * 19235 + c[..., 2] * 3735 + 16384) >> 15).astype(np.uint8)


class Session:
    def __init__(self, w, h):
        self.w, self.h = w, h
        self.labels = np.zeros((h, w), np.int32); self.labels_mask = np.zeros((h, w), np.int32)
        self.image = np.zeros((h, w, 3), np.uint8); self.mask = np.zeros((h, w, 3), np.uint8)
        self.grid = np.zeros((h, w, 3), np.uint8); self.selection = np.zeros((h, w, 3), np.uint8)

    def _plane(self, p):
        return [self.image, self.mask, self.grid, self.selection, self.labels, self.labels_mask][p]

    def begin(self, labels, contour_mask, colour):
        """mainwindow.cpp:2130-2138"""
        self.labels = labels.astype(np.int32).copy()
        self.labels_mask[:] = 0; self.mask[:] = 0; self.selection[:] = 0
        self.grid = np.repeat(contour_mask[:, :, None], 3, axis=2)           # cvtColor GRAY2RGB
        self.grid[np.any(self.grid != 0, axis=2)] = np.asarray(colour, np.uint8)   # grid.setTo(colour, grid)

    def recolourGrid(self, colour):
        """mainwindow.cpp:1099-1103"""
        self.grid[rgb2gray8(self.grid) > 0] = np.asarray(colour, np.uint8)

    def eqMask(self, plane, value):
        return np.where(self._plane(plane) == value, 255, 0).astype(np.uint8)

    def whichCell(self, x, y, paint, colour=(0, 0, 0), label_id=0):
        """mainwindow.cpp:2153-2171"""
        cell = int(self.labels[y, x])
        m = self.labels == cell
        self.mask[m] = 0; self.labels_mask[m] = 0
        if paint:
            self.mask[m] = np.asarray(colour, np.uint8); self.labels_mask[m] = label_id
        return cell

    def setWhereEq(self, key_plane, value, p3=-1, colour=(0, 0, 0), p1=-1, newval=0):
        m = self._plane(key_plane) == value
        if p3 >= 0:
            self._plane(p3)[m] = np.asarray(colour, np.uint8)
        if p1 >= 0:
            self._plane(p1)[m] = newval

    def newCellFromWhite(self, colour, label_id):
        """mainwindow.cpp:313-331 (the contour drawing stays with the caller)"""
        white = np.all(self.mask == 255, axis=2)
        cell = int(self.labels.max()) + 1
        hits = int(white.sum())
        if hits:
            self.grid[white] = 0; self.labels[white] = cell; self.labels_mask[white] = label_id
            self.mask[white] = np.asarray(colour, np.uint8)
        return cell, hits

    def maxLabel(self):
        return int(self.labels.max())

    def cellIndex(self):
        k = int(self.labels.max()) + 1
        boxes = np.empty((k, 4), np.int32); boxes[:, 0] = self.w; boxes[:, 1] = self.h; boxes[:, 2] = -1; boxes[:, 3] = -1
        ys, xs = np.indices(self.labels.shape)
        l = self.labels.ravel()
        np.minimum.at(boxes[:, 0], l, xs.ravel()); np.minimum.at(boxes[:, 1], l, ys.ravel())
        np.maximum.at(boxes[:, 2], l, xs.ravel()); np.maximum.at(boxes[:, 3], l, ys.ravel())
        return boxes, np.bincount(l, minlength=k).astype(np.int32)

    def compose(self, viewport, image=None, mask=None, grid=None, selection=None, holes=False):
        """ShowSegmentation, mainwindow.cpp:1936-1969"""
        vx, vy, vw, vh = viewport
        crop = lambda a: a[vy:vy + vh, vx:vx + vw]
        disp = np.zeros((vh, vw, 3), np.uint8)
        if image is not None:
            disp = add_weighted8(disp, 1 - float(image) / 100, crop(self.image), float(image) / 100)
        if mask is not None:
            disp = add_weighted8(disp, 1, crop(self.mask), float(mask) / 100)
        if grid is not None:
            disp = add_weighted8(disp, 1, crop(self.grid), float(grid) / 100)
        if selection is not None:
            s = dilate_rect(crop(self.selection), selection)
            disp = np.where(s != 0, s, disp)
        n = None
        if holes:
            hm = np.all(self.mask == 0, axis=2)
            white = np.where(hm[:, :, None], 255, 0).astype(np.uint8)
            disp = add_weighted8(disp, 1, crop(white), 1)
            n = int(hm.sum())
        return disp, n
