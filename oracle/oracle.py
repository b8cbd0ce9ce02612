"""ctypes binding of the CPU ORACLE (oracle/libspx_oracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  The product (superpixels-segmentation-gui-opencv_b200) never does.

The classes mirror cv::ximgproc::SuperpixelSLIC / LSC / SEEDS as the reference GUI drives them
(mainwindow.cpp:2105-2127) so that parity tests read like the reference's call sequence.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SLIC, SLICO, MSLIC = 100, 101, 102


def build(force=False):
    so = os.path.join(_HERE, "libspx_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
        L.spxo_set_threads.argtypes = [i32]
        L.spxo_get_threads.restype = i32
        for f in (L.spxo_bgr2lab8, L.spxo_bgr2hsv8):
            f.argtypes = [vp, sz, vp, sz, i32, i32]
        L.spxo_slic_create.restype = vp
        L.spxo_slic_create.argtypes = [vp, sz, i32, i32, i32, i32, i32, f32]
        L.spxo_slic_destroy.argtypes = [vp]
        L.spxo_slic_iterate.argtypes = [vp, i32]
        L.spxo_slic_enforce_label_connectivity.argtypes = [vp, i32]
        L.spxo_slic_get_number_of_superpixels.argtypes = [vp]
        L.spxo_slic_get_number_of_superpixels.restype = i32
        L.spxo_slic_get_labels.argtypes = [vp, vp]
        L.spxo_slic_set_labels.argtypes = [vp, vp, i32]
        L.spxo_slic_get_centres.argtypes = [vp, vp]
        L.spxo_slic_get_label_contour_mask.argtypes = [vp, vp, i32]
        L.spxo_enforce_label_connectivity.argtypes = [vp, i32, i32, i32, i32]
        L.spxo_enforce_label_connectivity.restype = i32
        L.spxo_label_contour_mask.argtypes = [vp, i32, i32, i32, vp]
        L.spxo_label_contour_mask_seeds.argtypes = [vp, i32, i32, i32, vp]
        if hasattr(L, "spxo_lsc_create"):
            L.spxo_lsc_create.restype = vp
            L.spxo_lsc_create.argtypes = [vp, sz, i32, i32, i32, i32, f32]
            L.spxo_lsc_destroy.argtypes = [vp]
            L.spxo_lsc_iterate.argtypes = [vp, i32]
            L.spxo_lsc_enforce_label_connectivity.argtypes = [vp, i32]
            L.spxo_lsc_get_number_of_superpixels.argtypes = [vp]
            L.spxo_lsc_get_number_of_superpixels.restype = i32
            L.spxo_lsc_get_labels.argtypes = [vp, vp]
            L.spxo_lsc_set_labels.argtypes = [vp, vp, i32]
            L.spxo_lsc_get_centres.argtypes = [vp, vp]
            L.spxo_lsc_set_merge_order.argtypes = [vp, i32]
            L.spxo_lsc_set_merge_algo.argtypes = [vp, i32]
            L.spxo_lsc_merge_iterations.argtypes = [vp]
            L.spxo_lsc_merge_iterations.restype = i32
            L.spxo_lsc_get_label_contour_mask.argtypes = [vp, vp, i32]
        if hasattr(L, "spxo_seeds_create"):
            L.spxo_seeds_create.restype = vp
            L.spxo_seeds_create.argtypes = [i32] * 8
            L.spxo_seeds_destroy.argtypes = [vp]
            L.spxo_seeds_iterate.argtypes = [vp, vp, sz, i32]
            L.spxo_seeds_get_number_of_superpixels.argtypes = [vp]
            L.spxo_seeds_get_number_of_superpixels.restype = i32
            L.spxo_seeds_get_labels.argtypes = [vp, vp]
            L.spxo_seeds_get_label_contour_mask.argtypes = [vp, vp, i32]
            L.spxo_seeds_set_order.argtypes = [vp, i32]
            L.spxo_seeds_energy.argtypes = [vp, vp]
            L.spxo_seeds_energy.restype = C.c_double
        _LIB = L
    return _LIB


def set_threads(n):
    lib().spxo_set_threads(int(n))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim == 2:
        a = a[:, :, None]
    return a


def bgr2lab8(img):
    img = _img(img); out = np.empty_like(img)
    lib().spxo_bgr2lab8(_p(img), img.shape[1] * 3, _p(out), img.shape[1] * 3, img.shape[1], img.shape[0])
    return out


def bgr2hsv8(img):
    img = _img(img); out = np.empty_like(img)
    lib().spxo_bgr2hsv8(_p(img), img.shape[1] * 3, _p(out), img.shape[1] * 3, img.shape[1], img.shape[0])
    return out


def enforce_label_connectivity(labels, numlabels, min_element_size=25):
    lab = np.ascontiguousarray(labels, dtype=np.int32).copy()
    h, w = lab.shape
    k = lib().spxo_enforce_label_connectivity(_p(lab), w, h, int(numlabels), int(min_element_size))
    return lab, k


def label_contour_mask(labels, thick_line=True, seeds_flavour=False):
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    h, w = lab.shape
    out = np.empty((h, w), np.uint8)
    f = lib().spxo_label_contour_mask_seeds if seeds_flavour else lib().spxo_label_contour_mask
    f(_p(lab), w, h, int(bool(thick_line)), _p(out))
    return out


class _Base:
    _pfx = None

    def __del__(self):
        if getattr(self, "_h", None):
            getattr(lib(), self._pfx + "destroy")(self._h)
            self._h = None

    def getNumberOfSuperpixels(self):
        return getattr(lib(), self._pfx + "get_number_of_superpixels")(self._h)

    def getLabels(self):
        out = np.empty((self.h, self.w), np.int32)
        getattr(lib(), self._pfx + "get_labels")(self._h, _p(out))
        return out

    def getLabelContourMask(self, thick_line=True):
        out = np.empty((self.h, self.w), np.uint8)
        getattr(lib(), self._pfx + "get_label_contour_mask")(self._h, _p(out), int(bool(thick_line)))
        return out


class SuperpixelSLIC(_Base):
    """createSuperpixelSLIC(image, algorithm, region_size, ruler) -- mainwindow.cpp:2105"""
    _pfx = "spxo_slic_"

    def __init__(self, image, algorithm=SLICO, region_size=10, ruler=10.0):
        image = _img(image)
        self.h, self.w, self.c = image.shape
        self._h = lib().spxo_slic_create(_p(image), self.w * self.c, self.w, self.h, self.c,
                                         int(algorithm), int(region_size), float(ruler))
        if not self._h:
            raise ValueError("spxo_slic_create: bad argument")

    def iterate(self, num_iterations=10):
        lib().spxo_slic_iterate(self._h, int(num_iterations))

    def enforceLabelConnectivity(self, min_element_size=25):
        lib().spxo_slic_enforce_label_connectivity(self._h, int(min_element_size))

    def setLabels(self, labels, numlabels=0):
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        lib().spxo_slic_set_labels(self._h, _p(lab), int(numlabels))

    def getCentres(self):
        out = np.empty((self.getNumberOfSuperpixels(), self.c + 2), np.float32)
        lib().spxo_slic_get_centres(self._h, _p(out))
        return out


class SuperpixelLSC(_Base):
    """createSuperpixelLSC(image, region_size, ratio) -- mainwindow.cpp:2114"""
    _pfx = "spxo_lsc_"

    def __init__(self, image, region_size=10, ratio=0.075):
        image = _img(image)
        self.h, self.w, self.c = image.shape
        self._h = lib().spxo_lsc_create(_p(image), self.w * self.c, self.w, self.h, self.c,
                                        int(region_size), float(ratio))
        if not self._h:
            raise ValueError("spxo_lsc_create: bad argument")

    def iterate(self, num_iterations=10):
        lib().spxo_lsc_iterate(self._h, int(num_iterations))

    def enforceLabelConnectivity(self, min_element_size=25):
        lib().spxo_lsc_enforce_label_connectivity(self._h, int(min_element_size))

    def setLabels(self, labels, numlabels=0):
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        lib().spxo_lsc_set_labels(self._h, _p(lab), int(numlabels))

    def setMergeOrder(self, order):
        """0 (default, here and in the CUDA path) = the authors' raster order of small regions, 1 = hashed order (opt-in)"""
        lib().spxo_lsc_set_merge_order(self._h, int(order))

    def setMergeAlgo(self, algo):
        """0 = the sequential loop (normative); 1 = fixed-point evaluation of the same rule (the CUDA path's form; same output)"""
        lib().spxo_lsc_set_merge_algo(self._h, int(algo))

    def mergeIterations(self):
        return lib().spxo_lsc_merge_iterations(self._h)

    def getCentres(self):
        """K rows of (2C+4 feature means, seed x, seed y)"""
        out = np.empty((self.getNumberOfSuperpixels(), 2 * self.c + 6), np.float32)
        lib().spxo_lsc_get_centres(self._h, _p(out))
        return out


class SuperpixelSEEDS(_Base):
    """createSuperpixelSEEDS(w, h, c, num_superpixels, num_levels, prior, bins, double_step) -- mainwindow.cpp:2123"""
    _pfx = "spxo_seeds_"

    def __init__(self, image_width, image_height, image_channels, num_superpixels, num_levels,
                 prior=2, histogram_bins=5, double_step=False):
        self.w, self.h, self.c = int(image_width), int(image_height), int(image_channels)
        self._h = lib().spxo_seeds_create(self.w, self.h, self.c, int(num_superpixels), int(num_levels),
                                          int(prior), int(histogram_bins), int(bool(double_step)))
        if not self._h:
            raise ValueError("spxo_seeds_create: bad argument")

    def iterate(self, img, num_iterations=4):
        img = _img(img)
        assert img.shape == (self.h, self.w, self.c)
        lib().spxo_seeds_iterate(self._h, _p(img), self.w * self.c, int(num_iterations))

    def setOrder(self, order):
        """0 = upstream's sequential raster order, 1 = the phase-parallel order of the CUDA path"""
        lib().spxo_seeds_set_order(self._h, int(order))

    def energy(self, labels):
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        return lib().spxo_seeds_energy(self._h, _p(lab))
