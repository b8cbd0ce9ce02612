"""superpixels-segmentation-gui-opencv_b200 -- B200-native superpixel engine (host-side Python mirror).

The product is the C-ABI shared library ``libspx_b200.so`` (hand-written sm_100a CUDA kernels, see
``csrc/`` and ``include/spx_b200.h``).  This module is a thin ctypes binding whose classes mirror the
``cv::ximgproc`` surface the reference GUI calls in ``MainWindow::on_button_compute_clicked``
(mainwindow.cpp:2096-2127): ``cvtColor`` (BGR2Lab / BGR2HSV), ``createSuperpixelSLIC`` /
``createSuperpixelLSC`` / ``createSuperpixelSEEDS`` and the objects' ``iterate``, ``getLabels``,
``getNumberOfSuperpixels``, ``getLabelContourMask``, ``enforceLabelConnectivity``.

There is NO CPU fallback: if the library is missing or no CUDA device is present, calls raise.
The directory name contains '-', so import it with
``importlib.import_module("superpixels-segmentation-gui-opencv_b200")``.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBPATH = os.path.join(_HERE, "libspx_b200.so")
_LIB = None

SLIC, SLICO, MSLIC = 100, 101, 102
COLOR_BGR2Lab, COLOR_BGR2HSV = 44, 40   # cv::ColorConversionCodes values


class SpxError(RuntimeError):
    """Raised for every non-zero status of the C ABI (plays the role of cv::Exception)."""

    def __init__(self, code, msg):
        super().__init__("spx_b200 error %d: %s" % (code, msg))
        self.code = code


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into libspx_b200.so (in-tree)."""
    subprocess.check_call(["make", "-C", os.path.join(_HERE, "csrc"), "-j8"] + ([] if verbose else ["-s"]))
    return _LIBPATH


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(_LIBPATH):
        raise SpxError(-2, "libspx_b200.so is not built (run __graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(_LIBPATH)
    vp, sz, i32, f32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float
    pp = C.POINTER(C.c_void_p)
    pi = C.POINTER(C.c_int)
    L.spx_last_error.restype = C.c_char_p
    sig = {
        "spx_device_count": [pi],
        "spx_release_cached_memory": [],
        "spx_bgr2lab8": [vp, sz, vp, sz, i32, i32, i32],
        "spx_bgr2hsv8": [vp, sz, vp, sz, i32, i32, i32],
        "spx_bgr2lab8_dev": [vp, sz, vp, sz, i32, i32, vp],
        "spx_bgr2hsv8_dev": [vp, sz, vp, sz, i32, i32, vp],
        "spx_slic_create": [vp, sz, i32, i32, i32, i32, i32, f32, i32, pp],
        "spx_slic_create_bgr": [vp, sz, i32, i32, i32, i32, i32, f32, i32, pp],
        "spx_lsc_create_bgr": [vp, sz, i32, i32, i32, i32, f32, i32, pp],
        "spx_slic_create_dev": [vp, sz, i32, i32, i32, i32, i32, f32, vp, pp],
        "spx_slic_reset": [vp, vp, sz],
        "spx_slic_reset_dev": [vp, vp, sz],
        "spx_slic_iterate": [vp, i32],
        "spx_slic_enforce_label_connectivity": [vp, i32],
        "spx_slic_get_number_of_superpixels": [vp, pi],
        "spx_slic_get_labels": [vp, vp, sz],
        "spx_slic_get_labels_dev": [vp, vp, sz],
        "spx_slic_get_label_contour_mask": [vp, vp, sz, i32],
        "spx_slic_get_label_contour_mask_dev": [vp, vp, sz, i32],
        "spx_slic_set_labels": [vp, vp, sz, i32],
        "spx_slic_get_centres": [vp, vp, i32],
        "spx_slic_synchronize": [vp],
        "spx_slic_launch_count": [vp, C.POINTER(C.c_longlong)],
        "spx_slic_time_assign": [vp, i32, i32, C.POINTER(C.c_float)],
        "spx_slic_time_iterations": [vp, i32, i32, i32, C.POINTER(C.c_float)],
        "spx_slic_destroy": [vp],
        "spx_session_create": [i32, i32, i32, pp],
        "spx_session_destroy": [vp],
        "spx_session_upload": [vp, i32, vp, sz],
        "spx_session_download": [vp, i32, vp, sz],
        "spx_session_begin": [vp, vp, sz, vp, sz, vp],
        "spx_session_recolour_grid": [vp, vp],
        "spx_session_eq_mask": [vp, i32, i32, vp, sz],
        "spx_session_which_cell": [vp, i32, i32, i32, vp, i32, pi],
        "spx_session_set_where_eq": [vp, i32, i32, i32, vp, i32, i32],
        "spx_session_new_cell_from_white": [vp, vp, i32, pi, pi],
        "spx_session_max_label": [vp, pi],
        "spx_session_cell_index": [vp, vp, vp, i32, pi],
        "spx_session_compose": [vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, sz, pi],
        "spx_preprocess_8uc3": [vp, sz, vp, sz, i32, i32, i32, i32, i32],
        "spx_session_xml_write": [C.c_char_p, i32, vp, i32, vp, sz, vp, sz, i32, i32],
        "spx_session_xml_read": [C.c_char_p, pi, vp, i32, pi, vp, sz, vp, sz, pi, pi],
        "spx_slic_batch_dev": [pp, i32, i32, i32, i32, i32, i32, f32, i32, i32, pp, pp, i32, i32, C.POINTER(C.c_longlong)],
        "spx_slic_batch_release": [],
        "spx_enforce_label_connectivity": [vp, sz, i32, i32, i32, i32, i32, pi],
        "spx_label_contour_mask": [vp, sz, i32, i32, i32, i32, vp, sz, i32],
        "spx_selftest_division": [f32, i32, i32, pi],
        "spx_lsc_create": [vp, sz, i32, i32, i32, i32, f32, i32, pp],
        "spx_lsc_iterate": [vp, i32],
        "spx_lsc_enforce_label_connectivity": [vp, i32],
        "spx_lsc_get_number_of_superpixels": [vp, pi],
        "spx_lsc_get_labels": [vp, vp, sz],
        "spx_lsc_get_label_contour_mask": [vp, vp, sz, i32],
        "spx_lsc_set_labels": [vp, vp, sz, i32],
        "spx_lsc_get_centres": [vp, vp, i32],
        "spx_lsc_launch_count": [vp, C.POINTER(C.c_longlong)],
        "spx_lsc_set_merge_order": [vp, i32],
        "spx_lsc_destroy": [vp],
        "spx_seeds_create": [i32, i32, i32, i32, i32, i32, i32, i32, i32, pp],
        "spx_seeds_iterate": [vp, vp, sz, i32],
        "spx_seeds_get_number_of_superpixels": [vp, pi],
        "spx_seeds_get_labels": [vp, vp, sz],
        "spx_seeds_get_label_contour_mask": [vp, vp, sz, i32],
        "spx_seeds_launch_count": [vp, C.POINTER(C.c_longlong)],
        "spx_seeds_destroy": [vp],
    }
    for name, args in sig.items():
        f = getattr(L, name)
        f.argtypes = args
        f.restype = i32
    _LIB = L
    return L


def _ck(rc):
    if rc != 0:
        raise SpxError(rc, lib().spx_last_error().decode("utf-8", "replace"))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _img(a):
    a = np.asarray(a)
    if a.dtype != np.uint8:
        raise SpxError(-5, "image must be 8-bit (CV_8U)")
    if a.ndim == 2:
        a = a[:, :, None]
    if a.ndim != 3:
        raise SpxError(-5, "image must be HxWxC")
    return np.ascontiguousarray(a)


def device_count():
    n = C.c_int(0)
    _ck(lib().spx_device_count(C.byref(n)))
    return n.value


def cvtColor(src, code, device=0):
    """cv::cvtColor(src, dst, COLOR_BGR2Lab | COLOR_BGR2HSV) for CV_8UC3 -- mainwindow.cpp:2096-2097."""
    img = _img(src)
    if img.shape[2] != 3:
        raise SpxError(-5, "cvtColor: 3-channel image required")
    out = np.empty_like(img)
    h, w = img.shape[:2]
    if code == COLOR_BGR2Lab:
        _ck(lib().spx_bgr2lab8(_p(img), w * 3, _p(out), w * 3, w, h, device))
    elif code == COLOR_BGR2HSV:
        _ck(lib().spx_bgr2hsv8(_p(img), w * 3, _p(out), w * 3, w, h, device))
    else:
        raise SpxError(-5, "cvtColor: only COLOR_BGR2Lab and COLOR_BGR2HSV are on the path")
    return out


def enforce_label_connectivity(labels, numlabels, min_element_size=25, device=0):
    """Stand-alone enforceLabelConnectivity on an injected label map (bit-exact stage test hook)."""
    lab = np.ascontiguousarray(labels, dtype=np.int32).copy()
    h, w = lab.shape
    k = C.c_int(0)
    _ck(lib().spx_enforce_label_connectivity(_p(lab), w * 4, w, h, int(numlabels), int(min_element_size), device, C.byref(k)))
    return lab, k.value


def label_contour_mask(labels, thick_line=True, seeds_flavour=False, device=0):
    """Stand-alone getLabelContourMask on an injected label map."""
    lab = np.ascontiguousarray(labels, dtype=np.int32)
    h, w = lab.shape
    out = np.empty((h, w), np.uint8)
    _ck(lib().spx_label_contour_mask(_p(lab), w * 4, w, h, int(bool(thick_line)), int(bool(seeds_flavour)), _p(out), w, device))
    return out


def selftest_division(w, n=1 << 22, device=0):
    """test hook: mismatches of the kernel's constant-divisor sequence against the IEEE division"""
    bad = C.c_int(0)
    _ck(lib().spx_selftest_division(float(w), int(n), device, C.byref(bad)))
    return bad.value


class _Superpixel:
    _pfx = None

    def __init__(self):
        self._h = C.c_void_p(None)

    def _f(self, name):
        return getattr(lib(), self._pfx + name)

    def release(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._f("destroy")(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def getNumberOfSuperpixels(self):
        n = C.c_int(0)
        _ck(self._f("get_number_of_superpixels")(self._h, C.byref(n)))
        return n.value

    def getLabels(self):
        out = np.empty((self.h, self.w), np.int32)
        _ck(self._f("get_labels")(self._h, _p(out), self.w * 4))
        return out

    def getLabelContourMask(self, thick_line=True):
        out = np.empty((self.h, self.w), np.uint8)
        _ck(self._f("get_label_contour_mask")(self._h, _p(out), self.w, int(bool(thick_line))))
        return out


class SuperpixelSLIC(_Superpixel):
    """cv::ximgproc::SuperpixelSLIC; construct with createSuperpixelSLIC."""
    _pfx = "spx_slic_"

    def __init__(self, image, algorithm=SLICO, region_size=10, ruler=10.0, device=0, colour_code=None):
        """colour_code = COLOR_BGR2Lab / COLOR_BGR2HSV: `image` is BGR and the GUI's cvtColor (mainwindow.cpp:2096-2097) runs on the
        device right after the upload (spx_slic_create_bgr)"""
        super().__init__()
        img = _img(image)
        self.h, self.w, self.c = img.shape
        if colour_code is None:
            _ck(lib().spx_slic_create(_p(img), self.w * self.c, self.w, self.h, self.c, int(algorithm), int(region_size),
                                      float(ruler), int(device), C.byref(self._h)))
        else:
            if self.c != 3:
                raise SpxError(-5, "createSuperpixelSLIC: the fused colour conversion needs a 3-channel BGR image")
            _ck(lib().spx_slic_create_bgr(_p(img), self.w * 3, self.w, self.h, int(colour_code), int(algorithm), int(region_size),
                                          float(ruler), int(device), C.byref(self._h)))

    def reset(self, image):
        img = _img(image)
        if img.shape != (self.h, self.w, self.c):
            raise SpxError(-5, "reset: geometry differs")
        _ck(lib().spx_slic_reset(self._h, _p(img), self.w * self.c))

    def iterate(self, num_iterations=10):
        _ck(lib().spx_slic_iterate(self._h, int(num_iterations)))

    def enforceLabelConnectivity(self, min_element_size=25):
        _ck(lib().spx_slic_enforce_label_connectivity(self._h, int(min_element_size)))

    def setLabels(self, labels, numlabels=0):
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        _ck(lib().spx_slic_set_labels(self._h, _p(lab), self.w * 4, int(numlabels)))

    def getCentres(self):
        k = self.getNumberOfSuperpixels()
        out = np.empty((k, self.c + 2), np.float32)
        _ck(lib().spx_slic_get_centres(self._h, _p(out), k))
        return out

    def launchCount(self):
        n = C.c_longlong(0)
        _ck(lib().spx_slic_launch_count(self._h, C.byref(n)))
        return n.value


class SuperpixelLSC(_Superpixel):
    """cv::ximgproc::SuperpixelLSC; construct with createSuperpixelLSC."""
    _pfx = "spx_lsc_"

    def __init__(self, image, region_size=10, ratio=0.075, device=0, colour_code=None):
        super().__init__()
        img = _img(image)
        self.h, self.w, self.c = img.shape
        if colour_code is None:
            _ck(lib().spx_lsc_create(_p(img), self.w * self.c, self.w, self.h, self.c, int(region_size), float(ratio),
                                     int(device), C.byref(self._h)))
        else:
            if self.c != 3:
                raise SpxError(-5, "createSuperpixelLSC: the fused colour conversion needs a 3-channel BGR image")
            _ck(lib().spx_lsc_create_bgr(_p(img), self.w * 3, self.w, self.h, int(colour_code), int(region_size), float(ratio),
                                         int(device), C.byref(self._h)))

    def iterate(self, num_iterations=10):
        _ck(lib().spx_lsc_iterate(self._h, int(num_iterations)))

    def enforceLabelConnectivity(self, min_element_size=25):
        _ck(lib().spx_lsc_enforce_label_connectivity(self._h, int(min_element_size)))

    def setLabels(self, labels, numlabels=0):
        lab = np.ascontiguousarray(labels, dtype=np.int32)
        _ck(lib().spx_lsc_set_labels(self._h, _p(lab), self.w * 4, int(numlabels)))

    def setMergeOrder(self, order):
        """1 (default) = hashed visiting order of small regions in enforceLabelConnectivity, 0 = the authors' raster order"""
        _ck(lib().spx_lsc_set_merge_order(self._h, int(order)))

    def getCentres(self):
        """test hook: rows of (2*channels+4 feature means, seed x, seed y); valid before enforceLabelConnectivity"""
        k = self.getNumberOfSuperpixels()
        out = np.empty((k, 2 * self.c + 6), np.float32)
        _ck(lib().spx_lsc_get_centres(self._h, _p(out), k))
        return out

    def launchCount(self):
        n = C.c_longlong(0)
        _ck(lib().spx_lsc_launch_count(self._h, C.byref(n)))
        return n.value


class SuperpixelSEEDS(_Superpixel):
    """cv::ximgproc::SuperpixelSEEDS; construct with createSuperpixelSEEDS."""
    _pfx = "spx_seeds_"

    def __init__(self, image_width, image_height, image_channels, num_superpixels, num_levels, prior=2,
                 histogram_bins=5, double_step=False, device=0):
        super().__init__()
        self.w, self.h, self.c = int(image_width), int(image_height), int(image_channels)
        _ck(lib().spx_seeds_create(self.w, self.h, self.c, int(num_superpixels), int(num_levels), int(prior),
                                   int(histogram_bins), int(bool(double_step)), int(device), C.byref(self._h)))

    def iterate(self, img, num_iterations=4):
        img = _img(img)
        if img.shape != (self.h, self.w, self.c):
            raise SpxError(-215, "iterate: image size/channels differ from the ones given at creation")
        _ck(lib().spx_seeds_iterate(self._h, _p(img), self.w * self.c, int(num_iterations)))

    def getLabelContourMask(self, thick_line=False):
        return super().getLabelContourMask(thick_line)

    def launchCount(self):
        n = C.c_longlong(0)
        _ck(lib().spx_seeds_launch_count(self._h, C.byref(n)))
        return n.value


def createSuperpixelSLIC(image, algorithm=SLICO, region_size=10, ruler=10.0, device=0, colour_code=None):
    """mainwindow.cpp:2105 (colour_code: fold the cvtColor of :2096-2097 into the upload)"""
    return SuperpixelSLIC(image, algorithm, region_size, ruler, device, colour_code)


def createSuperpixelLSC(image, region_size=10, ratio=0.075, device=0, colour_code=None):
    """mainwindow.cpp:2114"""
    return SuperpixelLSC(image, region_size, ratio, device, colour_code)


def createSuperpixelSEEDS(image_width, image_height, image_channels, num_superpixels, num_levels, prior=2,
                          histogram_bins=5, double_step=False, device=0):
    """mainwindow.cpp:2123"""
    return SuperpixelSEEDS(image_width, image_height, image_channels, num_superpixels, num_levels, prior,
                           histogram_bins, double_step, device)


def slic_batch_dev(frame_ptrs, width, height, channels, algorithm, region_size, ruler, num_iterations,
                   min_element_size, label_ptrs, mask_ptrs=None, thick_line=False, n_streams=4):
    """Batch of device-resident frames (BASELINE.json configs[4]); pointers are raw device addresses (ints).
    Returns the number of kernels launched."""
    n = len(frame_ptrs)
    FA = (C.c_void_p * n)(*frame_ptrs)
    LA = (C.c_void_p * n)(*label_ptrs)
    MA = (C.c_void_p * n)(*mask_ptrs) if mask_ptrs is not None else None
    launches = C.c_longlong(0)
    _ck(lib().spx_slic_batch_dev(FA, n, width, height, channels, int(algorithm), int(region_size), float(ruler),
                                 int(num_iterations), int(min_element_size), LA, MA, int(bool(thick_line)),
                                 int(n_streams), C.byref(launches)))
    return launches.value


class Session:
    """The Segmentation tab's working set on the device (include/spx_b200.h, spx_session_*): labels, labels_mask, mask, grid,
    selection and the image, with the passes the GUI runs over them (WhichCell, label list actions, ShowSegmentation)."""
    IMAGE, MASK, GRID, SELECTION, LABELS, LABELS_MASK = 0, 1, 2, 3, 4, 5

    def __init__(self, width, height, device=0):
        self._h = C.c_void_p(None)
        self.w, self.h = int(width), int(height)
        _ck(lib().spx_session_create(self.w, self.h, int(device), C.byref(self._h)))

    def release(self):
        if self._h:
            lib().spx_session_destroy(self._h)
            self._h = C.c_void_p(None)

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    def _colour(self, c):
        return np.ascontiguousarray(np.asarray(c, dtype=np.uint8).reshape(3))

    def upload(self, plane, a):
        a = np.ascontiguousarray(a, dtype=np.uint8 if plane < 4 else np.int32)
        assert a.shape[:2] == (self.h, self.w)
        _ck(lib().spx_session_upload(self._h, int(plane), _p(a), a.strides[0]))

    def download(self, plane):
        out = np.empty((self.h, self.w, 3), np.uint8) if plane < 4 else np.empty((self.h, self.w), np.int32)
        _ck(lib().spx_session_download(self._h, int(plane), _p(out), out.strides[0]))
        return out

    def begin(self, labels, contour_mask, colour):
        labels = np.ascontiguousarray(labels, dtype=np.int32); contour_mask = np.ascontiguousarray(contour_mask, dtype=np.uint8)
        c = self._colour(colour)
        _ck(lib().spx_session_begin(self._h, _p(labels), labels.strides[0], _p(contour_mask), contour_mask.strides[0], _p(c)))

    def recolourGrid(self, colour):
        c = self._colour(colour)
        _ck(lib().spx_session_recolour_grid(self._h, _p(c)))

    def eqMask(self, plane, value):
        out = np.empty((self.h, self.w), np.uint8)
        _ck(lib().spx_session_eq_mask(self._h, int(plane), int(value), _p(out), out.strides[0]))
        return out

    def whichCell(self, x, y, paint, colour=(0, 0, 0), label_id=0):
        c = self._colour(colour)
        cell = C.c_int(-1)
        _ck(lib().spx_session_which_cell(self._h, int(x), int(y), int(bool(paint)), _p(c), int(label_id), C.byref(cell)))
        return cell.value

    def setWhereEq(self, key_plane, value, p3=-1, colour=(0, 0, 0), p1=-1, newval=0):
        c = self._colour(colour)
        _ck(lib().spx_session_set_where_eq(self._h, int(key_plane), int(value), int(p3), _p(c), int(p1), int(newval)))

    def newCellFromWhite(self, colour, label_id):
        c = self._colour(colour)
        cell, hits = C.c_int(-1), C.c_int(0)
        _ck(lib().spx_session_new_cell_from_white(self._h, _p(c), int(label_id), C.byref(cell), C.byref(hits)))
        return cell.value, hits.value

    def maxLabel(self):
        v = C.c_int(0)
        _ck(lib().spx_session_max_label(self._h, C.byref(v)))
        return v.value

    def cellIndex(self):
        k = C.c_int(0)
        _ck(lib().spx_session_cell_index(self._h, None, None, 0, C.byref(k)))
        boxes = np.empty((k.value, 4), np.int32); counts = np.empty(k.value, np.int32)
        _ck(lib().spx_session_cell_index(self._h, _p(boxes), _p(counts), k.value, C.byref(k)))
        return boxes, counts

    def compose(self, viewport, image=None, mask=None, grid=None, selection=None, holes=False):
        """viewport = (x, y, w, h); image / mask / grid = blend slider value 0..100 or None (unchecked);
        selection = dilation size or None.  Returns (disp, holes_count or None)."""
        vx, vy, vw, vh = [int(v) for v in viewport]
        out = np.empty((vh, vw, 3), np.uint8)
        nh = C.c_int(0)
        _ck(lib().spx_session_compose(self._h, vx, vy, vw, vh, int(image is not None), int(image or 0), int(mask is not None), int(mask or 0),
                                      int(grid is not None), int(grid or 0), int(selection is not None), int(selection or 0), int(bool(holes)),
                                      _p(out), out.strides[0], C.byref(nh)))
        return out, (nh.value if holes else None)


class _Label(C.Structure):
    _fields_ = [("id", C.c_int), ("name", C.c_char * 128), ("colour", C.c_uint8 * 3)]


def writeSessionXml(path, grid_colour, label_list, labels, labels_mask):
    """<base>-segmentation-data.xml as the reference GUI writes it (mainwindow.cpp:673-706).
    label_list: [(id, name, (c0, c1, c2)), ...] in list order."""
    labels = np.ascontiguousarray(labels, dtype=np.int32); labels_mask = np.ascontiguousarray(labels_mask, dtype=np.int32)
    assert labels.shape == labels_mask.shape and labels.ndim == 2
    arr = (_Label * max(1, len(label_list)))()
    for i, (lid, name, col) in enumerate(label_list):
        arr[i].id = int(lid); arr[i].name = name.encode("utf-8")[:127]
        for k in range(3):
            arr[i].colour[k] = int(col[k])
    _ck(lib().spx_session_xml_write(str(path).encode(), int(grid_colour), C.cast(arr, C.c_void_p), len(label_list), _p(labels), labels.strides[0],
                                    _p(labels_mask), labels_mask.strides[0], labels.shape[1], labels.shape[0]))


def readSessionXml(path):
    """-> (grid_colour, [(id, name, (c0, c1, c2)), ...], labels, labels_mask)  (mainwindow.cpp:775-820)"""
    g, n, r, c = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    _ck(lib().spx_session_xml_read(str(path).encode(), C.byref(g), None, 0, C.byref(n), None, 0, None, 0, C.byref(r), C.byref(c)))
    arr = (_Label * max(1, n.value))()
    labels = np.empty((r.value, c.value), np.int32); mask = np.empty((r.value, c.value), np.int32)
    _ck(lib().spx_session_xml_read(str(path).encode(), C.byref(g), C.cast(arr, C.c_void_p), n.value, C.byref(n), _p(labels), labels.strides[0],
                                   _p(mask), mask.strides[0], C.byref(r), C.byref(c)))
    lst = [(arr[i].id, arr[i].name.decode("utf-8", "replace"), tuple(int(arr[i].colour[k]) for k in range(3))) for i in range(n.value)]
    return g.value, lst, labels, mask


PRE_EQUALIZE, PRE_NORMALIZE, PRE_COLOR_BALANCE, PRE_GAUSSIAN_BLUR = 1, 2, 4, 8


def preprocess(bgr, equalize=False, normalize=False, color_balance=None, gaussian_blur=False, device=0):
    """on_button_apply_clicked (mainwindow.cpp:1979-2047): equalize -> normalize -> colour balance(percent) -> 3x3 blur"""
    a = np.ascontiguousarray(bgr, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[2] == 3
    out = np.empty_like(a)
    flags = (1 if equalize else 0) | (2 if normalize else 0) | (4 if color_balance is not None else 0) | (8 if gaussian_blur else 0)
    _ck(lib().spx_preprocess_8uc3(_p(a), a.strides[0], _p(out), out.strides[0], a.shape[1], a.shape[0], flags, int(color_balance or 0), int(device)))
    return out
