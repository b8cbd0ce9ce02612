"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against the CPU oracle.
Integer / index stages must be bit-exact; SLIC labels and centres are compared exactly as well
(north-star tolerance: >= 99.9 % label agreement, centres within 1e-3)."""
import os

import ctypes as C

import numpy as np
import pytest

from conftest import GOLDEN, random_label_map

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------- colour conversion
def test_lab_hsv_exhaustive(spx, O):
    g, r = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    for b0 in range(0, 256, 32):
        img = np.concatenate([np.stack([np.full_like(g, b), g, r], -1) for b in range(b0, b0 + 32)], 0)
        img = np.ascontiguousarray(img)
        assert (spx.cvtColor(img, spx.COLOR_BGR2Lab) == O.bgr2lab8(img)).all()
        assert (spx.cvtColor(img, spx.COLOR_BGR2HSV) == O.bgr2hsv8(img)).all()


@pytest.mark.parametrize("w,h", [(1, 1), (3, 5), (17, 9), (1501, 7), (64, 64)])
def test_lab_ragged_sizes(spx, O, w, h):
    img = np.random.default_rng(w * 131 + h).integers(0, 256, (h, w, 3), dtype=np.uint8)
    assert (spx.cvtColor(img, spx.COLOR_BGR2Lab) == O.bgr2lab8(img)).all()
    assert (spx.cvtColor(img, spx.COLOR_BGR2HSV) == O.bgr2hsv8(img)).all()


def test_cvtcolor_matches_cv2(spx, example_bgr):
    import cv2
    assert (spx.cvtColor(example_bgr, spx.COLOR_BGR2Lab) == cv2.cvtColor(example_bgr, cv2.COLOR_BGR2Lab)).all()


# ---------------------------------------------------------------- connectivity / contour on injected label maps
@pytest.mark.parametrize("w,h,cell,noise,p", [(131, 97, 12, 0.03, 25), (64, 64, 5, 0.2, 50), (257, 129, 20, 0.01, 10),
                                               (1, 50, 3, 0.1, 25), (50, 1, 3, 0.1, 25), (300, 200, 40, 0.0, 100),
                                               (33, 31, 2, 0.3, 1)])
def test_connectivity_bit_exact(spx, O, w, h, cell, noise, p):
    lab = random_label_map(np.random.default_rng(w + h + p), h, w, cell, noise)
    k0 = int(lab.max()) + 1
    ref, kr = O.enforce_label_connectivity(lab, k0, p)
    out, kg = spx.enforce_label_connectivity(lab, k0, p)
    assert kg == kr
    assert (out == ref).all()


@pytest.mark.parametrize("thick", [False, True])
@pytest.mark.parametrize("seeds", [False, True])
@pytest.mark.parametrize("w,h,cell,noise", [(131, 97, 12, 0.03), (64, 64, 4, 0.3), (300, 5, 7, 0.1), (2, 2, 1, 0.5)])
def test_contour_bit_exact(spx, O, thick, seeds, w, h, cell, noise):
    lab = random_label_map(np.random.default_rng(w * 7 + h), h, w, cell, noise)
    ref = O.label_contour_mask(lab, thick, seeds_flavour=seeds)
    out = spx.label_contour_mask(lab, thick, seeds_flavour=seeds)
    assert (out == ref).all()


@pytest.mark.parametrize("thick", [False, True])
@pytest.mark.parametrize("w,h,cell,noise", [(9001, 160, 9, 0.05), (16384, 96, 20, 0.02), (16500, 70, 14, 0.1)])
def test_contour_wide_images(spx, O, thick, w, h, cell, noise, monkeypatch):
    """widths above 8192: the 1024-thread band kernel (up to 16384), the persistent sweep kernel beyond; and the sweep kernel
    forced at every width -- it must end on a full verification sweep (a list-only version left pixels unsettled at 16384^2)"""
    lab = random_label_map(np.random.default_rng(w + h), h, w, cell, noise)
    ref = O.label_contour_mask(lab, thick)
    assert (spx.label_contour_mask(lab, thick) == ref).all()
    monkeypatch.setenv("SPX_CONTOUR_SWEEPS", "1")
    assert (spx.label_contour_mask(lab, thick) == ref).all()


def test_contour_single_label(spx):
    assert spx.label_contour_mask(np.zeros((40, 50), np.int32), True).sum() == 0


# ---------------------------------------------------------------- SLIC family vs oracle
def _run_pair(spx, O, img, alg, S, ruler, iters):
    a = O.SuperpixelSLIC(img, alg, S, ruler)
    b = spx.createSuperpixelSLIC(img, alg, S, ruler)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels()
    assert (a.getCentres() == b.getCentres()).all(), "seeds differ"
    a.iterate(iters); b.iterate(iters)
    return a, b


@pytest.mark.parametrize("S,ruler,iters", [(20, 10.0, 10), (16, 30.0, 10), (9, 7.0, 3), (31, 45.0, 4)])
def test_slic_labels_and_centres_exact(spx, O, example_lab, S, ruler, iters):
    img = np.ascontiguousarray(example_lab[250:850, 380:1100])
    a, b = _run_pair(spx, O, img, spx.SLIC, S, ruler, iters)
    la, lb = a.getLabels(), b.getLabels()
    agree = float((la == lb).mean())
    assert agree >= 0.999, agree
    assert agree == 1.0, "labels expected identical, got %.6f" % agree
    ca, cb = a.getCentres(), b.getCentres()
    assert np.abs(ca - cb).max() <= 1e-3
    assert (ca == cb).all()


@pytest.mark.parametrize("w,h,S", [(37, 23, 5), (64, 48, 8), (101, 77, 10), (20, 20, 20), (19, 45, 3)])
def test_slic_ragged_sizes(spx, O, synthmod, w, h, S):
    img, _ = synthmod.synth(w, h, 7)
    a, b = _run_pair(spx, O, img, spx.SLIC, S, 10.0, 4)
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()


@pytest.mark.parametrize("channels", [1, 2, 4])
def test_slic_other_channel_counts(spx, O, channels):
    rng = np.random.default_rng(channels)
    base = np.kron(rng.integers(0, 256, (8, 10, channels)), np.ones((12, 12, 1))).astype(np.int64)
    img = np.clip(base + rng.integers(-6, 7, base.shape), 0, 255).astype(np.uint8)
    a, b = _run_pair(spx, O, img, spx.SLIC, 10, 10.0, 4)
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()


def test_slic_cfg1_full_pipeline(spx, O, example_bgr):
    """BASELINE.json configs[0]: the whole COMPUTE click (mainwindow.cpp:2096-2111) on the example image"""
    lab = spx.cvtColor(example_bgr, spx.COLOR_BGR2Lab)
    a, b = _run_pair(spx, O, lab, spx.SLIC, 20, 10.0, 10)
    assert (a.getLabels() == b.getLabels()).all()
    a.enforceLabelConnectivity(25); b.enforceLabelConnectivity(25)
    assert b.getNumberOfSuperpixels() == a.getNumberOfSuperpixels() == 5529
    assert (a.getLabels() == b.getLabels()).all()
    for thick in (False, True):
        assert (a.getLabelContourMask(thick) == b.getLabelContourMask(thick)).all()


def test_slic_golden_fixture_end_to_end(spx, example_bgr):
    """reference's shipped fixture straight through the CUDA path (no oracle involved)"""
    import cv2
    grid = cv2.imread(os.path.join(GOLDEN, "logo-absurdephoton-segmentation-grid.png"))
    lab = spx.cvtColor(example_bgr, spx.COLOR_BGR2Lab)
    s = spx.createSuperpixelSLIC(lab, spx.SLIC, 16, 30.0)
    s.iterate(10)
    s.enforceLabelConnectivity(25)
    assert s.getNumberOfSuperpixels() == 8924
    d = (s.getLabelContourMask(False) != 0) != (grid != 0).any(-1)
    assert int(d.sum()) == 6308
    box = np.zeros_like(d); box[200:1000, 400:1100] = True
    assert int((d & ~box).sum()) == 0


@pytest.mark.parametrize("alg", [101, 102])
def test_slico_mslic_match_oracle(spx, O, example_lab, alg):
    img = np.ascontiguousarray(example_lab[250:850, 380:1100])
    a, b = _run_pair(spx, O, img, alg, 18, 10.0, 6)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels()
    la, lb = a.getLabels(), b.getLabels()
    assert float((la == lb).mean()) == 1.0
    assert (a.getCentres() == b.getCentres()).all()


@pytest.mark.parametrize("S", [16, 25])
def test_slico_tiled_path_exact(spx, O, example_lab, S):
    """SLICO through the tiled kernel (S >= 14): power-of-two S*S (16) and the general divisor (25, BASELINE configs[1])"""
    img = np.ascontiguousarray(example_lab[200:900, 350:1150])
    a, b = _run_pair(spx, O, img, spx.SLICO, S, 10.0, 7)
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()
    a.iterate(3); b.iterate(3)                      # a second iterate() call restarts the adaptive maxima, as upstream
    assert (a.getLabels() == b.getLabels()).all()


def test_mslic_tiled_path_exact(spx, O, synthmod):
    """MSLIC through the tiled kernel: adaptive windows (up to 2S), distance / adaptk, cluster splits (K grows)"""
    img, _ = synthmod.synth(1280, 720, 4)
    lab = O.bgr2lab8(img)
    a, b = _run_pair(spx, O, lab, spx.MSLIC, 30, 10.0, 6)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels()
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()


def test_slico_cfg2_4k_equals_oracle(spx, O, synthmod):
    """BASELINE.json configs[1]: SLICO 3840x2160, region_size=25, 10 iterations, contour mask"""
    img, _ = synthmod.synth(3840, 2160, 0)
    lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
    a, b = _run_pair(spx, O, lab, spx.SLICO, 25, 10.0, 10)
    la, lb = a.getLabels(), b.getLabels()
    assert float((la == lb).mean()) == 1.0
    assert (a.getCentres() == b.getCentres()).all()
    assert (a.getLabelContourMask(False) == b.getLabelContourMask(False)).all()


def test_mslic_cfg3_8k_equals_oracle(spx, O, synthmod):
    """BASELINE.json configs[2], MSLIC half, at config size: 7680x4320, region_size=30, ruler=10, 10 iterations -- K (the
    adaptive splits grow it), labels and centres against the oracle"""
    img, _ = synthmod.synth(7680, 4320, 0)
    lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
    a, b = _run_pair(spx, O, lab, spx.MSLIC, 30, 10.0, 10)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels() > 256 * 144
    la, lb = a.getLabels(), b.getLabels()
    assert float((la == lb).mean()) == 1.0
    assert (a.getCentres() == b.getCentres()).all()


def test_slic_iterate_twice_equals_oracle(spx, O, synthmod):
    img, _ = synthmod.synth(160, 120, 3)
    a = O.SuperpixelSLIC(img, O.SLIC, 10, 10.0); b = spx.createSuperpixelSLIC(img, spx.SLIC, 10, 10.0)
    for n in (3, 2, 1):
        a.iterate(n); b.iterate(n)
    assert (a.getLabels() == b.getLabels()).all()


def test_slic_reset_reuses_object(spx, O, synthmod):
    i0, _ = synthmod.synth(160, 120, 1)
    i1, _ = synthmod.synth(160, 120, 2)
    b = spx.createSuperpixelSLIC(i0, spx.SLIC, 10, 10.0); b.iterate(5); b.enforceLabelConnectivity(25)
    b.reset(i1); b.iterate(5)
    a = O.SuperpixelSLIC(i1, O.SLIC, 10, 10.0); a.iterate(5)
    assert (a.getLabels() == b.getLabels()).all()


def test_slic_properties_4k(spx, synthmod):
    """BASELINE.json full size: size-independent properties (label range, determinism, connectivity)"""
    import cv2
    img, gt = synthmod.synth(3840, 2160, 0)
    lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
    s = spx.createSuperpixelSLIC(lab, spx.SLIC, 20, 10.0)
    K = s.getNumberOfSuperpixels()
    assert K == 192 * 108
    s.iterate(10)
    L1 = s.getLabels()
    assert L1.min() >= 0 and L1.max() < K
    s2 = spx.createSuperpixelSLIC(lab, spx.SLIC, 20, 10.0); s2.iterate(10)
    assert (s2.getLabels() == L1).all(), "run-to-run determinism"
    s.enforceLabelConnectivity(25)
    k2 = s.getNumberOfSuperpixels()
    L2 = s.getLabels()
    assert k2 == 18597            # SURVEY.md Appendix D known answer (oracle: K 20736 -> 18597)
    assert L2.min() == 0 and L2.max() == k2 - 1
    m = s.getLabelContourMask(False)
    assert set(np.unique(m)) <= {0, 255}
    br = synthmod.boundary_recall(L2, gt); ue = synthmod.undersegmentation_error(L2, gt)
    assert abs(br - 0.9982) < 1e-3 and abs(ue - 0.01543) < 1e-3, (br, ue)


@pytest.mark.parametrize("S,ruler", [(16, 30.0), (20, 10.0), (25, 7.0), (30, 30.0), (17, 3.0), (200, 1.0), (14, 100.0)])
def test_constant_division_is_ieee_exact(spx, S, ruler):
    q = np.float32(S) / np.float32(ruler)
    assert spx.selftest_division(float(q * q)) == 0


def test_slic_4k_equals_oracle(spx, O, synthmod):
    """BASELINE.json metric config at full size: labels and centres identical to the oracle"""
    img, _ = synthmod.synth(3840, 2160, 1)
    lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
    a, b = _run_pair(spx, O, lab, spx.SLIC, 20, 10.0, 10)
    la, lb = a.getLabels(), b.getLabels()
    assert float((la == lb).mean()) == 1.0
    assert (a.getCentres() == b.getCentres()).all()


@pytest.mark.parametrize("alg", ["SLIC", "SLICO"])
def test_throughput_mode_gives_identical_results(spx, O, synthmod, alg):
    """spx_slic_set_throughput_mode: the instruction-lean centre update (what spx_slic_batch_dev's pooled objects run) against the
    oracle and against the default latency-lean kernel"""
    img, _ = synthmod.synth(640, 480, 5)
    lab = O.bgr2lab8(img)
    A = getattr(spx, alg)
    a = O.SuperpixelSLIC(lab, getattr(O, alg), 16, 10.0); a.iterate(6)
    b = spx.createSuperpixelSLIC(lab, A, 16, 10.0)
    L = spx.lib()
    L.spx_slic_set_throughput_mode.argtypes = [C.c_void_p, C.c_int]
    assert L.spx_slic_set_throughput_mode(b._h, 1) == 0, L.spx_last_error()
    b.iterate(6)
    c = spx.createSuperpixelSLIC(lab, A, 16, 10.0); c.iterate(6)
    assert (a.getLabels() == b.getLabels()).all() and (b.getLabels() == c.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all() and (b.getCentres() == c.getCentres()).all()


def test_slic_many_centres_8k_equals_oracle(spx, O, synthmod):
    """7680x4320, S = 20: 82 944 centres -- above the 32 768 from which the warp-cooperative centre update takes over"""
    img, _ = synthmod.synth(7680, 4320, 2)
    lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
    a, b = _run_pair(spx, O, lab, spx.SLIC, 20, 10.0, 3)
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()


@pytest.mark.parametrize("alg,S", [("SLIC", 7), ("SLIC", 9), ("SLICO", 8), ("MSLIC", 11), ("SLIC", 13)])
def test_small_region_sizes_on_the_tiled_path(spx, O, synthmod, alg, S):
    """region sizes 7..13 run on the tiled kernel since round 2 (tiles whose candidate list overflows take the exact per-pixel path)"""
    img, _ = synthmod.synth(600, 400, 11)
    lab = O.bgr2lab8(img)
    a, b = _run_pair(spx, O, lab, getattr(spx, alg), S, 10.0, 5)
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()


# ---------------------------------------------------------------- error behaviour (cv::Exception analogue)
def test_errors(spx):
    img = np.zeros((16, 16, 3), np.uint8)
    with pytest.raises(spx.SpxError):
        spx.createSuperpixelSLIC(img, 99, 8, 10.0)
    with pytest.raises(spx.SpxError):
        spx.createSuperpixelSLIC(img, spx.SLIC, 0, 10.0)
    with pytest.raises(spx.SpxError):
        spx.createSuperpixelSLIC(np.zeros((16, 16, 5), np.uint8), spx.SLIC, 8, 10.0)
    s = spx.createSuperpixelSLIC(img, spx.SLIC, 8, 10.0)
    with pytest.raises(spx.SpxError):
        s.enforceLabelConnectivity(101)
    s.enforceLabelConnectivity(0)        # no-op, as upstream
    assert s.getNumberOfSuperpixels() == 4


# ---------------------------------------------------------------- LSC vs oracle (PARITY UNPINNED upstream; exact vs our restatement)
def _lsc_pair(spx, O, img, S, ratio):
    a = O.SuperpixelLSC(img, S, ratio)
    b = spx.createSuperpixelLSC(img, S, ratio)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels()
    assert (a.getCentres() == b.getCentres()).all(), "seeds / initial centres differ"
    return a, b


@pytest.mark.parametrize("w,h,S,ratio,iters", [(320, 240, 16, 0.075, 6), (203, 157, 10, 0.2, 4), (640, 480, 30, 0.075, 10)])
def test_lsc_labels_and_centres_exact(spx, O, synthmod, w, h, S, ratio, iters):
    bgr, _ = synthmod.synth(w, h, 5)
    lab = O.bgr2lab8(bgr)
    a, b = _lsc_pair(spx, O, lab, S, ratio)
    a.iterate(iters); b.iterate(iters)
    la, lb = a.getLabels(), b.getLabels()
    agree = float((la == lb).mean())
    assert agree >= 0.999, agree            # north-star tolerance
    assert agree == 1.0, agree              # the restatement is order-free, so the match is exact
    ca, cb = a.getCentres(), b.getCentres()
    assert np.abs(ca - cb).max() <= 1e-3
    assert (ca == cb).all()


@pytest.mark.parametrize("w,h,S,ratio,iters", [(333, 251, 16, 0.075, 5), (97, 65, 14, 0.075, 4), (257, 200, 20, 0.3, 6), (64, 33, 15, 0.075, 3),
                                               (1920, 1080, 30, 0.075, 3)])
def test_lsc_tiled_path_exact(spx, O, synthmod, w, h, S, ratio, iters):
    """the tiled assignment kernel (3 channels, steps >= 14): ragged tile grids, odd widths (scalar label stores), image
    borders (pixels no window covers keep their label), against the oracle"""
    bgr, _ = synthmod.synth(w, h, 9)
    lab = O.bgr2lab8(bgr)
    a, b = _lsc_pair(spx, O, lab, S, ratio)
    a.iterate(iters); b.iterate(iters)
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()


def test_lsc_tiled_equals_per_pixel_kernel_8k(spx, synthmod):
    """cfg3 size (7680x4320, S=30): the tiled kernel against the per-pixel gather kernel (SPX_LSC_NO_TILE=1), bit for bit"""
    bgr, _ = synthmod.synth(7680, 4320, 0)
    lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
    a = spx.createSuperpixelLSC(lab, 30, 0.075)
    a.iterate(3)
    os.environ["SPX_LSC_NO_TILE"] = "1"
    try:
        b = spx.createSuperpixelLSC(lab, 30, 0.075)
        b.iterate(3)
    finally:
        del os.environ["SPX_LSC_NO_TILE"]
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()


def test_lsc_cfg3_8k_equals_oracle(spx, O, synthmod):
    """BASELINE.json configs[2], LSC half, at config size: 7680x4320, region_size=30, ratio=0.075, 10 iterations, then
    enforceLabelConnectivity(25) in the default (= the authors') merge order and the contour mask, all against the oracle"""
    img, _ = synthmod.synth(7680, 4320, 0)
    lab = spx.cvtColor(img, spx.COLOR_BGR2Lab)
    a, b = _lsc_pair(spx, O, lab, 30, 0.075)
    a.iterate(10); b.iterate(10)
    la, lb = a.getLabels(), b.getLabels()
    assert float((la == lb).mean()) == 1.0
    assert (a.getCentres() == b.getCentres()).all()
    a.enforceLabelConnectivity(25); b.enforceLabelConnectivity(25)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels()
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getLabelContourMask(False) == b.getLabelContourMask(False)).all()


@pytest.mark.parametrize("order", [1, 0])
@pytest.mark.parametrize("p", [25, 5, 60])
def test_lsc_full_pipeline_exact(spx, O, example_lab, p, order):
    """order 0 = the authors' raster order of the small regions (default), order 1 = hashed visiting order (opt-in);
    either way the CUDA path must reproduce the oracle running the same order bit for bit"""
    img = np.ascontiguousarray(example_lab[300:780, 420:1060])
    a, b = _lsc_pair(spx, O, img, 20, 0.075)
    a.setMergeOrder(order); b.setMergeOrder(order)
    a.iterate(8); b.iterate(8)
    assert (a.getLabels() == b.getLabels()).all()
    a.enforceLabelConnectivity(p); b.enforceLabelConnectivity(p)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels()
    assert (a.getLabels() == b.getLabels()).all()
    for thick in (False, True):
        assert (a.getLabelContourMask(thick) == b.getLabelContourMask(thick)).all()


@pytest.mark.parametrize("channels", [1, 2, 4])
def test_lsc_other_channel_counts(spx, O, channels):
    rng = np.random.default_rng(channels + 10)
    base = np.kron(rng.integers(0, 256, (8, 10, channels)), np.ones((12, 12, 1))).astype(np.int64)
    img = np.clip(base + rng.integers(-6, 7, base.shape), 0, 255).astype(np.uint8)
    a, b = _lsc_pair(spx, O, img, 10, 0.1)
    a.iterate(4); b.iterate(4)
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()


@pytest.mark.parametrize("w,h,cell,noise,p,S", [(131, 97, 12, 0.05, 25, 8), (64, 64, 5, 0.2, 50, 8), (257, 129, 20, 0.02, 10, 8),
                                                 (50, 3, 4, 0.1, 25, 2), (1, 40, 3, 0.2, 25, 1)])
@pytest.mark.parametrize("order", [1, 0])
def test_lsc_connectivity_on_injected_labels(spx, O, synthmod, w, h, cell, noise, p, S, order):
    """enforceLabelConnectivity given the same label map: bit-exact (pre-pass adj rule + post-pass merge, both orders)"""
    img, _ = synthmod.synth(w, h, 11)
    lab = random_label_map(np.random.default_rng(w + h + p), h, w, cell, noise)
    k0 = int(lab.max()) + 1
    a = O.SuperpixelLSC(img, S, 0.075); b = spx.createSuperpixelLSC(img, S, 0.075)
    a.setMergeOrder(order); b.setMergeOrder(order)
    a.setLabels(lab, k0); b.setLabels(lab, k0)
    a.enforceLabelConnectivity(p); b.enforceLabelConnectivity(p)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels()
    assert (a.getLabels() == b.getLabels()).all()


def test_lsc_errors(spx):
    with pytest.raises(spx.SpxError):
        spx.createSuperpixelLSC(np.zeros((8, 8, 3), np.uint8), 50, 0.075)     # no seed fits
    with pytest.raises(spx.SpxError):
        spx.createSuperpixelLSC(np.zeros((16, 16, 5), np.uint8), 4, 0.075)


# ---------------------------------------------------------------- SEEDS (PARITY UNPINNED upstream)
# the CUDA path must equal the oracle's phase-parallel order (order 1) bit for bit; order 1 vs upstream's sequential
# order (order 0) is a statistical comparison (tests/test_oracle.py and below)
@pytest.mark.parametrize("w,h,num,levels,prior,bins,dstep,iters", [
    (320, 240, 200, 3, 2, 5, False, 4), (203, 157, 150, 2, 0, 4, True, 3), (256, 256, 64, 4, 5, 3, False, 5),
    (160, 120, 300, 1, 1, 5, False, 2), (640, 360, 400, 4, 3, 5, True, 6),
    (128, 96, 40, 2, 2, 50, False, 2),          # the GUI slider's maximum (mainwindow.ui:2490): 50^3 = 125000 bins per histogram
    (200, 150, 60, 3, 2, 20, True, 3)])
def test_seeds_equals_phase_parallel_oracle(spx, O, synthmod, w, h, num, levels, prior, bins, dstep, iters):
    bgr, _ = synthmod.synth(w, h, 9)
    lab = O.bgr2lab8(bgr)
    a = O.SuperpixelSEEDS(w, h, 3, num, levels, prior, bins, dstep); a.setOrder(1)
    b = spx.createSuperpixelSEEDS(w, h, 3, num, levels, prior, bins, dstep)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels()
    assert (a.getLabels() == b.getLabels()).all()                 # the top-level grid before iterate()
    a.iterate(lab, iters); b.iterate(lab, iters)
    la, lb = a.getLabels(), b.getLabels()
    assert (la == lb).all(), float((la == lb).mean())
    for thick in (False, True):
        assert (a.getLabelContourMask(thick) == b.getLabelContourMask(thick)).all()
    a.iterate(lab, 2); b.iterate(lab, 2)                          # a second COMPUTE click on the same object
    assert (a.getLabels() == b.getLabels()).all()


def test_seeds_cfg4_1080p_equals_phase_parallel_oracle(spx, O, synthmod):
    """BASELINE.json configs[3] at config size (1920x1080, 2000 superpixels, 4 levels, 5 bins, prior 2, 10 iterations):
    labels and both contour flavours bit-exact against the oracle running the GPU's phase-parallel order"""
    bgr, _ = synthmod.synth(1920, 1080, 0)
    lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
    a = O.SuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False); a.setOrder(1)
    b = spx.createSuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)
    a.iterate(lab, 10); b.iterate(lab, 10)
    assert a.getNumberOfSuperpixels() == b.getNumberOfSuperpixels() == 1296
    assert (a.getLabels() == b.getLabels()).all()
    for thick in (False, True):
        assert (a.getLabelContourMask(thick) == b.getLabelContourMask(thick)).all()


def test_seeds_cfg4_vs_sequential_order(spx, O, synthmod):
    """BASELINE.json configs[3]: 1080p, 2000 superpixels, 4 levels, 5 bins, 10 iterations.  Quality of the GPU's
    phase-parallel order against upstream's sequential order on the same image (statistical, not label parity)."""
    bgr, gt = synthmod.synth(1920, 1080, 0)
    lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
    g = spx.createSuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False)
    assert g.getNumberOfSuperpixels() == 1296
    g.iterate(lab, 10)
    Lg = g.getLabels()
    s = O.SuperpixelSEEDS(1920, 1080, 3, 2000, 4, 2, 5, False); s.setOrder(0)
    s.iterate(lab, 10)
    Ls = s.getLabels()
    assert Lg.min() >= 0 and Lg.max() < 1296
    br_g, br_s = synthmod.boundary_recall(Lg, gt), synthmod.boundary_recall(Ls, gt)
    ue_g, ue_s = synthmod.undersegmentation_error(Lg, gt), synthmod.undersegmentation_error(Ls, gt)
    e_g, e_s = s.energy(Lg), s.energy(Ls)
    # Two-sided: the phase-parallel order is a DIFFERENT optimiser order from upstream's Gauss-Seidel sweep, so the labels
    # differ and so do the quality figures -- measured: boundary recall +0.016, undersegmentation error
    # -0.014, energy +0.007.  That is OUTSIDE the north-star's 0.1 % band (on the better side); it is a documented
    # deviation, and the bands below pin it from both sides so that a drift in either direction is caught.
    d_br, d_ue, d_e = br_g - br_s, ue_g - ue_s, e_g - e_s
    print("SEEDS cfg4 phase-parallel minus sequential: dBR %+.4f dUE %+.4f dE %+.4f (BR %.4f/%.4f UE %.4f/%.4f E %.4f/%.4f)"
          % (d_br, d_ue, d_e, br_g, br_s, ue_g, ue_s, e_g, e_s))
    assert -0.005 <= d_br <= 0.03 and -0.03 <= d_ue <= 0.005 and -0.002 <= d_e <= 0.02, (br_g, br_s, ue_g, ue_s, e_g, e_s)
    assert g.launchCount() > 0


def test_seeds_screenshot_sizing(spx):
    """the reference's screenshot: 3888x5184, 1000 requested, 4 levels -> 972 (LCD shows nb_cells+1 = 973)"""
    assert spx.createSuperpixelSEEDS(3888, 5184, 3, 1000, 4).getNumberOfSuperpixels() == 972


def test_seeds_errors(spx):
    with pytest.raises(spx.SpxError):                                    # 8K, 50^3 bins, 8 levels: terabytes of histograms --
        spx.createSuperpixelSEEDS(7680, 4320, 3, 2000, 8, 2, 50, False)  # fails with StsNoMem, as upstream's allocation would
    s = spx.createSuperpixelSEEDS(64, 48, 3, 20, 2)
    with pytest.raises(spx.SpxError):
        s.iterate(np.zeros((48, 60, 3), np.uint8), 2)


def test_cpp_adapter_compute_click():
    """the reference's COMPUTE click compiled against include/spx_ximgproc.hpp: all five algorithms + the HSV branch"""
    import subprocess
    from conftest import ROOT
    cpp = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-C", cpp, "-s"])
    r = subprocess.run([os.path.join(cpp, "compute_click")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("nb_cells") == 6


@pytest.mark.gpu
def test_golden_hashes_on_gpu(spx, example_bgr):
    """the CUDA path reproduces the committed golden hashes byte for byte (no oracle involved)"""
    import importlib.util, json
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(os.path.dirname(__file__), "..", "tools", "gen_golden.py"))
    gg = importlib.util.module_from_spec(spec); spec.loader.exec_module(gg)
    want = json.load(open(gg.OUT))
    got = gg.compute(spx, example_bgr, lambda bgr: spx.cvtColor(bgr, spx.COLOR_BGR2Lab),
                     lambda lab, S, r: spx.createSuperpixelSLIC(lab, spx.SLIC, S, r))
    assert got == want


@pytest.mark.gpu
def test_golden_hashes_other_algorithms_on_gpu(spx, example_bgr):
    """SLICO, MSLIC, LSC and SEEDS reproduce tests/golden/algs_hashes.json byte for byte on the CUDA path (no oracle involved)"""
    import importlib.util, json
    spec = importlib.util.spec_from_file_location("gen_golden_algs", os.path.join(os.path.dirname(__file__), "..", "tools", "gen_golden_algs.py"))
    gg = importlib.util.module_from_spec(spec); spec.loader.exec_module(gg)
    want = json.load(open(gg.OUT))
    got = gg.compute(spx.cvtColor(example_bgr, spx.COLOR_BGR2Lab), lambda lab, alg, S, r: spx.createSuperpixelSLIC(lab, alg, S, r),
                     spx.createSuperpixelLSC, spx.createSuperpixelSEEDS)
    assert got == want


@pytest.mark.gpu
def test_sanity_sweep():
    """ragged / degenerate sizes of all five algorithms (tools/sanity_sweep.py): bit-exact against the oracle or rejected with
    SPX_ERR_BAD_ARG -- never a CUDA error, never a mismatch"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("sanity_sweep", os.path.join(os.path.dirname(__file__), "..", "tools", "sanity_sweep.py"))
    sw = importlib.util.module_from_spec(spec); spec.loader.exec_module(sw)
    ok, rejected, failed = sw.run(sw.SMALL)
    assert not failed, failed[:5]
    assert ok > 100


@pytest.mark.gpu
@pytest.mark.parametrize("tool,cases", [("fuzz_lsc.py", 80), ("fuzz_slic.py", 120)])
def test_randomised_parity(tool, cases):
    """tools/fuzz_lsc.py / tools/fuzz_slic.py: random sizes, region sizes, ratios / rulers, iteration counts and image kinds
    through create -> iterate -> enforceLabelConnectivity -> contour, every stage bit-exact against the oracle"""
    import subprocess, sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "..", "tools", tool), str(cases), "11"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


@pytest.mark.gpu
@pytest.mark.parametrize("code", ["COLOR_BGR2Lab", "COLOR_BGR2HSV"])
def test_create_with_fused_colour_conversion(spx, synthmod, code):
    """spx_slic_create_bgr / spx_lsc_create_bgr == cvtColor on the host side + create (mainwindow.cpp:2096-2105)"""
    bgr, _ = synthmod.synth(333, 211, 5)
    cc = getattr(spx, code)
    conv = spx.cvtColor(bgr, cc)
    for alg, S in ((spx.SLIC, 20), (spx.SLICO, 9)):
        a = spx.createSuperpixelSLIC(conv, alg, S, 10.0); b = spx.createSuperpixelSLIC(bgr, alg, S, 10.0, colour_code=cc)
        a.iterate(4); b.iterate(4)
        assert (a.getLabels() == b.getLabels()).all() and (a.getCentres() == b.getCentres()).all()
    a = spx.createSuperpixelLSC(conv, 15, 0.075); b = spx.createSuperpixelLSC(bgr, 15, 0.075, colour_code=cc)
    a.iterate(3); b.iterate(3)
    assert (a.getLabels() == b.getLabels()).all()
    with pytest.raises(spx.SpxError):
        spx.createSuperpixelSLIC(bgr, spx.SLIC, 20, 10.0, colour_code=7)


@pytest.mark.gpu
def test_blocking_sync_policy_roundtrip(spx, synthmod):
    """spx_set_blocking_sync switches the host waiting policy of the process without disturbing results"""
    import ctypes as C
    L = spx.lib()
    L.spx_set_blocking_sync.argtypes = [C.c_int, C.c_int]
    bgr, _ = synthmod.synth(200, 120, 2)
    lab = spx.cvtColor(bgr, spx.COLOR_BGR2Lab)
    ref = spx.createSuperpixelSLIC(lab, spx.SLIC, 16, 10.0); ref.iterate(3)
    try:
        assert L.spx_set_blocking_sync(0, 1) == 0, L.spx_last_error()
        s = spx.createSuperpixelSLIC(lab, spx.SLIC, 16, 10.0); s.iterate(3)
        assert (s.getLabels() == ref.getLabels()).all()
    finally:
        assert L.spx_set_blocking_sync(0, 0) == 0, L.spx_last_error()
    assert L.spx_set_blocking_sync(99, 1) != 0          # no such device
