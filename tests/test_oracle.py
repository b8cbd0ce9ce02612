"""CPU-only: pins the oracle against the reference's own artefacts."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_lab_hsv_match_cv2_sampled(O):
    """full 2^24 exhaustive check runs in tools/gen_color_tables.py; here every 5th level per channel + noise"""
    import cv2
    v = np.arange(0, 256, 5, dtype=np.uint8)
    b, g, r = np.meshgrid(v, v, v, indexing="ij")
    img = np.stack([b, g, r], -1).reshape(52, -1, 3)
    rng = np.random.default_rng(1)
    img = np.concatenate([img, rng.integers(0, 256, (52, 4096, 3), dtype=np.uint8)], 1)
    img = np.ascontiguousarray(img)
    assert (O.bgr2lab8(img) == cv2.cvtColor(img, cv2.COLOR_BGR2Lab)).all()
    assert (O.bgr2hsv8(img) == cv2.cvtColor(img, cv2.COLOR_BGR2HSV)).all()


def test_color_tables_identical_for_oracle_and_product():
    root = os.path.dirname(GOLDEN)
    root = os.path.dirname(root)
    a = open(os.path.join(root, "oracle", "spx_color_tables.h")).read().split("#define", 1)[1].split("\n", 1)[1]
    b = open(os.path.join(root, "superpixels-segmentation-gui-opencv_b200", "csrc", "color_tables.inc")).read().split("#define", 1)[1].split("\n", 1)[1]
    assert a == b


def test_slic_fixture_known_answer(O, example_bgr, example_lab):
    """example/*-grid.png = getLabelContourMask(SLIC S=16 ruler=30 10 it, connectivity 25, thick_line=false)
    plus hand-drawn outlines inside x[400,1100) y[200,1000)  (SURVEY.md section 4, Appendix D)."""
    import cv2
    grid = cv2.imread(os.path.join(GOLDEN, "logo-absurdephoton-segmentation-grid.png"))
    mask = cv2.imread(os.path.join(GOLDEN, "logo-absurdephoton-segmentation-mask.png"))
    s = O.SuperpixelSLIC(example_lab, O.SLIC, 16, 30.0)
    assert s.getNumberOfSuperpixels() == 8836
    s.iterate(10)
    s.enforceLabelConnectivity(25)
    assert s.getNumberOfSuperpixels() == 8924
    m = s.getLabelContourMask(False)
    d = (m != 0) != (grid != 0).any(-1)
    assert int(d.sum()) == 6308
    box = np.zeros_like(d); box[200:1000, 400:1100] = True
    assert int((d & ~box).sum()) == 0
    L = s.getLabels()
    code = (mask[..., 0].astype(np.int64) << 16) | (mask[..., 1].astype(np.int64) << 8) | mask[..., 2]
    uc, inv = np.unique(code, return_inverse=True)
    hist = np.zeros((L.max() + 1, len(uc)), np.int64)
    np.add.at(hist, (L.ravel(), inv.ravel()), 1)
    assert int((hist.sum(1) - hist.max(1)).sum()) == 6202


def test_slic_cfg1_counts(O, example_lab):
    """BASELINE.json configs[0]: SLIC S=20 ruler=10 10 it + connectivity on the example image"""
    s = O.SuperpixelSLIC(example_lab, O.SLIC, 20, 10.0)
    assert s.getNumberOfSuperpixels() == 5625
    s.iterate(10)
    s.enforceLabelConnectivity(25)
    assert s.getNumberOfSuperpixels() == 5529


def test_threads_do_not_change_results(O, example_lab):
    sub = np.ascontiguousarray(example_lab[300:700, 400:900])
    O.set_threads(1)
    a = O.SuperpixelSLIC(sub, O.SLIC, 12, 15.0); a.iterate(5)
    O.set_threads(4)
    b = O.SuperpixelSLIC(sub, O.SLIC, 12, 15.0); b.iterate(5)
    assert (a.getLabels() == b.getLabels()).all()
    assert (a.getCentres() == b.getCentres()).all()


def test_connectivity_properties(O):
    from conftest import random_label_map
    rng = np.random.default_rng(3)
    lab = random_label_map(rng, 97, 131)
    out, k = O.enforce_label_connectivity(lab, int(lab.max()) + 1, 25)
    assert out.min() == 0 and out.max() == k - 1
    # every label is one 4-connected component
    import cv2
    for l in range(k):
        n, _ = cv2.connectedComponents((out == l).astype(np.uint8), connectivity=4)
        assert n == 2


def test_contour_flavours(O):
    from conftest import random_label_map
    lab = random_label_map(np.random.default_rng(5), 64, 80)
    for thick in (False, True):
        m = O.label_contour_mask(lab, thick)
        assert set(np.unique(m)) <= {0, 255}
    thin = O.label_contour_mask(lab, False)
    thick = O.label_contour_mask(lab, True)
    assert thick.sum() <= thin.sum()          # threshold 2 marks fewer pixels than threshold 1
    # SEEDS flavour with thick_line=false is the same recurrence as SLIC's threshold 1
    assert (O.label_contour_mask(lab, False, seeds_flavour=True) == thin).all()


def test_lsc_oracle_properties(O, synthmod):
    """LSC restatement (PARITY UNPINNED upstream): seed grid, label range, connectivity after enforcement,
    independence from the thread count"""
    import cv2
    bgr, gt = synthmod.synth(320, 240, 2)
    lab = O.bgr2lab8(bgr)
    s = O.SuperpixelLSC(lab, 16, 0.075)
    k0 = s.getNumberOfSuperpixels()
    assert k0 == 20 * 15                      # ColNum = int(sqrt(300*320/240)) = 20, RowNum = 300/20
    s.iterate(6)
    L = s.getLabels()
    assert L.min() >= 0 and L.max() < k0
    O.set_threads(1)
    s1 = O.SuperpixelLSC(lab, 16, 0.075); s1.iterate(6)
    O.set_threads(4)
    assert (s1.getLabels() == L).all() and (s1.getCentres() == s.getCentres()).all()
    s.enforceLabelConnectivity(25)
    k = s.getNumberOfSuperpixels()
    L2 = s.getLabels()
    assert L2.min() == 0 and L2.max() == k - 1
    for l in range(0, k, 7):
        n, _ = cv2.connectedComponents((L2 == l).astype(np.uint8), connectivity=4)
        assert n == 2
    assert synthmod.boundary_recall(L2, gt) > 0.95


def test_seeds_sizing_pinned_by_screenshot(O):
    """screenshots/screenshot-segmentation.jpg: 3888x5184, Superpixels 1000, Levels 4 -> LCD 973 = nb_cells+1"""
    assert O.SuperpixelSEEDS(3888, 5184, 3, 1000, 4).getNumberOfSuperpixels() == 972
    assert O.SuperpixelSEEDS(1920, 1080, 3, 2000, 4).getNumberOfSuperpixels() == 1296


def test_seeds_orders_agree_statistically(O, synthmod):
    """the phase-parallel visiting order (what the GPU runs) vs upstream's sequential order: same decision rules,
    comparable quality; both keep every superpixel 4-connected"""
    import cv2
    bgr, gt = synthmod.synth(640, 360, 4)
    lab = O.bgr2lab8(bgr)
    res = []
    for order in (0, 1):
        s = O.SuperpixelSEEDS(640, 360, 3, 400, 3, 2, 5, False); s.setOrder(order)
        k = s.getNumberOfSuperpixels()
        e0 = s.energy(s.getLabels()) if False else None
        s.iterate(lab, 6)
        L = s.getLabels()
        assert L.min() >= 0 and L.max() < k
        for l in range(0, k, 5):
            n, _ = cv2.connectedComponents((L == l).astype(np.uint8), connectivity=4)
            assert n <= 2
        res.append((synthmod.boundary_recall(L, gt), synthmod.undersegmentation_error(L, gt), s.energy(L)))
    (br0, ue0, en0), (br1, ue1, en1) = res
    assert abs(br0 - br1) < 0.05 and abs(ue0 - ue1) < 0.03 and abs(en0 - en1) < 0.02, res


def test_lsc_merge_orders_agree_statistically(O, synthmod):
    """post-pass merge of small regions: the hashed visiting order (CUDA opt-in, spx_lsc_set_merge_order(h, 1)) vs the authors' raster order (default) --
    same rule, different order; the partitions must be close and both fully connected"""
    bgr, gt = synthmod.synth(480, 360, 6)
    lab = O.bgr2lab8(bgr)
    out = []
    for order in (0, 1):
        s = O.SuperpixelLSC(lab, 16, 0.075); s.setMergeOrder(order)
        s.iterate(8); s.enforceLabelConnectivity(25)
        L = s.getLabels()
        out.append((s.getNumberOfSuperpixels(), synthmod.boundary_recall(L, gt), synthmod.undersegmentation_error(L, gt), L))
    (k0, br0, ue0, L0), (k1, br1, ue1, L1) = out
    assert abs(k0 - k1) <= 0.03 * k0 and abs(br0 - br1) < 0.02 and abs(ue0 - ue1) < 0.01, (k0, k1, br0, br1, ue0, ue1)
    # same-segment relation of horizontal neighbours agrees almost everywhere
    same0 = L0[:, 1:] == L0[:, :-1]; same1 = L1[:, 1:] == L1[:, :-1]
    assert float((same0 == same1).mean()) > 0.98


@pytest.mark.parametrize("w,h,S,it,p", [(320, 200, 12, 3, 25), (640, 360, 16, 5, 25), (480, 270, 8, 4, 60), (257, 129, 20, 2, 5)])
def test_lsc_fixed_point_merge_equals_the_sequential_loop(O, synthmod, w, h, S, it, p):
    """the authors' sequential merge of small regions (visiting order 0) evaluated as a fixed point -- the data-parallel
    form lsc_fixpoint_kernel runs (oracle/spx_oracle_lsc.c: lsc_merge_fixedpoint) -- gives the sequential loop's output"""
    bgr, _ = synthmod.synth(w, h, 3)
    lab = O.bgr2lab8(bgr)
    out = []
    for algo in (0, 1):
        s = O.SuperpixelLSC(lab, S, 0.075); s.setMergeAlgo(algo)
        s.iterate(it); s.enforceLabelConnectivity(p)
        out.append((s.getNumberOfSuperpixels(), s.getLabels(), s.mergeIterations()))
    assert out[0][0] == out[1][0] and (out[0][1] == out[1][1]).all()
    assert out[1][2] >= 1


def test_golden_hashes_reproduce(O, example_bgr):
    """tests/golden/cfg1_hashes.json (tools/gen_golden.py): every stage of configs[0] and of the reference's fixture run"""
    import importlib.util, json, os
    spec = importlib.util.spec_from_file_location("gen_golden", os.path.join(os.path.dirname(__file__), "..", "tools", "gen_golden.py"))
    gg = importlib.util.module_from_spec(spec); spec.loader.exec_module(gg)
    want = json.load(open(gg.OUT))
    got = gg.compute(O, example_bgr, O.bgr2lab8, lambda lab, S, r: O.SuperpixelSLIC(lab, O.SLIC, S, r))
    assert got == want


def test_golden_hashes_other_algorithms_reproduce(O, example_bgr):
    """tests/golden/algs_hashes.json (tools/gen_golden_algs.py): SLICO, MSLIC, LSC, SEEDS on a crop of the example image --
    regression anchors of our restatement where upstream ships no vector (parity unpinned)"""
    import importlib.util, json, os
    spec = importlib.util.spec_from_file_location("gen_golden_algs", os.path.join(os.path.dirname(__file__), "..", "tools", "gen_golden_algs.py"))
    gg = importlib.util.module_from_spec(spec); spec.loader.exec_module(gg)
    want = json.load(open(gg.OUT))
    got = gg.compute(O.bgr2lab8(example_bgr), lambda lab, alg, S, r: O.SuperpixelSLIC(lab, alg, S, r), O.SuperpixelLSC, O.SuperpixelSEEDS)
    assert got == want
