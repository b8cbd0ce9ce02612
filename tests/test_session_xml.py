"""The reference's session data file (OpenCV FileStorage XML, mainwindow.cpp:673-706 / 775-820): the library's host-side writer
and reader against cv2.FileStorage 4.13, both directions.  No GPU involved (host code of libspx_b200.so)."""
import os

import numpy as np
import pytest

LABELS = [(1, "sky", (255, 0, 10)), (2, "my tree & <leaf>", (0, 128, 0)), (7, 'say "hi" \\ there', (1, 2, 3)), (9, "1abc", (9, 9, 9))]


def _maps(rng, h, w):
    lab = rng.integers(0, 5000, (h, w)).astype(np.int32)
    lab[0, 0] = -7; lab[-1, -1] = 2147483647; lab[0, -1] = -2147483648
    return lab, rng.integers(0, 12, (h, w)).astype(np.int32)


def test_cv2_reads_what_we_write(spx, tmp_path):
    import cv2
    rng = np.random.default_rng(1)
    lab, lm = _maps(rng, 37, 53)
    p = str(tmp_path / "a-segmentation-data.xml")
    spx.writeSessionXml(p, 3, LABELS, lab, lm)
    fs = cv2.FileStorage(p, cv2.FileStorage_READ)
    assert fs.isOpened()
    assert int(fs.getNode("GridColor").real()) == 3 and int(fs.getNode("LabelsCount").real()) == len(LABELS)
    for i, (lid, name, col) in enumerate(LABELS):
        assert int(fs.getNode("LabelId%d" % i).real()) == lid
        assert fs.getNode("LabelName%d" % i).string() == name
        c = fs.getNode("LabelColor%d" % i)
        assert [int(c.at(k).real()) for k in range(3)] == list(col)
    assert np.array_equal(fs.getNode("Labels").mat(), lab) and np.array_equal(fs.getNode("LabelsMask").mat(), lm)
    fs.release()


def test_we_read_what_cv2_writes(spx, tmp_path):
    import cv2
    rng = np.random.default_rng(2)
    lab, lm = _maps(rng, 41, 29)
    p = str(tmp_path / "b-segmentation-data.xml")
    fs = cv2.FileStorage(p, cv2.FileStorage_WRITE)
    fs.write("GridColor", 5); fs.write("LabelsCount", len(LABELS))
    for i, (lid, name, col) in enumerate(LABELS):
        fs.write("LabelId%d" % i, lid); fs.write("LabelName%d" % i, name)
        fs.write("LabelColor%d" % i, np.array(col, np.uint8))       # Python writes a 3x1 matrix; C++ `fs << Vec3b` a plain sequence
    fs.write("Labels", lab); fs.write("LabelsMask", lm)
    fs.release()
    g, lst, a, b = spx.readSessionXml(p)
    assert g == 5 and lst == LABELS and np.array_equal(a, lab) and np.array_equal(b, lm)


def test_round_trip_and_errors(spx, tmp_path):
    rng = np.random.default_rng(3)
    lab, lm = _maps(rng, 64, 200)
    p = str(tmp_path / "c-segmentation-data.xml")
    spx.writeSessionXml(p, 0, [], lab, lm)
    g, lst, a, b = spx.readSessionXml(p)
    assert g == 0 and lst == [] and np.array_equal(a, lab) and np.array_equal(b, lm)
    assert max(len(l) for l in open(p).read().split("\n")) <= 80        # FileStorage's line wrapping
    with pytest.raises(spx.SpxError):
        spx.readSessionXml(str(tmp_path / "missing.xml"))
    bad = str(tmp_path / "bad.xml")
    open(bad, "w").write("<?xml version=\"1.0\"?>\n<opencv_storage>\n<GridColor>1</GridColor>\n</opencv_storage>\n")
    with pytest.raises(spx.SpxError):
        spx.readSessionXml(bad)
