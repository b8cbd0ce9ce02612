"""CPU-only: the C-ABI library loads, exports every symbol include/spx_b200.h declares, and fails loudly
(no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def _declared():
    txt = open(os.path.join(ROOT, "include", "spx_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(spx_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported(spx):
    L = spx.lib()
    names = _declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing


def test_version_and_error_string(spx):
    L = spx.lib()
    assert L.spx_version() >= 100
    assert isinstance(L.spx_last_error(), bytes)


def test_no_cpu_fallback_without_gpu(spx):
    if spx.device_count() > 0:
        pytest.skip("a CUDA device is present")
    img = np.zeros((32, 32, 3), np.uint8)
    with pytest.raises(spx.SpxError) as ei:
        spx.cvtColor(img, spx.COLOR_BGR2Lab)
    assert ei.value.code == -217
    with pytest.raises(spx.SpxError):
        spx.createSuperpixelSLIC(img, spx.SLIC, 8, 10.0)
    with pytest.raises(spx.SpxError):
        spx.label_contour_mask(np.zeros((8, 8), np.int32))


def test_product_does_not_reference_oracle():
    pkg = os.path.join(ROOT, "superpixels-segmentation-gui-opencv_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".inc")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "spx_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def _sass_functions():
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    so = os.path.join(ROOT, "superpixels-segmentation-gui-opencv_b200", "libspx_b200.so")
    txt = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in txt.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
            funcs[cur] = []
        elif cur is not None:
            funcs[cur].append(line)
    return funcs


def test_no_stray_fma_contraction_in_tile_kernel():
    """ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 despite -fmad=false; the kernel works around it.
    The power-of-two variant must contain no FFMA2 at all; the general one only the Markstein division's
    (two per evaluated pixel pair, so an even count) and never more FFMA2 than a quarter of its FMUL2."""
    funcs = _sass_functions()
    tile = {k: v for k, v in funcs.items() if "slic_tile_kernel" in k}
    assert len(tile) >= 4          # {power-of-two xywt, general} x {SLIC, SLICO}
    for name, body in tile.items():
        n_ffma2 = sum(1 for l in body if " FFMA2 " in l)
        n_fmul2 = sum(1 for l in body if " FMUL2" in l)
        assert n_fmul2 >= 3, name
        if "ILb1ELi100E" in name:      # SLIC, POW2 = true: no division sequence at all
            assert n_ffma2 == 0, (name, n_ffma2)
        else:                          # Markstein sequences only: two FFMA2 per division per evaluated pixel pair
            assert n_ffma2 >= 2 and n_ffma2 % 2 == 0, (name, n_ffma2)


def test_no_fma_contraction_in_lsc_tile_kernel():
    """the tiled LSC assignment evaluates the 10-D distance with FADD2 / FMUL2 only: 2 rows x (10 sub + 10 mul + 9 add + 2
    penalty adds) = 62 packed instructions in the candidate loop, and no FFMA2 anywhere (the oracle never contracts)"""
    funcs = _sass_functions()
    k = [v for n, v in funcs.items() if "lsc_tile_kernel" in n]
    assert len(k) == 1
    assert sum(1 for l in k[0] if " FFMA2 " in l) == 0
    assert sum(1 for l in k[0] if " FADD2" in l or " FMUL2" in l) == 62


def test_cpp_adapter_builds_and_fails_loudly_without_gpu(spx):
    """include/spx_ximgproc.hpp (the cv::ximgproc-shaped adapter) compiles against a mock <opencv2/core.hpp>, links to
    libspx_b200.so, and the reference's COMPUTE click written against it throws cv::Exception(-217) on a CPU box."""
    import subprocess
    cpp = os.path.join(ROOT, "tests", "cpp")
    subprocess.check_call(["make", "-C", cpp, "-s"])
    r = subprocess.run([os.path.join(cpp, "compute_click")], capture_output=True, text=True)
    if spx.device_count() > 0:
        assert r.returncode == 0, r.stdout + r.stderr
    else:
        assert r.returncode == 3, r.stdout + r.stderr
        assert "no CUDA device" in r.stdout
