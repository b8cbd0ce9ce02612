"""Row-strip SLIC over 2 GPUs against the single-GPU object (bit-identical labels and centres).  Needs two CUDA
devices: skipped on a one-GPU box (the recorded 2-GPU run is profiles/r1_strip_mode_2gpu.json)."""
import json
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _strip_check(extra_env=None):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "strip_check.py"), "2048", "1536"]
    env = dict(os.environ)
    env.update(extra_env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["labels_identical"] and d["centres_identical"] and d["connectivity_and_contour_identical"] and d["n_gpus"] == 2
    return d


def test_strip_mode_two_gpus_bit_identical():
    """all three exchange modes (neighbour = partitioned cluster state, peer = all-to-all, NCCL) against the single-GPU object"""
    d = _strip_check()
    assert d["neighbour_mode_ran_as"] == "neighbour"


def test_strip_neighbour_mode_reruns_when_the_drift_check_fires():
    """the device-side drift check of the partitioned mode is raised artificially in the 3rd update: iterate() must notice,
    rebuild the object with replicated state, rerun -- and still match the single-GPU object bit for bit"""
    d = _strip_check({"SPX_STRIP_FORCE_DRIFT": "3"})
    assert d["neighbour_mode_ran_as"] == "peer"
