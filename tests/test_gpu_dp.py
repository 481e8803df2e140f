"""Data-parallel critic on two real GPUs: W = 2 ranks with one NCCL all-reduce of the packed gradient per update equal a
single device on the same global batch (tools/check_dp.py asserts parameters within 1e-4 relative L2, loss within 1e-4).
Skipped on a single-GPU box; the host logic is covered on CPU with gloo in test_dp_gloo.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_ranks_match_one_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "check_dp.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    parity = [l for l in lines if l["check"] == "dp_parity"][0]
    assert parity["ok"] and parity["world"] == 2 and parity["worst_param_rel_l2"] <= 1e-4
