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
    for parity in [l for l in lines if l["check"] == "dp_parity"]:
        assert parity["ok"] and parity["world"] == 2 and parity["worst_param_rel_l2"] <= 1e-4 and parity["replicas_bit_identical"]


def test_dp_allreduce_grads_sums_peer_buffers_in_rank_order():
    """The isolated reduction (C ABI dp_allreduce_grads) on one GPU: the `peers` are four local buffers; the result is the
    rank-ordered fp32 sum, bit for bit, ragged tail included."""
    import ctypes as C
    from td3_b200 import _lib
    lib = _lib.require_cuda()
    for n in (262_148, 1023, 5):
        g = torch.Generator(device="cuda")
        g.manual_seed(n)
        bufs = [torch.randn(n, device="cuda", generator=g) for _ in range(4)]
        out = torch.empty(n, device="cuda")
        arr = (C.c_void_p * 4)(*[C.c_void_p(b.data_ptr()) for b in bufs])
        _lib.check(lib.dp_allreduce_grads(C.c_void_p(out.data_ptr()), arr, 4, n, _lib.stream_ptr()))
        want = ((bufs[0] + bufs[1]) + bufs[2]) + bufs[3]
        assert torch.equal(out, want)


def test_fused_reduce_with_one_rank_equals_the_plain_update():
    """World size 1 through the data-parallel driver in p2p mode (gradient in symmetric memory, Adam reading it through the
    peer table, flag exchanges with itself, CUDA-graph replay) against the plain single-device update: same Philox key,
    strict fp32 -> bit-identical parameters."""
    import torch.distributed as dist
    from helpers import make_featured
    from td3_b200.data_parallel import DataParallelTD3
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29541")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        _, _, plain, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=4096, lr=1e-3, precision="fp32")
        plain.exec_mode = "launches"
        import pytest as _pt
        _, _, rep, rb2 = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=4096, lr=1e-3, precision="fp32")
        dp = DataParallelTD3(rep, mode="p2p")
        if dp.mode != "p2p":
            _pt.skip(f"symmetric memory unavailable: {getattr(dp, 'p2p_unavailable', '')}")
        import os as _os
        _os.environ["TD3_NO_TAIL_FUSION"] = "1"          # the data-parallel driver runs the unfused sequences
        try:
            _, _, plain, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=4096, lr=1e-3, precision="fp32")
            plain.exec_mode = "launches"
            plain.train(rb, 256, iterations=8)
        finally:
            _os.environ.pop("TD3_NO_TAIL_FUSION", None)
        for _ in range(8):
            dp.train(rb2, 256)
        torch.cuda.synchronize()
        for k in ("actor", "critic", "actor_target", "critic_target"):
            for (n, v), w in zip(getattr(rep, k).state_dict().items(), getattr(plain, k).state_dict().values()):
                assert torch.equal(v, w), f"{k}.{n}: max |d| {(v - w).abs().max().item():.3e}"
    finally:
        dist.destroy_process_group()
