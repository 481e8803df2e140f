"""The particle-set encoder in isolation, through the C ABI (set_encoder_fwd / set_encoder_bwd), against the
oracle's _encode (TD3_particles.py:29-32, 53-58 restated in oracle/td3_oracle.py:_SetEncoderMixin) and torch autograd:

  use_tc = 0   strict fp32 tiles: pooled features rel 1e-5, gradients rel 1e-4 (summation order over B*N particles)
  use_tc = 1   the fused tcgen05 kernel (csrc/enc.cuh, K6): TF32 operands, fp32 accumulation; pooled features
               |d| <= 2e-3 * max(1, |x|) (SURVEY 8d's TF32 bound), gradients rel 2e-2 / cosine >= 0.999 (ReLU masks computed in
               TF32 flip for pre-activations within ~1e-4 of zero: tests/test_gpu_tf32.py header)
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import td3_oracle as O

pytestmark = pytest.mark.gpu


class _Enc(torch.nn.Module, O._SetEncoderMixin):
    def __init__(self, n, d):
        super().__init__()
        self._build_encoder((O.Space(4), O.Space(n, d)))


def _run(B, N, D, use_tc, seed=0):
    from td3_b200 import _lib
    lib = _lib.require_cuda()
    torch.manual_seed(seed)
    enc = _Enc(N, D)
    parts = torch.randn(B, N, D)
    pooled_ref = enc._encode(parts)                                   # [B, 128]
    gen = torch.Generator().manual_seed(seed + 1)
    d_pooled = torch.randn(B, O.ENC_OUT, generator=gen)
    pooled_ref.backward(d_pooled)
    H, E = O.ENC_HIDDEN, O.ENC_OUT
    w1 = enc.conv1.weight.detach().reshape(H, D).contiguous().cuda()
    b1 = enc.conv1.bias.detach().contiguous().cuda()
    w2 = enc.conv2.weight.detach().reshape(E, H).contiguous().cuda()
    b2 = enc.conv2.bias.detach().contiguous().cuda()
    P = parts.reshape(B * N, D).contiguous().cuda()
    rows = B * N
    ws_n = int(lib.set_encoder_workspace_floats(B, N, D, H, E))
    ws = torch.empty(ws_n, device="cuda")
    ld = E + 4                                                        # the trunk input is wider than the pooled block
    pooled = torch.full((B, ld), float("nan"), device="cuda")
    h1 = torch.empty(rows, H, device="cuda")
    h2 = torch.empty(rows, E, device="cuda")
    _lib.check(lib.set_encoder_fwd(P.data_ptr(), B, N, D, H, E, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                   pooled.data_ptr(), ld, h1.data_ptr(), h2.data_ptr(), ws.data_ptr(), ws_n, int(use_tc),
                                   _lib.stream_ptr()))
    dp = torch.zeros(B, ld, device="cuda")
    dp[:, :E] = d_pooled.cuda()
    g = {k: torch.full_like(v, float("nan")) for k, v in dict(w1=w1, b1=b1, w2=w2, b2=b2).items()}
    _lib.check(lib.set_encoder_bwd(P.data_ptr(), B, N, D, H, E, w2.data_ptr(), h1.data_ptr(), h2.data_ptr(), pooled.data_ptr(), ld,
                                   dp.data_ptr(), ld, g["w1"].data_ptr(), g["b1"].data_ptr(), g["w2"].data_ptr(),
                                   g["b2"].data_ptr(), ws.data_ptr(), ws_n, int(use_tc), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.isnan(pooled[:, E:]).all()                           # nothing written past the pooled block
    want = {"w1": enc.conv1.weight.grad.reshape(H, D), "b1": enc.conv1.bias.grad, "w2": enc.conv2.weight.grad.reshape(E, H),
            "b2": enc.conv2.bias.grad}
    return pooled[:, :E].cpu(), pooled_ref.detach(), {k: v.cpu() for k, v in g.items()}, want


@pytest.mark.parametrize("B,N,D", [(4, 128, 6), (3, 64, 6), (2, 200, 3)])
def test_encoder_fp32_tiles_match_the_oracle(B, N, D):
    got, want, g, gw = _run(B, N, D, use_tc=False)
    rel = float((got - want).abs().max() / want.abs().max())
    assert rel <= 1e-5, rel
    for k in g:
        r = float((g[k].double() - gw[k].double()).norm() / gw[k].double().norm())
        assert r <= 1e-4, (k, r)


@pytest.mark.parametrize("B,N,D", [(4, 128, 6), (16, 1024, 6), (2, 256, 3)])
def test_fused_tcgen05_encoder_meets_the_tf32_bound(B, N, D):
    got, want, g, gw = _run(B, N, D, use_tc=True)
    err = float(((got - want).abs() / want.abs().clamp(min=1.0)).max())
    print(f"[enc tf32 B={B} N={N} D={D}] pooled |d| {err:.2e}")
    assert err <= 2e-3, err
    for k in g:
        a, b = g[k].double().reshape(-1), gw[k].double().reshape(-1)
        r = float((a - b).norm() / b.norm())
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        print(f"   grad {k}: rel {r:.2e} cos {cos:.6f}")
        assert r <= 2e-2 and cos >= 0.999, (k, r, cos)


def _run_fused(B, N, D, seed=0):
    """set_encoder_fwd_bits + set_encoder_bwd_fused (csrc/encbwd.cuh): the pair TD3_particles.train runs in TF32 mode."""
    from td3_b200 import _lib
    lib = _lib.require_cuda()
    torch.manual_seed(seed)
    enc = _Enc(N, D)
    parts = torch.randn(B, N, D)
    pooled_ref = enc._encode(parts)
    gen = torch.Generator().manual_seed(seed + 1)
    d_pooled = torch.randn(B, O.ENC_OUT, generator=gen)
    pooled_ref.backward(d_pooled)
    H, E = O.ENC_HIDDEN, O.ENC_OUT
    w1 = enc.conv1.weight.detach().reshape(H, D).contiguous().cuda()
    b1 = enc.conv1.bias.detach().contiguous().cuda()
    w2 = enc.conv2.weight.detach().reshape(E, H).contiguous().cuda()
    b2 = enc.conv2.bias.detach().contiguous().cuda()
    P = parts.reshape(B * N, D).contiguous().cuda()
    rows = B * N
    ws_n = int(lib.set_encoder_workspace_floats(B, N, D, H, E))
    ws = torch.empty(ws_n, device="cuda")
    ld = E + 4
    pooled = torch.full((B, ld), float("nan"), device="cuda")
    bits = torch.zeros(16 * rows, dtype=torch.int32, device="cuda")
    _lib.check(lib.set_encoder_fwd_bits(P.data_ptr(), B, N, D, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                        pooled.data_ptr(), ld, bits.data_ptr(), ws.data_ptr(), ws_n, _lib.stream_ptr()))
    dp = torch.zeros(B, ld, device="cuda")
    dp[:, :E] = d_pooled.cuda()
    g = {k: torch.full_like(v, float("nan")) for k, v in dict(w1=w1, b1=b1, w2=w2, b2=b2).items()}
    _lib.check(lib.set_encoder_bwd_fused(P.data_ptr(), B, N, D, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), bits.data_ptr(),
                                         pooled.data_ptr(), ld, dp.data_ptr(), ld, g["w1"].data_ptr(), g["b1"].data_ptr(),
                                         g["w2"].data_ptr(), g["b2"].data_ptr(), ws.data_ptr(), ws_n, _lib.stream_ptr()))
    torch.cuda.synchronize()
    want = {"w1": enc.conv1.weight.grad.reshape(H, D), "b1": enc.conv1.bias.grad, "w2": enc.conv2.weight.grad.reshape(E, H),
            "b2": enc.conv2.bias.grad}
    # the bitmaps against the oracle's activations: bits2 [rows][4] natural, then the two per-tile transposed blocks
    with torch.no_grad():
        h1 = torch.relu(parts.reshape(rows, D) @ enc.conv1.weight.reshape(H, D).T + enc.conv1.bias)
        h2 = torch.relu(h1 @ enc.conv2.weight.reshape(E, H).T + enc.conv2.bias)
    words = bits.cpu().numpy().view(np.uint32)
    nat = words[:4 * rows].reshape(rows, 4)
    got2 = ((nat[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(rows, E).astype(bool)
    t2 = words[4 * rows:8 * rows].reshape(rows // 128, E, 4)
    got2t = ((t2[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(rows // 128, E, 128).transpose(0, 2, 1).reshape(rows, E)
    t1 = words[8 * rows:].reshape(rows // 128, H, 4)
    got1t = ((t1[:, :, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(rows // 128, H, 128).transpose(0, 2, 1).reshape(rows, H)
    flips = {"h2": float((got2 != (h2.numpy() > 0)).mean()), "h2T_vs_h2": float((got2t.astype(bool) != got2).mean()),
             "h1": float((got1t.astype(bool) != (h1.numpy() > 0)).mean())}
    return pooled[:, :E].cpu(), pooled_ref.detach(), {k: v.cpu() for k, v in g.items()}, want, flips


@pytest.mark.parametrize("B,N,D", [(4, 128, 6), (16, 1024, 6), (2, 256, 3), (5, 384, 7), (300, 128, 6)])
def test_fused_backward_kernels_match_autograd(B, N, D):
    got, want, g, gw, flips = _run_fused(B, N, D)
    err = float(((got - want).abs() / want.abs().clamp(min=1.0)).max())
    print(f"[enc fused fwd+bwd B={B} N={N} D={D}] pooled |d| {err:.2e}  bitmap mismatches {flips}")
    assert err <= 2e-3, err
    # the two copies of the h2 bitmap are the same bits; against an fp32 forward only pre-activations within TF32 error of
    # zero may differ
    assert flips["h2T_vs_h2"] == 0.0 and flips["h2"] <= 2e-3 and flips["h1"] <= 2e-3, flips
    for k in g:
        a, b = g[k].double().reshape(-1), gw[k].double().reshape(-1)
        r = float((a - b).norm() / b.norm())
        cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
        print(f"   grad {k}: rel {r:.2e} cos {cos:.6f}")
        assert r <= 2e-2 and cos >= 0.999, (k, r, cos)


def test_fused_backward_is_deterministic():
    a = _run_fused(8, 256, 6, seed=3)[2]
    b = _run_fused(8, 256, 6, seed=3)[2]
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_fused_pair_stress_many_tiles_per_cta_bit_identical_replays():
    """The kernels' cross-warp protocols (mbarrier phases over a CTA's run of tiles, TMEM / shared-memory buffer reuse,
    per-CTA partials) at the real cfg4 shape -- 2048 tiles, ~14 per CTA in the dW2 kernel and ~28 per CTA in the dX
    kernel -- and at shapes whose tile count does not divide the CTA count: 40 replays must be bit-identical (a lost
    barrier phase or a buffer reused too early shows up as a changed low bit long before it breaks a tolerance)."""
    from td3_b200 import _lib
    lib = _lib.require_cuda()
    H, E = O.ENC_HIDDEN, O.ENC_OUT
    for (B, N, D) in [(256, 1024, 6), (37, 384, 5), (149, 128, 7)]:
        torch.manual_seed(B)
        rows = B * N
        P = torch.randn(rows, D, device="cuda")
        w1 = (torch.randn(H, D, device="cuda") * 0.4).contiguous()
        b1 = torch.randn(H, device="cuda") * 0.1
        w2 = (torch.randn(E, H, device="cuda") * 0.06).contiguous()
        b2 = torch.randn(E, device="cuda") * 0.1
        ws_n = int(lib.set_encoder_workspace_floats(B, N, D, H, E))
        ws = torch.empty(ws_n, device="cuda")
        pooled = torch.zeros(B, E, device="cuda")
        bits = torch.zeros(16 * rows, dtype=torch.int32, device="cuda")
        dp = torch.randn(B, E, device="cuda")
        first = None
        for it in range(40):
            g = [torch.full_like(t, float("nan")) for t in (w1, b1, w2, b2)]
            bits.zero_()
            _lib.check(lib.set_encoder_fwd_bits(P.data_ptr(), B, N, D, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                                pooled.data_ptr(), E, bits.data_ptr(), ws.data_ptr(), ws_n, _lib.stream_ptr()))
            _lib.check(lib.set_encoder_bwd_fused(P.data_ptr(), B, N, D, w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), bits.data_ptr(),
                                                 pooled.data_ptr(), E, dp.data_ptr(), E, g[0].data_ptr(), g[1].data_ptr(),
                                                 g[2].data_ptr(), g[3].data_ptr(), ws.data_ptr(), ws_n, _lib.stream_ptr()))
            torch.cuda.synchronize()
            snap = [t.clone() for t in g] + [pooled.clone(), bits.clone()]
            assert all(torch.isfinite(t).all() for t in snap[:5])
            if first is None:
                first = snap
            else:
                for a, b in zip(first, snap):
                    assert torch.equal(a, b), (B, N, D, it)


def test_fused_encoder_refuses_shapes_it_does_not_cover():
    from td3_b200 import _lib
    lib = _lib.require_cuda()
    x = torch.zeros(1024, device="cuda")
    rc = lib.set_encoder_fwd(x.data_ptr(), 1, 100, 6, 256, 128, x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(), x.data_ptr(),
                             128, None, None, x.data_ptr(), 10 ** 9, 1, _lib.stream_ptr())
    assert rc != 0 and b"128" in lib.td3_last_error()
