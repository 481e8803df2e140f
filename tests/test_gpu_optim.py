"""K4/K5 fused Adam + Polyak over packed buffers vs torch.optim.Adam / the reference's Polyak loop.
Tolerance: moments <= 2 ulp (torch CPU lerp/addcmul kernels may or may not contract to fma depending on
the vector path taken); parameters <= 2 ulp per step measured at the magnitude of the operands of the final add,
ulp(max(|p_before|, |p_after|, lr)) -- an element that is almost cancelled by its +-lr update has a tiny value whose own
ulp says nothing about the rounding of the add that produced it.  Polyak bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ulp_diff(a, b):
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    return np.abs(ai - bi)


@pytest.mark.parametrize("n", [1, 7, 2048, 100_003])
def test_adam_matches_torch(n):
    from td3_b200 import _lib
    lib = _lib.require_cuda()
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3)
    p = p0.clone().cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    for t in range(1, 6):
        grad = torch.randn(n, generator=g) * (10.0 ** float(torch.randint(-6, 2, (1,), generator=g)))
        ref.grad = grad.clone()
        p_before = ref.detach().numpy().copy()
        opt.step()
        gd = grad.cuda()
        _lib.check(lib.adam_polyak_step(C.c_void_p(p.data_ptr()), C.c_void_p(gd.data_ptr()), C.c_void_p(m.data_ptr()),
                                        C.c_void_p(v.data_ptr()), None, n, t, 1e-3, 0.9, 0.999, 1e-8, 0.0, _lib.stream_ptr()))
        st = opt.state[ref]
        assert _ulp_diff(m.cpu().numpy(), st["exp_avg"].numpy()).max() <= 2
        assert _ulp_diff(v.cpu().numpy(), st["exp_avg_sq"].numpy()).max() <= 2
        got, want = p.cpu().numpy(), ref.detach().numpy()
        scale = np.spacing(np.maximum(np.maximum(np.abs(p_before), np.abs(want)), np.float32(1e-3)))
        assert np.all(np.abs(got - want) <= 2 * t * scale), (t, float((np.abs(got - want) / scale).max()))


def test_polyak_bit_exact_and_fused_order():
    from td3_b200 import _lib
    lib = _lib.require_cuda()
    n, tau = 50_001, 0.005
    g = torch.Generator().manual_seed(1)
    p, tgt = torch.randn(n, generator=g), torch.randn(n, generator=g)
    want = tau * p + (1 - tau) * tgt                                         # TD3_featured.py:167-171
    pd, td = p.cuda(), tgt.clone().cuda()
    _lib.check(lib.adam_polyak_step(C.c_void_p(pd.data_ptr()), None, None, None, C.c_void_p(td.data_ptr()), n, 0, 0.0,
                                    0.9, 0.999, 1e-8, tau, _lib.stream_ptr()))
    assert torch.equal(td.cpu(), want)
    assert torch.equal(pd.cpu(), p)
    # fused Adam + Polyak: the target sees the *stepped* parameters (actor step order, :164-171)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=1e-2)
    grad = torch.randn(n, generator=g)
    ref.grad = grad.clone()
    opt.step()
    want_t = tau * ref.detach() + (1 - tau) * tgt
    pd, td = p.cuda(), tgt.clone().cuda()
    m, v = torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    gd = grad.cuda()
    _lib.check(lib.adam_polyak_step(C.c_void_p(pd.data_ptr()), C.c_void_p(gd.data_ptr()), C.c_void_p(m.data_ptr()),
                                    C.c_void_p(v.data_ptr()), C.c_void_p(td.data_ptr()), n, 1, 1e-2, 0.9, 0.999, 1e-8, tau,
                                    _lib.stream_ptr()))
    assert _ulp_diff(td.cpu().numpy(), want_t.numpy()).max() <= 2


def test_bad_arguments_raise_value_error():
    from td3_b200 import _lib
    lib = _lib.require_cuda()
    p = torch.zeros(4, device="cuda")
    with pytest.raises(ValueError):
        _lib.check(lib.adam_polyak_step(C.c_void_p(p.data_ptr()), None, None, None, None, 4, 1, 1e-3, 0.9, 0.999, 1e-8,
                                        0.005, _lib.stream_ptr()))
