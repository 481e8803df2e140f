"""CPU-only checks of the C-ABI boundary: the library loads without a GPU and exports every symbol
include/td3_b200.h declares; ctypes mirrors match the C structs; host-side layout logic."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT
from oracle import td3_oracle as O


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "td3_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:td3|rb|adam|dp|set_encoder)_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from td3_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/td3_b200.h but not exported by libtd3b200.so"
    assert set(declared) == set(_lib.SIGNATURES), set(declared) ^ set(_lib.SIGNATURES)
    assert lib.td3_abi_version() == 1


def test_struct_mirrors_match_c_sizes():
    from td3_b200 import _lib
    lib = _lib.load()
    sizes = (C.c_int64 * 4)()
    lib.td3_struct_sizes(sizes)
    assert list(sizes) == [C.sizeof(_lib.NetLayout), C.sizeof(_lib.ParamSet), C.sizeof(_lib.AgentConfig), C.sizeof(_lib.ReplayView)]


def test_argument_errors_need_no_gpu():
    from td3_b200 import _lib
    lib = _lib.load()
    with pytest.raises(ValueError):
        _lib.check(lib.td3_agent_create(None, None))
    cfg = _lib.AgentConfig()
    cfg.n_q, cfg.n_agents = 3, 1
    h = C.c_void_p()
    with pytest.raises(ValueError, match="n_q"):
        _lib.check(lib.td3_agent_create(C.byref(cfg), C.byref(h)))
    with pytest.raises(ValueError, match="empty"):
        _lib.check(lib.rb_philox_indices(C.c_void_p(16), 4, 0, 0, 0, 0, None))


def test_product_refuses_to_run_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from td3_b200.TD3_featured import TD3
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        TD3(O.Space(17), O.Space(6))


@pytest.mark.parametrize("norm", [None, "layer"])
def test_packed_layout_follows_reference_parameter_order(norm):
    """state_dict keys/order/shapes of the shells == the oracle's (== the reference's, test_oracle_golden)."""
    from td3_b200 import packing as P
    a = P.MlpActor(17, 6, 1.0, norm, (500, 400, 300))
    c = P.MlpCritic(17, 6, norm, (500, 400, 200))
    oa, oc = O.MlpActor(17, 6, 1.0, norm), O.MlpCritic(17, 6, norm)
    for mine, ref in ((a, oa), (c, oc)):
        assert [(k, tuple(v.shape)) for k, v in mine.state_dict().items()] == [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    lay = P.net_layout(c.q1)
    assert lay.n_linear == 4 and list(lay.dims)[:5] == [23, 500, 400, 200, 1]
    table = P.param_table(c.q1)
    offs = [off for k, (off, _) in table.items() if k != "__total__"]
    assert all(o % 4 == 0 for o in offs) and offs == sorted(offs) and lay.n_floats % 64 == 0
    sa = P.SetActor(8, 64, 6, 3, norm, (500, 400, 300))
    osa = O.SetActor((O.Space(8), O.Space(64, 6)), O.Space(3), norm)
    assert [(k, tuple(v.shape)) for k, v in sa.state_dict().items()] == [(k, tuple(v.shape)) for k, v in osa.state_dict().items()]
    sl = P.net_layout(sa)
    assert (sl.enc_hidden, sl.enc_out, sl.dims[0]) == (256, 128, 136)


def test_dropin_module_names_resolve():
    import importlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "dropin"))
    try:
        for name in ("TD3_base", "TD3_featured", "TD3_particles", "my_replay_buffer"):
            sys.modules.pop(name, None)
            importlib.import_module(name)
        import TD3_featured, TD3_particles, my_replay_buffer, TD3_base
        assert TD3_featured.TD3.__mro__[1] is TD3_base.TD3_base
        assert my_replay_buffer.ReplayBuffer is my_replay_buffer.ReplayBuffer_particles
        assert hasattr(TD3_particles.TD3, "_actor_learn") and hasattr(TD3_particles.TD3, "select_action")
    finally:
        sys.path.remove(os.path.join(ROOT, "dropin"))
        for name in ("TD3_base", "TD3_featured", "TD3_particles", "my_replay_buffer"):
            sys.modules.pop(name, None)


@pytest.mark.parametrize("S,A,aw,qw", [(17, 6, (400, 300), (400, 300)), (17, 6, (500, 400, 300), (500, 400, 200)),
                                        (32, 0, (64, 64), (64, 64)), (3, 1, (400, 300), (400, 300))])
def test_plain_mlp_layouts_meet_the_tail_fusion_preconditions(S, A, aw, qw):
    """engine.cu fuses the first layer's dW into the optimiser launch only when the first layer sits at the head of each
    network's packed block, 16-byte aligned, with K <= 32 and a width that is a multiple of 4 (plan_agent: l0_fusable).
    The packing of the plain MLPs must keep satisfying that, or the fast path silently turns itself off."""
    from td3_b200.packing import MlpActor, MlpCritic, net_layout
    A = max(A, 1)
    for net, k0 in ((MlpActor(S, A, 1.0, None, aw), S), (MlpCritic(S, A, None, qw).q1, S + A)):
        lay = net_layout(net)
        n, d0, d1 = lay.n_linear, lay.dims[0], lay.dims[1]
        assert d0 == k0 and n >= 2
        fusable = (d0 <= 32 and d1 % 4 == 0 and lay.w_off[0] == 0 and lay.w_off[1] % 4 == 0 and lay.b_off[0] >= d0 * d1 and
                   lay.w_off[1] >= lay.b_off[0] + d1 and lay.n_floats % 4 == 0)
        assert fusable == (k0 <= 32), (S, A, d0, d1, lay.w_off[0], lay.b_off[0], lay.w_off[1], lay.n_floats)
