"""Oracle vs the committed golden fixtures (made from the real reference by
oracle/make_golden.py) and, where /root/reference exists, vs the live reference."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, has_reference
from oracle import td3_oracle as O
from oracle import make_golden as MG

CASES = sorted(MG.CASES)


def _load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return z, ast.literal_eval(str(z["case"]))


def _build_oracle(case):
    kw = dict(norm=case["norm"], policy_freq=case["policy_freq"], lr=1e-3, **MG.hyper(case))
    torch.manual_seed(0)
    if case["kind"] == "featured":
        obs, act = O.Space(case["S"]), O.Space(case["A"])
        ag = O.TD3Featured(obs, act, **kw)
        rb = O.ReplayFeatured(obs, act, case["rows"])
        O.fill_featured(rb, O.synthetic_transitions_featured(case["rows"], case["S"], case["A"], seed=0))
    else:
        kw.pop("max_action", None)
        obs, act = (O.Space(case["F"]), O.Space(case["N"], case["D"])), O.Space(case["A"])
        ag = O.TD3Particles(obs, act, CDQ=case["CDQ"], **kw)
        rb = O.ReplayParticles(obs, act, case["rows"])
        O.fill_particles(rb, O.synthetic_transitions_particles(case["rows"], case["F"], case["N"], case["D"], case["A"], seed=0))
    return ag, rb


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_golden(name):
    """Tolerance: 2e-5 relative.  The fixtures were produced single-threaded on this
    image's torch; another host CPU may pick a different GEMM micro-kernel (summation
    order), so exact equality is only asserted in test_oracle_matches_live_reference."""
    torch.set_num_threads(1)
    z, case = _load(name)
    assert case == MG.CASES[name], "fixture is stale: re-run oracle/make_golden.py"
    ag, rb = _build_oracle(case)
    for t in range(case["steps"]):
        ag.train(rb, case["B"], indices=z["indices"][t], noise=z["noise"][t])
        np.testing.assert_allclose(ag.trace["critic_loss"], z["critic_loss"][t], rtol=2e-5)
        np.testing.assert_allclose(ag.trace["q1"].numpy(), z["q1"][t], rtol=2e-4, atol=2e-5)
        np.testing.assert_allclose(ag.trace["target_q"].numpy(), z["target_q"][t], rtol=2e-4, atol=2e-5)
        for k in ("actor", "critic", "actor_target", "critic_target"):
            np.testing.assert_allclose(O.param_digest(getattr(ag, k)), z["digest_" + k][t], rtol=1e-4, atol=1e-4)


def test_replay_sample_golden():
    z = np.load(os.path.join(GOLDEN, "replay_sample.npz"))
    rs = np.random.RandomState(int(z["seed"]))
    obs, act = O.Space(5), O.Space(2)
    rb = O.ReplayFeatured(obs, act, 37)
    for _ in range(50):
        rb.add(rs.standard_normal(5), rs.uniform(-1, 1, 2), rs.standard_normal(5).astype(np.float32),
               float(rs.standard_normal()), float(rs.uniform() < 0.2))
    ind = rs.randint(0, 37, size=64)
    assert np.array_equal(ind, z["feat_indices"])
    for k, v in zip(rb.fields, rb.sample(64, ind)):
        assert v.dtype == torch.float32
        assert np.array_equal(v.numpy(), z["feat_" + k]), k          # bit-exact
    assert [rb.ptr, rb.size] == list(z["feat_ptr_size"])


def test_sample_empty_buffer_raises():
    rb = O.ReplayFeatured(O.Space(3), O.Space(1), 8)
    with pytest.raises(ValueError):
        rb.sample(4)                                                  # np.random.randint(0, 0) -> ValueError


@pytest.mark.skipif(not has_reference(), reason="/root/reference only exists in the build container")
@pytest.mark.parametrize("name", ["featured_layer", "particles_nocdq", "particles_wn"])
def test_oracle_matches_live_reference(name):
    RF, RP, RB = MG.import_reference()
    res = MG.run_case(name, MG.CASES[name], RF, RP, RB)               # asserts bit-equality internally
    z, _ = _load(name)
    assert np.array_equal(res["critic_loss"], z["critic_loss"])


def test_package_and_oracle_generate_the_same_synthetic_inputs():
    """bench.py feeds the CUDA arm from td3_b200/synthetic.py and the CPU arm from the oracle's own generators (the
    measured path never imports oracle/): both must be the same bytes."""
    from oracle import td3_oracle as O
    from td3_b200 import synthetic as Syn
    a, b = O.synthetic_transitions_featured(64, 17, 6, seed=3), Syn.transitions_featured(64, 17, 6, seed=3)
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
    a, b = O.synthetic_transitions_particles(8, 8, 16, 6, 3, seed=1), Syn.transitions_particles(8, 8, 16, 6, 3, seed=1)
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) and a[k].dtype == b[k].dtype for k in a)
