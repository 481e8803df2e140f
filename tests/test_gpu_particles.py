"""TD3_particles update parity (set encoder, A-wide Q head, unclamped target action, CDQ switch).
Tolerances as in test_gpu_featured.py; the pooled mean over N particles adds one more reduction, so Q
values get 5e-5 and parameters 5e-4 relative."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import compare_nets, make_particles
from oracle import make_golden as MG

pytestmark = pytest.mark.gpu


def _run(ora, orb, ours, rb, B, steps, A, rows, lr, seed=7):
    rs = np.random.RandomState(seed)
    for t in range(steps):
        idx = rs.randint(0, rows, size=B)
        nz = rs.standard_normal((B, A)).astype(np.float32)
        ora.train(orb, B, indices=idx, noise=nz)
        ours.train(rb, B, indices=idx, noise=nz)
        want = ora.trace["critic_loss"]
        got = float(ours.last_critic_loss[0].item())
        assert abs(got - want) <= 5e-5 * max(1.0, abs(want)), (t, got, want)
        dbg = ours.debug_tensors()
        q1 = dbg["q"][0, 0].cpu().numpy()
        assert q1.shape == (B, A)
        assert np.all(np.abs(q1 - ora.trace["q1"].numpy()) <= 5e-5 * np.maximum(1.0, np.abs(q1)))
        tq = dbg["target_q"][0].cpu().numpy()
        assert np.all(np.abs(tq - ora.trace["target_q"].numpy()) <= 5e-5 * np.maximum(1.0, np.abs(tq)))
        if ora.trace["actor_loss"] is not None:
            al = float(ours.last_actor_loss[0].item())
            assert abs(al - ora.trace["actor_loss"]) <= 5e-5 * max(1.0, abs(ora.trace["actor_loss"]))
        compare_nets(ours, ora, tol_rel=5e-4, max_abs=0.2 * lr * (t + 1), label=f"step {t}")


@pytest.mark.parametrize("norm", [None, "layer"])
@pytest.mark.parametrize("CDQ", [True, False])
def test_trajectory_matches_oracle(norm, CDQ):
    ora, orb, ours, rb = make_particles(norm=norm, CDQ=CDQ)
    _run(ora, orb, ours, rb, B=8, steps=4, A=3, rows=64, lr=1e-3)


@pytest.mark.parametrize("CDQ", [True, False])
@pytest.mark.parametrize("exec_mode", ["graph", "persistent", "launches"])
def test_weight_normalization_trajectory(CDQ, exec_mode):
    """norm="weight_normalization" (TD3_particles.py:48-50): parameters are (bias, weight_g, weight_v); the gradient
    reaches g and v through W = g v / ||v||, Adam and Polyak run on g and v themselves."""
    if exec_mode != "graph" and not CDQ:
        pytest.skip("one CDQ setting per alternative executor is enough")
    ora, orb, ours, rb = make_particles(norm="weight_normalization", CDQ=CDQ)
    ours.exec_mode = exec_mode
    assert [n for n, _ in ours.critic.named_parameters()][:5] == [n for n, _ in ora.critic.named_parameters()][:5]
    assert any(n.endswith("weight_g") for n, _ in ours.actor.named_parameters())
    _run(ora, orb, ours, rb, B=8, steps=4, A=3, rows=64, lr=1e-3)
    # B = 1 surface reads the current (g, v), not a stale effective weight
    rs = np.random.RandomState(3)
    st = (rs.standard_normal(8), rs.standard_normal((64, 6)))
    np.testing.assert_allclose(ours.select_action(st), ora.select_action(st), rtol=1e-4, atol=1e-5)
    ac = np.linspace(-0.5, 0.5, 3)
    np.testing.assert_allclose(np.stack(ours.eval_q(st, ac)), np.stack(ora.eval_q(st, ac)), rtol=1e-4, atol=1e-4)


def test_weight_normalization_ragged_policy_freq_1_tf32():
    ora, orb, ours, rb = make_particles(F=5, N=37, D=3, A=2, rows=50, policy_freq=1, norm="weight_normalization",
                                        precision="tf32")
    rs = np.random.RandomState(7)
    for t in range(3):
        idx = rs.randint(0, 50, size=19)
        nz = rs.standard_normal((19, 2)).astype(np.float32)
        ora.train(orb, 19, indices=idx, noise=nz)
        ours.train(rb, 19, indices=idx, noise=nz)
        want = ora.trace["critic_loss"]
        assert abs(float(ours.last_critic_loss[0].item()) - want) <= 3e-2 * max(1.0, abs(want))
    # an Adam step moves an element by at most ~lr whatever the gradient's size, so a TF32-rounded gradient near zero can
    # flip an element's direction: bound the absolute difference by the distance two opposite trajectories can reach
    compare_nets(ours, ora, tol_rel=5e-2, max_abs=2 * 1e-3 * 3, label="tf32 weight norm")


def test_ragged_shapes_and_policy_freq_1():
    # N not a multiple of 8/32, D = 3, F = 5, A = 2, batch not a multiple of the tile
    ora, orb, ours, rb = make_particles(F=5, N=37, D=3, A=2, rows=50, policy_freq=1)
    _run(ora, orb, ours, rb, B=19, steps=3, A=2, rows=50, lr=1e-3)


def test_split_k_path_large_particle_count():
    # B*N = 16384 rows -> the encoder weight gradients take the split-K + reduce path
    ora, orb, ours, rb = make_particles(N=1024, rows=32)
    _run(ora, orb, ours, rb, B=16, steps=2, A=3, rows=32, lr=1e-3)


@pytest.mark.parametrize("name", ["particles_none", "particles_layer", "particles_nocdq", "particles_wn"])
def test_matches_reference_golden_fixture(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = ast.literal_eval(str(z["case"]))
    ora, orb, ours, rb = make_particles(F=case["F"], N=case["N"], D=case["D"], A=case["A"], rows=case["rows"],
                                        norm=case["norm"], CDQ=case["CDQ"], policy_freq=case["policy_freq"], lr=1e-3)
    for t in range(case["steps"]):
        ours.train(rb, case["B"], indices=z["indices"][t], noise=z["noise"][t])
        got = float(ours.last_critic_loss[0].item())
        assert abs(got - z["critic_loss"][t]) <= 5e-5 * max(1.0, abs(z["critic_loss"][t])), (t, got, z["critic_loss"][t])
    # B=1 surface against the reference's own outputs
    rs = np.random.RandomState(3)
    st = (rs.standard_normal(case["F"]), rs.standard_normal((case["N"], case["D"])))
    ac = np.linspace(-0.5, 0.5, case["A"])
    np.testing.assert_allclose(ours.select_action(st), z["select_action"], rtol=1e-4, atol=1e-5)
    q = ours.eval_q(st, ac)
    assert len(q) == (2 if case["CDQ"] else 1)
    np.testing.assert_allclose(np.stack(q), z["eval_q"], rtol=1e-4, atol=1e-4)


def test_actor_learn_on_caller_tensors():
    """evaluate_model.py:49 calls policy._actor_learn(features, particles) directly."""
    ora, orb, ours, rb = make_particles()
    rs = np.random.RandomState(5)
    f = torch.as_tensor(rs.standard_normal((8, 8)).astype(np.float32))
    p = torch.as_tensor(rs.standard_normal((8, 64, 6)).astype(np.float32))
    ora._actor_learn(f, p)
    ours._actor_learn(f.cuda(), p.cuda())
    compare_nets(ours, ora, tol_rel=5e-4, max_abs=0.2 * 1e-3, label="_actor_learn")
