"""API surface the reference's drivers use: select_action / eval_q, save / load in the
reference's checkpoint format (both directions), seeded construction."""
import os

import numpy as np
import pytest
import torch

from helpers import compare_nets, make_featured
from oracle import td3_oracle as O

pytestmark = pytest.mark.gpu


def test_same_seed_same_initial_weights_as_reference_construction():
    from td3_b200.TD3_featured import TD3
    obs, act = O.Space(17), O.Space(6)
    torch.manual_seed(123)
    ora = O.TD3Featured(obs, act)
    torch.manual_seed(123)
    ours = TD3(obs, act, seed=0)
    for k in ("actor", "critic", "actor_target", "critic_target"):
        for (n1, a), (n2, b) in zip(getattr(ours, k).state_dict().items(), getattr(ora, k).state_dict().items()):
            assert n1 == n2 and torch.equal(a.cpu(), b), (k, n1)


def test_select_action_and_eval_q():
    ora, orb, ours, rb = make_featured(norm="layer", max_action=2.0)
    st, ac = np.linspace(-1, 1, 17), np.linspace(-0.5, 0.5, 6)
    a = ours.select_action(st)
    assert a.shape == (6,) and a.dtype == np.float32
    np.testing.assert_allclose(a, ora.select_action(st), rtol=1e-5, atol=1e-6)
    q, oq = ours.eval_q(st, ac), ora.eval_q(st, ac)
    assert isinstance(q, list) and len(q) == 2 and q[0].shape == (1,)
    np.testing.assert_allclose(np.stack(q), np.stack(oq), rtol=1e-5, atol=1e-5)
    # nn.Module-style calls used by evaluate_model.py
    ours.actor.eval(); ours.critic.eval()
    x = torch.randn(5, 17)
    np.testing.assert_allclose(ours.actor(x.cuda()).cpu().numpy(), ora.actor(x).detach().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ours.actor_target(x.cuda()).cpu().numpy(), ora.actor_target(x).detach().numpy(), rtol=1e-5, atol=1e-6)
    u = torch.rand(5, 6) * 2 - 1
    q1 = ours.critic.Q1(x.cuda(), u.cuda())
    np.testing.assert_allclose(q1.cpu().numpy(), ora.critic.Q1(x, u).detach().numpy(), rtol=1e-5, atol=1e-5)


def test_checkpoint_roundtrip_with_reference_format(tmp_path):
    """our save -> oracle(torch) load, and oracle save -> our load, then both continue identically."""
    from td3_b200.TD3_base import TD3_base
    ora, orb, ours, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300))
    rs = np.random.RandomState(1)
    for _ in range(4):
        idx, nz = rs.randint(0, 512, size=32), rs.standard_normal((32, 6)).astype(np.float32)
        ora.train(orb, 32, indices=idx, noise=nz)
        ours.train(rb, 32, indices=idx, noise=nz)
    d1 = str(tmp_path / "ours")
    ours.save(d1)
    assert sorted(os.listdir(d1)) == ["actor", "actor_optimizer", "actor_target", "critic", "critic_optimizer", "critic_target"]
    # a torch-only consumer (the reference's TD3_base.load) reads our files
    torch.manual_seed(99)
    ora2 = O.TD3Featured(O.Space(17), O.Space(6), actor_widths=(400, 300), q_widths=(400, 300), lr=1e-3)
    for k in ("critic", "critic_target", "actor", "actor_target"):
        getattr(ora2, k).load_state_dict(torch.load(os.path.join(d1, k), map_location="cpu"))
    for k in ("critic_optimizer", "actor_optimizer"):
        getattr(ora2, k).load_state_dict(torch.load(os.path.join(d1, k), map_location="cpu"))
    ora2.total_it = ora.total_it
    sd = torch.load(os.path.join(d1, "critic_optimizer"), map_location="cpu")
    assert set(sd.keys()) == {"state", "param_groups"} and float(sd["state"][0]["step"]) == 4.0
    assert sd["param_groups"][0]["params"] == list(range(len(list(ora.critic.parameters()))))
    # reference-format files written by torch -> our load
    d2 = str(tmp_path / "ref")
    os.makedirs(d2)
    for k in ("critic", "critic_target", "critic_optimizer", "actor", "actor_target", "actor_optimizer"):
        torch.save(getattr(ora, k).state_dict(), os.path.join(d2, k))
    from td3_b200.TD3_featured import TD3
    ours2 = TD3(O.Space(17), O.Space(6), actor_widths=(400, 300), q_widths=(400, 300), lr=1e-3, seed=5)
    ours2.load(d2)
    ours2.total_it = ora.total_it
    for _ in range(3):
        idx, nz = rs.randint(0, 512, size=32), rs.standard_normal((32, 6)).astype(np.float32)
        ora.train(orb, 32, indices=idx, noise=nz)
        ora2.train(orb, 32, indices=idx, noise=nz)
        ours2.train(rb, 32, indices=idx, noise=nz)
    compare_nets(ours2, ora, tol_rel=2e-4, max_abs=0.2 * 1e-3 * 7, label="ref->ours load")
    for a, b in zip(ora2.critic.parameters(), ora.critic.parameters()):
        np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=2e-3, atol=2e-5)
    # missing *_target files fall back to a copy of the online net (TD3_base.py:40-50)
    os.remove(os.path.join(d2, "critic_target")); os.remove(os.path.join(d2, "actor_target"))
    ours3 = TD3(O.Space(17), O.Space(6), actor_widths=(400, 300), q_widths=(400, 300), lr=1e-3, seed=5)
    ours3.load(d2)
    for a, b in zip(ours3.critic_target.parameters(), ours3.critic.parameters()):
        assert torch.equal(a, b)


def test_wrong_buffer_type_and_weight_norm_raise():
    from td3_b200.TD3_featured import TD3
    ora, orb, ours, rb = make_featured()
    with pytest.raises(TypeError):
        ours.train(orb, 8)
    with pytest.raises(NotImplementedError):
        TD3(O.Space(4), O.Space(2), norm="weight_normalization")
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    empty = ReplayBuffer_featured(O.Space(17), O.Space(6), max_size=16)
    with pytest.raises(ValueError):
        ours.train(empty, 8)


def test_wait_critic_loss_matches_device_value():
    """The host mirror of the critic loss (8-byte word written by the fused head kernel to pinned memory) is the value
    the device holds, for single updates, multi-iteration calls and after a re-plan."""
    _, _, ours, rb = make_featured(rows=1024, actor_widths=(400, 300), q_widths=(400, 300))
    for step in range(6):
        ours.train(rb, 64)
        got = ours.wait_critic_loss()
        torch.cuda.synchronize()
        assert got == float(ours.last_critic_loss[0])
    ours.train(rb, 64, iterations=5)
    got = ours.wait_critic_loss()
    torch.cuda.synchronize()
    assert got == float(ours.last_critic_loss[0])
    ours.train(rb, 32)                       # new batch size: new plan, the update count restarts
    got = ours.wait_critic_loss()
    torch.cuda.synchronize()
    assert got == float(ours.last_critic_loss[0])
