"""On-disk compatibility with artefacts WRITTEN BY THE REFERENCE CLASSES: tests/golden/ondisk/ holds replay
buffer folders saved by my_replay_buffer.ReplayBuffer_featured / _particles (my_replay_buffer.py:28-36,91-99) and policy
checkpoints saved by TD3_base.save (TD3_base.py:26-34) from the reference's TD3_featured (norm="layer") and TD3_particles
(norm="weight_normalization": weight_g / weight_v keys) after three updates, so the Adam files carry real state.  The
generator (oracle/make_golden.py: ondisk_cases) runs the reference with its hard-coded hidden widths edited to small ones
in memory so that the files are small; everything that defines the format is the reference's own code.
expect.npz = what the reference computes from those files."""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import td3_oracle as O

pytestmark = pytest.mark.gpu
DISK = os.path.join(GOLDEN, "ondisk")


@pytest.fixture(scope="module")
def expect():
    return np.load(os.path.join(DISK, "expect.npz"))


def test_reference_written_featured_buffer_loads_and_samples_bit_exactly(expect):
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    # default max_size, like experience_injection.py: the stored 37-row ring is adopted
    rb = ReplayBuffer_featured(O.Space(5), O.Space(2), load_folder=os.path.join(DISK, "buffer_featured"))
    assert (rb.ptr, rb.size) == tuple(expect["bf_ptr_size"]) and rb.max_size == 37
    got = rb.sample(32, indices=expect["bf_indices"])
    for k, v in zip(O.ReplayFeatured.fields, got):
        assert np.array_equal(v.cpu().numpy(), expect["bf_" + k]), k
    rb.set_rows("reward", 3, [[7.5]])
    assert rb.get_rows("reward", 3, 4)[0, 0] == 7.5 and rb.reward[3, 0] == 7.5


def test_reference_written_particles_buffer_loads_and_samples_bit_exactly(expect, tmp_path):
    from td3_b200.my_replay_buffer import ReplayBuffer_particles
    obs = (O.Space(3), O.Space(6, 4))
    rb = ReplayBuffer_particles(obs, O.Space(2), max_size=9, load_folder=os.path.join(DISK, "buffer_particles"))
    assert (rb.ptr, rb.size) == tuple(expect["bp_ptr_size"])
    got = rb.sample(16, indices=expect["bp_indices"])
    for k, v in zip(O.ReplayParticles.fields, got):
        assert np.array_equal(v.cpu().numpy(), expect["bp_" + k]), k
    # and back: what we save has the reference's files, dtypes and shapes; the payloads are the reference's float64 arrays
    # rounded to float32 (the device ring stores what sample() would have produced: FloatTensor(float64) at :61-69)
    rb.save(str(tmp_path))
    names = sorted(os.listdir(os.path.join(DISK, "buffer_particles")))
    assert names == sorted(os.listdir(str(tmp_path)))
    for name in names:
        fa, fb = os.path.join(DISK, "buffer_particles", name), os.path.join(str(tmp_path), name)
        if name in ("ptr.pkl", "size.pkl"):
            assert open(fa, "rb").read() == open(fb, "rb").read(), name
        else:
            a, b = np.load(open(fa, "rb")), np.load(open(fb, "rb"))
            assert a.dtype == b.dtype == np.float64 and a.shape == b.shape, name
            assert np.array_equal(a.astype(np.float32), b.astype(np.float32)), name


def _agent_and_buffer(tag, meta, widths):
    if meta["kind"] == "featured":
        from td3_b200.TD3_featured import TD3
        from td3_b200.my_replay_buffer import ReplayBuffer_featured
        obs, act = O.Space(meta["S"]), O.Space(meta["A"])
        agent = TD3(obs, act, norm=meta["norm"], lr=meta["lr"], actor_widths=widths["actor"], q_widths=widths["q"], precision="fp32", seed=1)
        rb = ReplayBuffer_featured(obs, act, max_size=meta["rows"])
        rb.add_batch(**O.synthetic_transitions_featured(meta["rows"], meta["S"], meta["A"], seed=meta["data_seed"]))
        st, ac = np.linspace(-1, 1, meta["S"]), np.linspace(-0.5, 0.5, meta["A"])
    else:
        from td3_b200.TD3_particles import TD3
        from td3_b200.my_replay_buffer import ReplayBuffer_particles
        obs, act = (O.Space(meta["F"]), O.Space(meta["N"], meta["D"])), O.Space(meta["A"])
        agent = TD3(obs, act, norm=meta["norm"], lr=meta["lr"], actor_widths=widths["pq"], q_widths=widths["pq"], precision="fp32", seed=1)
        rb = ReplayBuffer_particles(obs, act, max_size=meta["rows"])
        rb.add_batch(**O.synthetic_transitions_particles(meta["rows"], meta["F"], meta["N"], meta["D"], meta["A"], seed=meta["data_seed"]))
        r2 = np.random.RandomState(3)
        st, ac = (r2.standard_normal(meta["F"]), r2.standard_normal((meta["N"], meta["D"]))), np.linspace(-0.5, 0.5, meta["A"])
    return agent, rb, st, ac


@pytest.mark.parametrize("tag", ["ckpt_featured_layer", "ckpt_particles_wn"])
def test_reference_written_checkpoint_loads_and_continues_like_the_reference(expect, tag, tmp_path):
    meta = ast.literal_eval(str(expect[tag + "_meta"]))
    widths = ast.literal_eval(str(expect["widths"]))
    agent, rb, st, ac = _agent_and_buffer(tag, meta, widths)
    agent.load(os.path.join(DISK, tag))
    agent.total_it = meta["total_it_at_save"]                    # the reference does not persist total_it either (TD3_base.py:26-34)
    np.testing.assert_allclose(agent.select_action(st), expect[tag + "_select_action"], rtol=2e-5, atol=2e-6)
    for got, want in zip(agent.eval_q(st, ac), expect[tag + "_eval_q"]):
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-6)
    # the update the reference made next (a policy step) -- needs the imported Adam moments AND step counts to be right:
    # with a wrong step the bias corrections, with wrong moments every parameter would move differently by O(lr)
    agent.train(rb, meta["B"], indices=expect[tag + "_indices"], noise=expect[tag + "_noise"])
    for net in ("actor", "critic", "actor_target", "critic_target"):
        for k, v in getattr(agent, net).state_dict().items():
            want = expect[f"{tag}_after_{net}.{k}"]
            got = v.detach().cpu().numpy()
            err = np.abs(got - want).max()
            assert err <= 0.05 * meta["lr"], f"{net}.{k}: max |d| {err:.3e} after the continued update (lr {meta['lr']})"
    # and back: our checkpoint of this state loads into the reference-shaped CPU modules (oracle) key for key
    agent.save(str(tmp_path))
    torch.manual_seed(0)
    if meta["kind"] == "featured":
        ora = O.TD3Featured(O.Space(meta["S"]), O.Space(meta["A"]), norm=meta["norm"], lr=meta["lr"], actor_widths=widths["actor"], q_widths=widths["q"])
    else:
        ora = O.TD3Particles((O.Space(meta["F"]), O.Space(meta["N"], meta["D"])), O.Space(meta["A"]), norm=meta["norm"], lr=meta["lr"],
                             actor_widths=widths["pq"], q_widths=widths["pq"])
    for name in ("actor", "critic", "actor_target", "critic_target"):
        sd = torch.load(os.path.join(str(tmp_path), name), map_location="cpu")
        getattr(ora, name).load_state_dict(sd)                   # strict: keys and shapes must match
    for name in ("actor_optimizer", "critic_optimizer"):
        sd = torch.load(os.path.join(str(tmp_path), name), map_location="cpu")
        getattr(ora, name).load_state_dict(sd)
        assert all(int(float(s["step"])) > 0 for s in sd["state"].values())
    np.testing.assert_allclose(ora.select_action(st), agent.select_action(st), rtol=2e-5, atol=2e-6)
