"""Populations (n_agents > 1): N independent agents stepped in lock-step by the same launches -- the unit that is
sharded over GPUs with no communication (BASELINE config 5: 64 agents, 8 per GPU).
Bar: every member's losses and parameters are BIT-IDENTICAL to a standalone agent given the same weights, data,
indices and noise (both precisions, persistent kernel and graph), and the on-device Philox streams of member i equal
those of a standalone agent keyed seed + i * 0x9E3779B97F4A7C15."""
import numpy as np
import pytest
import torch

from oracle import td3_oracle as O

pytestmark = pytest.mark.gpu

S, A, B, ROWS, N = 17, 6, 128, 1024, 3
KW = dict(actor_widths=(400, 300), q_widths=(400, 300), lr=1e-3)
NETS = ("actor", "critic", "actor_target", "critic_target")


def _build(precision, mode, seed=5):
    from td3_b200.TD3_featured import TD3
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    obs, act = O.Space(S), O.Space(A)
    torch.manual_seed(3)
    pop = TD3(obs, act, n_agents=N, precision=precision, seed=seed, **KW)
    pop.exec_mode = mode
    rbp = ReplayBuffer_featured(obs, act, max_size=ROWS, n_agents=N)
    singles = []
    for i in range(N):
        a = TD3(obs, act, precision=precision, seed=(seed + i * 0x9E3779B97F4A7C15) % (1 << 64), **KW)
        a.exec_mode = mode
        for k in NETS:
            getattr(a, k).load_state_dict({n: v.clone() for n, v in pop.agent_state_dict(k, i).items()})
        data = O.synthetic_transitions_featured(ROWS, S, A, seed=10 + i)
        rb = ReplayBuffer_featured(obs, act, max_size=ROWS)
        rb.add_batch(**data)
        rbp.add_batch(agent=i, **data)
        singles.append((a, rb))
    return pop, rbp, singles


def _assert_same(pop, singles):
    for i, (a, _) in enumerate(singles):
        assert float(pop.last_critic_loss[i].item()) == float(a.last_critic_loss[0].item()), i
        for k in NETS:
            for (n, v), w in zip(pop.agent_state_dict(k, i).items(), getattr(a, k).state_dict().values()):
                assert torch.equal(v, w), (i, k, n)


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
@pytest.mark.parametrize("mode", ["persistent", "graph"])
def test_members_match_standalone_agents_bit_exactly(precision, mode):
    pop, rbp, singles = _build(precision, mode)
    assert not torch.equal(pop.agent_state_dict("actor", 0)["linears.0.weight"], pop.agent_state_dict("actor", 1)["linears.0.weight"])
    rs = np.random.RandomState(1)
    for t in range(4):
        idx = rs.randint(0, ROWS, size=(N, B))
        nz = rs.standard_normal((N, B, A)).astype(np.float32)
        pop.train(rbp, B, indices=idx, noise=nz)
        for i, (a, rb) in enumerate(singles):
            a.train(rb, B, indices=idx[i], noise=nz[i])
    _assert_same(pop, singles)
    assert pop.total_it == 4


def test_member_philox_streams_equal_standalone_streams():
    pop, rbp, singles = _build("tf32", "persistent")
    pop.train(rbp, B, iterations=6)
    for a, rb in singles:
        a.train(rb, B, iterations=6)
    torch.cuda.synchronize()
    _assert_same(pop, singles)
    idx = pop.debug_tensors()["indices"].cpu().numpy()
    assert not np.array_equal(idx[0], idx[1])            # members draw different batches


def test_member_inference_addresses_its_own_weights():
    pop, rbp, singles = _build("fp32", "persistent")
    s = np.random.RandomState(0).standard_normal(S).astype(np.float32)
    for i, (a, _) in enumerate(singles):
        np.testing.assert_array_equal(pop.select_action(s, agent=i), a.select_action(s))
        u = pop.select_action(s, agent=i)
        for q_pop, q_one in zip(pop.eval_q(s, u, agent=i), a.eval_q(s, u)):
            np.testing.assert_array_equal(q_pop, q_one)


@pytest.mark.parametrize("mode", ["graph", "persistent", "launches"])
def test_host_mirror_of_the_losses_covers_every_member(mode):
    """wait_critic_loss() on a population returns every member's critic loss of the last enqueued update -- the
    8-byte words the fused head kernel stores to pinned host memory -- and they equal the device values, in every
    execution mode."""
    pop, rbp, _ = _build("tf32", mode)
    for it in (1, 1, 3):
        pop.train(rbp, B, iterations=it)
        got = pop.wait_critic_loss()
        torch.cuda.synchronize()
        want = pop.last_critic_loss.cpu().numpy()
        assert got.shape == (N,) and np.array_equal(got, want), (mode, it, got, want)


def test_sharded_population_equals_the_unsharded_one():
    """A population of 4 on one GPU and the same 4 global agents as two shards of 2 (keyed by population.shard_seed)
    give bit-identical members: what makes `64 agents over 8 GPUs` independent of the number of GPUs."""
    from td3_b200.TD3_featured import TD3
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    from td3_b200.population import shard_range, shard_seed
    obs, act = O.Space(S), O.Space(A)
    seed, total = 77, 4

    def run(first, count, init):
        torch.manual_seed(3)
        pop = TD3(obs, act, n_agents=count, precision="tf32", seed=shard_seed(seed, first), **KW)
        rb = ReplayBuffer_featured(obs, act, max_size=ROWS, n_agents=count)
        for i in range(count):
            for k in NETS:
                pop.load_agent_state_dict(k, i, init[first + i][k])
            rb.add_batch(agent=i, **O.synthetic_transitions_featured(ROWS, S, A, seed=20 + first + i))
        pop.train(rb, B, iterations=6)
        torch.cuda.synchronize()
        return [{k: {n: v.clone() for n, v in pop.agent_state_dict(k, i).items()} for k in NETS} for i in range(count)], \
            pop.last_critic_loss.cpu().numpy().copy()

    torch.manual_seed(11)
    donor = TD3(obs, act, n_agents=total, precision="tf32", seed=1, **KW)
    init = [{k: {n: v.clone() for n, v in donor.agent_state_dict(k, i).items()} for k in NETS} for i in range(total)]
    whole, loss_whole = run(0, total, init)
    parts, losses = [], []
    for r in range(2):
        a, b = shard_range(total, 2, r)
        p, l = run(a, b - a, init)
        parts += p
        losses += list(l)
    assert np.array_equal(np.asarray(losses), loss_whole)
    for g in range(total):
        for k in NETS:
            for n in whole[g][k]:
                assert torch.equal(whole[g][k][n], parts[g][k][n]), (g, k, n)


def test_particles_population_members_match_standalone_agents():
    """TD3_particles populations (round 2): two members stepped in lock-step equal two standalone agents bit for bit, with
    per-member add() into the members' own rings."""
    from td3_b200.TD3_particles import TD3 as PT
    from td3_b200.my_replay_buffer import ReplayBuffer_particles
    F, Np, D, Ap, rows, Bp = 4, 32, 3, 2, 48, 16
    obs, act = (O.Space(F), O.Space(Np, D)), O.Space(Ap)
    torch.manual_seed(9)
    pop = PT(obs, act, n_agents=2, precision="fp32", seed=5, lr=1e-3, actor_widths=(48, 32), q_widths=(48, 32))
    rbp = ReplayBuffer_particles(obs, act, max_size=rows, n_agents=2)
    singles = []
    for i in range(2):
        a = PT(obs, act, precision="fp32", seed=(5 + i * 0x9E3779B97F4A7C15) % (1 << 64), lr=1e-3, actor_widths=(48, 32), q_widths=(48, 32))
        for k in NETS:
            getattr(a, k).load_state_dict({n: v.clone() for n, v in pop.agent_state_dict(k, i).items()})
        data = O.synthetic_transitions_particles(rows, F, Np, D, Ap, seed=20 + i)
        rb = ReplayBuffer_particles(obs, act, max_size=rows)
        for r in range(rows):       # one row at a time through add(..., agent=i): the per-member ingest path
            row = ((data["state_features"][r], data["state_particles"][r]), data["action"][r],
                   (data["next_state_features"][r], data["next_state_particles"][r]), float(data["reward"][r]), float(data["done"][r]))
            rb.add(*row)
            rbp.add(*row, agent=i)
        singles.append((a, rb))
    rs = np.random.RandomState(2)
    for t in range(4):
        idx = rs.randint(0, rows, size=(2, Bp))
        nz = rs.standard_normal((2, Bp, Ap)).astype(np.float32)
        pop.train(rbp, Bp, indices=idx, noise=nz)
        for i, (a, rb) in enumerate(singles):
            a.train(rb, Bp, indices=idx[i], noise=nz[i])
    _assert_same(pop, singles)
    assert np.allclose(pop.select_action((data["state_features"][0], data["state_particles"][0]), agent=1),
                       singles[1][0].select_action((data["state_features"][0], data["state_particles"][0])))


def test_run_population_trains_every_member_on_the_device():
    """The population driver loop (td3_b200.population.run_population, main.py:240-289 generalised) on a toy environment:
    three members, own rings, lock-step updates after the random phase; every member's parameters move and stay finite."""
    from td3_b200.TD3_featured import TD3
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    from td3_b200.population import run_population

    class Toy:
        def __init__(self, seed):
            self.rs, self.t = np.random.RandomState(seed), 0

        def reset(self):
            self.t, self.s = 0, self.rs.standard_normal(S)
            return self.s

        def step(self, a):
            self.t += 1
            self.s = 0.9 * self.s + 0.1 * self.rs.standard_normal(S)
            return self.s, float(-np.square(a).sum()), self.t >= 25, {}

    obs, act = O.Space(S), O.Space(A)
    torch.manual_seed(1)
    pop = TD3(obs, act, n_agents=3, seed=3, **KW)
    rb = ReplayBuffer_featured(obs, act, max_size=512, n_agents=3)
    before = [pop.agent_state_dict("critic", i)["q1.linears.0.weight"].clone() for i in range(3)]
    rets = run_population(pop, rb, [Toy(i) for i in range(3)], max_timesteps=120, start_timesteps=64, batch_size=32)
    torch.cuda.synchronize()
    assert pop.total_it == 56 and all(len(r) == 4 for r in rets)
    assert rb.size == 120 and rb._sizes[1] == 120 and rb._sizes[2] == 120
    for i in range(3):
        after = pop.agent_state_dict("critic", i)["q1.linears.0.weight"]
        assert torch.isfinite(after).all() and not torch.equal(after, before[i])
    assert torch.isfinite(pop.last_critic_loss).all()
