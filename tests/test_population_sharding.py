"""Host logic of sharding independent agents over ranks (td3_b200/population.py): CPU only."""
import pytest

from td3_b200.population import GOLDEN, member_seed, shard_range, shard_seed


@pytest.mark.parametrize("n,world", [(64, 8), (64, 1), (10, 4), (3, 8), (8, 8)])
def test_shards_partition_the_population(n, world):
    owned = []
    for r in range(world):
        a, b = shard_range(n, world, r)
        assert 0 <= a <= b <= n and b - a in (n // world, n // world + 1)
        owned += list(range(a, b))
    assert owned == list(range(n))


def test_member_keys_do_not_depend_on_the_number_of_ranks():
    seed, n = 1234567, 64
    one_gpu = [member_seed(seed, g) for g in range(n)]
    for world in (2, 4, 8):
        keys = []
        for r in range(world):
            a, b = shard_range(n, world, r)
            s = shard_seed(seed, a)
            keys += [(s + i * GOLDEN) % (1 << 64) for i in range(b - a)]     # what member i of the shard is keyed with
        assert keys == one_gpu
    assert len(set(one_gpu)) == n


def test_bad_requests_raise():
    with pytest.raises(ValueError):
        shard_range(0, 1, 0)
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


# ---- the N > 1 launch shape of bench.py on CPU: one process per rank, gloo, world size 2 -------------------------------
def _free_port():
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_worker(rank, world, port, out):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_pop, seed = 8, 1001
    first, last = shard_range(world * n_pop, world, rank)
    keys = torch.tensor([(shard_seed(seed, first) + i * GOLDEN) % (1 << 64) - (1 << 63) for i in range(last - first)],
                        dtype=torch.int64)                                   # shifted into int64 range for the collective
    gathered = [torch.zeros_like(keys) for _ in range(world)]
    dist.all_gather(gathered, keys)
    # the timing reduction bench.py does: whole-job throughput = all ranks' units / max-over-ranks time
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put(([int(k) + (1 << 63) for g in gathered for k in g], float(t[0])))
    dist.destroy_process_group()


def test_two_ranks_own_disjoint_agents_with_rank_independent_keys_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    keys, t_max = out.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert keys == [member_seed(1001, g) for g in range(16)]
    assert t_max == 2.0


def test_run_population_follows_the_reference_loop():
    """Control flow of td3_b200.population.run_population (main.py:240-289 for a population) with duck-typed stand-ins:
    random actions then policy actions + noise, one add per member per step into its own ring, one lock-step train per
    step after start_timesteps, episode bookkeeping per member."""
    import numpy as np
    from td3_b200.population import run_population

    class Env:
        def __init__(self, horizon):
            self.h, self.t = horizon, 0

        def reset(self):
            self.t = 0
            return np.zeros(3)

        def step(self, a):
            self.t += 1
            return np.full(3, self.t, dtype=np.float64), 1.0, self.t >= self.h, {}

    class Policy:
        n_agents, max_action = 2, 1.0

        def __init__(self):
            self.sel, self.trains = [], 0

        def select_action(self, s, agent=0):
            self.sel.append(agent)
            return np.array([0.5, -0.5])

        def train(self, rb, batch):
            self.trains += 1
            assert batch == 8 and all(c >= 5 for c in rb.count)

    class Buffer:
        def __init__(self):
            self.count, self.rows = [0, 0], []

        def add(self, s, a, s2, r, d, agent=0):
            assert -1.0 <= a.min() and a.max() <= 1.0 and len(a) == 2
            self.count[agent] += 1
            self.rows.append((agent, float(s2[0]), d))

    pol, rb = Policy(), Buffer()
    rets = run_population(pol, rb, [Env(4), Env(6)], max_timesteps=12, start_timesteps=5, batch_size=8, action_dim=2)
    assert rb.count == [12, 12] and pol.trains == 7
    assert pol.sel == [0, 1] * 7                      # policy actions only after the random phase, members in order
    assert rets == [[4.0, 4.0, 4.0], [6.0, 6.0]]
    assert [r for r in rb.rows if r[0] == 0][3] == (0, 4.0, 1.0)      # member 0's fourth transition ends its episode
