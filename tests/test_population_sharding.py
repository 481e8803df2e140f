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
