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
