"""Sharding a population of independent agents over GPUs (BASELINE config 5: 64 agents, 8 per GPU).

Independent agents/seeds are the unit that shards naturally: no data-path collective, one process per GPU, every
process steps its members in lock-step (``TD3(..., n_agents=k)``).  What has to be global is the identity of an agent:
member ``i`` of a population keyed ``seed`` draws its replay indices and smoothing noise from the Philox stream
``seed + i * GOLDEN`` (csrc/misc.cuh), so a shard that starts at global agent ``a0`` must be keyed
``seed + a0 * GOLDEN`` for agent ``a0 + i`` to see the same stream whatever the number of GPUs -- which makes the
sharded run bit-identical, agent by agent, to the same population on one GPU
(tests/test_gpu_population.py::test_sharded_population_equals_the_unsharded_one).
"""
from __future__ import annotations

from typing import Tuple

GOLDEN = 0x9E3779B97F4A7C15          # per-member key stride of the device RNG (csrc/misc.cuh: gather_index)
_MASK = (1 << 64) - 1


def shard_range(n_agents_total: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[first, last) global agent ids owned by ``rank``: contiguous blocks, the first ``n % world`` ranks one larger."""
    n, w, r = int(n_agents_total), int(world_size), int(rank)
    if n < 1 or w < 1 or not 0 <= r < w:
        raise ValueError(f"bad shard request: {n} agents, world {w}, rank {r}")
    base, extra = divmod(n, w)
    first = r * base + min(r, extra)
    return first, first + base + (1 if r < extra else 0)


def shard_seed(seed: int, first_agent: int) -> int:
    """Philox key of a shard whose member 0 is global agent ``first_agent`` of the population keyed ``seed``."""
    return (int(seed) + int(first_agent) * GOLDEN) & _MASK


def member_seed(seed: int, agent: int) -> int:
    """Key a standalone (n_agents = 1) agent needs to reproduce global agent ``agent`` of the population keyed ``seed``."""
    return shard_seed(seed, agent)
