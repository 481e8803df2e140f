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

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

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


def run_population(policy, replay_buffer, envs: Sequence, max_timesteps: int, start_timesteps: int = 1000, batch_size: int = 256,
                   expl_noise: float = 0.1, seed: int = 0, on_episode: Optional[Callable] = None,
                   action_dim: Optional[int] = None) -> List[List[float]]:
    """The reference's collect/train loop (main.py:240-289) for a population: member ``i`` of ``policy`` (a td3_b200 agent
    built with ``n_agents = len(envs)``) acts in ``envs[i]`` and stores its transitions in ring ``i`` of ``replay_buffer``;
    once every member has ``start_timesteps`` transitions, ONE lock-step ``policy.train`` per timestep updates all
    members (each from its own ring, with its own Philox stream).  Members never exchange anything, so a larger
    population is sharded over GPUs by giving each process its ``shard_range`` of environments and a ``shard_seed`` key.

    ``envs[i]`` needs ``reset() -> state`` and ``step(action) -> (next_state, reward, done, info)`` (the gym API main.py
    uses); exploration is uniform random actions for the first ``start_timesteps`` steps (main.py:245) and Gaussian
    action noise afterwards (the reference's OU process, utils/noise.py, is a caller-side choice and can be applied by
    wrapping the environment).  Returns the list of episode returns of every member.
    """
    n = len(envs)
    if n != policy.n_agents:
        raise ValueError(f"{n} environments for a population of {policy.n_agents}")
    rs = np.random.RandomState(seed)
    max_action = float(policy.max_action)
    a_dim = int(action_dim if action_dim is not None else policy._cfg.action_dim)
    states = [e.reset() for e in envs]
    ep_ret = [0.0] * n
    returns: List[List[float]] = [[] for _ in range(n)]
    for t in range(int(max_timesteps)):
        for i, env in enumerate(envs):
            if t < start_timesteps:                                   # main.py:244-245
                action = rs.uniform(-max_action, max_action, size=a_dim)
            else:                                                     # main.py:247-252
                action = np.asarray(policy.select_action(states[i], agent=i), dtype=np.float64)
                action = (action + rs.normal(0.0, max_action * expl_noise, size=action.shape)).clip(-max_action, max_action)
            nxt, reward, done, _ = env.step(action)                    # main.py:254
            replay_buffer.add(states[i], action, nxt, reward, float(done), agent=i)     # main.py:261
            states[i] = nxt
            ep_ret[i] += float(reward)
            if done:                                                  # main.py:273-283
                returns[i].append(ep_ret[i])
                if on_episode is not None:
                    on_episode(i, t, ep_ret[i])
                states[i], ep_ret[i] = env.reset(), 0.0
        if t >= start_timesteps:                                      # main.py:268-269
            policy.train(replay_buffer, batch_size)
    return returns
