"""Large-batch data-parallel TD3 update (BASELINE config 5b: one 400-300 critic at global batch 8192).

Every rank holds a full replica (parameters, optimiser state, replay buffer) and computes the update on its slice of
the global batch: global row j is drawn by Philox element j whatever the world size, rank r takes rows
[r*B/W, (r+1)*B/W).  Losses are normalised by 1/B_global on every rank, so ONE sum-all-reduce of the packed critic
gradient (and, on policy steps, one of the packed actor gradient -- it depends on the stepped critic, so the two cannot be
merged) makes every rank take the identical Adam step.  NCCL in-stream over NVLink; no other communication.

The reference has no distributed path (single process, single device); this is the natural sharding of
TD3.train (TD3_featured.py:123-171) over the batch dimension.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(global_batch: int, world_size: int, rank: int):
    """[begin, end) of rank's rows in the global batch; the batch must divide evenly (as torch DDP requires)."""
    if global_batch % world_size:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    per = global_batch // world_size
    return rank * per, (rank + 1) * per


class DataParallelTD3(object):
    """Wraps a td3_b200 agent (one per rank, same seed and initial weights on every rank)."""

    def __init__(self, agent, process_group=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.agent, self.group = agent, process_group
        self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        self._configured = None

    def _configure(self, global_batch):
        if self._configured == global_batch:
            return
        lo, hi = shard_bounds(global_batch, self.world, self.rank)
        a = self.agent
        a._ensure_plan(hi - lo)
        _lib.check(a._lib.td3_agent_set_global_batch(a._handle, global_batch, lo))
        self._configured, self.local_batch = global_batch, hi - lo

    def train(self, replay_buffer, global_batch):
        """One update on the global batch; returns None.  Collectives: one all-reduce(sum) of the critic gradient, plus
        one of the actor gradient on every policy_freq-th call."""
        a = self.agent
        self._configure(int(global_batch))
        view = a._rb_view(replay_buffer)
        if view.size <= 0:
            raise ValueError("high <= 0")
        s = _lib.stream_ptr()
        lib, h = a._lib, a._handle
        a.total_it += 1
        _lib.check(lib.td3_sample_batch(h, C.byref(view), _lib.RNG_PHILOX, s))
        _lib.check(lib.td3_target_step(h, s))
        _lib.check(lib.td3_critic_step(h, 0, s))                      # forward, loss, backward: local gradient / B_global
        dist.all_reduce(a._critic_family.grad, op=dist.ReduceOp.SUM, group=self.group)
        _lib.check(lib.td3_critic_apply(h, s))
        if a.total_it % a.policy_freq == 0:
            _lib.check(lib.td3_actor_step(h, 0, s))
            dist.all_reduce(a._actor_family.grad, op=dist.ReduceOp.SUM, group=self.group)
            _lib.check(lib.td3_actor_apply(h, s))

    def global_critic_loss(self) -> torch.Tensor:
        """Sum over ranks of the local partial losses (each already divided by B_global)."""
        loss = self.agent.last_critic_loss.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss
