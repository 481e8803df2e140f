"""Large-batch data-parallel TD3 update (BASELINE config 5b: one 400-300 critic at global batch 8192).

Every rank holds a full replica (parameters, optimiser state, replay buffer) and computes the update on its slice of
the global batch: global row j is drawn by Philox element j whatever the world size, rank r takes rows
[r*B/W, (r+1)*B/W).  Losses are normalised by 1/B_global on every rank, so ONE sum of the packed critic gradient over
the ranks (and, on policy steps, one of the packed actor gradient -- it depends on the stepped critic, so the two
cannot be merged) makes every rank take the identical Adam step.  No other communication.

Two ways to form that sum (``mode``):

* ``"p2p"`` (default when the ranks can map each other's memory: one NVSwitch domain) -- the gradient buffers live in
  symmetric memory; the Adam kernel of every rank READS the W peer gradients over NVLink and adds them in rank order
  while it steps (csrc/misc.cuh: adam_polyak_body, EwRange::g_peers), so there is no all-reduce pass, no second copy
  of the gradient and every replica computes the bit-identical sum.  Ranks meet at two flag exchanges per reduction
  (td3::dp_signal_wait_kernel): "my gradient is complete" before the reads and "I am done reading" before the next
  backward pass overwrites it.
* ``"nccl"`` -- ``ncclAllReduce`` on the packed gradient, in-stream.

Either way the whole update -- sampling, target step, critic step, reduction, Adam, and the actor half on policy
steps -- is captured once into two CUDA graphs (critic-only / policy update) and replayed: one graph launch per update
instead of ~40 stage launches from Python.

The reference has no distributed path (single process, single device); this is the natural sharding of
TD3.train (TD3_featured.py:123-171) over the batch dimension.
"""
from __future__ import annotations

import ctypes as C
import os

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

    def __init__(self, agent, process_group=None, mode=None, use_graph=True):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.agent, self.group = agent, process_group
        agent._dp_owner = True           # the agent's own train() now refuses (its plan carries a global batch / shard offset);
                                         # a flag, not a back-reference: a cycle would keep the symmetric-memory buffers alive past `del`
        self.world, self.rank = dist.get_world_size(process_group), dist.get_rank(process_group)
        self._configured = None
        self.communicate = True          # False: skip the reduction (timing of the compute alone; replicas diverge)
        self.use_graph = bool(use_graph) and dist.get_backend(process_group) == "nccl"
        self._graphs = {}
        mode = mode or os.environ.get("TD3_DP_MODE") or ("p2p" if dist.get_backend(process_group) == "nccl" else "nccl")
        if mode not in ("p2p", "nccl"):
            raise ValueError("mode must be 'p2p' or 'nccl'")
        self.mode = mode
        self._symm = None
        if mode == "p2p":
            try:
                self._bind_symmetric_gradients()
            except Exception as exc:     # no peer access / no symmetric memory on this system: the collective path still works
                self.mode, self.p2p_unavailable = "nccl", repr(exc)[:200]

    # ------------------------------------------------------------------ p2p: gradients in symmetric memory
    def _bind_symmetric_gradients(self):
        import torch.distributed._symmetric_memory as symm_mem
        a = self.agent
        fa, fc = a._actor_family, a._critic_family
        na, nc = fa.grad.numel(), fc.grad.numel()
        group = self.group or dist.group.WORLD
        # one allocation: [critic gradient | actor gradient | 64 flag words]
        try:
            symm_mem.enable_symm_mem_for_group(group.group_name)
        except Exception:
            pass
        buf = symm_mem.empty(nc + na + 64, dtype=torch.float32, device=a._device)
        buf.zero_()
        hdl = symm_mem.rendezvous(buf, group.group_name)
        self._symm = (buf, hdl)
        fc.grad, fa.grad = buf[:nc], buf[nc:nc + na]
        ptrs = [int(p) for p in hdl.buffer_ptrs]
        W = self.world
        arr = lambda vals: (C.c_void_p * W)(*[C.c_void_p(v) for v in vals])
        a_ps, c_ps = fa.param_set(), fc.param_set()
        torch.cuda.synchronize()
        _lib.check(a._lib.td3_agent_bind_params(a._handle, C.byref(a_ps), C.byref(c_ps)))
        _lib.check(a._lib.td3_dp_bind_peers(a._handle, W, self.rank, arr(ptrs), arr([p + 4 * nc for p in ptrs]),
                                            arr([p + 4 * (nc + na) for p in ptrs])))
        a._planned_batch = 0             # bind_params dropped the plan
        dist.barrier(self.group)

    # ------------------------------------------------------------------ plan
    def _configure(self, global_batch):
        a = self.agent
        if self._configured == global_batch and a._planned_batch == self.local_batch:
            return
        lo, hi = shard_bounds(global_batch, self.world, self.rank)
        a._ensure_plan(hi - lo)
        _lib.check(a._lib.td3_agent_set_global_batch(a._handle, global_batch, lo))
        self._configured, self.local_batch = global_batch, hi - lo
        self._graphs = {}

    local_batch = -1

    def _enqueue(self, view, policy_step: bool):
        """One update as stream-ordered launches on the current stream (also what gets captured)."""
        a = self.agent
        s = _lib.stream_ptr()
        lib, h = a._lib, a._handle
        p2p = self.mode == "p2p" and self.communicate
        _lib.check(lib.td3_sample_batch(h, C.byref(view), _lib.RNG_PHILOX, s))
        _lib.check(lib.td3_target_step(h, s))
        _lib.check(lib.td3_critic_step(h, 0, s))                      # forward, loss, backward: local gradient / B_global
        if self.communicate and not p2p:
            dist.all_reduce(a._critic_family.grad, op=dist.ReduceOp.SUM, group=self.group)
        _lib.check(lib.td3_critic_apply(h, s))                        # p2p: peers' gradients are summed inside the Adam kernel
        if policy_step:
            _lib.check(lib.td3_actor_step(h, 0, s))
            if self.communicate and not p2p:
                dist.all_reduce(a._actor_family.grad, op=dist.ReduceOp.SUM, group=self.group)
            _lib.check(lib.td3_actor_apply(h, s))

    def train(self, replay_buffer, global_batch):
        """One update on the global batch; returns None.  Reductions: the critic gradient every call, the actor gradient
        on every policy_freq-th call."""
        a = self.agent
        self._configure(int(global_batch))
        view = a._rb_view(replay_buffer)
        if view.size <= 0:
            raise ValueError("high <= 0")
        a.total_it += 1
        policy_step = a.total_it % a.policy_freq == 0
        if self.mode == "p2p":           # re-plans (outside any capture) when the setting changes
            _lib.check(a._lib.td3_dp_set_fused_reduce(a._handle, 1 if self.communicate else 0))
        if not self.use_graph:
            self._enqueue(view, policy_step)
            return
        key = (policy_step, self.communicate, self.mode, int(view.rows or 0))
        # what an update settles lazily on entry (replay size word, TF32 copies) happens here, outside the graph
        _lib.check(a._lib.td3_agent_prepare(a._handle, C.byref(view), _lib.stream_ptr()))
        g = self._graphs.get(key)
        if g is None:
            # the first update runs eagerly (lazy allocations, NCCL channels); later ones are captured once and replayed
            if not self._graphs:
                self._enqueue(view, policy_step)
                torch.cuda.synchronize()
                self._graphs["warm"] = None
                return
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue(view, policy_step)
            self._graphs[key] = g
        g.replay()

    def global_critic_loss(self) -> torch.Tensor:
        """Sum over ranks of the local partial losses (each already divided by B_global)."""
        loss = self.agent.last_critic_loss.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss
