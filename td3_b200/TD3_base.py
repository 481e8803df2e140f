"""Base agent: hyper-parameters, checkpoint I/O and the C-ABI plumbing shared by both variants.

Public surface == the reference's TD3_base (TD3_base.py:6-50): the six hyper-parameters,
``total_it``, ``save(folder)`` / ``load(folder)`` writing/reading six ``torch.save``d state_dicts.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Optional

import numpy as np
import torch

from . import _lib
from .packing import PackedAdam, PackedFamily


class TD3_base(object):
    def __init__(self, max_action=1, discount=0.99, tau=0.005, policy_noise=0.2, noise_clip=0.5, policy_freq=2):
        # TD3_base.py:17-24
        self.max_action = max_action
        self.discount = discount
        self.tau = tau
        self.policy_noise = policy_noise
        self.noise_clip = noise_clip
        self.policy_freq = policy_freq
        self.total_it = 0

    # ------------------------------------------------------------------ checkpoint I/O (TD3_base.py:26-50)
    def save(self, folder):
        os.makedirs(folder, exist_ok=True)
        torch.save(self.critic.state_dict(), os.path.join(folder, "critic"))
        torch.save(self.critic_target.state_dict(), os.path.join(folder, "critic_target"))
        torch.save(self.critic_optimizer.state_dict(), os.path.join(folder, "critic_optimizer"))
        torch.save(self.actor.state_dict(), os.path.join(folder, "actor"))
        torch.save(self.actor_target.state_dict(), os.path.join(folder, "actor_target"))
        torch.save(self.actor_optimizer.state_dict(), os.path.join(folder, "actor_optimizer"))

    def load(self, folder):
        dev = self._device
        rd = lambda name: torch.load(os.path.join(folder, name), map_location=dev)
        self.critic.load_state_dict(rd("critic"))
        self.critic_optimizer.load_state_dict(rd("critic_optimizer"))
        if os.path.isfile(os.path.join(folder, "critic_target")):
            self.critic_target.load_state_dict(rd("critic_target"))
        else:                                            # :43 -- target := copy of the online net
            self._critic_family.rebind_target(self.critic)
        self.actor.load_state_dict(rd("actor"))
        self.actor_optimizer.load_state_dict(rd("actor_optimizer"))
        if os.path.isfile(os.path.join(folder, "actor_target")):
            self.actor_target.load_state_dict(rd("actor_target"))
        else:
            self._actor_family.rebind_target(self.actor)
        self.params_changed()

    def params_changed(self):
        """Tell the engine that parameter values were written from outside its kernels (``load_state_dict`` and
        ``load`` call this themselves; call it after editing a ``state_dict()`` view in place): in TF32 mode the
        tensor cores read round-to-nearest copies of the weights, which are rebuilt before the next update."""
        if getattr(self, "_handle", None) is not None:
            _lib.check(self._lib.td3_agent_params_changed(self._handle))

    # ------------------------------------------------------------------ engine plumbing
    def _engine_init(self, cfg: _lib.AgentConfig, actor_family: PackedFamily, critic_family: PackedFamily,
                     lr: float, rng: str):
        """Create the C agent, bind buffers.  Called by the variant constructors once hyper-parameters are set."""
        self._lib = _lib.require_cuda()
        self._device = actor_family.params.device
        self._actor_family, self._critic_family = actor_family, critic_family
        self._cfg = cfg
        if rng not in ("device", "host"):
            raise ValueError("rng must be 'device' (on-device Philox) or 'host' (reference's NumPy/torch CPU streams)")
        self.rng = rng
        # how td3_train_n executes an update: "graph" = CUDA-graph replay of one kernel per stage (default: measured
        # fastest on B200, 103 vs 118 us per cfg2 update), "persistent" = one cooperative kernel walking the same stage
        # program with device-wide barriers, "launches" = plain stage-by-stage launches (debugging)
        self.exec_mode = os.environ.get("TD3_EXEC_MODE", "graph")
        n_agents = cfg.n_agents
        self._state = torch.zeros(16 + n_agents, dtype=torch.int64, device=self._device)
        self._losses = self._state[16:].view(torch.float32)          # critic_loss[nA], actor_loss[nA]
        handle = C.c_void_p()
        _lib.check(self._lib.td3_agent_create(C.byref(cfg), C.byref(handle)))
        self._handle = handle
        a_ps, c_ps = actor_family.param_set(), critic_family.param_set()
        _lib.check(self._lib.td3_agent_bind_params(handle, C.byref(a_ps), C.byref(c_ps)))
        _lib.check(self._lib.td3_agent_bind_state(handle, C.c_void_p(self._state.data_ptr()), self._state.numel() * 8))
        # host mirror of the critic loss: 8-byte words {fp32 loss, update count} in pinned memory that the fused critic
        # head writes over PCIe as soon as the loss exists (wait_critic_loss)
        self._host_status = torch.zeros(n_agents, dtype=torch.int64).pin_memory()
        self._host_words = self._host_status.numpy()
        _lib.check(self._lib.td3_agent_bind_host_status(handle, C.c_void_p(self._host_status.data_ptr())))
        self._seq_expected = 0
        self._status_live = False
        self._workspace: Optional[torch.Tensor] = None
        self._planned_batch = 0
        self._global_batch = 0
        self.critic_optimizer = PackedAdam(critic_family, lr, lambda: self._state[1], lambda t: self._set_adam_step(0, t, lr))
        self.actor_optimizer = PackedAdam(actor_family, lr, lambda: self._state[2], lambda t: self._set_adam_step(1, t, lr))

    def _set_adam_step(self, which: int, step: int, lr: float, betas=(0.9, 0.999)):
        """Checkpoint restore: set the device-resident Adam step and the scalars derived from it (the layout
        adam_tick maintains: csrc/misc.cuh)."""
        step = int(step)
        self._state[1 + which] = step
        pw = self._state.view(torch.float64)
        pw[4 + 2 * which], pw[5 + 2 * which] = betas[0] ** step, betas[1] ** step
        if step > 0:
            sc = self._state.view(torch.float32)
            sc[20 + 2 * which] = lr / (1.0 - betas[0] ** step)
            sc[21 + 2 * which] = (1.0 - betas[1] ** step) ** 0.5

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and getattr(self, "_lib", None) is not None:
            try:
                torch.cuda.synchronize()
                self._lib.td3_agent_destroy(h)
            except Exception:
                pass
            self._handle = None

    def _ensure_plan(self, batch: int):
        """(Re)allocate the workspace and build the launch plan when the batch size grows."""
        batch = int(batch)
        if batch <= 0:
            raise ValueError("batch_size must be positive")
        if batch == self._planned_batch:
            return
        need = self._lib.td3_agent_workspace_floats(self._handle, batch)
        if need < 0:
            _lib.check(-1)
        if self._workspace is None or self._workspace.numel() < need:
            torch.cuda.synchronize()
            self._workspace = None
            self._workspace = torch.zeros(int(need), dtype=torch.float32, device=self._device)
        _lib.check(self._lib.td3_agent_plan(self._handle, batch, C.c_void_p(self._workspace.data_ptr()),
                                            self._workspace.numel(), _lib.stream_ptr()))
        self._planned_batch = batch
        torch.cuda.synchronize()
        self._host_status.zero_()                      # the plan starts its update count from zero
        self._seq_expected = 0
        self._status_live = bool(self._lib.td3_agent_host_status_live(self._handle))

    def _forward_chunks(self, B: int):
        """Row ranges for a forward pass over B caller rows that never re-plans an agent that already trains: the planned
        batch is kept (a re-plan re-captures the update graphs and resets the host-status sequence) and larger inputs go
        through in slices of it.  An agent without a plan gets one for B rows."""
        if self._planned_batch <= 0:
            self._ensure_plan(B)
        step = self._planned_batch
        return [(lo, min(B, lo + step)) for lo in range(0, B, step)]

    def _region(self, name: str) -> torch.Tensor:
        off, n = C.c_int64(), C.c_int64()
        _lib.check(self._lib.td3_agent_region(self._handle, name.encode(), C.byref(off), C.byref(n)))
        return self._workspace[off.value: off.value + n.value]

    def _rb_view(self, replay_buffer) -> _lib.ReplayView:
        view = getattr(replay_buffer, "_view", None)
        if view is None:
            raise TypeError("train() needs a device-resident td3_b200.my_replay_buffer.ReplayBuffer_* instance")
        return view()

    def _inject(self, batch, indices, noise):
        """Fill the injection regions with this update's indices / N(0,1) draws (parity mode)."""
        nA, A = self._cfg.n_agents, self._cfg.action_dim
        idx = torch.as_tensor(np.asarray(indices), dtype=torch.int64).reshape(nA, batch)
        nz = torch.as_tensor(np.asarray(noise), dtype=torch.float32).reshape(nA, batch, A)
        self._region("indices_in").view(torch.int64)[: nA * batch].copy_(idx.reshape(-1))
        self._region("noise_in")[: nA * batch * A].copy_(nz.reshape(-1))

    _EXEC_MODES = {"launches": 0, "graph": 1, "persistent": 2}

    def _train_common(self, replay_buffer, batch_size, iterations, indices, noise, use_graph):
        batch_size, iterations = int(batch_size), int(iterations)
        use_graph = self._EXEC_MODES[self.exec_mode] if use_graph else 0
        view = self._rb_view(replay_buffer)
        if view.size <= 0:
            raise ValueError("high <= 0")        # what np.random.randint(0, 0) raises (my_replay_buffer.py:59,120)
        if getattr(self, "_dp_owner", None) is not None:
            raise RuntimeError("this agent is driven by DataParallelTD3 (global batch, gradient sums): call its train()")
        self._ensure_plan(batch_size)
        s = _lib.stream_ptr()
        injected = indices is not None or noise is not None or self.rng == "host"
        if not injected:
            rc = self._lib.td3_train_n(self._handle, C.byref(view), self.total_it, iterations, _lib.RNG_PHILOX, use_graph, s)
            if rc:
                _lib.check(rc)
            self.total_it += iterations
            self._seq_expected += iterations
            return
        if (indices is not None or noise is not None) and iterations != 1:
            raise ValueError("indices/noise injection covers exactly one update (iterations=1)")
        nA, A = self._cfg.n_agents, self._cfg.action_dim
        for _ in range(iterations):
            # same draw order as the reference: indices inside sample() (my_replay_buffer.py:120), then noise
            # (TD3_featured.py:132), both from the global host generators
            idx = np.random.randint(0, view.size, size=(nA, batch_size)) if indices is None else indices
            nz = torch.randn(nA, batch_size, A) if noise is None else noise
            self._inject(batch_size, idx, nz)
            _lib.check(self._lib.td3_train_n(self._handle, C.byref(view), self.total_it, 1, _lib.RNG_INJECTED,
                                             int(use_graph), s))
            self.total_it += 1
            self._seq_expected += 1

    # ------------------------------------------------------------------ population members (n_agents > 1)
    @property
    def n_agents(self) -> int:
        return self._cfg.n_agents

    def agent_state_dict(self, net: str, agent: int):
        """Reference-keyed views of one population member's ``net`` in {"actor","critic","actor_target","critic_target"}."""
        fam = self._actor_family if net.startswith("actor") else self._critic_family
        return fam.agent_state_dict(int(agent), target=net.endswith("_target"))

    def load_agent_state_dict(self, net: str, agent: int, state_dict):
        fam = self._actor_family if net.startswith("actor") else self._critic_family
        fam.load_agent(int(agent), state_dict, which="target" if net.endswith("_target") else "online")
        self.params_changed()

    # ------------------------------------------------------------------ diagnostics (device tensors, no sync)
    @property
    def last_critic_loss(self) -> torch.Tensor:
        return self._losses[: self._cfg.n_agents]

    @property
    def last_actor_loss(self) -> torch.Tensor:
        return self._losses[self._cfg.n_agents: 2 * self._cfg.n_agents]

    def wait_critic_loss(self, timeout: float = 2.0):
        """Critic loss of the most recently enqueued update, as host numbers (float for one agent, float32 array for
        a population).  Blocks only until that loss has landed in pinned host memory -- the fused critic-head kernel
        stores it there together with the update count -- not until the optimiser kernels queued behind it have
        drained: the next ``add``/``train`` can be enqueued while they run.  Falls back to a stream synchronise and a
        D2H copy when the planned update has no fused head or other entry points advanced the device-side count."""
        nA = self._cfg.n_agents
        if self._status_live and self._seq_expected > 0:
            words, exp = self._host_words, self._seq_expected & 0xFFFFFFFF
            t_end = None
            while True:
                done, ahead = True, False
                for i in range(nA):
                    seq = (int(words[i]) >> 32) & 0xFFFFFFFF
                    if seq != exp:
                        done = False
                        ahead = 0 < ((seq - exp) & 0xFFFFFFFF) < 0x80000000   # other entry points ran updates too
                        break
                if ahead:
                    break
                if done:
                    bits = np.array([int(words[i]) & 0xFFFFFFFF for i in range(nA)], dtype=np.uint32).view(np.float32)
                    return float(bits[0]) if nA == 1 else bits
                if t_end is None:
                    t_end = time.perf_counter() + timeout
                elif time.perf_counter() > t_end:
                    self._status_live = False            # the mirror is not arriving on this system: stop relying on it
                    break
        torch.cuda.current_stream().synchronize()
        if self._status_live:                            # re-base on what the device has counted
            self._seq_expected = (int(self._host_words[0]) >> 32) & 0xFFFFFFFF
        out = self.last_critic_loss.cpu().numpy()
        return float(out[0]) if nA == 1 else out

    def chain_active(self) -> bool:
        """True when the planned update runs as the layer-fused chain launches (csrc/chain.cuh, TD3_CHAIN=1)."""
        return bool(self._lib.td3_agent_chain_active(self._handle))

    def debug_tensors(self):
        """Views of the last update's Q-values / Bellman target (tests)."""
        B, nq, nA = self._planned_batch, self._cfg.n_q, self._cfg.n_agents
        qw = self._cfg.q.dims[self._cfg.q.n_linear]
        return dict(q=self._region("q").view(nA, nq, B, qw), target_q=self._region("target_q").view(nA, B, qw),
                    indices=self._region("indices").view(torch.int64)[: nA * B].view(nA, B),
                    eps=self._region("eps").view(nA, B, self._cfg.action_dim))
