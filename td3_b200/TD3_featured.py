"""TD3 with plain-MLP actor/critics on feature observations -- B200-native drop-in for the
reference's TD3_featured (TD3_featured.py:15-171).

Same constructor keywords and methods; extra keyword-only options:
  actor_widths / q_widths   hidden widths (defaults = the fork's hard-coded (500,400,300)/(500,400,200),
                            TD3_featured.py:19,54; BASELINE's "400-300" is widths=(400,300) for both)
  rng                       "device": Philox indices + noise generated on the GPU (default)
                            "host":   np.random.randint + torch.randn on the host, i.e. the reference's
                                      own random streams (seed-for-seed comparable with a CPU reference run)
  seed                      Philox key for rng="device" (default: drawn from torch's global generator)
  n_agents                  N > 1: a population of N independent agents (own weights, optimiser state, replay
                            buffer and Philox stream each) stepped in lock-step by the same launches; the module
                            shells (``policy.actor`` ...) and the reference-style methods address agent 0, the
                            ``agent=`` keyword and ``agent_state_dict`` the others (independent seeds)
  precision                 "tf32" (default; env TD3_PRECISION): the layer GEMMs with K >= 64 run on the tcgen05 tensor
                            cores in TF32 with fp32 accumulation; "fp32": every contraction in strict fp32 (parity mode)
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .TD3_base import TD3_base
from .packing import MlpActor, MlpCritic, PackedFamily, net_layout

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

Actor, Critic = MlpActor, MlpCritic


class TD3(TD3_base):
    def __init__(self, obs_space, action_space, max_action=1, lr=1e-4, norm=None, CDQ=True, *,
                 actor_widths=(500, 400, 300), q_widths=(500, 400, 200), rng="device", seed=None, precision=None,
                 n_agents=1, **kwargs):
        _lib.require_cuda()
        S, A = obs_space.shape[0], action_space.shape[0]
        n_agents = int(n_agents)
        if n_agents < 1:
            raise ValueError("n_agents must be >= 1")
        # Build on the CPU with torch's default initialisers in the reference's construction order
        # (actor, then critic q1, q2; TD3_featured.py:101-106) so that torch.manual_seed(s) gives the
        # same initial weights as the reference; targets start as copies (deepcopy at :102,107).
        actor = MlpActor(S, A, max_action, norm, actor_widths)
        actor_t = _clone_shell(actor)
        critic = MlpCritic(S, A, norm, q_widths)
        critic_t = _clone_shell(critic)
        super(TD3, self).__init__(max_action=max_action, **kwargs)      # :110 (CDQ accepted and ignored, :100)
        dev = torch.device("cuda", torch.cuda.current_device())
        fam_a = PackedFamily(actor, actor_t, [""], dev, n_agents)
        fam_c = PackedFamily(critic, critic_t, ["q1", "q2"], dev, n_agents)
        for i in range(1, n_agents):      # further members of the population: fresh default initialisations, in order
            fam_a.load_agent(i, MlpActor(S, A, max_action, norm, actor_widths).state_dict())
            fam_c.load_agent(i, MlpCritic(S, A, norm, q_widths).state_dict())
        self.actor, self.actor_target, self.critic, self.critic_target = actor, actor_t, critic, critic_t
        for m, w in ((actor, 0), (actor_t, 1), (critic, 0), (critic_t, 1)):
            m._attach(self, w)
        cfg = _lib.AgentConfig()
        cfg.variant, cfg.norm = _lib.VARIANT_FEATURED, (_lib.NORM_LAYER if norm == "layer" else _lib.NORM_NONE)
        cfg.n_q, cfg.state_dim, cfg.action_dim = 2, S, A
        cfg.n_particles = cfg.particle_dim = 0
        cfg.clamp_target_action, cfg.n_agents = 1, n_agents
        cfg.max_action, cfg.discount, cfg.tau = float(self.max_action), float(self.discount), float(self.tau)
        cfg.policy_noise, cfg.noise_clip = float(self.policy_noise), float(self.noise_clip)
        cfg.lr_actor = cfg.lr_critic = float(lr)
        cfg.beta1, cfg.beta2, cfg.adam_eps = 0.9, 0.999, 1e-8
        cfg.policy_freq = int(self.policy_freq)
        cfg.precision = _lib.PRECISIONS[precision or os.environ.get("TD3_PRECISION", "tf32")]
        cfg.seed = int(torch.randint(0, 2**62, (1,)).item()) if seed is None else int(seed)
        cfg.actor, cfg.q = net_layout(actor), net_layout(critic.q1)
        self.CDQ = CDQ
        self._engine_init(cfg, fam_a, fam_c, lr, rng)

    # ------------------------------------------------------------------ B=1 API (TD3_featured.py:113-121)
    def _b1_buffers(self):
        """Pinned host I/O of the batch-1 kernel (csrc/infer.cuh): allocated once, addressed directly by the device."""
        b = self.__dict__.get("_b1")
        if b is None:
            S, A = self._cfg.state_dim, self._cfg.action_dim
            pin_in = torch.zeros(S + A, dtype=torch.float32).pin_memory()
            pin_out = [torch.zeros(A + 4, dtype=torch.float32).pin_memory(),                 # actor: A results + 1 sequence word
                       torch.zeros(self._cfg.n_q + 4, dtype=torch.float32).pin_memory()]     # critics: n_q results + n_q words
            b = dict(pin_in=pin_in, pin_out=pin_out, in_np=pin_in.numpy(), out_np=[t.numpy() for t in pin_out],
                     in_ptr=C.c_void_p(pin_in.data_ptr()), out_ptr=[C.c_void_p(t.data_ptr()) for t in pin_out], seq=0,
                     fn=self._lib.td3_infer_b1, dev=self._device.index)
            self.__dict__["_b1"] = b
        return b

    def _b1_call(self, net, agent, n_nets, n_out):
        b = self._b1_buffers()
        b["seq"] = seq = (b["seq"] + 1) & 0x7FFFFFFF
        rc = b["fn"](self._handle, net, 0, agent, b["in_ptr"], b["out_ptr"][net], seq, 5_000_000,
                     torch._C._cuda_getCurrentRawStream(b["dev"]))
        if rc:
            _lib.check(rc)
        return b["out_np"][net]

    def select_action(self, state, agent=0):
        """actor(state) for one state (TD3_featured.py:113-115): the row goes into a pinned host buffer, ONE kernel reads
        it over PCIe, runs the whole network and writes the action back to pinned memory; no device tensors are made."""
        S, A = self._cfg.state_dim, self._cfg.action_dim
        b = self._b1_buffers()
        b["in_np"][:S] = np.asarray(state, dtype=np.float32).reshape(-1)
        return self._b1_call(0, agent, 1, A)[:A].copy()

    def eval_q(self, state, action, agent=0):
        """[Q1(s, a), Q2(s, a)] for one pair (TD3_featured.py:117-121), same single-kernel path."""
        S, A, nq = self._cfg.state_dim, self._cfg.action_dim, self._cfg.n_q
        b = self._b1_buffers()
        b["in_np"][:S] = np.asarray(state, dtype=np.float32).reshape(-1)
        b["in_np"][S:S + A] = np.asarray(action, dtype=np.float32).reshape(-1)
        out = self._b1_call(1, agent, nq, 1)
        return [out[i:i + 1].copy() for i in range(nq)]

    def _actor_forward(self, which, state, particles=None, agent=0):
        state = state.to(self._device, torch.float32).contiguous()
        B = state.shape[0]
        out = torch.empty(B, self._cfg.action_dim, device=self._device)
        for lo, hi in self._forward_chunks(B):
            _lib.check(self._lib.td3_actor_forward(self._handle, which, int(agent), state[lo:hi].data_ptr(), None, hi - lo,
                                                   out[lo:hi].data_ptr(), _lib.stream_ptr()))
        return out

    def _critic_forward(self, which, state, action, particles=None, agent=0):
        state = state.to(self._device, torch.float32).contiguous()
        action = action.to(self._device, torch.float32).contiguous()
        B = state.shape[0]
        outs = []
        for lo, hi in self._forward_chunks(B):
            o = torch.empty(self._cfg.n_q, hi - lo, 1, device=self._device)
            _lib.check(self._lib.td3_critic_forward(self._handle, which, int(agent), state[lo:hi].data_ptr(), None,
                                                    action[lo:hi].data_ptr(), hi - lo, o.data_ptr(), _lib.stream_ptr()))
            outs.append(o)
        out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
        return [out[i] for i in range(self._cfg.n_q)]

    # ------------------------------------------------------------------ the hot path (TD3_featured.py:123-171)
    def train(self, replay_buffer, batch_size=100, *, iterations=1, indices=None, noise=None, use_graph=True):
        """One TD3 update (``iterations`` of them back to back when given), entirely on the device:
        sample -> target step -> twin-critic step -> every policy_freq-th update actor step + Polyak.
        Returns None and never synchronises the host, like the reference.  ``indices`` ([batch] ints) and
        ``noise`` ([batch, A] N(0,1) draws) replace the two random draws for parity tests."""
        self._train_common(replay_buffer, batch_size, iterations, indices, noise, use_graph)


def _clone_shell(module):
    """Fresh shell of the same architecture without consuming the global RNG (the reference deep-copies)."""
    import copy
    return copy.deepcopy(module)
