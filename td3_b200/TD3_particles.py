"""TD3 with a particle-set encoder in every network -- B200-native drop-in for the reference's
TD3_particles (TD3_particles.py:19-224).

Observation = (feature vector [F], particle set [N, D]).  Every network encodes the particles with a
shared per-particle MLP D -> 256 -> 128 (the reference's Conv2d(1,256,(1,D)) + Conv1d(256,128,1)),
mean-pools over the N particles, concatenates the features (and the action for Q networks), and
runs a 500-400-300 trunk.  Reference quirks kept on purpose: the actor output is a
plain tanh (no max_action), the smoothed target action is NOT clamped, the Q head is action_dim wide
and the [B,1] reward broadcasts against it, ``CDQ=False`` drops the second critic.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import _lib
from .TD3_base import TD3_base
from .TD3_featured import _clone_shell
from .packing import PackedFamily, SetActor, SetCritic, net_layout

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

Actor, Critic = SetActor, SetCritic


class TD3(TD3_base):
    def __init__(self, obs_space, action_space, lr=1e-4, norm=None, CDQ=True, *,
                 actor_widths=(500, 400, 300), q_widths=(500, 400, 300), rng="device", seed=None, precision=None,
                 n_agents=1, **kwargs):
        _lib.require_cuda()
        F = obs_space[0].shape[0]
        N, D = obs_space[1].shape
        A = action_space.shape[0]
        # Construction order (and therefore global-RNG consumption) of TD3_particles.py:140-147:
        # actor, a throw-away freshly initialised actor_target, critic, a throw-away critic_target.
        actor = SetActor(F, N, D, A, norm, actor_widths)
        actor_t = SetActor(F, N, D, A, norm, actor_widths)
        actor_t.load_state_dict(actor.state_dict())
        critic = SetCritic(F, N, D, A, norm, CDQ, q_widths)
        critic_t = SetCritic(F, N, D, A, norm, CDQ, q_widths)
        critic_t.load_state_dict(critic.state_dict())
        super(TD3, self).__init__(**kwargs)                                     # :151
        dev = torch.device("cuda", torch.cuda.current_device())
        n_agents = int(n_agents)
        if n_agents < 1:
            raise ValueError("n_agents must be >= 1")
        fam_a = PackedFamily(actor, actor_t, [""], dev, n_agents)
        fam_c = PackedFamily(critic, critic_t, ["q1", "q2"] if CDQ else ["q1"], dev, n_agents)
        for i in range(1, n_agents):      # further members of the population: fresh default initialisations, in order
            fam_a.load_agent(i, SetActor(F, N, D, A, norm, actor_widths).state_dict())
            fam_c.load_agent(i, SetCritic(F, N, D, A, norm, CDQ, q_widths).state_dict())
        self.actor, self.actor_target, self.critic, self.critic_target = actor, actor_t, critic, critic_t
        for m, w in ((actor, 0), (actor_t, 1), (critic, 0), (critic_t, 1)):
            m._attach(self, w)
        self.CDQ = CDQ
        cfg = _lib.AgentConfig()
        cfg.variant = _lib.VARIANT_PARTICLES
        cfg.norm = {None: _lib.NORM_NONE, "layer": _lib.NORM_LAYER, "weight_normalization": _lib.NORM_WEIGHT}[norm]
        cfg.n_q, cfg.state_dim, cfg.action_dim = (2 if CDQ else 1), F, A
        cfg.n_particles, cfg.particle_dim = N, D
        cfg.clamp_target_action, cfg.n_agents = 0, n_agents
        cfg.max_action = 1.0
        cfg.discount, cfg.tau = float(self.discount), float(self.tau)
        cfg.policy_noise, cfg.noise_clip = float(self.policy_noise), float(self.noise_clip)
        cfg.lr_actor = cfg.lr_critic = float(lr)
        cfg.beta1, cfg.beta2, cfg.adam_eps = 0.9, 0.999, 1e-8
        cfg.policy_freq = int(self.policy_freq)
        cfg.precision = _lib.PRECISIONS[precision or os.environ.get("TD3_PRECISION", "tf32")]
        cfg.seed = int(torch.randint(0, 2**62, (1,)).item()) if seed is None else int(seed)
        cfg.actor, cfg.q = net_layout(actor), net_layout(critic.q1)
        self._engine_init(cfg, fam_a, fam_c, lr, rng)

    # ------------------------------------------------------------------ B=1 API (TD3_particles.py:153-164)
    def _state_to_device(self, state):
        feats = torch.as_tensor(np.array([state[0]], dtype=np.float32).reshape(1, -1), device=self._device)
        parts = np.asarray(state[1], dtype=np.float32)
        return feats, torch.as_tensor(parts.reshape(1, *parts.shape), device=self._device)

    def select_action(self, state, agent=0):
        feats, parts = self._state_to_device(state)
        return self._actor_forward(0, feats, parts, agent=agent).cpu().numpy().flatten()

    def eval_q(self, state, action, agent=0):
        feats, parts = self._state_to_device(state)
        action = torch.as_tensor(np.array(action, dtype=np.float32).reshape(1, -1), device=self._device)
        return [q.cpu().numpy().flatten() for q in self._critic_forward(0, feats, action, parts, agent=agent)]

    def _actor_forward(self, which, feats, particles, agent=0):
        feats = feats.to(self._device, torch.float32).contiguous()
        particles = particles.to(self._device, torch.float32).contiguous()
        B = feats.shape[0]
        out = torch.empty(B, self._cfg.action_dim, device=self._device)
        for lo, hi in self._forward_chunks(B):
            _lib.check(self._lib.td3_actor_forward(self._handle, which, int(agent), feats[lo:hi].data_ptr(),
                                                   particles[lo:hi].data_ptr(), hi - lo, out[lo:hi].data_ptr(), _lib.stream_ptr()))
        return out

    def _critic_forward(self, which, feats, action, particles, agent=0):
        feats = feats.to(self._device, torch.float32).contiguous()
        action = action.to(self._device, torch.float32).contiguous()
        particles = particles.to(self._device, torch.float32).contiguous()
        B = feats.shape[0]
        outs = []
        for lo, hi in self._forward_chunks(B):
            o = torch.empty(self._cfg.n_q, hi - lo, self._cfg.action_dim, device=self._device)
            _lib.check(self._lib.td3_critic_forward(self._handle, which, int(agent), feats[lo:hi].data_ptr(),
                                                    particles[lo:hi].data_ptr(), action[lo:hi].data_ptr(), hi - lo, o.data_ptr(),
                                                    _lib.stream_ptr()))
            outs.append(o)
        out = outs[0] if len(outs) == 1 else torch.cat(outs, dim=1)
        return [out[i] for i in range(self._cfg.n_q)]

    # ------------------------------------------------------------------ the hot path (TD3_particles.py:167-207)
    def train(self, replay_buffer, batch_size=100, *, iterations=1, indices=None, noise=None, use_graph=True):
        self._train_common(replay_buffer, batch_size, iterations, indices, noise, use_graph)

    def _actor_learn(self, state_features, state_particles):
        """Actor update + Polyak on caller-supplied device tensors (TD3_particles.py:209-224;
        called directly by evaluate_model.py:49)."""
        B = int(state_features.shape[0])
        self._ensure_plan(max(B, self._planned_batch))
        if B != self._planned_batch:
            self._ensure_plan(B)
        E, F, A = 128, self._cfg.state_dim, self._cfg.action_dim
        feats = state_features.to(self._device, torch.float32)
        parts = state_particles.to(self._device, torch.float32).reshape(B, -1)
        ld_a, ld_q = (E + F + 3) // 4 * 4, (E + F + A + 3) // 4 * 4
        self._region("x_actor").view(-1, ld_a)[:B, E:E + F].copy_(feats)
        self._region("x_q_pi").view(-1, ld_q)[:B, E:E + F].copy_(feats)
        self._region("particles").view(B, -1).copy_(parts)
        _lib.check(self._lib.td3_actor_step(self._handle, 1, _lib.stream_ptr()))
