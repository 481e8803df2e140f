"""CPU oracle for the TD3 update hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module, and only as the checker or the
timed CPU baseline.  Nothing under ``td3_b200/`` imports it; the product path
fails loudly when the CUDA library is missing.

What it restates
----------------
The reference (yannikkellerde/TD3, ``/root/reference``) is pure Python whose
arithmetic lives in third-party libraries that are *not* vendored in the
reference tree and are not version-pinned by it (no requirements file):
PyTorch (``nn.Linear``, ``nn.Conv2d/Conv1d``, ``nn.LayerNorm``,
``torch.optim.Adam``, ``F.mse_loss``) and NumPy (``np.random.randint``, fancy
indexing).  The versions this oracle is pinned on are the ones in this image:
torch 2.11.0+cu128, numpy 2.3.x.  The oracle therefore restates the reference's
*call structure* on top of the same library primitives, width-parametrised
(the reference hard-codes (500,400,300)/(500,400,200) at TD3_featured.py:19,54)
and with injection hooks for the two global RNG draws
(my_replay_buffer.py:59,120 and TD3_featured.py:132 / TD3_particles.py:176).

Parity pin
----------
The reference ships no tests or golden vectors.  The pin
is the reference itself, imported in the build container by
``oracle/make_golden.py``, which (a) asserts this oracle reproduces the real
classes bit-for-bit at the fork widths and (b) writes ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures on
every run and, where ``/root/reference`` exists, against the live reference.
"""
from __future__ import annotations

import copy
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

FORK_ACTOR_WIDTHS = (500, 400, 300)    # TD3_featured.py:19, TD3_particles.py:25
FORK_Q_WIDTHS = (500, 400, 200)        # TD3_featured.py:54
FORK_PQ_WIDTHS = (500, 400, 300)       # TD3_particles.py:76
VANILLA_WIDTHS = (400, 300)            # BASELINE.json "400-300 MLPs"
ENC_HIDDEN = 256                       # TD3_particles.py:27,29  (num_features*2)
ENC_OUT = 128                          # TD3_particles.py:27,30


class Space:
    """Minimal stand-in for gym.spaces.Box: only ``.shape`` is ever read."""

    def __init__(self, *shape):
        self.shape = tuple(shape)


# --------------------------------------------------------------------------- #
# Replay buffers (my_replay_buffer.py:6-128)
# --------------------------------------------------------------------------- #
class ReplayFeatured:
    """float64 ring buffer; follows my_replay_buffer.py:72-128."""

    fields = ("state", "action", "next_state", "reward", "not_done")

    def __init__(self, obs_space, action_space, max_size=int(1e6)):
        s, a = obs_space.shape[0], action_space.shape[0]
        self.max_size, self.ptr, self.size = int(max_size), 0, 0
        self.state = np.zeros((self.max_size, s))            # :81
        self.action = np.zeros((self.max_size, a))           # :82
        self.next_state = np.zeros((self.max_size, s))       # :83
        self.reward = np.zeros((self.max_size, 1))           # :84
        self.not_done = np.zeros((self.max_size, 1))         # :85

    def add(self, state, action, next_state, reward, done):  # :109-117
        p = self.ptr
        self.state[p] = state
        self.action[p] = action
        self.next_state[p] = next_state
        self.reward[p] = reward
        self.not_done[p] = 1.0 - done
        self.ptr = (p + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def sample(self, batch_size, indices=None):              # :119-128
        ind = np.random.randint(0, self.size, size=batch_size) if indices is None else np.asarray(indices)
        return tuple(torch.FloatTensor(getattr(self, f)[ind]) for f in self.fields)


class ReplayParticles:
    """float64 ring buffer with particle sets; follows my_replay_buffer.py:6-69."""

    fields = ("state_features", "state_particles", "action", "next_state_features",
              "next_state_particles", "reward", "not_done")

    def __init__(self, obs_space, action_space, max_size=int(1e6)):
        f = obs_space[0].shape[0]
        pshape = tuple(obs_space[1].shape)
        a = action_space.shape[0]
        self.max_size, self.ptr, self.size = int(max_size), 0, 0
        self.state_features = np.zeros((self.max_size, f))             # :16
        self.state_particles = np.zeros((self.max_size, *pshape))      # :17
        self.action = np.zeros((self.max_size, a))                     # :18
        self.next_state_features = np.zeros((self.max_size, f))        # :19
        self.next_state_particles = np.zeros((self.max_size, *pshape))  # :20
        self.reward = np.zeros((self.max_size, 1))                     # :21
        self.not_done = np.zeros((self.max_size, 1))                   # :22

    def add(self, state, action, next_state, reward, done):            # :46-56
        p = self.ptr
        self.state_features[p] = state[0]
        self.state_particles[p] = state[1]
        self.action[p] = action
        self.next_state_features[p] = next_state[0]
        self.next_state_particles[p] = next_state[1]
        self.reward[p] = reward
        self.not_done[p] = 1.0 - done
        self.ptr = (p + 1) % self.max_size
        self.size = min(self.size + 1, self.max_size)

    def sample(self, batch_size, indices=None):                        # :58-69
        ind = np.random.randint(0, self.size, size=batch_size) if indices is None else np.asarray(indices)
        return tuple(torch.FloatTensor(getattr(self, f)[ind]) for f in self.fields)


# --------------------------------------------------------------------------- #
# Networks
# --------------------------------------------------------------------------- #
def _stack(in_dim: int, widths: Sequence[int], out_dim: int) -> nn.ModuleList:
    dims = [in_dim, *widths, out_dim]
    return nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1))


def _apply_norm_option(mod: nn.Module, widths, norm, in_norm_dim=None):
    # TD3_featured.py:28-36 / TD3_particles.py:42-50: LN modules are registered
    # after the linears (lnorm1 before lnorms in the particle nets), weight-norm
    # wraps the linears only.
    mod.norm = norm
    if norm == "layer":
        if in_norm_dim is not None:
            mod.lnorm1 = nn.LayerNorm(in_norm_dim)
        mod.lnorms = nn.ModuleList(nn.LayerNorm(w) for w in widths)
    if norm == "weight_normalization":
        for i in range(len(mod.linears)):
            mod.linears[i] = nn.utils.weight_norm(mod.linears[i])


def _trunk(mod: nn.Module, x: torch.Tensor) -> torch.Tensor:
    # TD3_featured.py:41-46,75-80; TD3_particles.py:62-67,113-118
    last = len(mod.linears) - 1
    for i, lin in enumerate(mod.linears):
        x = lin(x)
        if i != last:
            x = F.relu(x)
            if mod.norm == "layer":
                x = mod.lnorms[i](x)
    return x


class MlpActor(nn.Module):
    """TD3_featured.py:15-48, widths parametrised."""

    def __init__(self, state_dim, action_dim, max_action, norm, widths=FORK_ACTOR_WIDTHS):
        super().__init__()
        self.linears = _stack(state_dim, widths, action_dim)
        _apply_norm_option(self, widths, norm)
        self.max_action = max_action

    def forward(self, state):
        return self.max_action * torch.tanh(_trunk(self, state))          # :47-48


class MlpQ(nn.Module):
    """TD3_featured.py:50-81, widths parametrised."""

    def __init__(self, state_dim, action_dim, norm, widths=FORK_Q_WIDTHS):
        super().__init__()
        self.linears = _stack(state_dim + action_dim, widths, 1)
        _apply_norm_option(self, widths, norm)

    def forward(self, state, action):
        return _trunk(self, torch.cat([state, action], 1))                # :74


class MlpCritic(nn.Module):
    """TD3_featured.py:84-96."""

    def __init__(self, state_dim, action_dim, norm, widths=FORK_Q_WIDTHS):
        super().__init__()
        self.q1 = MlpQ(state_dim, action_dim, norm, widths)
        self.q2 = MlpQ(state_dim, action_dim, norm, widths)

    def forward(self, state, action):
        return self.q1(state, action), self.q2(state, action)

    def Q1(self, state, action):
        return self.q1(state, action)


class _SetEncoderMixin:
    """Per-particle D->256->128 MLP, mean over N, ReLU (TD3_particles.py:29-32,53-58)."""

    def _build_encoder(self, obs_space):
        n, d = obs_space[1].shape
        self.conv1 = nn.Conv2d(1, ENC_HIDDEN, kernel_size=(1, d), stride=1)       # :29
        self.conv2 = nn.Conv1d(ENC_HIDDEN, ENC_OUT, kernel_size=1, stride=1)      # :30
        self.avg_pool = nn.AvgPool2d(kernel_size=(1, n))                          # :32

    def _encode(self, particles):
        h = F.relu(self.conv1(torch.unsqueeze(particles, 1)))     # :53-54
        h = torch.squeeze(h, dim=3)                               # :55
        h = F.relu(self.conv2(h))                                 # :56
        h = F.relu(self.avg_pool(h))                              # :57
        return torch.squeeze(h, dim=2)                            # :58


class SetActor(nn.Module, _SetEncoderMixin):
    """TD3_particles.py:19-69."""

    def __init__(self, obs_space, action_space, norm=None, widths=FORK_ACTOR_WIDTHS):
        super().__init__()
        self._build_encoder(obs_space)
        in_dim = ENC_OUT + obs_space[0].shape[0]
        self.linears = _stack(in_dim, widths, action_space.shape[0])
        _apply_norm_option(self, widths, norm, in_norm_dim=in_dim)

    def forward(self, feats, particles):
        x = torch.cat([self._encode(particles), feats], 1)        # :59
        if self.norm == "layer":
            x = self.lnorm1(x)                                    # :60-61
        return torch.tanh(_trunk(self, x))                        # :68-69 (no max_action)


class SetQ(nn.Module, _SetEncoderMixin):
    """TD3_particles.py:71-119; the head is ``action_dim`` wide (:91)."""

    def __init__(self, obs_space, action_space, norm=None, widths=FORK_PQ_WIDTHS):
        super().__init__()
        self._build_encoder(obs_space)
        in_dim = ENC_OUT + obs_space[0].shape[0] + action_space.shape[0]
        self.linears = _stack(in_dim, widths, action_space.shape[0])
        _apply_norm_option(self, widths, norm, in_norm_dim=in_dim)

    def forward(self, feats, particles, action):
        x = torch.cat([self._encode(particles), feats, action], 1)   # :110
        if self.norm == "layer":
            x = self.lnorm1(x)
        return _trunk(self, x)


class SetCritic(nn.Module):
    """TD3_particles.py:121-136."""

    def __init__(self, obs_space, action_space, norm=None, CDQ=True, widths=FORK_PQ_WIDTHS):
        super().__init__()
        self.q1 = SetQ(obs_space, action_space, norm, widths)
        self.CDQ = CDQ
        if CDQ:
            self.q2 = SetQ(obs_space, action_space, norm, widths)

    def forward(self, feats, particles, action):
        if not self.CDQ:
            return (self.q1(feats, particles, action),)
        return self.q1(feats, particles, action), self.q2(feats, particles, action)

    def Q1(self, feats, particles, action):
        return self.q1(feats, particles, action)


# --------------------------------------------------------------------------- #
# Agents
# --------------------------------------------------------------------------- #
class _AgentBase:
    """TD3_base.py:6-24 plus the parts of ``train`` both variants share."""

    def _hyper(self, max_action=1, discount=0.99, tau=0.005, policy_noise=0.2,
               noise_clip=0.5, policy_freq=2):
        self.max_action, self.discount, self.tau = max_action, discount, tau
        self.policy_noise, self.noise_clip, self.policy_freq = policy_noise, noise_clip, policy_freq
        self.total_it = 0
        self.trace = {}
        # parity tests read self.trace after every update; timing runs switch it off so that the timed work is exactly
        # the reference's (the trace adds clones and a float() per update that TD3_featured.py does not have)
        self.keep_trace = True

    def _smoothing_noise(self, action, noise):
        eps = torch.randn_like(action) if noise is None else torch.as_tensor(noise, dtype=action.dtype)
        return (eps * self.policy_noise).clamp(-self.noise_clip, self.noise_clip)

    def _polyak(self):
        # TD3_featured.py:167-171 / TD3_particles.py:220-224 (per-tensor loop kept:
        # it is part of what the CPU baseline costs)
        for p, tp in zip(self.critic.parameters(), self.critic_target.parameters()):
            tp.data.copy_(self.tau * p.data + (1 - self.tau) * tp.data)
        for p, tp in zip(self.actor.parameters(), self.actor_target.parameters()):
            tp.data.copy_(self.tau * p.data + (1 - self.tau) * tp.data)


class TD3Featured(_AgentBase):
    """TD3_featured.py:99-171 with parametrised widths and RNG injection."""

    def __init__(self, obs_space, action_space, max_action=1, lr=1e-4, norm=None, CDQ=True,
                 actor_widths=FORK_ACTOR_WIDTHS, q_widths=FORK_Q_WIDTHS, **kwargs):
        s, a = obs_space.shape[0], action_space.shape[0]
        self.actor = MlpActor(s, a, max_action, norm, actor_widths)
        self.actor_target = copy.deepcopy(self.actor)                                  # :102
        self.actor_optimizer = torch.optim.Adam(self.actor.parameters(), lr=lr)        # :104
        self.critic = MlpCritic(s, a, norm, q_widths)
        self.critic_target = copy.deepcopy(self.critic)                                # :107
        self.critic_optimizer = torch.optim.Adam(self.critic.parameters(), lr=lr)      # :108
        self._hyper(max_action=max_action, **kwargs)

    def select_action(self, state):                                                    # :113-115
        x = torch.FloatTensor(np.asarray(state).reshape(1, -1))
        return self.actor(x).data.numpy().flatten()

    def eval_q(self, state, action):                                                   # :117-121
        x = torch.FloatTensor(np.asarray(state).reshape(1, -1))
        u = torch.FloatTensor(np.asarray(action).reshape(1, -1))
        return [q.data.numpy().flatten() for q in self.critic(x, u)]

    def train(self, replay_buffer, batch_size=100, indices=None, noise=None):          # :123-171
        self.total_it += 1
        state, action, next_state, reward, not_done = replay_buffer.sample(batch_size, indices) \
            if indices is not None else replay_buffer.sample(batch_size)
        with torch.no_grad():
            eps = self._smoothing_noise(action, noise)                                 # :131-133
            next_action = (self.actor_target(next_state) + eps).clamp(-self.max_action, self.max_action)
            tq1, tq2 = self.critic_target(next_state, next_action)                     # :140
            target_q = reward + not_done * self.discount * torch.min(tq1, tq2)         # :141-142
        q1, q2 = self.critic(state, action)                                            # :145
        critic_loss = F.mse_loss(q1, target_q) + F.mse_loss(q2, target_q)              # :148
        self.critic_optimizer.zero_grad()
        critic_loss.backward()
        self.critic_optimizer.step()                                                   # :151-153
        if self.keep_trace:
            self.trace = dict(critic_loss=float(critic_loss.detach()), q1=q1.detach().clone(), q2=q2.detach().clone(),
                              target_q=target_q.clone(), next_action=next_action.clone(), actor_loss=None)
        if self.total_it % self.policy_freq == 0:                                      # :156
            actor_loss = -self.critic.Q1(state, self.actor(state)).mean()              # :159
            self.actor_optimizer.zero_grad()
            actor_loss.backward()
            self.actor_optimizer.step()                                                # :162-164
            if self.keep_trace:
                self.trace["actor_loss"] = float(actor_loss.detach())
            self._polyak()


class TD3Particles(_AgentBase):
    """TD3_particles.py:138-224 with parametrised widths and RNG injection."""

    def __init__(self, obs_space, action_space, lr=1e-4, norm=None, CDQ=True,
                 actor_widths=FORK_ACTOR_WIDTHS, q_widths=FORK_PQ_WIDTHS, **kwargs):
        self.actor = SetActor(obs_space, action_space, norm, actor_widths)
        self.actor_target = SetActor(obs_space, action_space, norm, actor_widths)      # :141 (fresh init, then overwritten)
        self.actor_target.load_state_dict(self.actor.state_dict())                     # :142
        self.actor_optimizer = torch.optim.Adam(self.actor.parameters(), lr=lr)
        self.critic = SetCritic(obs_space, action_space, norm, CDQ=CDQ, widths=q_widths)
        self.critic_target = SetCritic(obs_space, action_space, norm, CDQ=CDQ, widths=q_widths)
        self.critic_target.load_state_dict(self.critic.state_dict())                   # :147
        self.critic_optimizer = torch.optim.Adam(self.critic.parameters(), lr=lr)
        self.CDQ = CDQ
        self._hyper(**kwargs)                                                          # :151

    @staticmethod
    def _to_batch1(state):
        feats = torch.FloatTensor(np.array([state[0]]).reshape(1, -1))                 # :154
        parts = torch.FloatTensor(np.asarray(state[1]).reshape(1, *np.asarray(state[1]).shape))
        return feats, parts

    def select_action(self, state):                                                    # :153-157
        feats, parts = self._to_batch1(state)
        return self.actor(feats, parts).data.numpy().flatten()

    def eval_q(self, state, action):                                                   # :159-164
        feats, parts = self._to_batch1(state)
        u = torch.FloatTensor(np.array(action).reshape(1, -1))
        return [q.data.numpy().flatten() for q in self.critic(feats, parts, u)]

    def train(self, replay_buffer, batch_size=100, indices=None, noise=None):          # :167-207
        self.total_it += 1
        sf, sp, action, nsf, nsp, reward, not_done = replay_buffer.sample(batch_size, indices) \
            if indices is not None else replay_buffer.sample(batch_size)
        with torch.no_grad():
            eps = self._smoothing_noise(action, noise)                                 # :175-177
            next_action = self.actor_target(nsf, nsp) + eps                            # :179-181 (no clamp)
            tqs = self.critic_target(nsf, nsp, next_action)
            tq = torch.min(tqs[0], tqs[1]) if self.CDQ else tqs[0]                     # :184-188
            target_q = reward + not_done * self.discount * tq                          # :189 ([B,1] x [B,A] broadcast)
        qs = self.critic(sf, sp, action)
        critic_loss = F.mse_loss(qs[0], target_q)
        if self.CDQ:
            critic_loss = critic_loss + F.mse_loss(qs[1], target_q)                    # :195
        self.critic_optimizer.zero_grad()
        critic_loss.backward()
        self.critic_optimizer.step()                                                   # :201-203
        if self.keep_trace:
            self.trace = dict(critic_loss=float(critic_loss.detach()), q1=qs[0].detach().clone(),
                              q2=qs[1].detach().clone() if self.CDQ else None,
                              target_q=target_q.clone(), next_action=next_action.clone(), actor_loss=None)
        if self.total_it % self.policy_freq == 0:
            self._actor_learn(sf, sp)                                                  # :206-207

    def _actor_learn(self, state_features, state_particles):                           # :209-224
        action = self.actor(state_features, state_particles)
        actor_loss = -self.critic.Q1(state_features, state_particles, action).mean()
        self.actor_optimizer.zero_grad()
        actor_loss.backward()
        self.actor_optimizer.step()
        if self.keep_trace:
            self.trace["actor_loss"] = float(actor_loss.detach())
        self._polyak()


# --------------------------------------------------------------------------- #
# Synthetic workload (SURVEY.md section 8d "Synthetic inputs")
# --------------------------------------------------------------------------- #
def synthetic_transitions_featured(n, state_dim, action_dim, seed=0):
    """states/next ~ N(0,1), actions ~ U(-1,1), rewards ~ N(0,1), 1 % terminal."""
    rs = np.random.RandomState(seed)
    return dict(
        state=rs.standard_normal((n, state_dim)),
        action=rs.uniform(-1.0, 1.0, (n, action_dim)),
        next_state=rs.standard_normal((n, state_dim)),
        reward=rs.standard_normal((n, 1)),
        done=(rs.uniform(size=(n, 1)) < 0.01).astype(np.float64),
    )


def synthetic_transitions_particles(n, feat_dim, n_particles, particle_dim, action_dim, seed=0):
    rs = np.random.RandomState(seed)
    return dict(
        state_features=rs.standard_normal((n, feat_dim)),
        state_particles=rs.standard_normal((n, n_particles, particle_dim)).astype(np.float32),
        action=rs.uniform(-1.0, 1.0, (n, action_dim)),
        next_state_features=rs.standard_normal((n, feat_dim)),
        next_state_particles=rs.standard_normal((n, n_particles, particle_dim)).astype(np.float32),
        reward=rs.standard_normal((n, 1)),
        done=(rs.uniform(size=(n, 1)) < 0.01).astype(np.float64),
    )


def fill_featured(rb, data):
    """Bulk equivalent of ``n`` calls to ``add`` (same final contents, ptr, size)."""
    n = len(data["state"])
    assert n <= rb.max_size
    rb.state[:n], rb.action[:n], rb.next_state[:n] = data["state"], data["action"], data["next_state"]
    rb.reward[:n], rb.not_done[:n] = data["reward"], 1.0 - data["done"]
    rb.ptr, rb.size = n % rb.max_size, n


def fill_particles(rb, data):
    n = len(data["action"])
    assert n <= rb.max_size
    for k in ("state_features", "state_particles", "action", "next_state_features", "next_state_particles", "reward"):
        getattr(rb, k)[:n] = data[k]
    rb.not_done[:n] = 1.0 - data["done"]
    rb.ptr, rb.size = n % rb.max_size, n


def param_digest(module: nn.Module) -> np.ndarray:
    """Per-tensor (sum, abs-sum, first, last) in float64: a compact fingerprint of a net."""
    rows = []
    for p in module.parameters():
        v = p.detach().double().flatten()
        rows.append([float(v.sum()), float(v.abs().sum()), float(v[0]), float(v[-1])])
    return np.asarray(rows, dtype=np.float64)
