"""Packed parameter buffers and nn.Module shells.

Each network family (actor / critic) lives in five flat fp32 device buffers --
params, target, grad, exp_avg, exp_avg_sq -- laid out in the reference's
``nn.Module.parameters()`` order.  The ``nn.Module`` objects the reference's
callers touch (``policy.actor``, ``policy.critic_target`` ...: ``.eval()``,
``.parameters()``, ``.state_dict()``, ``__call__``) are thin shells
whose Parameters are *views* into those buffers, so ``torch.save(state_dict())``
and ``load_state_dict`` interoperate with reference checkpoints while the CUDA
kernels update the buffers in place.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib

ENC_HIDDEN, ENC_OUT = 256, 128            # TD3_particles.py:27-30
ALIGN = 4                                 # every tensor starts on a 16-byte boundary


def _linear_stack(in_dim: int, widths: Sequence[int], out_dim: int) -> nn.ModuleList:
    dims = [in_dim, *widths, out_dim]
    return nn.ModuleList(nn.Linear(dims[i], dims[i + 1]) for i in range(len(dims) - 1))


class _Net(nn.Module):
    """Common shell: holds reference-named sub-modules; forward goes through the C ABI."""

    def __init__(self):
        super().__init__()
        self._owner = None          # (agent, role) set by the agent; not a Module -> not in state_dict
        self._which = 0

    def _attach(self, owner, which):
        object.__setattr__(self, "_owner", owner)
        object.__setattr__(self, "_which", which)

    def load_state_dict(self, state_dict, *args, **kwargs):
        # the values land in the packed device buffers through the Parameter views; the engine's TF32 copies of
        # them (csrc/engine.cu: ensure_shadows) have to be rebuilt before the next update
        out = super().load_state_dict(state_dict, *args, **kwargs)
        if self._owner is not None:
            self._owner.params_changed()
        return out

    def _check_norm(self, norm, allow_weight_norm=False):
        if norm == "weight_normalization" and allow_weight_norm:
            return
        if norm not in (None, "layer"):
            # TD3_featured + weight_normalization cannot even be constructed in the reference under torch 2.x
            # (deepcopy of weight-normed modules raises): there is nothing to be a drop-in for.
            raise NotImplementedError(f"norm={norm!r} is not supported here (TD3_featured: None, 'layer'; "
                                      "TD3_particles: None, 'layer', 'weight_normalization')")

    def _apply_weight_norm(self):
        """TD3_particles.py:48-50: weight_norm on the `linears` only (the encoder convolutions stay plain)."""
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for i in range(len(self.linears)):
                self.linears[i] = nn.utils.weight_norm(self.linears[i])


class MlpActor(_Net):
    """Parameter container mirroring TD3_featured.Actor (TD3_featured.py:15-48)."""

    def __init__(self, state_dim, action_dim, max_action, norm, widths):
        super().__init__()
        self._check_norm(norm)
        self.linears = _linear_stack(state_dim, widths, action_dim)
        self.norm = norm
        if norm == "layer":
            self.lnorms = nn.ModuleList(nn.LayerNorm(w) for w in widths)
        self.max_action = max_action
        self.architecture = tuple(widths)

    def forward(self, state):
        return self._owner._actor_forward(self._which, state)


class MlpQ(_Net):
    """Mirrors TD3_featured.Q (TD3_featured.py:50-81)."""

    def __init__(self, state_dim, action_dim, norm, widths):
        super().__init__()
        self._check_norm(norm)
        self.linears = _linear_stack(state_dim + action_dim, widths, 1)
        self.norm = norm
        if norm == "layer":
            self.lnorms = nn.ModuleList(nn.LayerNorm(w) for w in widths)
        self.architecture = tuple(widths)


class MlpCritic(_Net):
    """Mirrors TD3_featured.Critic (TD3_featured.py:84-96)."""

    def __init__(self, state_dim, action_dim, norm, widths):
        super().__init__()
        self.q1 = MlpQ(state_dim, action_dim, norm, widths)
        self.q2 = MlpQ(state_dim, action_dim, norm, widths)

    def forward(self, state, action):
        return tuple(self._owner._critic_forward(self._which, state, action))

    def Q1(self, state, action):
        return self._owner._critic_forward(self._which, state, action)[0]


def _encoder(mod: nn.Module, n_particles: int, particle_dim: int):
    mod.conv1 = nn.Conv2d(1, ENC_HIDDEN, kernel_size=(1, particle_dim), stride=1)   # TD3_particles.py:29
    mod.conv2 = nn.Conv1d(ENC_HIDDEN, ENC_OUT, kernel_size=1, stride=1)             # :30
    mod.avg_pool = nn.AvgPool2d(kernel_size=(1, n_particles))                       # :32 (no parameters)


class SetActor(_Net):
    """Mirrors TD3_particles.Actor (TD3_particles.py:19-69)."""

    def __init__(self, feat_dim, n_particles, particle_dim, action_dim, norm, widths):
        super().__init__()
        self._check_norm(norm, allow_weight_norm=True)
        self.num_features = ENC_OUT
        _encoder(self, n_particles, particle_dim)
        in_dim = ENC_OUT + feat_dim
        self.linears = _linear_stack(in_dim, widths, action_dim)
        self.norm = norm
        if norm == "layer":
            self.lnorm1 = nn.LayerNorm(in_dim)
            self.lnorms = nn.ModuleList(nn.LayerNorm(w) for w in widths)
        if norm == "weight_normalization":
            self._apply_weight_norm()
        self.architecture = tuple(widths)

    def forward(self, state_features, state_particles):
        return self._owner._actor_forward(self._which, state_features, state_particles)


class SetQ(_Net):
    """Mirrors TD3_particles.Q_network (TD3_particles.py:71-119); head is action_dim wide (:91)."""

    def __init__(self, feat_dim, n_particles, particle_dim, action_dim, norm, widths):
        super().__init__()
        self._check_norm(norm, allow_weight_norm=True)
        self.num_features = ENC_OUT
        _encoder(self, n_particles, particle_dim)
        in_dim = ENC_OUT + feat_dim + action_dim
        self.linears = _linear_stack(in_dim, widths, action_dim)
        self.norm = norm
        if norm == "layer":
            self.lnorm1 = nn.LayerNorm(in_dim)
            self.lnorms = nn.ModuleList(nn.LayerNorm(w) for w in widths)
        if norm == "weight_normalization":
            self._apply_weight_norm()
        self.architecture = tuple(widths)


class SetCritic(_Net):
    """Mirrors TD3_particles.Critic (TD3_particles.py:121-136)."""

    def __init__(self, feat_dim, n_particles, particle_dim, action_dim, norm, CDQ, widths):
        super().__init__()
        self.q1 = SetQ(feat_dim, n_particles, particle_dim, action_dim, norm, widths)
        self.CDQ = CDQ
        if CDQ:
            self.q2 = SetQ(feat_dim, n_particles, particle_dim, action_dim, norm, widths)

    def forward(self, state_features, state_particles, action):
        return tuple(self._owner._critic_forward(self._which, state_features, action, state_particles))

    def Q1(self, state_features, state_particles, action):
        return self._owner._critic_forward(self._which, state_features, action, state_particles)[0]


# --------------------------------------------------------------------------- #
# layout
# --------------------------------------------------------------------------- #
def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def param_table(module: nn.Module) -> "OrderedDict[str, Tuple[int, torch.Size]]":
    """name -> (offset in floats, shape), in ``parameters()`` order, 16-byte aligned."""
    table, off = OrderedDict(), 0
    for name, p in module.named_parameters():
        table[name] = (off, p.shape)
        off += _round_up(p.numel(), ALIGN)
    table["__total__"] = (_round_up(off, 64), torch.Size([]))
    return table


def net_layout(single_net: nn.Module, prefix_table=None) -> _lib.NetLayout:
    """td3_net_layout of ONE network (an actor, or q1 of a critic) from its parameter table."""
    table = param_table(single_net)
    lay = _lib.NetLayout()
    linears = single_net.linears
    lay.n_linear = len(linears)
    if lay.n_linear > _lib.TD3_MAX_LINEAR:
        raise ValueError(f"at most {_lib.TD3_MAX_LINEAR} linear layers are supported")
    lay.dims[0] = linears[0].in_features
    wn = getattr(single_net, "norm", None) == "weight_normalization"
    for i, lin in enumerate(linears):
        lay.dims[i + 1] = lin.out_features
        if wn:       # parameters are (bias, weight_g, weight_v); the engine materialises W = g * v / ||v|| per row
            lay.w_off[i] = table[f"linears.{i}.weight_v"][0]
            lay.wg_off[i] = table[f"linears.{i}.weight_g"][0]
        else:
            lay.w_off[i] = table[f"linears.{i}.weight"][0]
        lay.b_off[i] = table[f"linears.{i}.bias"][0]
        if getattr(single_net, "norm", None) == "layer" and i < len(linears) - 1:
            lay.ln_g_off[i] = table[f"lnorms.{i}.weight"][0]
            lay.ln_b_off[i] = table[f"lnorms.{i}.bias"][0]
    if hasattr(single_net, "conv1"):
        lay.enc_hidden, lay.enc_out = single_net.conv1.out_channels, single_net.conv2.out_channels
        lay.c1w_off, lay.c1b_off = table["conv1.weight"][0], table["conv1.bias"][0]
        lay.c2w_off, lay.c2b_off = table["conv2.weight"][0], table["conv2.bias"][0]
        if single_net.norm == "layer":
            lay.ln_in_g_off, lay.ln_in_b_off = table["lnorm1.weight"][0], table["lnorm1.bias"][0]
    lay.n_floats = table["__total__"][0]
    return lay


class PackedFamily:
    """The five flat buffers of one network family + the online/target module shells."""

    def __init__(self, online: nn.Module, target: nn.Module, sub_nets: List[str], device, n_agents: int = 1):
        # sub_nets: [""] for an actor, ["q1", "q2"] (or ["q1"]) for a critic: every sub-net gets its own
        # 64-float-aligned slot so the twin networks sit a constant stride apart.
        self.sub_nets = sub_nets
        first = online if sub_nets == [""] else getattr(online, sub_nets[0])
        self.sub_table = param_table(first)
        self.stride = self.sub_table["__total__"][0]
        self.n_floats = self.stride * len(sub_nets)
        self.n_agents = n_agents
        total = self.n_floats * n_agents
        mk = lambda: torch.zeros(total, dtype=torch.float32, device=device)
        self.params, self.target, self.grad, self.exp_avg, self.exp_avg_sq = mk(), mk(), mk(), mk(), mk()
        self.online_module, self.target_module = online, target
        self._adopt(online, self.params)
        self._adopt(target, self.target)
        self.target.copy_(self.params)
        self.names = [n for n, _ in online.named_parameters()]

    def _slots(self, module):
        for si, sub in enumerate(self.sub_nets):
            net = module if sub == "" else getattr(module, sub)
            for name, p in net.named_parameters():
                off, shape = self.sub_table[name]
                yield (name if sub == "" else f"{sub}.{name}"), p, si * self.stride + off, shape

    def _adopt(self, module: nn.Module, flat: torch.Tensor):
        """Copy the module's (CPU-initialised) values into agent 0's slot and re-point its Parameters
        at views of the flat buffer."""
        with torch.no_grad():
            for _, p, off, shape in self._slots(module):
                view = flat[off:off + p.numel()].view(shape)
                view.copy_(p.data)
                p.data = view
                p.requires_grad_(False)

    def flat_views(self, flat: torch.Tensor, agent: int = 0) -> "OrderedDict[str, torch.Tensor]":
        base = agent * self.n_floats
        out = OrderedDict()
        for name, p, off, shape in self._slots(self.online_module):
            out[name] = flat[base + off: base + off + p.numel()].view(shape)
        return out

    def load_agent(self, agent: int, state_dict, which: str = "both"):
        """Write a reference-keyed state_dict into population member ``agent``'s slot of the online and/or target
        buffer (a fresh member starts with target == online, like the reference's deepcopy)."""
        with torch.no_grad():
            for flat in ([self.params, self.target] if which == "both" else [self.params] if which == "online" else [self.target]):
                views = self.flat_views(flat, agent)
                for name, v in views.items():
                    v.copy_(state_dict[name])

    def agent_state_dict(self, agent: int, target: bool = False) -> "OrderedDict[str, torch.Tensor]":
        """Views (not copies) of member ``agent``'s tensors under the reference's state_dict keys."""
        return self.flat_views(self.target if target else self.params, agent)

    def param_set(self) -> _lib.ParamSet:
        ps = _lib.ParamSet()
        for k in ("params", "target", "grad", "exp_avg", "exp_avg_sq"):
            setattr(ps, k, C.c_void_p(getattr(self, k).data_ptr()))
        return ps

    def rebind_target(self, new_target_module: nn.Module):
        """TD3_base.load's fallback replaces the target by a deepcopy of the online net (TD3_base.py:43,50):
        here the target buffer is overwritten instead, so the C side keeps its pointers."""
        self.target.copy_(self.params)


class PackedAdam:
    """torch.optim.Adam-compatible shell (state_dict layout, zero_grad) over the packed moments.

    The update itself runs inside the fused CUDA kernels; ``step`` counters live in the agent's
    device state block.  state_dict() follows torch 2.x's Adam: ``state[i] = {step, exp_avg,
    exp_avg_sq}`` in ``parameters()`` order, one param group.
    """

    def __init__(self, family: PackedFamily, lr: float, step_ref, set_step, betas=(0.9, 0.999), eps=1e-8):
        self.family, self._step_ref, self._set_step = family, step_ref, set_step   # step_ref: () -> 0-dim int64 device view
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                             capturable=False, differentiable=False, fused=None, decoupled_weight_decay=False)
        self.param_groups = [dict(self.defaults, params=list(family.online_module.parameters()))]

    def zero_grad(self, set_to_none: bool = True):
        self.family.grad.zero_()

    def state_dict(self):
        step = int(self._step_ref().item())
        names = self.family.names
        state = {}
        if step > 0:
            m, v = self.family.flat_views(self.family.exp_avg), self.family.flat_views(self.family.exp_avg_sq)
            for i, n in enumerate(names):
                state[i] = {"step": torch.tensor(float(step)), "exp_avg": m[n], "exp_avg_sq": v[n]}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(names)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        names = self.family.names
        st = sd["state"]
        if len(sd["param_groups"]) != 1 or len(sd["param_groups"][0]["params"]) != len(names):
            raise ValueError("loaded state dict contains a parameter group that doesn't match the size of optimizer's group")
        m, v = self.family.flat_views(self.family.exp_avg), self.family.flat_views(self.family.exp_avg_sq)
        steps = set()
        for i, n in enumerate(names):
            if i in st:
                m[n].copy_(st[i]["exp_avg"])
                v[n].copy_(st[i]["exp_avg_sq"])
                steps.add(int(float(st[i]["step"])))
            else:
                m[n].zero_()
                v[n].zero_()
                steps.add(0)
        if len(steps) > 1:
            raise ValueError("per-parameter Adam step counts differ; the packed optimiser keeps one step per network")
        self._set_step(steps.pop() if steps else 0)
        g = sd["param_groups"][0]
        if abs(g["lr"] - self.defaults["lr"]) > 0 or tuple(g["betas"]) != tuple(self.defaults["betas"]):
            raise ValueError("loading an optimizer with different lr/betas is not supported (hyper-parameters are fixed "
                             "at construction; build the agent with the checkpoint's lr)")
