"""Shared builders for the GPU parity tests: an oracle agent + a td3_b200 agent with identical
weights, buffers filled with identical synthetic transitions."""
import numpy as np
import torch

from oracle import td3_oracle as O


def philox4x32_10(counter, seed):
    """Reference Philox4x32-10 in pure Python (Salmon et al. 2011) for known-answer checks."""
    M0, M1, W0, W1, mask = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    c = list(counter)
    k = [seed & mask, (seed >> 32) & mask]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & mask, p1 & mask, ((p0 >> 32) ^ c[3] ^ k[1]) & mask, p0 & mask]
        k = [(k[0] + W0) & mask, (k[1] + W1) & mask]
    return c


def philox_index(seed, stream, step, elem, size):
    c = philox4x32_10([elem, stream, step & 0xFFFFFFFF, step >> 32], seed)
    return (((c[0] << 32) | c[1]) * size) >> 64


def make_featured(S=17, A=6, rows=512, norm=None, actor_widths=(500, 400, 300), q_widths=(500, 400, 200), lr=1e-3,
                  seed=0, precision=None, **hyper):
    from td3_b200.TD3_featured import TD3
    from td3_b200.my_replay_buffer import ReplayBuffer_featured
    obs, act = O.Space(S), O.Space(A)
    torch.manual_seed(seed)
    ora = O.TD3Featured(obs, act, norm=norm, lr=lr, actor_widths=actor_widths, q_widths=q_widths, **hyper)
    ours = TD3(obs, act, norm=norm, lr=lr, actor_widths=actor_widths, q_widths=q_widths, seed=1, precision=precision, **hyper)
    copy_weights(ora, ours)
    data = O.synthetic_transitions_featured(rows, S, A, seed=0)
    orb = O.ReplayFeatured(obs, act, rows)
    O.fill_featured(orb, data)
    rb = ReplayBuffer_featured(obs, act, max_size=rows)
    rb.add_batch(**data)
    return ora, orb, ours, rb


def make_particles(F=8, N=64, D=6, A=3, rows=64, norm=None, CDQ=True, lr=1e-3, seed=0, precision=None, **hyper):
    from td3_b200.TD3_particles import TD3
    from td3_b200.my_replay_buffer import ReplayBuffer_particles
    obs, act = (O.Space(F), O.Space(N, D)), O.Space(A)
    torch.manual_seed(seed)
    ora = O.TD3Particles(obs, act, norm=norm, CDQ=CDQ, lr=lr, **hyper)
    ours = TD3(obs, act, norm=norm, CDQ=CDQ, lr=lr, seed=1, precision=precision, **hyper)
    copy_weights(ora, ours)
    data = O.synthetic_transitions_particles(rows, F, N, D, A, seed=0)
    orb = O.ReplayParticles(obs, act, rows)
    O.fill_particles(orb, data)
    rb = ReplayBuffer_particles(obs, act, max_size=rows)
    rb.add_batch(**data)
    return ora, orb, ours, rb


def copy_weights(ora, ours):
    for k in ("actor", "critic", "actor_target", "critic_target"):
        getattr(ours, k).load_state_dict(getattr(ora, k).state_dict())


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def compare_nets(ours, ora, tol_rel, max_abs, label="", abs_floor=0.0):
    """Every parameter tensor of all four networks: relative L2 error <= tol_rel and max |diff| <= max_abs.
    abs_floor: per-element magnitude below which a tensor's own norm stops being the yardstick -- LayerNorm biases
    start at exactly 0 and have only moved by ~lr per update, so their error is measured against that movement."""
    worst = (0.0, 0.0, "")
    for k in ("actor", "critic", "actor_target", "critic_target"):
        osd = getattr(ora, k).state_dict()
        gsd = getattr(ours, k).state_dict()
        assert list(osd.keys()) == list(gsd.keys()), (k, list(osd.keys())[:3], list(gsd.keys())[:3])
        for name in osd:
            a, b = gsd[name].detach().cpu().numpy(), osd[name].numpy()
            m = float(np.abs(a - b).max())
            r = float(np.linalg.norm((a - b).astype(np.float64)) /
                      max(np.linalg.norm(b.astype(np.float64)), abs_floor * np.sqrt(b.size), 1e-30))
            if r > worst[0]:
                worst = (r, m, f"{k}.{name}")
            assert r <= tol_rel and m <= max_abs, f"{label} {k}.{name}: rel {r:.3e} (tol {tol_rel}) max|d| {m:.3e} (tol {max_abs})"
    return worst
