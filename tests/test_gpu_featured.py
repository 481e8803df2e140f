"""TD3_featured update parity: CUDA path (through the C ABI) vs the CPU oracle on identical weights,
indices and noise, and vs the committed golden fixtures made from the real reference.

Stated tolerances (strict-fp32 FFMA path; differences are summation order only):
  per-step critic loss        rel 2e-5
  Q1/Q2, Bellman target       |d| <= 2e-5 * max(1, |Q|)
    norm="layer": 5e-5 for the first two updates, 1e-3 up to the fifth, 5e-2 up to the tenth.  With lr = 1e-3 Adam moves every weight by
    ~lr per step whatever the gradient's magnitude, so elements whose gradient is at summation-order noise level
    take different +-lr steps; through LayerNorm's 1/sigma that difference grows ~2x per update (measured
    7.7e-5 at update 4, 2.5e-4 at update 6) while staying far below the effect of any logic error (>1e-2).
  parameters after N updates  per-tensor relative L2 <= 2e-4 (norm="layer": 1e-2 -- the LayerNorm biases start at
                              0 and have moved only ~lr*N, so their relative error is the Adam-step noise itself), max |d| <= 0.2 * lr * N
    (Adam divides by sqrt(v): an element whose gradient is at rounding-noise level can move by a
     fraction of lr in a different direction; such elements are rare and bounded by lr per step)
"""
import ast
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from helpers import compare_nets, make_featured
from oracle import make_golden as MG

pytestmark = pytest.mark.gpu


def _run(ora, orb, ours, rb, B, steps, A, rows, lr, seed=7, check_every=1, use_graph=True, tol=2e-5, tol_fn=None, tol_params=2e-4):
    rs = np.random.RandomState(seed)
    worst = None
    for t in range(steps):
        idx = rs.randint(0, rows, size=B)
        nz = rs.standard_normal((B, A)).astype(np.float32)
        ora.train(orb, B, indices=idx, noise=nz)
        ours.train(rb, B, indices=idx, noise=nz, use_graph=use_graph)
        if (t + 1) % check_every and t + 1 != steps:
            continue
        dbg = ours.debug_tensors()
        want = ora.trace["critic_loss"]
        got = float(ours.last_critic_loss[0].item())
        tol = tol_fn(t) if tol_fn else tol
        assert abs(got - want) <= tol * max(1.0, abs(want)), (t, got, want)
        q1 = dbg["q"][0, 0].cpu().numpy()
        q2 = dbg["q"][0, 1].cpu().numpy()
        tq = dbg["target_q"][0].cpu().numpy()
        for g_, w_ in ((q1, ora.trace["q1"].numpy()), (q2, ora.trace["q2"].numpy()), (tq, ora.trace["target_q"].numpy())):
            assert np.all(np.abs(g_ - w_) <= tol * np.maximum(1.0, np.abs(w_))), (t, np.abs(g_ - w_).max())
        assert np.array_equal(dbg["indices"][0].cpu().numpy(), idx)
        if ora.trace["actor_loss"] is not None:
            al = float(ours.last_actor_loss[0].item())
            assert abs(al - ora.trace["actor_loss"]) <= tol * max(1.0, abs(ora.trace["actor_loss"])), (t, al)
        worst = compare_nets(ours, ora, tol_rel=tol_params, max_abs=0.2 * lr * (t + 1), label=f"step {t}", abs_floor=lr * (t + 1))
    assert ours.total_it == ora.total_it == steps
    return worst


@pytest.mark.parametrize("norm", [None, "layer"])
@pytest.mark.parametrize("widths", ["fork", "vanilla"])
def test_trajectory_matches_oracle(norm, widths):
    aw, qw = ((500, 400, 300), (500, 400, 200)) if widths == "fork" else ((400, 300), (400, 300))
    ora, orb, ours, rb = make_featured(norm=norm, actor_widths=aw, q_widths=qw, lr=1e-3)
    worst = _run(ora, orb, ours, rb, B=64, steps=10, A=6, rows=512, lr=1e-3,
                 tol_fn=(lambda t: 5e-5 if t < 2 else 1e-3 if t < 5 else 5e-2) if norm == "layer" else None,
                 tol_params=1e-2 if norm == "layer" else 2e-4)
    print(f"featured norm={norm} widths={widths}: worst param rel err {worst}")


def test_layernorm_fork_widths_at_the_default_learning_rate():
    """main.py's default configuration (norm="layer", fork widths, lr = 1e-4): at the reference's own learning rate the
    update-to-update amplification is mild and the whole 10-update trajectory stays within 5e-4 of the oracle."""
    ora, orb, ours, rb = make_featured(norm="layer", lr=1e-4)
    _run(ora, orb, ours, rb, B=64, steps=10, A=6, rows=512, lr=1e-4, tol=5e-4, tol_params=1e-2)


@pytest.mark.parametrize("policy_freq", [1, 3])
def test_policy_freq_and_hyperparameters(policy_freq):
    ora, orb, ours, rb = make_featured(S=11, A=3, rows=300, policy_freq=policy_freq, max_action=2.0, discount=0.9,
                                       tau=0.05, policy_noise=0.3, noise_clip=0.4)
    _run(ora, orb, ours, rb, B=100, steps=7, A=3, rows=300, lr=1e-3)


@pytest.mark.parametrize("mode", ["launches", "graph", "persistent"])
def test_batch_256_every_exec_mode(mode):
    """The three ways td3_train_n can execute the same stage program give the same update."""
    ora, orb, ours, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=2048, lr=1e-4)
    ours.exec_mode = mode
    _run(ora, orb, ours, rb, B=256, steps=6, A=6, rows=2048, lr=1e-4)


def test_long_trajectory_200_updates():
    """N = 200: loss within 1 %, parameters relative L2 <= 1e-3."""
    ora, orb, ours, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=4096, lr=1e-4)
    rs = np.random.RandomState(3)
    for t in range(200):
        idx = rs.randint(0, 4096, size=128)
        nz = rs.standard_normal((128, 6)).astype(np.float32)
        ora.train(orb, 128, indices=idx, noise=nz)
        ours.train(rb, 128, indices=idx, noise=nz)
    got, want = float(ours.last_critic_loss[0].item()), ora.trace["critic_loss"]
    assert abs(got - want) <= 1e-2 * abs(want), (got, want)
    compare_nets(ours, ora, tol_rel=1e-3, max_abs=200 * 1e-4, label="N=200")


@pytest.mark.parametrize("name", ["featured_none", "featured_layer", "featured_pf3_maxact2"])
def test_matches_reference_golden_fixture(name):
    """The same runs the real reference produced in oracle/make_golden.py (no oracle in the loop)."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = ast.literal_eval(str(z["case"]))
    ora, orb, ours, rb = make_featured(S=case["S"], A=case["A"], rows=case["rows"], norm=case["norm"],
                                       policy_freq=case["policy_freq"], lr=1e-3, **MG.hyper(case))
    for t in range(case["steps"]):
        ours.train(rb, case["B"], indices=z["indices"][t], noise=z["noise"][t])
        got = float(ours.last_critic_loss[0].item())
        assert abs(got - z["critic_loss"][t]) <= 2e-5 * max(1.0, abs(z["critic_loss"][t])), (t, got)
        q1 = ours.debug_tensors()["q"][0, 0].cpu().numpy()
        assert np.all(np.abs(q1 - z["q1"][t]) <= 2e-5 * np.maximum(1.0, np.abs(z["q1"][t])))
    from oracle.td3_oracle import param_digest
    for k in ("actor", "critic", "actor_target", "critic_target"):
        got = param_digest(_cpu_module(getattr(ours, k)))
        np.testing.assert_allclose(got, z["digest_" + k][-1], rtol=2e-3, atol=2e-3)


def _cpu_module(mod):
    import copy
    m = copy.copy(mod)

    class _P:
        def __init__(self, ps):
            self._ps = ps

        def parameters(self):
            return self._ps
    return _P([p.detach().cpu() for p in mod.parameters()])


def test_philox_mode_trains_and_is_deterministic():
    """rng='device': two agents with the same seed and weights produce bit-identical trajectories, replayed
    through CUDA graphs with iterations=N; the loss goes down on a fixed synthetic problem."""
    outs = []
    for _ in range(2):
        ora, orb, ours, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=4096, lr=1e-3)
        ours.train(rb, 256, iterations=1)
        first = float(ours.last_critic_loss[0].item())
        ours.train(rb, 256, iterations=299)
        torch.cuda.synchronize()
        outs.append((first, float(ours.last_critic_loss[0].item()), ours.critic.state_dict()["q1.linears.0.weight"].clone()))
        assert ours.total_it == 300
        assert int(ours._state[0].item()) == 300 and int(ours._state[1].item()) == 300 and int(ours._state[2].item()) == 150
    assert outs[0][0] == outs[1][0] and outs[0][1] == outs[1][1] and torch.equal(outs[0][2], outs[1][2])
    assert np.isfinite(outs[0][1])
    idx = ours.debug_tensors()["indices"][0].cpu().numpy()
    assert idx.min() >= 0 and idx.max() < 4096 and len(np.unique(idx)) > 200
    eps = ours.debug_tensors()["eps"][0].cpu().numpy()
    assert np.abs(eps).max() <= 0.5 + 1e-7 and 0.15 < eps.std() < 0.25        # clipped N(0, 0.2^2)


def test_host_rng_mode_reproduces_seeded_reference_run():
    """rng='host' consumes np.random / torch RNG exactly like the reference: a seeded natural run of the
    oracle (no injection) and of the CUDA path stay together."""
    ora, orb, ours, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=1024, lr=1e-3)
    ours.rng = "host"
    np.random.seed(11); torch.manual_seed(11)
    for _ in range(5):
        ora.train(orb, 64)
    want = ora.trace["critic_loss"]
    np.random.seed(11); torch.manual_seed(11)
    for _ in range(5):
        ours.train(rb, 64)
    got = float(ours.last_critic_loss[0].item())
    assert abs(got - want) <= 5e-5 * max(1.0, abs(want)), (got, want)


@pytest.mark.parametrize("norm", [None, "layer"])
@pytest.mark.parametrize("mode", ["persistent", "graph"])
def test_tf32_tensor_core_path_tracks_the_fp32_oracle(norm, mode):
    """precision="tf32": the K >= 64 layer GEMMs run on tcgen05 (operands truncated to 10 mantissa bits, fp32
    accumulation in TMEM).  Stated tolerance against the fp32 CPU oracle over 10 updates at lr = 1e-3:
    Q-values / Bellman target |d| <= 2e-2 * max(1, |Q|), critic loss rel 3e-2, parameters per-tensor relative L2 <= 5e-2."""
    ora, orb, ours, rb = make_featured(norm=norm, actor_widths=(400, 300), q_widths=(400, 300), rows=2048, lr=1e-3,
                                       precision="tf32")
    ours.exec_mode = mode
    rs = np.random.RandomState(7)
    worst_q = worst_l = 0.0
    for t in range(10):
        idx = rs.randint(0, 2048, size=256)
        nz = rs.standard_normal((256, 6)).astype(np.float32)
        ora.train(orb, 256, indices=idx, noise=nz)
        ours.train(rb, 256, indices=idx, noise=nz)
        dbg = ours.debug_tensors()
        want, got = ora.trace["critic_loss"], float(ours.last_critic_loss[0].item())
        worst_l = max(worst_l, abs(got - want) / max(1.0, abs(want)))
        for g_, w_ in ((dbg["q"][0, 0], ora.trace["q1"]), (dbg["q"][0, 1], ora.trace["q2"]), (dbg["target_q"][0], ora.trace["target_q"])):
            g_, w_ = g_.cpu().numpy(), w_.numpy()
            worst_q = max(worst_q, float((np.abs(g_ - w_) / np.maximum(1.0, np.abs(w_))).max()))
    worst_p = compare_nets(ours, ora, tol_rel=5e-2, max_abs=1.0, label="tf32", abs_floor=1e-3 * 10)
    print(f"tf32 norm={norm} mode={mode}: worst |dQ| {worst_q:.2e}, loss rel {worst_l:.2e}, params {worst_p}")
    assert worst_q <= 2e-2 and worst_l <= 3e-2


def test_tail_fused_update_matches_the_unfused_sequences(monkeypatch):
    """The default execution of a plain-MLP update (first-layer dW + Adam inside the optimiser launch, actor forward
    riding with the target pass, one-wave tile widths: engine.cu plan_agent / layout_stage) and the unfused stage
    sequences are the same arithmetic up to summation order: strict-fp32 mode, identical seeds -> parameters agree to
    fp32 round-off after a policy_freq cycle and stay within 1e-5 relative after 20 updates."""
    def run(unfused):
        if unfused:
            monkeypatch.setenv("TD3_NO_TAIL_FUSION", "1")
            monkeypatch.setenv("TD3_NO_WAVE_FIT", "1")
        else:
            monkeypatch.delenv("TD3_NO_TAIL_FUSION", raising=False)
            monkeypatch.delenv("TD3_NO_WAVE_FIT", raising=False)
        _, _, ours, rb = make_featured(rows=2048, actor_widths=(400, 300), q_widths=(400, 300), precision="fp32")
        out = []
        for n in (2, 18):
            ours.train(rb, 256, iterations=n)
            torch.cuda.synchronize()
            out.append({k: {name: t.detach().clone() for name, t in getattr(ours, k).state_dict().items()}
                        for k in ("actor", "critic", "actor_target", "critic_target")})
        return out
    fused, unfused = run(False), run(True)
    for tol, a, b in ((2e-6, fused[0], unfused[0]), (1e-5, fused[1], unfused[1])):
        for k in a:
            for name in a[k]:
                x, y = a[k][name].double(), b[k][name].double()
                rel = float((x - y).norm() / y.norm().clamp_min(1e-30))
                assert rel <= tol, f"{k}.{name}: fused vs unfused rel {rel:.3e} > {tol}"


@pytest.mark.parametrize("unfused", [False, True])
@pytest.mark.parametrize("aw,qw", [((64,), (48,)), ((96, 64, 48, 32), (80, 64, 48, 32))])
def test_shallow_and_deep_networks_match_oracle(monkeypatch, unfused, aw, qw):
    """Two-layer networks (one hidden layer: the output layer's dW rides with the first layer's stage) and five-layer
    ones, through the tail-fused and the unfused sequences, against the fp32 oracle."""
    if unfused:
        monkeypatch.setenv("TD3_NO_TAIL_FUSION", "1")
    else:
        monkeypatch.delenv("TD3_NO_TAIL_FUSION", raising=False)
    ora, orb, ours, rb = make_featured(norm=None, actor_widths=aw, q_widths=qw, lr=1e-3, precision="fp32")
    worst = _run(ora, orb, ours, rb, B=64, steps=8, A=6, rows=512, lr=1e-3, tol_params=2e-4)
    print(f"featured widths={aw}/{qw} unfused={unfused}: worst param rel err {worst}")


def test_ragged_batch_and_small_dims_through_the_fused_path():
    """B = 37 (partial row blocks in the front / head / first-layer tiles), S = 5, A = 2, widths that are not
    multiples of 16, policy update every step: the tail-fused default path against the fp32 oracle."""
    ora, orb, ours, rb = make_featured(S=5, A=2, rows=300, norm=None, actor_widths=(40, 24), q_widths=(44, 20), lr=1e-3,
                                       precision="fp32", policy_freq=1)
    worst = _run(ora, orb, ours, rb, B=37, steps=6, A=2, rows=300, lr=1e-3, tol_params=2e-4)
    print(f"ragged featured: worst param rel err {worst}")


@pytest.mark.parametrize("precision", ["fp32", "tf32"])
def test_persistent_kernel_many_updates_per_launch_is_bit_identical_to_plain_launches(precision, monkeypatch):
    """the persistent kernel runs `iterations` updates inside ONE cooperative launch; the first stage
    of an update (the front sampling kernel) reads parameters the previous update's optimiser stage writes, so the
    inter-update barrier must stay.  60 updates in one launch against 60 updates as plain stage-by-stage launches, same
    Philox key: every parameter bit-identical (a race would show up as run-to-run differences long before that).
    With TD3_CHAIN=1 the TF32 launch sequence would be the layer-fused chain (chain.cuh), a different program from the
    stage program the persistent kernel walks: the persistent kernel is always compared with the stage-by-stage
    launches (TD3_NO_CHAIN), and the graph with the plain launches of whatever sequence is the default."""
    outs = {}
    for mode in ("launches", "persistent", "graph", "launches_stage_program"):
        if mode == "launches_stage_program":
            monkeypatch.setenv("TD3_NO_CHAIN", "1")
        _, _, ours, rb = make_featured(actor_widths=(400, 300), q_widths=(400, 300), rows=4096, lr=1e-3, precision=precision)
        ours.exec_mode = "launches" if mode == "launches_stage_program" else mode
        ours.train(rb, 256, iterations=60)
        torch.cuda.synchronize()
        outs[mode] = {k: {n: t.detach().clone() for n, t in getattr(ours, k).state_dict().items()}
                      for k in ("actor", "critic", "actor_target", "critic_target")}
        assert int(ours._state[1].item()) == 60 and int(ours._state[2].item()) == 30
    for mode, ref in (("persistent", "launches_stage_program"), ("graph", "launches")):
        for k, sd in outs[ref].items():
            for n, t in sd.items():
                assert torch.equal(t, outs[mode][k][n]), f"{mode} vs {ref}: {k}.{n} differs (max |d| {(t - outs[mode][k][n]).abs().max().item():.3e})"
