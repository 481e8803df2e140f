"""Pin the oracle on the real reference and write tests/golden/*.npz.

Run in the build container (the only place ``/root/reference`` exists):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

For every case it runs the *unmodified* reference classes (imported from
``/root/reference``; ``gym`` / ``pysplishsplash`` stubbed because
TD3_particles.py:2,11 import them without using them on this path) and the
oracle restatement from identical seeds, pre-drawn indices and pre-drawn noise,
asserts the two are bit-identical (losses, Q-values, every parameter of all
four networks after every step) and then stores the reference's outputs.
The fixtures hold seeds + injected draws + results; initial weights and buffer
contents are regenerated from the seeds by the tests.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REFERENCE = os.environ.get("TD3_REFERENCE", "/root/reference")

from oracle import td3_oracle as O  # noqa: E402


def import_reference():
    sys.dont_write_bytecode = True
    for name in ("gym", "pysplishsplash"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if REFERENCE not in sys.path:
        sys.path.insert(0, REFERENCE)
    with contextlib.redirect_stdout(io.StringIO()):
        import TD3_featured as RF
        import TD3_particles as RP
        import my_replay_buffer as RB
    return RF, RP, RB


class _Inject:
    """Replace the two global RNG draws with pre-drawn values."""

    def __init__(self, indices, noise):
        self.indices, self.noise = iter(indices), iter(noise)

    def __enter__(self):
        self._ri, self._rn = np.random.randint, torch.randn_like
        np.random.randint = lambda lo, hi, size=None: next(self.indices)
        torch.randn_like = lambda t: torch.as_tensor(next(self.noise), dtype=t.dtype)
        return self

    def __exit__(self, *a):
        np.random.randint, torch.randn_like = self._ri, self._rn


CASES = {
    # name: dict(kind, dims, norm, policy_freq, batch, steps, rows, CDQ)
    "featured_none": dict(kind="featured", S=17, A=6, norm=None, policy_freq=2, B=32, steps=6, rows=512),
    "featured_layer": dict(kind="featured", S=17, A=6, norm="layer", policy_freq=2, B=32, steps=6, rows=512),
    "featured_pf3_maxact2": dict(kind="featured", S=11, A=3, norm=None, policy_freq=3, B=16, steps=7, rows=300,
                                 max_action=2.0, discount=0.9, tau=0.05),
    "particles_none": dict(kind="particles", F=8, N=64, D=6, A=3, norm=None, policy_freq=2, B=8, steps=4, rows=64, CDQ=True),
    "particles_layer": dict(kind="particles", F=8, N=64, D=6, A=3, norm="layer", policy_freq=2, B=8, steps=4, rows=64, CDQ=True),
    "particles_nocdq": dict(kind="particles", F=5, N=32, D=4, A=2, norm=None, policy_freq=1, B=8, steps=3, rows=64, CDQ=False),
    "particles_wn": dict(kind="particles", F=8, N=64, D=6, A=3, norm="weight_normalization", policy_freq=2, B=8, steps=4,
                         rows=64, CDQ=True),
}


def draws(case, seed=7):
    rs = np.random.RandomState(seed)
    idx = rs.randint(0, case["rows"], size=(case["steps"], case["B"]))
    noise = rs.standard_normal((case["steps"], case["B"], case["A"])).astype(np.float32)
    return idx, noise


def hyper(case):
    return {k: case[k] for k in ("discount", "tau", "max_action") if k in case}


def build_pair(case, RF, RP, RB):
    """(reference agent, reference buffer, oracle agent, oracle buffer) from identical seeds."""
    kw = dict(norm=case["norm"], policy_freq=case["policy_freq"], lr=1e-3, **hyper(case))
    if case["kind"] == "featured":
        obs, act = O.Space(case["S"]), O.Space(case["A"])
        data = O.synthetic_transitions_featured(case["rows"], case["S"], case["A"], seed=0)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):      # TD3_featured.py:103 prints every weight
            ref = RF.TD3(obs, act, **kw)
        torch.manual_seed(0)
        ora = O.TD3Featured(obs, act, **kw)
        rrb, orb = RB.ReplayBuffer_featured(obs, act, max_size=case["rows"]), O.ReplayFeatured(obs, act, case["rows"])
        O.fill_featured(rrb, data)
        O.fill_featured(orb, data)
    else:
        obs, act = (O.Space(case["F"]), O.Space(case["N"], case["D"])), O.Space(case["A"])
        kw.pop("max_action", None)
        data = O.synthetic_transitions_particles(case["rows"], case["F"], case["N"], case["D"], case["A"], seed=0)
        torch.manual_seed(0)
        ref = RP.TD3(obs, act, CDQ=case["CDQ"], **kw)
        torch.manual_seed(0)
        ora = O.TD3Particles(obs, act, CDQ=case["CDQ"], **kw)
        rrb, orb = RB.ReplayBuffer_particles(obs, act, max_size=case["rows"]), O.ReplayParticles(obs, act, case["rows"])
        O.fill_particles(rrb, data)
        O.fill_particles(orb, data)
    return ref, rrb, ora, orb


def nets(agent):
    return dict(actor=agent.actor, critic=agent.critic, actor_target=agent.actor_target, critic_target=agent.critic_target)


def assert_same_params(ref, ora, where):
    for name, rn in nets(ref).items():
        on = nets(ora)[name]
        rsd, osd = rn.state_dict(), on.state_dict()
        assert list(rsd.keys()) == list(osd.keys()), (where, name, list(rsd.keys())[:4], list(osd.keys())[:4])
        for k in rsd:
            assert torch.equal(rsd[k], osd[k]), f"{where}: {name}.{k} differs between reference and oracle"


def run_case(name, case, RF, RP, RB):
    torch.set_num_threads(1)
    idx, noise = draws(case)
    ref, rrb, ora, orb = build_pair(case, RF, RP, RB)
    assert_same_params(ref, ora, f"{name} init")
    out = dict(indices=idx, noise=noise, critic_loss=[], q1=[], target_q=[], digest_actor=[], digest_critic=[],
               digest_actor_target=[], digest_critic_target=[])
    with _Inject(list(idx), list(noise)):
        for t in range(case["steps"]):
            ref.train(rrb, case["B"])                         # the unmodified reference, hooks feed its RNG calls
            ora.train(orb, case["B"], indices=idx[t], noise=noise[t])
            assert_same_params(ref, ora, f"{name} step {t}")
            out["critic_loss"].append(ora.trace["critic_loss"])
            out["q1"].append(ora.trace["q1"].numpy())
            out["target_q"].append(ora.trace["target_q"].numpy())
            for k, net in nets(ref).items():
                out["digest_" + k].append(O.param_digest(net))
    # B=1 API surface
    if case["kind"] == "featured":
        st = np.linspace(-1, 1, case["S"])
        ac = np.linspace(-0.5, 0.5, case["A"])
    else:
        rs = np.random.RandomState(3)
        st = (rs.standard_normal(case["F"]), rs.standard_normal((case["N"], case["D"])))
        ac = np.linspace(-0.5, 0.5, case["A"])
    ra, oa = ref.select_action(st), ora.select_action(st)
    assert np.array_equal(ra, oa)
    rq, oq = ref.eval_q(st, ac), ora.eval_q(st, ac)
    assert all(np.array_equal(a, b) for a, b in zip(rq, oq)) and len(rq) == len(oq)
    out["select_action"] = ra
    out["eval_q"] = np.stack(rq)
    out = {k: np.asarray(v) for k, v in out.items()}
    out["case"] = np.array(repr(case))
    out["versions"] = np.array(f"torch {torch.__version__} numpy {np.__version__}")
    return out


def sample_case(RB):
    """Bit-exact gather fixture: add() path incl. wrap-around, then sample with fixed indices."""
    out = {}
    rs = np.random.RandomState(11)
    obs, act = O.Space(5), O.Space(2)
    rrb, orb = RB.ReplayBuffer_featured(obs, act, max_size=37), O.ReplayFeatured(obs, act, 37)
    n_add = 50                                                    # wraps: ptr = 13, size = 37
    stream = [(rs.standard_normal(5), rs.uniform(-1, 1, 2), rs.standard_normal(5).astype(np.float32),
               float(rs.standard_normal()), float(rs.uniform() < 0.2)) for _ in range(n_add)]
    for row in stream:
        rrb.add(*row)
        orb.add(*row)
    ind = rs.randint(0, 37, size=64)
    r, o = None, orb.sample(64, ind)
    with _Inject([ind], []):
        r = rrb.sample(64)
    for a, b in zip(r, o):
        assert torch.equal(a, b)
    assert (rrb.ptr, rrb.size) == (orb.ptr, orb.size) == (13, 37)
    out["feat_indices"] = ind
    for k, v in zip(O.ReplayFeatured.fields, r):
        out["feat_" + k] = v.numpy()
    out["feat_ptr_size"] = np.array([rrb.ptr, rrb.size])
    # particles
    pobs = (O.Space(3), O.Space(6, 4))
    rrb, orb = RB.ReplayBuffer_particles(pobs, act, max_size=9), O.ReplayParticles(pobs, act, 9)
    for _ in range(14):
        row = ((rs.standard_normal(3), rs.standard_normal((6, 4))), rs.uniform(-1, 1, 2),
               (rs.standard_normal(3), rs.standard_normal((6, 4))), float(rs.standard_normal()), float(rs.uniform() < 0.3))
        rrb.add(*row)
        orb.add(*row)
    ind = rs.randint(0, 9, size=16)
    o = orb.sample(16, ind)
    with _Inject([ind], []):
        r = rrb.sample(16)
    for a, b in zip(r, o):
        assert torch.equal(a, b)
    out["part_indices"] = ind
    for k, v in zip(O.ReplayParticles.fields, r):
        out["part_" + k] = v.numpy()
    out["part_ptr_size"] = np.array([rrb.ptr, rrb.size])
    out["seed"] = np.array(11)
    return out


SMALL_ACTOR, SMALL_Q, SMALL_PQ = (24, 16, 12), (24, 16, 8), (24, 16, 12)


def import_reference_small():
    """The reference's TD3_featured / TD3_particles with ONE edit made in memory: the hard-coded hidden widths
    (TD3_featured.py:19,54; TD3_particles.py:25,77) are replaced by small ones, so that the checkpoints the reference
    classes write (TD3_base.save, unmodified: TD3_base.py:26-34) are a few tens of KB instead of 15-20 MB and can be
    committed.  Everything else -- module structure, state_dict keys, weight-norm reparametrisation, optimiser
    construction, save/load -- is the reference's own code."""
    import importlib.util
    mods = {}
    for name, repl in (("TD3_featured", {"(500,400,300)": repr(SMALL_ACTOR), "(500,400,200)": repr(SMALL_Q)}),
                       ("TD3_particles", {"(500,400,300)": repr(SMALL_PQ)})):
        src = open(os.path.join(REFERENCE, name + ".py")).read()
        for k, v in repl.items():
            assert k in src, (name, k)
            src = src.replace(k, v)
        src = src.split('if __name__ == "__main__":')[0].split("if __name__ == '__main__':")[0]
        mod = types.ModuleType(name + "_small")
        mod.__file__ = os.path.join(REFERENCE, name + ".py")
        with contextlib.redirect_stdout(io.StringIO()):
            exec(compile(src, mod.__file__, "exec"), mod.__dict__)
        mods[name] = mod
    return mods["TD3_featured"], mods["TD3_particles"]


def ondisk_cases(RB):
    """Artefacts WRITTEN BY THE REFERENCE CLASSES (tests/golden/ondisk/): replay-buffer folders of both buffer classes
    (my_replay_buffer.py:28-36,91-99) and policy checkpoints of both agents (TD3_base.py:26-34) incl. LayerNorm and
    weight-normalisation keys and non-empty Adam state, plus what the reference computes after loading them."""
    import shutil
    RFs, RPs = import_reference_small()
    base = os.path.join(ROOT, "tests", "golden", "ondisk")
    shutil.rmtree(base, ignore_errors=True)
    os.makedirs(base)
    expect = {}
    rs = np.random.RandomState(21)
    # ---- buffers ----
    obs, act = O.Space(5), O.Space(2)
    rrb = RB.ReplayBuffer_featured(obs, act, max_size=37)
    for _ in range(50):
        rrb.add(rs.standard_normal(5), rs.uniform(-1, 1, 2), rs.standard_normal(5), float(rs.standard_normal()), float(rs.uniform() < 0.2))
    rrb.save(os.path.join(base, "buffer_featured"))
    ind = rs.randint(0, 37, size=32)
    with _Inject([ind], []):
        got = rrb.sample(32)
    expect["bf_indices"] = ind
    for k, v in zip(O.ReplayFeatured.fields, got):
        expect["bf_" + k] = v.numpy()
    expect["bf_ptr_size"] = np.array([rrb.ptr, rrb.size])
    pobs = (O.Space(3), O.Space(6, 4))
    prb = RB.ReplayBuffer_particles(pobs, act, max_size=9)
    for _ in range(14):
        prb.add((rs.standard_normal(3), rs.standard_normal((6, 4))), rs.uniform(-1, 1, 2),
                (rs.standard_normal(3), rs.standard_normal((6, 4))), float(rs.standard_normal()), float(rs.uniform() < 0.3))
    prb.save(os.path.join(base, "buffer_particles"))
    ind = rs.randint(0, 9, size=16)
    with _Inject([ind], []):
        got = prb.sample(16)
    expect["bp_indices"] = ind
    for k, v in zip(O.ReplayParticles.fields, got):
        expect["bp_" + k] = v.numpy()
    expect["bp_ptr_size"] = np.array([prb.ptr, prb.size])
    # ---- checkpoints: train 3 updates (Adam state populated, one policy step), save, then one more update ----
    torch.set_num_threads(1)
    for tag, kind, norm in (("ckpt_featured_layer", "featured", "layer"), ("ckpt_particles_wn", "particles", "weight_normalization")):
        torch.manual_seed(5)
        if kind == "featured":
            S, A, rows, B = 7, 3, 64, 16
            o, a = O.Space(S), O.Space(A)
            data = O.synthetic_transitions_featured(rows, S, A, seed=2)
            with contextlib.redirect_stdout(io.StringIO()):
                ref = RFs.TD3(o, a, norm=norm, lr=1e-3)
            rb = RB.ReplayBuffer_featured(o, a, max_size=rows)
            O.fill_featured(rb, data)
            st, ac = np.linspace(-1, 1, S), np.linspace(-0.5, 0.5, A)
            dims = dict(S=S, A=A)
        else:
            F, N, D, A, rows, B = 4, 16, 3, 2, 32, 8
            o, a = (O.Space(F), O.Space(N, D)), O.Space(A)
            data = O.synthetic_transitions_particles(rows, F, N, D, A, seed=2)
            ref = RPs.TD3(o, a, norm=norm, lr=1e-3)
            rb = RB.ReplayBuffer_particles(o, a, max_size=rows)
            O.fill_particles(rb, data)
            r2 = np.random.RandomState(3)
            st, ac = (r2.standard_normal(F), r2.standard_normal((N, D))), np.linspace(-0.5, 0.5, A)
            dims = dict(F=F, N=N, D=D, A=A)
        idx = rs.randint(0, rows, size=(4, B))
        noise = rs.standard_normal((4, B, A)).astype(np.float32)
        with _Inject(list(idx), list(noise)):
            for t in range(3):
                ref.train(rb, B)
            ref.save(os.path.join(base, tag))
            expect[tag + "_select_action"] = ref.select_action(st)
            expect[tag + "_eval_q"] = np.stack(ref.eval_q(st, ac))
            ref.train(rb, B)                       # update 4 (a policy step) from the saved state
        for net in ("actor", "critic", "actor_target", "critic_target"):
            for k, v in getattr(ref, net).state_dict().items():
                expect[f"{tag}_after_{net}.{k}"] = v.numpy()
        expect[tag + "_indices"], expect[tag + "_noise"] = idx[3], noise[3]
        expect[tag + "_meta"] = np.array(repr(dict(kind=kind, norm=norm, rows=rows, B=B, total_it_at_save=3, data_seed=2, lr=1e-3, **dims)))
    expect["widths"] = np.array(repr(dict(actor=SMALL_ACTOR, q=SMALL_Q, pq=SMALL_PQ)))
    np.savez_compressed(os.path.join(base, "expect.npz"), **expect)
    return base


def main():
    only = sys.argv[1:]                     # optional: names of the cases to (re)generate
    RF, RP, RB = import_reference()
    gold = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold, exist_ok=True)
    if not only:
        np.savez_compressed(os.path.join(gold, "replay_sample.npz"), **sample_case(RB))
        print("replay_sample: oracle == reference (bit-exact); fixture written")
    if not only or "ondisk" in only:
        print("reference-written buffer folders and checkpoints ->", ondisk_cases(RB))
    for name, case in CASES.items():
        if only and name not in only:
            continue
        res = run_case(name, case, RF, RP, RB)
        np.savez_compressed(os.path.join(gold, f"{name}.npz"), **res)
        print(f"{name}: oracle == reference bit-exact over {case['steps']} steps; "
              f"loss[0]={res['critic_loss'][0]:.6f} loss[-1]={res['critic_loss'][-1]:.6f}")


if __name__ == "__main__":
    main()
