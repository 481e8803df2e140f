#!/usr/bin/env python
"""bench.py -- TD3 gradient updates/sec (batch 256), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2]

A "step" is one complete TD3 update: sample 256 transitions from the replay buffer, target step,
twin-critic forward/backward + Adam, and on every policy_freq-th step the actor update + Polyak
(TD3_featured.py:123-171).  Default workload = BASELINE configs[1] ("cfg2"): S=17, A=6, 400-300 MLPs,
batch 256, 1M-row device-resident replay buffer, one agent per GPU.

  value  updates/s with everything resident in HBM (the K updates are K CUDA-graph replays of the update's
         stage kernels -- or one cooperative launch of the persistent update kernel with --exec-mode
         persistent -- CUDA-event timed, max over ranks).
  e2e    the same metric through the public Python API the reference's main.py loop uses, per step:
         replay_buffer.add(one host transition -> pinned -> H2D), policy.train(replay_buffer, 256), and
         policy.wait_critic_loss(): the host blocks until THIS step's critic loss (an 8-byte word the critic-head
         kernel stores to pinned host memory) has landed, then enqueues the next step behind the optimiser kernels
         still running; the timed region ends with a full device synchronise.  e2e.drain_value is the same loop
         with a D2H copy behind the whole update and a stream synchronise every step.
  roofline  the dominant kernel (the tcgen05 GEMM stage kernel): algorithmic flops of its launches in one
         policy_freq cycle / their in-situ duration (CUDA events around graph replays of the first k launches,
         td3_debug_prefix_times), against the TF32 peak from MEASURED_PEAKS.json; whole-update floors beside it.
  torch_eager_gpu  the oracle port of the reference with its networks on the GPU (eager PyTorch): a comparator.
  population  8 independent agents per GPU in lock-step (BASELINE config 5) + the HBM-side roofline of its
         optimiser launch.
  particles_cfg4  (default workload, rank 0) BASELINE config 4 -- TD3_particles, 1024 particles, batch 256 -- timed the same
         way (device-resident graph replays) with the fused set-encoder launches' share and their algorithmic TFLOP/s
         against the TF32 peak measured in this run; `--workload cfg4` gives the full line for it.
  --impl reference times the CPU oracle port of the reference (oracle/td3_oracle.py, torch CPU ops --
         /root/reference is not present on the GPU box) on the host cores with the same config.
With N > 1 (torchrun) every rank runs an independent agent/seed on its own GPU (weak scaling, no
data-path collective: SURVEY.md 8e); value = total updates of all ranks / max-over-ranks time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: description of the synthetic job
    "cfg1": dict(kind="featured", S=17, A=6, aw=(400, 300), qw=(400, 300), norm=None, B=100, rows=10_000),
    "cfg2": dict(kind="featured", S=17, A=6, aw=(400, 300), qw=(400, 300), norm=None, B=256, rows=1_000_000),
    "cfg2_S32": dict(kind="featured", S=32, A=6, aw=(400, 300), qw=(400, 300), norm=None, B=256, rows=200_000),
    "cfg5b": dict(kind="featured", S=17, A=6, aw=(400, 300), qw=(400, 300), norm=None, B=8192, rows=1_000_000),
    "cfg3": dict(kind="featured", S=17, A=6, aw=(500, 400, 300), qw=(500, 400, 200), norm=None, B=256, rows=1_000_000),
    "cfg3_layer": dict(kind="featured", S=17, A=6, aw=(500, 400, 300), qw=(500, 400, 200), norm="layer", B=256,
                       rows=1_000_000),
    "cfg4": dict(kind="particles", F=8, N=1024, D=6, A=3, norm=None, B=256, rows=8192),
    "cfg4_layer": dict(kind="particles", F=8, N=1024, D=6, A=3, norm="layer", B=256, rows=8192),
}
HYPER = dict(discount=0.99, tau=0.005, policy_noise=0.2, noise_clip=0.5, policy_freq=2)


def algorithmic_per_update(w):
    """(GFLOP, MB) per update averaged over the policy_freq cycle -- SURVEY.md 8d formulas."""
    pf = HYPER["policy_freq"]
    B = w["B"]

    def macs(dims):
        return sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1))

    def nparams(dims):
        return sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
    if w["kind"] == "featured":
        da = [w["S"], *w["aw"], w["A"]]
        dq = [w["S"] + w["A"], *w["qw"], 1]
        Wa, Wq, Wa0, Wq0, E, Eb = macs(da), macs(dq), da[0] * da[1], dq[0] * dq[1], 0, 0
        Pa, Pc = nparams(da), 2 * nparams(dq)
        if w["norm"] == "layer":
            Pa += 2 * sum(w["aw"])
            Pc += 4 * sum(w["qw"])
        row_bytes = (2 * w["S"] + w["A"] + 2) * 4
    else:
        da = [128 + w["F"], 500, 400, 300, w["A"]]
        dq = [128 + w["F"] + w["A"], 500, 400, 300, w["A"]]
        Wa, Wq, Wa0, Wq0 = macs(da), macs(dq), da[0] * da[1], dq[0] * dq[1]
        E = w["N"] * w["D"] * 256 + w["N"] * 256 * 128
        Eb = w["N"] * w["D"] * 256 + 2 * w["N"] * 256 * 128
        enc_p = w["D"] * 256 + 256 + 256 * 128 + 128
        Pa, Pc = nparams(da) + enc_p, 2 * (nparams(dq) + enc_p)
        row_bytes = (2 * (w["F"] + w["N"] * w["D"]) + w["A"] + 2) * 4
    critic = 2 * B * ((Wa + E) + 2 * (Wq + E) + 2 * (Wq + E) + 2 * (2 * Wq - Wq0 + Eb))
    actor = 2 * B * ((Wa + E) + (Wq + E) + Wq + (2 * Wa - Wa0 + Eb))
    flops = critic + actor / pf
    nbytes = 28 * Pc + (28 * Pa + 12 * (Pa + Pc)) / pf + 2 * row_bytes * B
    return flops / 1e9, nbytes / 1e6


def stage_gflop_per_cycle(w):
    """GFLOP of the contractions the stage kernels run in one policy_freq cycle (policy_freq - 1 critic-only updates +
    one policy update): everything of algorithmic_per_update except first layers and output heads (featured), which
    live in the front / head / apply kernels."""
    pf, B = HYPER["policy_freq"], w["B"]
    if w["kind"] != "featured":
        return algorithmic_per_update(w)[0] * pf

    def hidden(dims):
        return sum(dims[i] * dims[i + 1] for i in range(1, len(dims) - 2))
    Ha, Hq = hidden([w["S"], *w["aw"], w["A"]]), hidden([w["S"] + w["A"], *w["qw"], 1])
    critic = 2 * B * (Ha + 2 * Hq + 2 * Hq + 2 * (2 * Hq))          # target actor, target critics, critics fwd, critics bwd (dW + dX)
    actor = 2 * B * (Ha + Hq + Hq + 2 * Ha)                         # actor fwd, Q1 fwd, Q1 bwd (dX only), actor bwd (dW + dX)
    return (critic * pf + actor) / 1e9


def encoder_gflop_per_cycle(w):
    """GFLOP of the particle-set encoder passes in one policy_freq cycle (the E / Eb terms of algorithmic_per_update)."""
    pf, B = HYPER["policy_freq"], w["B"]
    E = w["N"] * w["D"] * 256 + w["N"] * 256 * 128
    Eb = w["N"] * w["D"] * 256 + 2 * w["N"] * 256 * 128
    critic = 2 * B * (E + 2 * E + 2 * E + 2 * Eb)          # target actor, target critics, critics forward, critics backward
    actor = 2 * B * (E + E + Eb)                           # actor forward, Q1 forward, actor backward
    return (critic * pf + actor) / 1e9


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]), bf16_sustained=float(p["bf16_tflops_sustained"]),
                    source="measured (MEASURED_PEAKS.json)")
    except Exception:
        return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


# fp32 FFMA peak of the chip: 148 SMs x 128 FMA lanes x 2 flop x 1.965 GHz (nominal, at the maximum SM clock)
FFMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12


def measure_tf32_peak():
    """TF32 tensor-core peak of THIS GPU, measured the way MEASURED_PEAKS.json measures bf16: cuBLAS matmul 8192^3 with
    allow_tf32, best of 5, CUDA events (a library GEMM used as a yardstick only; nothing on the product path calls it)."""
    import torch
    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        n = 8192
        a = torch.randn(n, n, device="cuda")
        b = torch.randn(n, n, device="cuda")
        best = 0.0
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
        torch.cuda.empty_cache()
        return best
    except Exception:
        return None


def hbm_kernel_rooflines(peaks):
    """Achieved HBM GB/s of the bandwidth-bound kernels of the path, each launched through the C ABI on buffers larger
    than the 126 MB L2 and timed with CUDA events on the launching stream: replay gather (cfg2 and cfg4 row shapes,
    algorithmic bytes = 2 x row bytes x batch: every sampled row is read once and written once), Adam (28 B / parameter),
    Polyak (12 B / parameter)."""
    import ctypes as C
    import numpy as np
    import torch
    from td3_b200 import _lib as L_, synthetic as O
    from td3_b200.my_replay_buffer import ReplayBuffer_featured, ReplayBuffer_particles
    lib = L_.require_cuda()
    out = {}

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps            # us per launch

    # ---- replay gather ----
    for name, rb, B in (("gather_cfg2_rows", ReplayBuffer_featured(O.Space(17), O.Space(6), max_size=1_000_000), 256),
                        ("gather_cfg4_rows", ReplayBuffer_particles((O.Space(8), O.Space(1024, 6)), O.Space(3), max_size=8192), 256)):
        rb._all_rows.normal_()
        rb.size = rb.max_size
        g = torch.Generator(device="cuda")
        g.manual_seed(1)
        idxs = [torch.randint(0, rb.max_size, (B,), device="cuda", generator=g) for _ in range(16)]
        it = [0]

        def one():
            rb.sample(B, indices=idxs[it[0] % 16])
            it[0] += 1
        us = timed(one, 16)
        nbytes = 2.0 * rb.row_floats * 4 * B
        out[name] = {"kernel": "td3::gather_kernel" + (" (cp.async.bulk staging of the particle sets)" if "cfg4" in name else ""),
                     "bound": "hbm", "us_per_launch": us, "algorithmic_mb_per_launch": nbytes / 1e6,
                     "achieved": nbytes / (us * 1e-6) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": nbytes / (us * 1e-6) / 1e9 / peaks["hbm"],
                     "buffer_mb": rb._all_rows.numel() * 4 / 1e6, "rows_per_launch": B}
        del rb
    # ---- Adam / Polyak over a packed buffer of 48 M parameters (192 MB per array) ----
    n = 48 * 1024 * 1024
    p, gr, m, v, t = (torch.randn(n, device="cuda") * 0.01 for _ in range(5))
    v.abs_()
    s = L_.stream_ptr
    vp = lambda x: C.c_void_p(x.data_ptr())
    for name, args, bpp in (("adam", (vp(p), vp(gr), vp(m), vp(v), None), 28.0), ("polyak", (vp(p), None, None, None, vp(t)), 12.0),
                            ("adam_polyak_fused", (vp(p), vp(gr), vp(m), vp(v), vp(t)), 40.0)):
        def one():
            L_.check(lib.adam_polyak_step(*args, n, 10, 1e-4, 0.9, 0.999, 1e-8, 0.005, s()))
        us = timed(one, 4)
        nbytes = bpp * n
        out[name] = {"kernel": "td3::adam_polyak_kernel", "bound": "hbm", "us_per_launch": us,
                     "algorithmic_mb_per_launch": nbytes / 1e6, "achieved": nbytes / (us * 1e-6) / 1e9, "peak": peaks["hbm"],
                     "unit": "GB/s", "frac": nbytes / (us * 1e-6) / 1e9 / peaks["hbm"], "parameters": n}
    del p, gr, m, v, t
    torch.cuda.empty_cache()
    return out


def time_dp_critic(rank, world, K):
    """BASELINE config 5b: ONE 400-300 agent at global batch 8192, the batch sharded over the ranks, one sum-all-reduce of
    the packed critic gradient per update (+ one of the actor gradient on policy steps).  Strong scaling: total work is
    fixed.  Returns per-update times with and without the collectives (the latter is timing only: replicas diverge)."""
    import torch
    import torch.distributed as dist
    from td3_b200.data_parallel import DataParallelTD3
    w = WORKLOADS["cfg5b"]
    if w["B"] % world:
        return None
    agent, rb = build_ours(w, seed=4242, rows=200_000)        # same seed on every rank: replicas
    dp = DataParallelTD3(agent)
    res = {}
    for label, comm in (("with_allreduce", True), ("compute_only", False)):
        dp.communicate = comm
        for _ in range(10):
            dp.train(rb, w["B"])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            dp.train(rb, w["B"])
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[label] = float(t[0]) / K                       # ms per update, max over ranks
    pc, pa = int(agent._critic_family.params.numel()), int(agent._actor_family.params.numel())
    out = {"value": 1000.0 / res["with_allreduce"], "unit": "updates/s", "global_batch": w["B"], "rows_per_rank": w["B"] // world,
           "ms_per_update": res["with_allreduce"], "ms_per_update_compute_only": res["compute_only"],
           "collective_share": max(0.0, 1.0 - res["compute_only"] / res["with_allreduce"]), "steps": K, "scaling": "strong",
           "allreduce_bytes_per_update": 4 * pc + 4 * pa / HYPER["policy_freq"], "mode": dp.mode,
           "what": "one agent, global batch 8192 split over the ranks (world-size-invariant Philox batch), gradients summed "
                   "across ranks before identical Adam steps; value = updates/s of the whole job (max over ranks)"}
    del agent, rb, dp
    torch.cuda.empty_cache()
    return out


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    """Polls NVML for SM clock + throttle reasons during the timed region."""

    BAD = {0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}
    NOTE = {0x4: "sw_power_cap"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in {**self.BAD, **self.NOTE}.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# --------------------------------------------------------------------------- builders
def build_ours(w, seed, rows=None, n_agents=1, rng_seed=None):
    import torch
    from td3_b200 import synthetic as O             # inputs of the named shapes; nothing on this arm touches oracle/
    rows = rows or w["rows"]
    if w["kind"] == "featured":
        from td3_b200.TD3_featured import TD3
        from td3_b200.my_replay_buffer import ReplayBuffer_featured
        obs, act = O.Space(w["S"]), O.Space(w["A"])
        torch.manual_seed(seed)
        agent = TD3(obs, act, lr=1e-4, norm=w["norm"], actor_widths=w["aw"], q_widths=w["qw"],
                    seed=seed + 1 if rng_seed is None else rng_seed, max_action=1, n_agents=n_agents, **HYPER)
        rb = ReplayBuffer_featured(obs, act, max_size=rows, n_agents=n_agents)
        for i in range(n_agents):
            rb.add_batch(agent=i, **O.transitions_featured(rows, w["S"], w["A"], seed=i))
    else:
        from td3_b200.TD3_particles import TD3
        from td3_b200.my_replay_buffer import ReplayBuffer_particles
        obs, act = (O.Space(w["F"]), O.Space(w["N"], w["D"])), O.Space(w["A"])
        torch.manual_seed(seed)
        agent = TD3(obs, act, lr=1e-4, norm=w["norm"], seed=seed + 1, **HYPER)
        rb = ReplayBuffer_particles(obs, act, max_size=rows)
        rb.add_batch(**O.transitions_particles(rows, w["F"], w["N"], w["D"], w["A"], seed=0))
    return agent, rb


def build_oracle(w, seed, rows):
    import torch
    from oracle import td3_oracle as O
    torch.manual_seed(seed)
    if w["kind"] == "featured":
        obs, act = O.Space(w["S"]), O.Space(w["A"])
        agent = O.TD3Featured(obs, act, lr=1e-4, norm=w["norm"], actor_widths=w["aw"], q_widths=w["qw"], max_action=1, **HYPER)
        rb = O.ReplayFeatured(obs, act, rows)
        O.fill_featured(rb, O.synthetic_transitions_featured(rows, w["S"], w["A"], seed=0))
    else:
        obs, act = (O.Space(w["F"]), O.Space(w["N"], w["D"])), O.Space(w["A"])
        agent = O.TD3Particles(obs, act, lr=1e-4, norm=w["norm"], **HYPER)
        rb = O.ReplayParticles(obs, act, rows)
        O.fill_particles(rb, O.synthetic_transitions_particles(rows, w["F"], w["N"], w["D"], w["A"], seed=0))
    return agent, rb


def time_cpu(w, steps, warmup, threads_list, rows):
    """Reference arm / cpu_baseline: oracle port on the host cores; best over thread counts."""
    import numpy as np
    import torch
    best = None
    for th in threads_list:
        torch.set_num_threads(th)
        agent, rb = build_oracle(w, 0, rows)
        agent.keep_trace = False                       # time the reference's work only (no parity bookkeeping)
        np.random.seed(7)
        torch.manual_seed(7)
        for _ in range(warmup):
            agent.train(rb, w["B"])
        t0 = time.perf_counter()
        for _ in range(steps):
            agent.train(rb, w["B"])
        dt = time.perf_counter() - t0
        ups = steps / dt
        if best is None or ups > best[0]:
            best = (ups, th, dt)
    return best


def time_torch_gpu(w, steps, warmup, rows):
    """The same oracle port run the way the reference runs on a GPU box (TD3_featured.py:10 picks cuda when present):
    networks and optimiser state on the device, float64 host replay arrays, five (seven) pageable H2D copies per sample
    (my_replay_buffer.py:122-128), eager PyTorch kernels, TF32 off (torch default).  An 'existing Blackwell path'
    comparator beside the CPU arm; not the reference arm."""
    import numpy as np
    import torch
    agent, rb = build_oracle(w, 0, rows)
    agent.keep_trace = False
    for name in ("actor", "actor_target", "critic", "critic_target"):
        getattr(agent, name).cuda()
    host_sample = rb.sample
    rb.sample = lambda *a, **k: tuple(t.cuda() for t in host_sample(*a, **k))
    np.random.seed(7)
    torch.manual_seed(7)
    for _ in range(warmup):
        agent.train(rb, w["B"])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        agent.train(rb, w["B"])
    torch.cuda.synchronize()
    return steps / (time.perf_counter() - t0)


def cpu_sample_size(w):
    # ~10-30 s of CPU work in total
    return (5, 1) if w["kind"] == "particles" else (150, 15)


# --------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dp", action="store_true", help="skip the batch-8192 data-parallel update (config 5b) at N = 1")
    ap.add_argument("--precision", default=os.environ.get("TD3_PRECISION", "tf32"), choices=["tf32", "fp32"],
                    help="tf32: layer GEMMs on tcgen05 tensor cores (default); fp32: strict-fp32 FFMA tiles")
    ap.add_argument("--population", type=int, default=8,
                    help="also time a population of this many independent agents per GPU stepped in lock-step (0 = skip)")
    ap.add_argument("--exec-mode", default=os.environ.get("TD3_EXEC_MODE", "graph"), choices=["persistent", "graph", "launches"])
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    os.environ["TD3_PRECISION"], os.environ["TD3_EXEC_MODE"] = args.precision, args.exec_mode
    dtype = "tf32" if args.precision == "tf32" else "f32"
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 3)
    gflop, mbytes = algorithmic_per_update(w)
    config = {"workload": f"{args.workload}: " + (
        f"TD3_{w['kind']} S={w['S']} A={w['A']} actor {w['aw']} critic {w['qw']} norm={w['norm']}" if w["kind"] == "featured"
        else f"TD3_particles F={w['F']} N={w['N']} D={w['D']} A={w['A']} norm={w['norm']}"),
        "batch": w["B"], "replay_rows": w["rows"], "policy_freq": HYPER["policy_freq"], "agents_per_gpu": 1,
        "precision": args.precision + (" (tcgen05 kind::tf32 GEMMs, fp32 accumulate and storage)" if args.precision == "tf32" else " (FFMA)"),
        "exec_mode": args.exec_mode,
        "algorithmic_gflop_per_update": round(gflop, 4), "algorithmic_mb_per_update": round(mbytes, 3),
        "l2": "inputs larger than L2: the replay buffer (192 MB at 1M rows) is sampled uniformly at random every step; "
              "the parameters are re-used across steps by the algorithm itself"}

    ncpu = os.cpu_count() or 1
    threads_list = sorted({t for t in (1, 4, 8, 16, ncpu) if t <= ncpu})

    if args.impl == "reference":
        if rank != 0:
            return
        cpu_rows = min(w["rows"], 100_000)            # float64 host buffer; sampling cost does not depend on rows
        steps = max(150, min(K, 300)) if w["kind"] == "featured" else min(max(K, 3), 5)     # >= 150 updates: a 20-step sample is noise
        warm = min(W, 30 if w["kind"] == "featured" else 1)
        ups, th, dt = time_cpu(w, steps, warm, threads_list, cpu_rows)
        line = {"impl": "reference", "metric": "TD3 gradient updates/sec (batch 256)", "value": ups, "unit": "updates/s",
                "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": 1000.0 / ups, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": th, "kind": "port",
                                 "sample": f"{steps} updates after {warm} warm-up, oracle port of the reference "
                                           f"(torch {__import__('torch').__version__} CPU), best of threads {threads_list}, "
                                           f"os.cpu_count()={ncpu}"},
                "e2e": {"value": ups, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    agent, rb = build_ours(w, seed=100 + rank)
    lib = agent._lib
    B = w["B"]
    peaks = measured_peaks()

    # ---------------- device-resident throughput ----------------
    agent.train(rb, B, iterations=W)
    barrier()
    launches0 = lib.td3_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        barrier()
        ev0.record()
        agent.train(rb, B, iterations=K)
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = lib.td3_launch_count() - launches0
    clocks = clk.summary()

    # ---------------- end to end through the public API ----------------
    rs = np.random.RandomState(1234 + rank)
    if w["kind"] == "featured":
        new_rows = [(rs.standard_normal(w["S"]), rs.uniform(-1, 1, w["A"]), rs.standard_normal(w["S"]),
                     float(rs.standard_normal()), 0.0) for _ in range(64)]
    else:
        new_rows = [((rs.standard_normal(w["F"]), rs.standard_normal((w["N"], w["D"])).astype(np.float32)),
                     rs.uniform(-1, 1, w["A"]),
                     (rs.standard_normal(w["F"]), rs.standard_normal((w["N"], w["D"])).astype(np.float32)),
                     float(rs.standard_normal()), 0.0) for _ in range(8)]
    loss_host = torch.zeros(1).pin_memory()
    loss_dev = agent.last_critic_loss                                # device view of the step's result
    stream = torch.cuda.current_stream()
    Ke = min(K, 2000)

    def e2e_step(i):
        rb.add(*new_rows[i % len(new_rows)])                       # host transition -> pinned slot -> H2D
        agent.train(rb, B)                                          # the call main.py:269 makes
        return agent.wait_critic_loss()                             # D2H: the step's loss, stored to pinned host memory
                                                                    # by the critic-head kernel; the host blocks until it lands

    def e2e_step_drain(i):
        rb.add(*new_rows[i % len(new_rows)])
        agent.train(rb, B)
        loss_host.copy_(loss_dev, non_blocking=True)                # D2H copy behind the whole update
        stream.synchronize()                                        # ... and drain the stream before the next step
        return float(loss_host[0])

    for i in range(min(W, 50)):
        e2e_step(i)
    with ClockSampler(local_rank) as clk2:
        barrier()
        t0 = time.perf_counter()
        last = 0.0
        for i in range(Ke):
            last = e2e_step(i)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        barrier()
    clocks_e2e = clk2.summary()
    agent_status_live = bool(getattr(agent, "_status_live", False))
    # the step's loss must be the one the device holds once everything has drained
    mirror_ok = abs(last - float(loss_dev[0])) <= 1e-6 * max(1.0, abs(last))
    if not mirror_ok:
        print(f"WARNING: host mirror of the critic loss {last} != device value {float(loss_dev[0])}", file=sys.stderr)
    Kd = min(Ke, 500)
    for i in range(20):
        e2e_step_drain(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(Kd):
        e2e_step_drain(i)
    torch.cuda.synchronize()
    e2e_drain_s = time.perf_counter() - t0

    # ---------------- dominant kernel, measured in situ ----------------
    # td3_debug_prefix_times replays the first k launches of an update as a CUDA graph for k = 1..n and times each
    # with CUDA events on the replay stream: consecutive differences are what every launch adds to the chain.
    stage_info = None
    if rank == 0 and args.exec_mode == "graph":
        import ctypes as C
        from td3_b200 import _lib as L_
        view = agent._rb_view(rb)
        tot_us, stage_us, n_stage, n_all, enc_us, n_enc = 0.0, 0.0, 0, 0, 0.0, 0
        for with_actor in (0, 1):                      # one policy_freq = 2 cycle: a critic-only and a policy update
            us, kinds, n = (C.c_float * 128)(), (C.c_int32 * 128)(), C.c_int32()
            L_.check(lib.td3_debug_prefix_times(agent._handle, C.byref(view), with_actor, 200, us, kinds, 128, C.byref(n)))
            prev = 0.0
            for k in range(n.value):
                if kinds[k] == 0:
                    stage_us += us[k] - prev
                    n_stage += 1
                elif kinds[k] in (9, 11, 12):          # K6: fused set-encoder forward / backward launches (enc.cuh, encbwd.cuh)
                    enc_us += us[k] - prev
                    n_enc += 1
                prev = us[k]
            tot_us += us[n.value - 1]
            n_all += n.value
        stage_info = dict(stage_us=stage_us, cycle_us=tot_us, n_stage=n_stage, n_all=n_all, enc_us=enc_us, n_enc=n_enc)
        torch.cuda.synchronize()

    # ---------------- population: several independent agents per GPU in lock-step ----------------
    pop_ms, n_pop = 0.0, (args.population if w["kind"] == "featured" else 0)
    h2d_bytes = int(rb.row_floats * 4)
    if n_pop > 1:
        del agent, rb
        torch.cuda.empty_cache()
        # global agent ids: rank r owns agents [r * n_pop, (r + 1) * n_pop) of one population keyed 1001, so that the
        # members' random streams do not depend on how many GPUs the population is spread over (td3_b200/population.py)
        from td3_b200.population import shard_range, shard_seed
        first_agent, _ = shard_range(world * n_pop, world, rank)
        pop, prb = build_ours(w, seed=1000 + 64 * rank, rows=min(w["rows"], 100_000), n_agents=n_pop,
                              rng_seed=shard_seed(1001, first_agent))
        Kp = max(200, K // 4)                          # >= 200 timed lock-step updates whatever --steps says
        pop.train(prb, B, iterations=max(20, W // 4))
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        pop.train(prb, B, iterations=Kp)
        p1.record()
        barrier()
        pop_ms = p0.elapsed_time(p1)
        # HBM-bound kernel of the path: the optimiser launch over the population's packed critic buffers (Adam reads
        # p, g, m, v and writes p, m, v = 28 B/param; 8 agents x 0.25 M params x 5 buffers do not fit the L2), timed in
        # situ as the last launch of a critic-only update
        pop_adam = None
        if rank == 0 and args.exec_mode == "graph":
            import ctypes as C
            from td3_b200 import _lib as L_
            us, kinds, n = (C.c_float * 128)(), (C.c_int32 * 128)(), C.c_int32()
            pview = pop._rb_view(prb)
            L_.check(lib.td3_debug_prefix_times(pop._handle, C.byref(pview), 0, 300, us, kinds, 128, C.byref(n)))
            if n.value >= 2 and kinds[n.value - 1] == 3:
                t_us = us[n.value - 1] - us[n.value - 2]
                pc = sum(int(t.numel()) for t in pop.critic.parameters()) if n_pop == 1 else int(pop._critic_family.params.numel())
                nbytes = 28.0 * pc
                pop_adam = {"kernel": "td3::apply_kernel (first-layer dW tiles + Adam over the packed critic buffers of all agents)",
                            "bound": "hbm", "us_per_launch": t_us, "algorithmic_mb_per_launch": nbytes / 1e6,
                            "achieved": nbytes / (t_us * 1e-6) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
                            "frac": nbytes / (t_us * 1e-6) / 1e9 / peaks["hbm"],
                            "how": "td3_debug_prefix_times on the population agent: prefix(7 launches) - prefix(6 launches), "
                                   "CUDA events, 300 graph replays each (a replay is quantised to ~2 us by the GPU front end)"}
            torch.cuda.synchronize()

    if n_pop > 1:
        del pop, prb
    else:
        del agent, rb
    torch.cuda.empty_cache()

    # ---------------- the same workload in strict fp32 (FFMA tiles; what the parity suite runs by default) ----------------
    fp32_ms = fp32_e2e_s = 0.0
    Kf = min(K, 500)
    if args.precision != "fp32" and args.workload == "cfg2":
        os.environ["TD3_PRECISION"] = "fp32"
        fa, frb = build_ours(w, seed=100 + rank, rows=200_000)
        os.environ["TD3_PRECISION"] = args.precision
        fa.train(frb, B, iterations=max(W, 20))
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        fa.train(frb, B, iterations=Kf)
        f1.record()
        barrier()
        fp32_ms = f0.elapsed_time(f1)
        for i in range(20):
            frb.add(*new_rows[i % len(new_rows)]); fa.train(frb, B); fa.wait_critic_loss()
        barrier()
        t0 = time.perf_counter()
        for i in range(Kf):
            frb.add(*new_rows[i % len(new_rows)]); fa.train(frb, B); fa.wait_critic_loss()
        torch.cuda.synchronize()
        fp32_e2e_s = time.perf_counter() - t0
        barrier()
        del fa, frb
        torch.cuda.empty_cache()

    # ---------------- a wider observation (S = 32): S + A > 32 takes the first layers off the row-local front kernels ----------------
    wide = None
    if args.workload == "cfg2":
        ww = WORKLOADS["cfg2_S32"]
        wa, wrb = build_ours(ww, seed=300 + rank)
        wa.train(wrb, ww["B"], iterations=max(W, 20))
        barrier()
        l0 = lib.td3_launch_count()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        wa.train(wrb, ww["B"], iterations=Kf)
        w1.record()
        torch.cuda.synchronize()
        wide = {"workload": "cfg2 with S = 32 (SURVEY 8d's wider synthetic observation)", "ms_per_step": w0.elapsed_time(w1) / Kf,
                "value": Kf / (w0.elapsed_time(w1) / 1000.0), "unit": "updates/s (this rank)", "steps": Kf,
                "launches_per_update": (lib.td3_launch_count() - l0) / Kf}
        del wa, wrb
        torch.cuda.empty_cache()

    # ---------------- BASELINE config 4 (particles, N = 1024, batch 256): the tensor-bound workload, on rank 0 ----------------
    # the driver runs this file with the default workload only; the particle variant's number (fused set-encoder forward and
    # backward, csrc/enc.cuh + encbwd.cuh) rides along as a key.  `python bench.py --workload cfg4` is the full line.
    cfg4 = None
    if rank == 0 and args.workload == "cfg2" and args.precision == "tf32" and args.exec_mode == "graph":
        try:
            import ctypes as C
            from td3_b200 import _lib as L_
            w4 = WORKLOADS["cfg4"]
            a4, rb4 = build_ours(w4, seed=400)
            a4.train(rb4, w4["B"], iterations=10)
            torch.cuda.synchronize()
            n4 = max(20, min(K, 100))
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            a4.train(rb4, w4["B"], iterations=n4)
            c1.record()
            torch.cuda.synchronize()
            ms4 = c0.elapsed_time(c1) / n4
            view4 = a4._rb_view(rb4)
            enc_us, n_enc, cyc_us = 0.0, 0, 0.0
            for with_actor in (0, 1):
                us, kinds, n = (C.c_float * 128)(), (C.c_int32 * 128)(), C.c_int32()
                L_.check(lib.td3_debug_prefix_times(a4._handle, C.byref(view4), with_actor, 20, us, kinds, 128, C.byref(n)))
                prev = 0.0
                for k in range(n.value):
                    if kinds[k] in (9, 11, 12):
                        enc_us += us[k] - prev
                        n_enc += 1
                    prev = us[k]
                cyc_us += us[n.value - 1]
            gf4, mb4 = algorithmic_per_update(w4)
            cfg4 = {"workload": "cfg4: TD3_particles F=8 N=1024 D=6 A=3, batch 256, 8192-row buffer (1.6 GB of particle sets)",
                    "value": 1000.0 / ms4, "unit": "updates/s (this rank)", "ms_per_step": ms4, "steps": n4,
                    "algorithmic_gflop_per_update": gf4,
                    "encoder_kernels": {"us_per_cycle": enc_us, "launches_per_cycle": n_enc, "share_of_update_time": enc_us / max(cyc_us, 1e-9),
                                        "algorithmic_gflop_per_cycle": encoder_gflop_per_cycle(w4),
                                        "kernel": "td3::enc_fwd_kernel + enc_bwd_w2_kernel + enc_bwd_x_kernel (tcgen05 kind::tf32)"}}
            del a4, rb4
        except Exception as exc:                              # a side measurement, never a reason to lose the bench line
            cfg4 = {"unavailable": repr(exc)[:300]}
        torch.cuda.empty_cache()

    # ---------------- B = 1 latency of select_action / eval_q (the calls main.py makes every env step) ----------------
    b1 = None
    if rank == 0 and args.workload == "cfg2":
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import latency_b1
            b1 = latency_b1.measure("cfg2", n=1000, with_eager=not args.no_cpu_baseline)
            b1["what"] = ("host call -> host result of policy.select_action(state) / policy.eval_q(state, action) at batch 1: one kernel "
                          "per call reading the row from pinned host memory and writing the result back (csrc/infer.cuh); "
                          "torch_eager_gpu = the oracle port of the reference with its networks on this GPU")
        except Exception as exc:
            b1 = {"unavailable": repr(exc)[:300]}
        torch.cuda.empty_cache()

    # ---------------- BASELINE config 5b: data-parallel update at global batch 8192 over all ranks ----------------
    dp_line = None
    if args.workload == "cfg2" and (world > 1 or not args.no_dp):
        if world == 1 and not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29531")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", local_rank))
        try:
            dp_line = time_dp_critic(rank, world, max(50, min(K, 200)))
        except Exception as exc:
            dp_line = {"unavailable": repr(exc)[:300]}

    # ---------------- reduce over ranks ----------------
    if world > 1:
        t = torch.tensor([ms, e2e_s, pop_ms, e2e_drain_s, fp32_ms, fp32_e2e_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s, pop_ms, e2e_drain_s, fp32_ms, fp32_e2e_s = (float(x) for x in t)
    if dist.is_initialized():
        # the other ranks are done: they leave here instead of spinning in an NCCL barrier while rank 0 times the CPU arms
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    value = world * K / (ms / 1000.0)
    e2e = world * Ke / e2e_s
    t_update_us = ms * 1000.0 / K

    line = {"metric": "TD3 gradient updates/sec (batch 256)", "value": value, "unit": "updates/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": e2e, "unit": "updates/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 8,
                    "steps": Ke, "clocks": clocks_e2e,
                    "what": "per step: rb.add(host row -> pinned slot -> H2D copy) + policy.train(rb, 256) + "
                            "policy.wait_critic_loss(): the host blocks until THIS step's critic loss -- an 8-byte "
                            "{fp32 loss, update count} word the critic-head kernel stores to pinned host memory -- "
                            "has landed, then enqueues the next step behind the optimiser kernels still running; the "
                            "timed region ends with a full device synchronise",
                    "loss_mirror_matches_device": bool(mirror_ok), "loss_mirror_live": bool(agent_status_live),
                    "drain_value": world * Kd / e2e_drain_s, "drain_steps": Kd,
                    "drain_what": "same loop with a D2H copy of the loss behind the whole update and a stream "
                                  "synchronise every step (the GPU idles while the host prepares the next step)"},
            "gpu_launches": int(launches)}
    if fp32_ms > 0:
        line["fp32_mode"] = {"value": world * Kf / (fp32_ms / 1000.0), "e2e": world * Kf / fp32_e2e_s, "unit": "updates/s", "steps": Kf,
                             "ffma_peak_tflops": FFMA_PEAK_TFLOPS,
                             "frac_of_ffma_peak": gflop * 1e9 / (fp32_ms / Kf * 1e-3) / 1e12 / FFMA_PEAK_TFLOPS,
                             "what": "the same workload with precision='fp32': every contraction on fp32 FFMA tiles (the mode "
                                     "the parity suite compares with the CPU oracle at 2e-5); peak = 148 SMs x 128 lanes x 2 x 1.965 GHz"}
    if b1:
        line["b1_latency"] = b1
    if wide:
        line["wide_state"] = wide
    if cfg4:
        if "encoder_kernels" in cfg4:     # against the TF32 peak measured in this run
            ek = cfg4["encoder_kernels"]
            ek["achieved_tflops"] = ek["algorithmic_gflop_per_cycle"] * 1e9 / (ek["us_per_cycle"] * 1e-6) / 1e12
        line["particles_cfg4"] = cfg4
    if dp_line:
        line["dp_critic"] = dp_line
    if n_pop > 1 and pop_ms > 0:
        line["population"] = {"adam_kernel_roofline": pop_adam, "agents_per_gpu": n_pop, "value": world * n_pop * Kp / (pop_ms / 1000.0), "unit": "updates/s",
                              "ms_per_lockstep_update": pop_ms / Kp, "steps": Kp,
                              "what": f"{n_pop} independent agents per GPU (own weights, optimiser state, 100k-row replay buffer and "
                                      "Philox stream each) stepped by the same launches; value = agent-updates/s over all GPUs "
                                      "(BASELINE config 5: independent seeds sharded with no communication)"}

    if rank == 0:
        hbm_floor_us = mbytes * 1e6 / (peaks["hbm"] * 1e9) * 1e6
        tf32_measured = measure_tf32_peak()
        tf32_peak = tf32_measured if tf32_measured else peaks["bf16_sustained"] / 2.0
        tf32_src = ("measured in this run: cuBLAS TF32 matmul 8192^3, best of 6 (MEASURED_PEAKS.json bf16 sustained / 2 = "
                    f"{peaks['bf16_sustained'] / 2.0:.1f})") if tf32_measured else peaks["source"] + " (bf16 sustained / 2 for kind::tf32)"
        if args.precision == "fp32":
            tf32_peak, tf32_src = FFMA_PEAK_TFLOPS, "fp32 FFMA peak: 148 SMs x 128 lanes x 2 flop x 1.965 GHz"
        if "particles_cfg4" in line and "encoder_kernels" in line["particles_cfg4"]:
            ek = line["particles_cfg4"]["encoder_kernels"]
            ek.update({"bound": "tensor", "peak_tflops": tf32_peak, "frac": ek["achieved_tflops"] / tf32_peak, "peak_source": tf32_src})
            line["particles_cfg4"]["whole_update_frac_of_tf32_peak"] = (line["particles_cfg4"]["algorithmic_gflop_per_update"] * 1e9 /
                                                                        (line["particles_cfg4"]["ms_per_step"] * 1e-3) / 1e12 / tf32_peak)
        try:
            line["hbm_kernels"] = hbm_kernel_rooflines(peaks)
        except Exception as exc:
            line["hbm_kernels"] = {"unavailable": repr(exc)[:300]}
        tensor_floor_us = gflop * 1e9 / (tf32_peak * 1e12) * 1e6
        bound = "hbm" if hbm_floor_us >= tensor_floor_us else "tensor"
        if bound == "hbm":
            achieved, peak, unit = mbytes * 1e6 / (t_update_us * 1e-6) / 1e9, peaks["hbm"], "GB/s"
        else:
            achieved, peak, unit = gflop * 1e9 / (t_update_us * 1e-6) / 1e12, tf32_peak, "TFLOP/s"
        traffic = None
        try:       # dram__bytes_read+write per update from the committed ncu --set full capture of this workload, if any
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(f"{args.workload}:{args.precision}:{args.exec_mode}")
        except Exception:
            pass
        kernel = {"persistent": "td3::persistent_update_kernel (one cooperative launch = all K updates; per-update figures)",
                  "graph": "whole update = 7 (critic-only) / 14 (policy) dependent graph nodes + 1 / 2 head_finish nodes on a fork beside the backward stage (per-update figures)",
                  "launches": "td3::stage_kernel launches (per-update figures)"}[args.exec_mode]
        whole = {"bound": bound, "achieved": achieved, "peak": peak, "unit": unit, "frac": achieved / peak,
                 "us_per_update": t_update_us, "hbm_floor_us": hbm_floor_us, "tensor_floor_us": tensor_floor_us, "kernel": kernel}
        note = ("single-agent MLP updates are bound by the dependency chain of 7-14 launches (cold code, first-operand "
                "latency, a K loop that is bound by the ~50-cycle issue rate of tcgen05.mma from one thread plus barrier rounds), not by HBM or tensor throughput "
                "(SURVEY.md 8d, DESIGN.md 5); fractions are reported against the floors anyway")
        if stage_info and stage_info["n_enc"] > 0:
            # particles: the dominant kernels are the fused set-encoder launches.  Algorithmic flops = the encoder terms of
            # SURVEY 8d's formula (forward E per pass, backward Eb per pass with an encoder gradient); the h1 recomputation
            # of the backward kernels and the padded K = 8 layer 1 are NOT counted.
            eg = encoder_gflop_per_cycle(w)
            ach = eg * 1e9 / (stage_info["enc_us"] * 1e-6) / 1e12
            line["roofline"] = {
                "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak, "traffic": traffic,
                "kernel": "td3::enc_fwd_kernel + td3::enc_bwd_w2_kernel + td3::enc_bwd_x_kernel (K6: fused particle-set encoder, tcgen05 kind::tf32)",
                "launches_per_cycle": stage_info["n_enc"], "us_per_launch": stage_info["enc_us"] / stage_info["n_enc"],
                "share_of_update_time": stage_info["enc_us"] / stage_info["cycle_us"],
                "algorithmic_gflop_per_launch": eg / stage_info["n_enc"],
                "how": "CUDA events around graph replays of the first k launches of a critic-only and a policy update, "
                       "k = 1..n (td3_debug_prefix_times, on the replay stream); a launch's duration = prefix(k) - prefix(k-1)",
                "peak_source": tf32_src, "whole_update": whole, "note": "cfg4 is tensor-bound: whole_update carries the update-level fraction"}
        elif stage_info and stage_info["n_stage"] > 0:
            # dominant kernel: td3::stage_kernel<true>, the tcgen05 TF32 GEMM stages (hidden-layer contractions of all
            # passes).  Algorithmic flops = every contraction of the cycle that is not a first layer (front / apply
            # kernels) or an output head (head / front kernels).
            sg = stage_gflop_per_cycle(w)
            st_s = stage_info["stage_us"] * 1e-6
            ach = sg * 1e9 / st_s / 1e12
            line["roofline"] = {
                "bound": "tensor", "achieved": ach, "peak": tf32_peak, "unit": "TFLOP/s", "frac": ach / tf32_peak,
                "traffic": traffic,
                "kernel": "td3::stage_kernel<true> (tcgen05 kind::tf32 GEMM stages)" if args.precision == "tf32"
                          else "td3::stage_kernel<false> (fp32 FFMA GEMM stages)",
                "launches_per_cycle": stage_info["n_stage"], "us_per_launch": stage_info["stage_us"] / stage_info["n_stage"],
                "share_of_update_time": stage_info["stage_us"] / stage_info["cycle_us"],
                "algorithmic_gflop_per_launch": sg / stage_info["n_stage"],
                "how": "CUDA events around graph replays of the first k launches of a critic-only and a policy update, "
                       "k = 1..n (td3_debug_prefix_times, on the replay stream); a launch's duration = prefix(k) - prefix(k-1)",
                "peak_source": tf32_src, "whole_update": whole, "note": note}
        else:
            whole.update({"traffic": traffic, "per": "update", "peak_source": peaks["source"], "note": note})
            line["roofline"] = whole
        if not args.no_cpu_baseline:
            try:
                n_e = 300 if w["kind"] == "featured" else 20
                line["torch_eager_gpu"] = {
                    "value": time_torch_gpu(w, n_e, 30 if w["kind"] == "featured" else 2, min(w["rows"], 100_000)),
                    "unit": "updates/s", "steps": n_e,
                    "what": "oracle port of the reference with its networks on this GPU (eager PyTorch, fp32, host replay "
                            "arrays + per-sample H2D copies as my_replay_buffer.py does): what the reference gets from a B200 "
                            "unchanged"}
            except Exception as exc:                          # a comparator, never a reason to lose the bench line
                line["torch_eager_gpu"] = {"unavailable": repr(exc)[:200]}
            steps_cpu, warm_cpu = cpu_sample_size(w)
            ups, th, dt = time_cpu(w, steps_cpu, warm_cpu, threads_list, min(w["rows"], 100_000))
            line["cpu_baseline"] = {"value": ups, "unit": "updates/s", "cores": th, "kind": "port",
                                    "sample": f"{steps_cpu} updates after {warm_cpu} warm-up of the same workload, oracle "
                                              f"port of the reference on torch CPU, best of threads {threads_list}, "
                                              f"os.cpu_count()={ncpu}"}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
