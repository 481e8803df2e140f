"""Data-parallel critic on real GPUs (run under torchrun, one rank per GPU):
   1. parity: W ranks x (256/W) rows with gradient all-reduce == one device on the same 256-row global batch;
   2. timing: 400-300 networks at global batch 8192 (BASELINE config 5b), updates/s, CUDA-event timed, max over ranks.
     python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_dp.py"""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from td3_b200 import synthetic as O
from td3_b200.TD3_featured import TD3
from td3_b200.my_replay_buffer import ReplayBuffer_featured
from td3_b200.data_parallel import DataParallelTD3

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
S, A, ROWS = 17, 6, 65536
obs, act = O.Space(S), O.Space(A)
data = O.transitions_featured(ROWS, S, A, seed=0)


def make(precision):
    torch.manual_seed(0)                       # identical initial weights on every rank
    a = TD3(obs, act, actor_widths=(400, 300), q_widths=(400, 300), lr=1e-3, seed=11, precision=precision)
    rb = ReplayBuffer_featured(obs, act, max_size=ROWS)
    rb.add_batch(**data)
    return a, rb

# ---- 1. parity (fp32 tiles: differences are summation order only), both ways of forming the gradient sum ----
ref_state = None
for mode in ("p2p", "nccl"):
    a, rb = make("fp32")
    dp = DataParallelTD3(a, mode=mode)
    for _ in range(6):
        dp.train(rb, 256)
    loss = float(dp.global_critic_loss()[0].item())
    # replicas must hold bit-identical parameters after the updates (identical sums on every rank)
    flat = torch.cat([a._critic_family.params, a._actor_family.params])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    replicas_identical = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        if ref_state is None:
            ref, rrb = make("fp32")
            ref.exec_mode = "launches"
            for _ in range(6):
                ref.train(rrb, 256)
            ref_state = ({k: [w.clone() for w in getattr(ref, k).state_dict().values()] for k in ("actor", "critic", "actor_target", "critic_target")},
                         float(ref.last_critic_loss[0].item()))
        worst = 0.0
        for k in ("actor", "critic", "actor_target", "critic_target"):
            for (n, v), w in zip(getattr(a, k).state_dict().items(), ref_state[0][k]):
                worst = max(worst, float((v - w).norm() / max(float(w.norm()), 1e-12)))
        want = ref_state[1]
        ok = worst <= 1e-4 and abs(loss - want) <= 1e-4 * max(1.0, abs(want)) and replicas_identical
        print(json.dumps({"check": "dp_parity", "mode": dp.mode, "requested_mode": mode, "world": world, "global_batch": 256, "updates": 6,
                          "worst_param_rel_l2": worst, "critic_loss_dp": loss, "critic_loss_single": want,
                          "replicas_bit_identical": replicas_identical, "graph_replay": dp.use_graph,
                          "p2p_unavailable": getattr(dp, "p2p_unavailable", None), "ok": ok}))
        assert ok
    dist.barrier()
    del dp, a, rb

# ---- 2. timing at global batch 8192 ----
for mode in ("p2p", "nccl"):
    a, rb = make("tf32")
    dp = DataParallelTD3(a, mode=mode)
    for _ in range(20):
        dp.train(rb, 8192)
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    K = 200
    e0.record()
    for _ in range(K):
        dp.train(rb, 8192)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "dp_timing", "mode": dp.mode, "workload": "cfg5b: 400-300 networks, global batch 8192, policy_freq 2",
                          "world": world, "updates_per_s": K / (float(t[0]) / 1e3), "ms_per_update": float(t[0]) / K, "scaling": "strong",
                          "reductions_per_update": "1 x critic gradient (1.04 MB) + 0.5 x actor gradient (0.52 MB)"}))
    dist.barrier()
    del dp, a, rb
if os.environ.get("CHECK_DP_TRACE"):
    print(f"[rank {rank}] all checks done, tearing down", file=sys.stderr, flush=True)
# Teardown.  The agents and their driver reference each other (agent._dp_owner <-> dp.agent), so `del` alone leaves the
# symmetric-memory gradient buffers alive until the cycle collector runs; with them alive destroy_process_group() has
# been seen to wait forever on a 2-GPU box.  Collect first, and never let a stuck teardown turn a finished check into
# a hung process: every result line has been printed by now.
import gc, threading
gc.collect()
torch.cuda.synchronize()
if os.environ.get("CHECK_DP_TRACE"):
    print(f"[rank {rank}] device idle", file=sys.stderr, flush=True)
dist.barrier()
th = threading.Thread(target=dist.destroy_process_group, daemon=True)
th.start()
th.join(30.0)
if os.environ.get("CHECK_DP_TRACE"):
    print(f"[rank {rank}] process group {'destroyed' if not th.is_alive() else 'teardown still waiting after 30 s: leaving'}",
          file=sys.stderr, flush=True)
sys.stdout.flush()
sys.stderr.flush()
os._exit(0)
