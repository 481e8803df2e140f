"""Host-side logic of the data-parallel critic (td3_b200/data_parallel.py) on CPU with the gloo backend, world size 2:
the shard arithmetic, and the identity the scheme rests on -- per-shard gradients of a loss normalised by 1/B_global,
summed by one all-reduce, equal the full-batch gradient (checked with the CPU oracle's networks; no GPU involved)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import td3_oracle as O
from td3_b200.data_parallel import shard_bounds


def test_shard_bounds_partition_the_global_batch():
    for B, W in ((8192, 8), (256, 2), (100, 4)):
        spans = [shard_bounds(B, W, r) for r in range(W)]
        assert spans[0][0] == 0 and spans[-1][1] == B
        assert all(spans[i][1] == spans[i + 1][0] for i in range(W - 1))
    with pytest.raises(ValueError):
        shard_bounds(100, 3, 0)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    S, A, B = 17, 6, 64
    torch.manual_seed(0)
    agent = O.TD3Featured(O.Space(S), O.Space(A), actor_widths=(64, 48), q_widths=(64, 48), lr=1e-3)
    data = O.synthetic_transitions_featured(256, S, A, seed=0)
    idx = np.random.RandomState(3).randint(0, 256, size=B)
    lo, hi = shard_bounds(B, world, rank)

    def critic_grad(rows, norm):
        s = torch.tensor(data["state"][rows], dtype=torch.float32)
        a = torch.tensor(data["action"][rows], dtype=torch.float32)
        y = torch.tensor(data["reward"][rows], dtype=torch.float32).reshape(-1, 1)
        agent.critic.zero_grad()
        q1, q2 = agent.critic(s, a)
        (((q1 - y) ** 2).sum() / norm + ((q2 - y) ** 2).sum() / norm).backward()
        return torch.cat([p.grad.reshape(-1) for p in agent.critic.parameters()])

    full = critic_grad(idx, B)                       # what one device computes on the whole batch (mean over B)
    part = critic_grad(idx[lo:hi], B)                # this rank's shard, normalised by the GLOBAL batch
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((float((part - full).abs().max()), float(full.abs().max())))
    dist.destroy_process_group()


def test_sharded_gradients_sum_to_the_full_batch_gradient_gloo_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    err, scale = out.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert err <= 1e-6 * max(1.0, scale), (err, scale)
