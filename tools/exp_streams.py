"""Experiment: N independent cfg2 agents on one GPU as G lock-step populations of N / G agents, one CUDA stream per
population, against one lock-step population of N (the launches of different streams overlap on the idle SMs).

    python tools/exp_streams.py [n_agents_total] [timed_updates]
"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def run(total, groups, K, chunk=10):
    w = bench.WORKLOADS["cfg2"]
    per = total // groups
    pops = []
    for g in range(groups):
        a, rb = bench.build_ours(w, seed=1000 + g, rows=100_000, n_agents=per)
        pops.append((a, rb, torch.cuda.Stream()))
    torch.cuda.synchronize()
    for a, rb, s in pops:
        with torch.cuda.stream(s):
            a.train(rb, w["B"], iterations=20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for a, rb, s in pops:
        s.wait_event(e0)
    for c in range(K // chunk):
        for a, rb, s in pops:
            with torch.cuda.stream(s):
                a.train(rb, w["B"], iterations=chunk)
    for a, rb, s in pops:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    out = {"agents": total, "streams": groups, "agents_per_stream": per, "agent_updates_per_s": total * (K // chunk) * chunk / (ms * 1e-3),
           "us_per_update_of_all": ms * 1e3 / ((K // chunk) * chunk)}
    print(json.dumps(out), flush=True)
    del pops
    torch.cuda.empty_cache()


def main():
    total = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 400
    g = 1
    while g <= total:
        run(total, g, K)
        g *= 2


if __name__ == "__main__":
    main()
