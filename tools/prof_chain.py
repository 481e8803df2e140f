"""Clock stamps of the layer-fused chain kernels (chain.cuh, TD3_CHAIN_PROF=1): per role, the cycles at which every
step's accumulator was complete and its epilogue done (tile 0 of agent 0), plus wall-clock (globaltimer) start / end.

    python tools/prof_chain.py [cfg2] [agents_per_gpu]
"""
import os, sys
os.environ["TD3_CHAIN_PROF"] = "1"
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def dump(prof, base, names):
    t0 = None
    for i, name in enumerate(names):
        p = prof[(base + i) * 48:(base + i + 1) * 48]
        if p[1] == 0:
            continue
        if t0 is None:
            t0 = p[0]
        line = f"{name:8s} start +{(p[0] - t0) / 1e3:7.2f} us  end +{(p[47] - t0) / 1e3:7.2f} us | cycles since start:"
        for s in range(22):
            a, b = p[2 + 2 * s], p[3 + 2 * s]
            if a == 0:
                break
            line += f"  [{s}] acc {a - p[1]:6d} epi {b - p[1]:6d}"
        print(line)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    n_agents = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    w = bench.WORKLOADS[name]
    agent, rb = bench.build_ours(w, seed=100, rows=min(w["rows"], 100_000), n_agents=n_agents)
    agent.train(rb, w["B"], iterations=20)
    torch.cuda.synchronize()
    for _ in range(2):
        agent._region("prof").zero_()
        torch.cuda.synchronize()
        agent.train(rb, w["B"], iterations=1)
        torch.cuda.synchronize()
        prof = agent._region("prof").view(torch.int64).cpu().numpy()
        print(f"--- update {agent.total_it}")
        dump(prof, 0, ["target", "actor", "critic"])
        dump(prof, 4, ["policy"])


if __name__ == "__main__":
    main()
