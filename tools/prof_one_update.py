"""One update of a workload between cudaProfilerStart / Stop, for `ncu --profile-from-start off` captures of every kernel
of a single critic-only (or policy) update:

    ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:stage_kernel \
        -o gpurun_out/x python tools/prof_one_update.py cfg4 [policy] [agents_per_gpu]
"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    policy = len(sys.argv) > 2 and sys.argv[2] == "policy"
    n_agents = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    w = bench.WORKLOADS[name]
    agent, rb = bench.build_ours(w, seed=100, rows=min(w["rows"], 100_000), n_agents=n_agents)
    agent.exec_mode = "launches"                      # plain launches: every kernel is visible to the profiler by name
    agent.train(rb, w["B"], iterations=4 if not policy else 5)     # policy_freq = 2: the next update is critic-only / policy
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    agent.train(rb, w["B"], iterations=1)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()
