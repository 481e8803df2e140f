"""A few lock-step updates of an n-agent population as plain launches (profiler target):
     ncu --set full --clock-control none -k regex:front_kernel --launch-skip 12 -c 3 python tools/prof_pop.py [n_agents] [workload]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
w = bench.WORKLOADS[sys.argv[2] if len(sys.argv) > 2 else "cfg2"]
agent, rb = bench.build_ours(w, seed=100, rows=min(w["rows"], 100_000) if n > 1 else None, n_agents=n)
agent.exec_mode = "launches"
agent.train(rb, w["B"], iterations=int(sys.argv[3]) if len(sys.argv) > 3 else 6)
torch.cuda.synchronize()
print("ok", float(agent.last_critic_loss[0]))
