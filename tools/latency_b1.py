"""Latency of the B=1 API the reference's env loop calls every step (select_action, eval_q: TD3_featured.py:113-121)."""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
agent, rb = bench.build_ours(w, 100, rows=4096)
agent.train(rb, w["B"], iterations=4)
if w["kind"] == "featured":
    s = np.random.RandomState(0).standard_normal(w["S"])
else:
    rs = np.random.RandomState(0)
    s = (rs.standard_normal(w["F"]), rs.standard_normal((w["N"], w["D"])).astype(np.float32))
for _ in range(50):
    a = agent.select_action(s)
    q = agent.eval_q(s, a)
torch.cuda.synchronize()
n = 500
t0 = time.perf_counter()
for _ in range(n):
    a = agent.select_action(s)
t1 = time.perf_counter()
for _ in range(n):
    q = agent.eval_q(s, a)
t2 = time.perf_counter()
print(json.dumps({"workload": sys.argv[1] if len(sys.argv) > 1 else "cfg2", "select_action_us": (t1 - t0) / n * 1e6, "eval_q_us": (t2 - t1) / n * 1e6}))
