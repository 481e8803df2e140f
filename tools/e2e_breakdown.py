"""Host-side timeline of the end-to-end step bench.py times (add -> train -> loss D2H -> sync), cfg2.
Prints mean microseconds per segment; the `sync` segment is the GPU's remaining work plus the wake-up latency."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

def main():
    w = bench.WORKLOADS["cfg2"] if hasattr(bench, "WORKLOADS") else None
    agent, rb = bench.build_ours(w, seed=100)
    B = w["B"]
    rs = np.random.RandomState(0)
    rows = [(rs.standard_normal(w["S"]), rs.uniform(-1, 1, w["A"]), rs.standard_normal(w["S"]), float(rs.standard_normal()), 0.0)
            for _ in range(64)]
    loss_host = torch.zeros(1).pin_memory()
    loss_dev = agent.last_critic_loss
    stream = torch.cuda.current_stream()
    agent.train(rb, B, iterations=50)
    torch.cuda.synchronize()
    n = 3000
    seg = np.zeros(4)
    for i in range(n + 200):
        t0 = time.perf_counter()
        rb.add(*rows[i % 64])
        t1 = time.perf_counter()
        agent.train(rb, B)
        t2 = time.perf_counter()
        loss_host.copy_(loss_dev, non_blocking=True)
        t3 = time.perf_counter()
        stream.synchronize()
        t4 = time.perf_counter()
        if i >= 200:
            seg += (t1 - t0, t2 - t1, t3 - t2, t4 - t3)
    seg *= 1e6 / n
    print("add %.1f us | train call %.1f us | loss copy issue %.1f us | sync %.1f us | total %.1f us" % (*seg, seg.sum()))
    # device time per update for comparison
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); agent.train(rb, B, iterations=2000); ev1.record(); torch.cuda.synchronize()
    print("device-resident: %.1f us/update" % (ev0.elapsed_time(ev1) * 1000 / 2000))

if __name__ == "__main__":
    main()
