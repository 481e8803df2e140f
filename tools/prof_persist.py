"""Per-stage clock64 stamps of the persistent kernel's last iteration (TD3_PERSIST_PROF=1): work vs barrier wait."""
import os, sys
os.environ["TD3_PERSIST_PROF"] = "1"
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
agent, rb = bench.build_ours(w, 100)
for iters in (2, 1):   # last iteration = policy step (total_it even) then critic-only
    agent.train(rb, w["B"], iterations=200 + iters)
    torch.cuda.synchronize()
    prof = agent._region("prof").view(torch.int64).cpu().numpy().reshape(2, 128, 3)
    print(f"--- total_it={agent.total_it} last iteration, cycles (CTA0 | last CTA): work, barrier")
    for s in range(128):
        if prof[0, s, 0] == 0: break
        a, b = prof[0, s], prof[1, s]
        print(f"stage {s:2d}: work {a[1]-a[0]:7d} bar {a[2]-a[1]:7d} | work {b[1]-b[0]:7d} bar {b[2]-b[1]:7d}")
    tot = prof[0, :s, 2].max() - prof[0, 0, 0]
    print("iteration cycles", tot, "=", tot / 1.965e3, "us")
    if os.environ.get("TD3_LIB_NAME"):
        import ctypes as C
        buf = (C.c_longlong * (128 * 16))()
        agent._lib.td3_debug_tile_prof(buf)
        tp = np.array(buf).reshape(128, 2, 8)
        for s2 in range(s):
            p0, p1 = tp[s2, 0], tp[s2, 1]
            if p0[0]:
                t0 = p0[0]
                print(f"  TC tile stage {s2:2d} (cycles from tile start) producer: setup {p0[1]-t0} last-TMA-issued {p0[2]-t0} | "
                      f"mma: first-data {p1[1]-t0} last-MMA-issued {p1[2]-t0} | all: epilogue-start {p0[3]-t0} accum-done {p0[4]-t0} "
                      f"epilogue-end {p0[5]-t0} tile-end {p0[6]-t0}")
    agent._region("prof").zero_()
