"""In-situ cost of every launch of one update (td3_debug_prefix_times): the update is captured as graphs of its first
k launches, k = 1..n, each timed with CUDA events over back-to-back replays; consecutive differences are what each
launch adds to the dependency chain (launch boundary + cold code/parameter fetch + work), with no profiler attached.

    python tools/prefix_times.py [cfg2] [reps] [agents_per_gpu]
"""
import ctypes as C
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from td3_b200 import _lib

KINDS = {0: "stage", 1: "gather", 2: "loss", 3: "adam/apply", 4: "tick", 5: "head", 6: "wn", 7: "front", 8: "dpsync", 9: "enc", 10: "chain", 11: "encbwd_w2", 12: "encbwd_x"}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    w = bench.WORKLOADS[name]
    n_agents = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    agent, rb = bench.build_ours(w, seed=100, rows=min(w["rows"], 100_000) if n_agents > 1 else None, n_agents=n_agents)
    agent.train(rb, w["B"], iterations=20)
    torch.cuda.synchronize()
    view = agent._rb_view(rb)
    for with_actor in (0, 1):
        us = (C.c_float * 64)()
        kinds = (C.c_int32 * 64)()
        n = C.c_int32()
        _lib.check(agent._lib.td3_debug_prefix_times(agent._handle, C.byref(view), with_actor, reps, us, kinds, 64, C.byref(n)))
        print(f"--- {name} {'policy' if with_actor else 'critic-only'} update: {n.value} launches, {us[n.value - 1]:.1f} us")
        prev = 0.0
        for k in range(n.value):
            print(f"  {k + 1:2d} {KINDS.get(kinds[k], '?'):10s} +{us[k] - prev:6.2f} us   (prefix {us[k]:6.1f})")
            prev = us[k]


if __name__ == "__main__":
    main()
