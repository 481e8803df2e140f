"""Latency of the B=1 API the reference's env loop calls every step (select_action, eval_q: TD3_featured.py:113-121),
host call to host result, p50 / p99 in microseconds; beside it the oracle port of the reference with its networks on the
same GPU (eager PyTorch: FloatTensor -> .to(device) -> forward -> .cpu().numpy(), as TD3_featured.py:113-121 does).
    python tools/latency_b1.py [cfg2|cfg3_layer|cfg4]"""
import os, sys, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def percentiles(fn, n=2000, warm=200):
    for _ in range(warm):
        fn()
    ts = np.empty(n)
    for i in range(n):
        t0 = time.perf_counter()
        fn()
        ts[i] = time.perf_counter() - t0
    return {"p50_us": float(np.percentile(ts, 50) * 1e6), "p99_us": float(np.percentile(ts, 99) * 1e6), "mean_us": float(ts.mean() * 1e6), "calls": n}


def measure(name="cfg2", n=2000, with_eager=True):
    w = bench.WORKLOADS[name]
    agent, rb = bench.build_ours(w, 100, rows=4096)
    agent.train(rb, w["B"], iterations=4)
    torch.cuda.synchronize()
    rs = np.random.RandomState(0)
    if w["kind"] == "featured":
        s = rs.standard_normal(w["S"])
    else:
        s = (rs.standard_normal(w["F"]), rs.standard_normal((w["N"], w["D"])).astype(np.float32))
        n = min(n, 300)
    a = agent.select_action(s)
    out = {"workload": name, "select_action": percentiles(lambda: agent.select_action(s), n),
           "eval_q": percentiles(lambda: agent.eval_q(s, a), n)}
    if with_eager:
        ora, _ = bench.build_oracle(w, 0, 256)
        for k in ("actor", "critic"):
            getattr(ora, k).cuda()
        if w["kind"] == "featured":
            def sel():
                st = torch.FloatTensor(np.asarray(s).reshape(1, -1)).to("cuda")
                return ora.actor(st).cpu().data.numpy().flatten()

            def evq():
                st = torch.FloatTensor(np.asarray(s).reshape(1, -1)).to("cuda")
                u = torch.FloatTensor(np.asarray(a).reshape(1, -1)).to("cuda")
                return [q.cpu().data.numpy().flatten() for q in ora.critic(st, u)]
            out["torch_eager_gpu"] = {"select_action": percentiles(sel, min(n, 1000)), "eval_q": percentiles(evq, min(n, 1000))}
    return out


if __name__ == "__main__":
    print(json.dumps(measure(sys.argv[1] if len(sys.argv) > 1 else "cfg2")))
