"""Roofline of the two GEMM tiles on their own (td3_gemm in the C ABI): TFLOP/s against the measured tensor peak
(TF32 = half of MEASURED_PEAKS.json's bf16 figure) for the shapes the TD3 update contains.
    python tools/bench_gemm.py [--reps 20]
One JSON line per shape: {"shape": [M,N,K], "orient": "fwd|dx|dw", "tile": "tcgen05_tf32|ffma_fp32", "us": t, "tflops": x, "frac_of_tf32_peak": f}"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from td3_b200 import _lib

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--only-tc", action="store_true")
ap.add_argument("--shapes", default="", help="comma-separated indices into SHAPES (default: all)")
args = ap.parse_args()
lib = _lib.require_cuda()
try:
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    tf32_peak = peaks["bf16_tflops"] / 2
except Exception:
    tf32_peak = 1590.0 / 2
SHAPES = [  # (M, N, K, a_rc, b_rc, what)
    (256, 300, 400, 1, 1, "cfg2 layer-2 forward"),
    (256, 400, 300, 1, 0, "cfg2 layer-2 dX"),
    (300, 400, 256, 0, 0, "cfg2 layer-2 dW"),
    (8192, 300, 400, 1, 1, "batch-8192 critic layer-2 forward"),
    (300, 400, 8192, 0, 0, "batch-8192 critic layer-2 dW"),
    (262144, 128, 256, 1, 1, "particle encoder layer-2 forward (B*N = 256*1024 rows)"),
    (262144, 256, 128, 1, 0, "particle encoder layer-2 dX"),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > L2 (126 MB): cold operands between repetitions
sel = [int(x) for x in args.shapes.split(",")] if args.shapes else range(len(SHAPES))
for M, N, K, a_rc, b_rc, what in [SHAPES[i] for i in sel]:
    A = torch.randn((M, K) if a_rc else (K, M), device="cuda")
    B = torch.randn((N, K) if b_rc else (K, N), device="cuda")
    Cm = torch.empty(M, N, device="cuda")
    for use_tc in ((1,) if args.only_tc else (1, 0)):
        def run():
            _lib.check(lib.td3_gemm(M, N, K, A.data_ptr(), A.stride(0), a_rc, B.data_ptr(), B.stride(0), b_rc, Cm.data_ptr(), N, None,
                                    0, use_tc, _lib.stream_ptr()))
        for _ in range(3):
            run()
        ts = []
        for _ in range(args.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = sorted(ts)[len(ts) // 2]
        tf = 2.0 * M * N * K / (us * 1e-6) / 1e12
        print(json.dumps({"shape": [M, N, K], "what": what, "orient": {(1, 1): "fwd", (1, 0): "dx", (0, 0): "dw", (0, 1): "other"}[(a_rc, b_rc)],
                          "tile": "tcgen05_tf32" if use_tc else "ffma_fp32", "us": round(us, 2), "tflops": round(tf, 2),
                          "frac_of_tf32_peak": round(tf / tf32_peak, 4), "l2": "flushed between repetitions (256 MB memset)"}))
