"""Stamps of block 0's tcgen05 tile for one td3_gemm call (debug build: TD3_LIB_NAME=libtd3b200_prof.so)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from td3_b200 import _lib
lib = _lib.require_cuda()
shapes = [(256, 304, 400, 1, 1), (256, 304, 400, 0, 0), (256, 16, 400, 1, 1), (256, 304, 1600, 1, 1)]
for M, N, K, a_rc, b_rc in shapes:
    A = torch.randn((M, K) if a_rc else (K, M), device="cuda"); B = torch.randn((N, K) if b_rc else (K, N), device="cuda"); Cm = torch.empty(M, N, device="cuda")
    for rep in range(3):
        _lib.check(lib.td3_gemm(M, N, K, A.data_ptr(), A.stride(0), a_rc, B.data_ptr(), B.stride(0), b_rc, Cm.data_ptr(), N, None, 0, 1, _lib.stream_ptr()))
    buf = (C.c_longlong * (128 * 16))()
    lib.td3_debug_tile_prof(buf)
    tp = np.array(buf).reshape(128, 2, 8)
    p0, p1 = tp[0, 0], tp[0, 1]
    t0 = p0[0]
    n_chunks = (K + 31) // 32
    print(f"[{M}x{N}x{K} a_rc={a_rc} b_rc={b_rc}] cluster={'off' if os.environ.get('TD3_NO_CLUSTER') else 'on'} chunks={n_chunks}: setup {p0[1]-t0} last-TMA {p0[2]-t0} | first-data {p1[1]-t0} "
          f"last-MMA {p1[2]-t0} ({(p1[2]-p1[1])/max(1,n_chunks-1):.0f} cyc/chunk) | accum-done {p0[4]-t0} epi-end {p0[5]-t0}")
