"""Read out how the tcgen05 tile interprets an MN-major operand: multiply by an identity and print what comes back."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from td3_b200 import _lib
lib = _lib.require_cuda()
np.set_printoptions(linewidth=250, threshold=100000)

def gemm(A, a_rc, B, b_rc, M, N, K, bias=None, relu=0):
    Cm = torch.full((M, N), float("nan"), device="cuda")
    _lib.check(lib.td3_gemm(M, N, K, A.data_ptr(), A.stride(0), a_rc, B.data_ptr(), B.stride(0), b_rc, Cm.data_ptr(), Cm.stride(0),
                            bias.data_ptr() if bias is not None else None, relu, 1, _lib.stream_ptr()))
    torch.cuda.synchronize()
    return Cm.cpu().numpy()

M, N, K = 128, 32, 64
A = torch.zeros(M, K); A[:K, :K] = torch.eye(K)
kk = torch.arange(K, dtype=torch.float32)[:, None].expand(K, N).contiguous()
nn = torch.arange(N, dtype=torch.float32)[None, :].expand(K, N).contiguous()
for name, Bm in (("B=k", kk), ("B=n", nn)):
    got = gemm(A.cuda(), 1, Bm.cuda(), 0, M, N, K)
    print(f"--- B MN-major [{K}x{N}] {name}: C[i,j] should be {name[2:]} (i=k<64)")
    print(got[:K].astype(int))
# A MN-major: A stored [K, M]; B = identity K-major [N, K]
Bi = torch.zeros(N, K); Bi[:, :N] = torch.eye(N)
ka = torch.arange(K, dtype=torch.float32)[:, None].expand(K, M).contiguous()
ma = torch.arange(M, dtype=torch.float32)[None, :].expand(K, M).contiguous()
for name, Am in (("A=k", ka), ("A=m", ma)):
    got = gemm(Am.cuda(), 0, Bi.cuda(), 1, M, N, K)
    print(f"--- A MN-major [{K}x{M}] {name}: C[i,j] should be {'j' if name=='A=k' else 'i'}")
    print(got.astype(int)[::4])
# epilogue check on the K-major path
g = torch.Generator().manual_seed(1)
A = torch.randn(256, 400, generator=g); B = torch.randn(300, 400, generator=g); bias = torch.randn(300, generator=g)
got = gemm(A.cuda(), 1, B.cuda(), 1, 256, 300, 400, bias.cuda(), 1)
want = torch.relu(A.double() @ B.double().t() + bias.double()).numpy()
print("K-major bias+relu max err", np.abs(got - want).max(), "nan count", np.isnan(got).sum())
