"""Accuracy and speed of the split-fp16 tcgen05 GEMM (b200mnn_dev_debug_gemm); usage: time_gemm.py M N K"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batchelor_b200 import _lib

M, N, K = (int(x) for x in sys.argv[1:4])
cuda = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(3)
A = torch.randn((M, K), dtype=torch.float64, device=cuda, generator=g); A /= A.norm(dim=1, keepdim=True)
B = torch.randn((N, K), dtype=torch.float64, device=cuda, generator=g); B /= B.norm(dim=1, keepdim=True)
B[: N // 2] = 0.7 * B[: N // 2] + 0.7 * A[0]      # correlated rows: dot products ~0.7, same sign (biased accumulation shows)
ldo = (N + 255) // 256 * 256
out = torch.zeros((M, ldo), dtype=torch.float32, device=cuda)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
ms = min(M, 2048); ns = min(N, 4096)
ref = A[:ms] @ B[:ns].T
for terms in (3, 1):
    for cb in (0, 32, 8, 4, 2):
        _lib.call("b200mnn_dev_debug_gemm", C.c_void_p(A.data_ptr()), M, C.c_void_p(B.data_ptr()), N, K, terms, cb, C.c_void_p(out.data_ptr()), ldo, s)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        _lib.call("b200mnn_dev_debug_gemm", C.c_void_p(A.data_ptr()), M, C.c_void_p(B.data_ptr()), N, K, terms, cb, C.c_void_p(out.data_ptr()), ldo, s)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        err = (out[:ms, :ns].double() - ref)
        print(f"terms {terms} chunk_boxes {cb:3d}: {dt*1e3:8.2f} ms incl. operand prep  {2*M*N*K/dt/1e12:7.1f} algorithmic TFLOP/s   "
              f"max abs err {err.abs().max().item():.3e}  mean err (bias) {err.mean().item():+.3e}  rms {err.pow(2).mean().sqrt().item():.3e}")
