"""Times the trailing-update kernel alone: C[MxN] -= A[MxK] B[KxN] on device buffers (for ncu and tuning)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import scalapack_b200 as S

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
K = int(sys.argv[3]) if len(sys.argv) > 3 else 512
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
cplx = int(sys.argv[5]) if len(sys.argv) > 5 else 0
es = 2 if cplx else 1
A = torch.rand(K * M * es, dtype=torch.float64, device="cuda") - 0.5
B = torch.rand(N * K * es, dtype=torch.float64, device="cuda") - 0.5
Cm = torch.rand(N * M * es, dtype=torch.float64, device="cuda")
I64 = C.c_int64
L = S.lib()
L.slb200_test_gemm(I64(M), I64(N), K, S.api._ptr(A), I64(M), S.api._ptr(B), I64(K), S.api._ptr(Cm), I64(M), cplx, 1)
ms = L.slb200_test_gemm(I64(M), I64(N), K, S.api._ptr(A), I64(M), S.api._ptr(B), I64(K), S.api._ptr(Cm), I64(M), cplx, reps)
fl = 2.0 * M * N * K * (4 if cplx else 1)
print(f"gemm M={M} N={N} K={K} cplx={cplx}: {ms:.3f} ms/launch  {fl / ms / 1e9:.2f} TFLOP/s")
