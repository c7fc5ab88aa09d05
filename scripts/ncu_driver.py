"""One kernel family of the LU path at BASELINE config 2's width, alone, for `ncu` (scripts/gpu/r2_d.sh).

  python scripts/ncu_driver.py laswp  [m n jb]     row interchanges of one block: swap_plan / swap_pack / swap_unpack_out / copy2d
  python scripts/ncu_driver.py panel  [m jb]       panel factorisation (leaf kernels + recursion) of an m x jb panel
  python scripts/ncu_driver.py solve  [n nb]       PDGETRF then PDGETRS 'N' on a 1x1 grid (diag_solve / gemv_rows kernels)
"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import scalapack_b200 as S

mode = sys.argv[1]
arg = [int(x) for x in sys.argv[2:]]
L = S.lib()
I64 = C.c_int64
ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 1, 1)
if mode == "laswp":
    m, n, jb = (arg + [65536, 32768, 512])[:3] if len(arg) >= 3 else (65536, 32768, 512)
    A = torch.empty(n * m, dtype=torch.float64, device="cuda")
    S.matgen64(ctx, m, n, 512, 512, A, m, 7)
    rng = np.random.default_rng(1)
    piv = np.array([rng.integers(t + 1, m + 1) for t in range(jb)], dtype=np.int32)      # 1-based rows >= own row (j0 = 0)
    for _ in range(2):
        L.slb200_test_laswp(m, I64(n), S.api._ptr(A), I64(m), 0, jb, piv.ctypes.data_as(C.c_void_p))
    torch.cuda.synchronize()
    print(f"laswp m={m} n={n} jb={jb}: algorithmic bytes = 32*jb*n = {32 * jb * n / 1e9:.3f} GB")
elif mode == "panel":
    m, jb = (arg + [65536, 512])[:2] if len(arg) >= 2 else (65536, 512)
    W = torch.empty(jb * m, dtype=torch.float64, device="cuda")
    ipiv = np.zeros(jb, np.int32); info = C.c_int(0)
    for _ in range(2):
        S.matgen64(ctx, m, jb, 512, 512, W, m, 11)
        ms = L.slb200_test_panel(m, jb, S.api._ptr(W), I64(m), ipiv.ctypes.data_as(C.c_void_p), C.byref(info), 0)
    print(f"panel m={m} jb={jb}: {ms:.3f} ms  ({ms * 1e3 / jb:.2f} us per column), algorithmic bytes 16*m*jb = {16 * m * jb / 1e9:.3f} GB")
elif mode == "solve":
    n, nb = (arg + [65536, 512])[:2] if len(arg) >= 2 else (65536, 512)
    desca, _ = S.descinit(n, n, nb, nb, 0, 0, ctx, n); descb, _ = S.descinit(n, 1, nb, 1, 0, 0, ctx, n)
    A = torch.empty(n * n, dtype=torch.float64, device="cuda"); X = torch.empty(n, dtype=torch.float64, device="cuda")
    ipiv = np.zeros(n + nb, np.int32)
    S.matgen64(ctx, n, n, nb, nb, A, n, 20261017)
    assert S.pdgetrf(n, n, A, 1, 1, desca, ipiv) == 0
    for _ in range(2):
        S.matgen64(ctx, n, 1, nb, 1, X, n, 777)
        assert S.pdgetrs("N", n, 1, A, 1, 1, desca, ipiv, X, 1, 1, descb) == 0
    print(f"solve n={n}: {S.last_solve_ms():.3f} ms = {8.0 * n * n / S.last_solve_ms() / 1e6:.1f} GB/s of 8 N^2 bytes; sresid {S.pdlaschk(ctx, n, 1, X, descb, desca, 20261017, 777, gen=64):.2e}")
elif mode == "probe":
    n, nb = (arg + [16384, 512])[:2] if len(arg) >= 2 else (16384, 512)
    desca, _ = S.descinit(n, n, nb, nb, 0, 0, ctx, n)
    A = torch.empty(n * n, dtype=torch.float64, device="cuda"); ipiv = np.zeros(n + nb, np.int32)
    S.matgen64(ctx, n, n, nb, nb, A, n, 20261017)
    assert S.pdgetrf(n, n, A, 1, 1, desca, ipiv) == 0
    L.slb200_test_solve_probe.restype = C.c_double
    for ns in (0, 40, 200):
        S.set_option("solve_poll_ns", ns)
        print("poll", ns, "ns:", " ".join("%s %.2f us" % (name, L.slb200_test_solve_probe(which, nb, I64(0), S.api._ptr(A), I64(n), n, 50)) for which, name in ((0, "diag fwd"), (1, "diag bwd"), (2, "top"))))
    S.set_option("solve_poll_ns", 40)
    L.slb200_test_solve_probe(4, nb, I64(0), S.api._ptr(A), I64(n), n, 1)
    for nr in (n - 1024, n // 2, n // 4, n // 16):
        us = L.slb200_test_solve_probe(3, nb, I64(nr), S.api._ptr(A), I64(n), n, 50)
        print(f"bulk nr={nr}: {us:.2f} us = {nr * nb * 8 / us / 1e3:.1f} GB/s")
