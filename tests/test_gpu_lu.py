"""GPU parity tests of the drop-in entry points on a 1x1 grid, through the C-ABI, against the CPU oracle:
IPIV bit-exact, LU factors within tolerance, the reference's FRESID / SRESID below its threshold 1.0
(TESTING/traditional/LU.dat:17), guard zones intact (pdludriver.f:340-358,397-402)."""
import os

import numpy as np
import pytest

from tests.helpers import lu_err, first_mismatch, load_example_6x6, PADVAL

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def run_getrf(S, O, ctx, a0, nb, pad=0, device=False):
    """PDGETRF on a 1x1 grid with `pad` guard rows (LLD = M + pad) filled with PADVAL."""
    m, n = a0.shape
    lld = max(1, m) + pad
    al = np.full((lld, max(n, 1)), PADVAL, dtype=a0.dtype, order="F")
    al[:m, :n] = a0
    desc, info = S.descinit(m, n, nb, nb, 0, 0, ctx, lld)
    assert info == 0
    ipiv = np.full(m + nb + 4, -77, np.int32)
    f = S.pzgetrf if a0.dtype == np.complex128 else S.pdgetrf
    if device:
        import torch
        t = torch.from_numpy(np.ascontiguousarray(al.T)).cuda()       # (n, lld) row-major == (lld, n) column-major
        info = f(m, n, t, 1, 1, desc, ipiv)
        al = np.asfortranarray(t.cpu().numpy().T)
    else:
        info = f(m, n, al, 1, 1, desc, ipiv)
    assert np.all(al[m:, :] == PADVAL), "guard zone below A overwritten"
    assert np.all(ipiv[m + nb:] == -77), "IPIV written beyond LOCr(M_A)+MB_A"
    return al[:m, :n], ipiv[:min(m, n)].copy(), info, desc


def check_against_oracle(O, a0, lu, ipiv, info, nb, thresh=1.0):
    ref = a0.copy(order="F")
    ipr, infr = O.getrf(ref, nb)
    assert info == infr
    assert np.array_equal(ipiv, ipr), ("first IPIV mismatch at row", first_mismatch(ipiv, ipr))
    assert lu_err(lu, ref, a0) < 1.0
    fres = O.fresid(np.asfortranarray(lu), ipiv, a0)
    assert fres < thresh and fres - fres == 0.0
    return fres


@pytest.mark.parametrize("mn", [(4, 4), (10, 12), (17, 13), (13, 13)])
@pytest.mark.parametrize("nb", [2, 3, 4])
def test_lu_dat_grid_1x1(S, O, ctx11, mn, nb):
    """The reference's own test inputs (TESTING/traditional/LU.dat), grid 1x1."""
    m, n = mn
    a0 = O.pdmatgen(m, n, 100)
    lu, ipiv, info, desc = run_getrf(S, O, ctx11, a0, nb, pad=3)
    check_against_oracle(O, a0, lu, ipiv, info, nb)
    if m == n:
        for nrhs, nbrhs in [(1, 1), (3, 3), (9, 5)]:
            b0 = O.pdmatgen(n, nrhs, 200)
            descb, _ = S.descinit(n, nrhs, nb, nbrhs, 0, 0, ctx11, n)
            x = b0.copy(order="F")
            ipl = np.zeros(n + nb, np.int32); ipl[:n] = ipiv
            assert S.pdgetrs("N", n, nrhs, np.asfortranarray(lu), 1, 1, desc[:8] + [n], ipl, x, 1, 1, descb) == 0
            assert O.sresid(a0, x, b0) < 1.0


def test_example_pdgesv_6x6(S, O, ctx11):
    """EXAMPLE/pdscaex.f:155 on the shipped 6x6 system."""
    A, B = load_example_6x6(os.path.join(G, "DSCAEXMAT.dat"), os.path.join(G, "DSCAEXRHS.dat"))
    gold = np.load(os.path.join(G, "golden.npz"))
    a = A.copy(order="F"); b = B.copy(order="F")
    da, _ = S.descinit(6, 6, 2, 2, 0, 0, ctx11, 6); db, _ = S.descinit(6, 1, 2, 2, 0, 0, ctx11, 6)
    ipiv = np.zeros(8, np.int32)
    assert S.pdgesv(6, 1, a, 1, 1, da, ipiv, b, 1, 1, db) == 0
    assert np.array_equal(ipiv[:6], gold["ex6_ipiv"])
    assert np.allclose(b, gold["ex6_x"], rtol=1e-13)
    assert O.sresid(A, b, B) < 10.0


@pytest.mark.parametrize("n,nb,device", [(64, 8, False), (200, 64, False), (500, 32, True), (1000, 128, False), (2000, 64, True),
                                         (2000, 64, False), (1536, 512, True), (4096, 512, True), (3000, 500, False)])
def test_pdgetrf_square(S, O, ctx11, n, nb, device):
    """includes BASELINE config 1's matrix (PDMATGEN N=2000 NB=64 seed 100) on a 1x1 grid."""
    a0 = O.pdmatgen(n, n, 100)
    lu, ipiv, info, desc = run_getrf(S, O, ctx11, a0, nb, pad=2 if not device else 0, device=device)
    fres = check_against_oracle(O, a0, lu, ipiv, info, nb)
    b0 = O.pdmatgen(n, 1, 200)
    descb, _ = S.descinit(n, 1, nb, 1, 0, 0, ctx11, n)
    x = b0.copy(order="F")
    ipl = np.zeros(n + nb, np.int32); ipl[:n] = ipiv
    assert S.pdgetrs("N", n, 1, np.asfortranarray(lu), 1, 1, desc[:8] + [n], ipl, x, 1, 1, descb) == 0
    sres = O.sresid(a0, x, b0)
    assert sres < 1.0
    print(f"N={n} NB={nb} FRESID={fres:.4f} SRESID={sres:.5f} factor_ms={S.last_factor_ms():.2f}")


@pytest.mark.parametrize("m,n,nb", [(300, 200, 64), (200, 300, 64), (1000, 64, 64), (64, 1000, 32), (777, 513, 100)])
def test_pdgetrf_rectangular(S, O, ctx11, m, n, nb):
    a0 = O.pdmatgen(m, n, 100)
    lu, ipiv, info, _ = run_getrf(S, O, ctx11, a0, nb, pad=1)
    check_against_oracle(O, a0, lu, ipiv, info, nb)


def test_zero_pivot_info(S, O, ctx11):
    a0 = O.pdmatgen(96, 96, 100); a0[:, 40] = 0.0
    lu, ipiv, info, _ = run_getrf(S, O, ctx11, a0, 32)
    ref = a0.copy(order="F"); ipr, infr = O.getrf(ref, 32)
    assert info == infr == 41 and np.array_equal(ipiv, ipr)
    # PDGESV must skip the solve when INFO > 0 (pdgesv.f:231)
    a = a0.copy(order="F"); b0 = O.pdmatgen(96, 1, 200); b = b0.copy(order="F")
    da, _ = S.descinit(96, 96, 32, 32, 0, 0, ctx11, 96); db, _ = S.descinit(96, 1, 32, 1, 0, 0, ctx11, 96)
    assert S.pdgesv(96, 1, a, 1, 1, da, np.zeros(128, np.int32), b, 1, 1, db) == 41
    assert np.array_equal(b, b0)


@pytest.mark.parametrize("n,nb,nrhs", [(60, 8, 3), (500, 64, 1), (1024, 256, 2)])
def test_pzgetrf_pzgetrs(S, O, ctx11, n, nb, nrhs):
    a0 = O.pzmatgen(n, n, 100)
    lu, ipiv, info, desc = run_getrf(S, O, ctx11, a0, nb, pad=1)
    check_against_oracle(O, a0, lu, ipiv, info, nb)
    b0 = O.pzmatgen(n, nrhs, 200)
    descb, _ = S.descinit(n, nrhs, nb, 1, 0, 0, ctx11, n)
    x = b0.copy(order="F")
    ipl = np.zeros(n + nb, np.int32); ipl[:n] = ipiv
    assert S.pzgetrs("N", n, nrhs, np.asfortranarray(lu), 1, 1, desc[:8] + [n], ipl, x, 1, 1, descb) == 0
    assert O.sresid(a0, x, b0) < 1.0


def test_pdgesv_matches_oracle_solution(S, O, ctx11):
    n, nb, nrhs = 1500, 128, 4
    a0 = O.pdmatgen(n, n, 100); b0 = O.pdmatgen(n, nrhs, 200)
    a = a0.copy(order="F"); b = b0.copy(order="F")
    da, _ = S.descinit(n, n, nb, nb, 0, 0, ctx11, n); db, _ = S.descinit(n, nrhs, nb, 2, 0, 0, ctx11, n)
    ipiv = np.zeros(n + nb, np.int32)
    assert S.pdgesv(n, nrhs, a, 1, 1, da, ipiv, b, 1, 1, db) == 0
    ref = a0.copy(order="F"); ipr, _ = O.getrf(ref, nb)
    xr = b0.copy(order="F"); O.getrs(ref, ipr, xr)
    assert np.array_equal(ipiv[:n], ipr)
    assert np.abs(b - xr).max() / np.abs(xr).max() < 1e-9
    assert O.sresid(a0, b, b0) < 1.0


def test_device_generators_match_oracle(S, O, ctx11):
    """Device PDMATGEN / 64-bit generator == oracle (bit exact), and the device residual check == oracle's."""
    n, nb = 300, 32
    a = np.zeros((n, n), order="F")
    S.pdmatgen(ctx11, n, n, nb, nb, a, n, iseed=100)
    assert np.array_equal(a, O.pdmatgen(n, n, 100))
    a64 = np.zeros((n, n), order="F")
    S.matgen64(ctx11, n, n, nb, nb, a64, n, seed=42)
    assert np.array_equal(a64, O.matgen64_tile(n, 42, 0, n, 0, n))
    z64 = np.zeros((n, n), dtype=np.complex128, order="F")
    S.zmatgen64(ctx11, n, n, nb, nb, z64, n, seed=42)
    assert np.array_equal(z64, O.matgen64_tile(n, 42, 0, n, 0, n, complex_=True))
    # solve + device residual (b = first column of the generator with seed 43)
    b0 = O.matgen64_tile(n, 43, 0, n, 0, 1)
    lu = a64.copy(order="F"); x = b0.copy(order="F")
    da, _ = S.descinit(n, n, nb, nb, 0, 0, ctx11, n); db, _ = S.descinit(n, 1, nb, 1, 0, 0, ctx11, n)
    assert S.pdgesv(n, 1, lu, 1, 1, da, np.zeros(n + nb, np.int32), x, 1, 1, db) == 0
    r_dev = S.pdlaschk(ctx11, n, 1, x, db, da, 42, 43, gen=64)
    r_orc = O.sresid(a64, x, b0)
    assert r_dev < 1.0 and abs(r_dev - r_orc) <= 0.05 * r_orc + 1e-6, (r_dev, r_orc)


def test_large_properties_n8192(S, O, ctx11):
    """Size-independent properties at a size the oracle still finishes: N=8192 NB=512 (the bench block size)."""
    n, nb = 8192, 512
    a0 = O.matgen64_tile(n, 2024, 0, n, 0, n)
    lu, ipiv, info, desc = run_getrf(S, O, ctx11, a0, nb, device=True)
    ref = a0.copy(order="F"); ipr, infr = O.getrf(ref, nb)
    assert info == infr == 0 and np.array_equal(ipiv, ipr), first_mismatch(ipiv, ipr)
    assert lu_err(lu, ref, a0) < 1.0
    assert np.all(ipiv >= np.arange(1, n + 1)) and np.all(ipiv <= n)
    assert np.abs(np.tril(lu, -1)).max() <= 1.0 + 1e-12        # partial pivoting bound |l_ij| <= 1


def test_lookahead_is_bit_identical(S, O, ctx11):
    """Look-ahead only reorders independent work (next panel's columns first, left swaps on a side stream): the
    factors and pivots must be bit-identical to the serial schedule."""
    n, nb = 3072, 256
    a0 = O.matgen64_tile(n, 99, 0, n, 0, n)
    outs = []
    for la in (0, 1):
        S.set_option("lookahead", la)
        lu, ipiv, info, _ = run_getrf(S, O, ctx11, a0, nb, pad=0, device=True)
        outs.append((np.array(lu), ipiv, info))
    S.set_option("lookahead", 1)
    assert outs[0][2] == outs[1][2] == 0
    assert np.array_equal(outs[0][1], outs[1][1])
    assert np.array_equal(outs[0][0], outs[1][0])
    check_against_oracle(O, a0, outs[1][0], outs[1][1], 0, nb)


@pytest.mark.parametrize("n,nb,split_min", [(3072, 256, 512), (2560, 128, 256), (4096, 512, 1024)])
def test_pipelined_schedule_is_bit_identical(S, O, ctx11, n, nb, split_min):
    """The two-half pipeline (prep of one column half under the update of the other, moving boundary, re-splits) only
    reorders independent work: same bits as the serial schedule, pivots equal to the oracle's."""
    a0 = O.matgen64_tile(n, 1234 + n, 0, n, 0, n)
    outs = []
    for mode in ("serial", "pipe", "nopipe"):
        S.set_option("lookahead", 0 if mode == "serial" else 1)
        S.set_option("la_pipeline", 0 if mode == "nopipe" else 1)
        S.set_option("la_split_min", split_min)
        S.set_option("lookahead_min_us", 0)
        try:
            lu, ipiv, info, _ = run_getrf(S, O, ctx11, a0, nb, pad=0, device=True)
        finally:
            S.set_option("lookahead", 1); S.set_option("la_pipeline", 1); S.set_option("la_split_min", 6144)
            S.set_option("lookahead_min_us", 4000)
        outs.append((np.array(lu), ipiv, info))
    for o in outs[1:]:
        assert o[2] == outs[0][2] == 0
        assert np.array_equal(o[1], outs[0][1])
        assert np.array_equal(o[0], outs[0][0])
    check_against_oracle(O, a0, outs[1][0], outs[1][1], 0, nb)


def _gemm_in_subprocess(env_extra, M, N, K, tmp_path, tag):
    """C - A*B through slb200_test_gemm in a fresh process (kernel variant options are read once per process)."""
    import subprocess, sys, os
    out = str(tmp_path / f"gemm_{tag}.npy")
    code = (
        "import numpy as np, ctypes as C, scalapack_b200 as S\n"
        f"rng=np.random.default_rng(5); M,N,K={M},{N},{K}\n"
        "A=np.asfortranarray(rng.uniform(-1,1,(M,K))); B=np.asfortranarray(rng.uniform(-1,1,(K,N))); Cm=np.asfortranarray(rng.uniform(-1,1,(M,N)))\n"
        "ref=Cm-A@B; out=Cm.copy(order='F'); I=C.c_int64\n"
        "S.lib().slb200_test_gemm(I(M),I(N),K,S.api._ptr(A),I(M),S.api._ptr(B),I(K),S.api._ptr(out),I(M),0,1)\n"
        f"np.save(r'{out}', out)\n"
        "print('ERR', float(np.abs(out-ref).max()))\n")
    env = dict(os.environ, **{k: str(v) for k, v in env_extra.items()})
    p = subprocess.run([sys.executable, "-c", code], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))), env=env,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-2000:]
    err = float([l for l in p.stdout.splitlines() if l.startswith("ERR")][0].split()[1])
    return np.load(out), err


@pytest.mark.parametrize("opts", [{}, {"SLB200_GEMM_LAG": 6000}, {"SLB200_GEMM_EPI": 1}, {"SLB200_GEMM_TEST_CHUNK": 3}])
@pytest.mark.parametrize("shape", [(777, 1030, 200), (2048, 2304, 512), (130, 5000, 37)])
def test_packed_update_kernel_is_bit_identical_to_cp_async_kernel(S, opts, shape, tmp_path):
    """gemm_packed.cu (fragment-ordered operands, bulk-copy ring, mbarriers) vs the cp.async kernel: same bits, on ragged
    shapes (zero-padded blocks), with the late-start lag off, with the red.add epilogue and with chunked CTAs."""
    M, N, K = shape
    ref, err7 = _gemm_in_subprocess({"SLB200_GEMM_VARIANT": 7}, M, N, K, tmp_path, "v7")
    env = {"SLB200_GEMM_VARIANT": 9, "SLB200_GEMM_PACKED_MIN": 1}
    env.update(opts)
    out, err9 = _gemm_in_subprocess(env, M, N, K, tmp_path, "v9")
    assert err7 < 1e-11 and err9 < 1e-11, (err7, err9)
    assert np.array_equal(out, ref)


def test_full_size_properties_n65536(S, ctx11):
    """BASELINE config 2 at its full size (N=65536, NB=512, 32 GiB of A in HBM), where the oracle is out of reach:
    size-independent properties only -- INFO = 0, every pivot within [i, N] (PDGETRF's IPIV contract, pdgetrf.f:118-121),
    and the reference's solve residual ||Ax-b|| / (||A|| ||x|| N eps) (pdlaschk.f:187,296) of PDGETRS with the factors,
    evaluated on the device against the regenerated matrix, below its threshold 1.0.  Mirrors what bench.py does."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    n, nb = 65536, 512
    if free < 48 * 2**30:
        pytest.skip("needs 48 GiB of free HBM")
    lld = n
    desca, info = S.descinit(n, n, nb, nb, 0, 0, ctx11, lld)
    assert info == 0
    A = torch.empty(n * lld, dtype=torch.float64, device="cuda")
    ipiv = np.zeros(n + nb, np.int32)
    S.matgen64(ctx11, n, n, nb, nb, A, lld, 20261017)
    torch.cuda.synchronize()
    assert S.pdgetrf(n, n, A, 1, 1, desca, ipiv) == 0
    rows = np.arange(1, n + 1)
    assert np.all(ipiv[:n] >= rows) and np.all(ipiv[:n] <= n)
    descb, _ = S.descinit(n, 1, nb, 1, 0, 0, ctx11, lld)
    X = torch.zeros(lld, dtype=torch.float64, device="cuda")
    S.matgen64(ctx11, n, 1, nb, 1, X, lld, 777)
    assert S.pdgetrs("N", n, 1, A, 1, 1, desca, ipiv, X, 1, 1, descb) == 0
    sresid = S.pdlaschk(ctx11, n, 1, X, descb, desca, 20261017, 777, gen=64)
    assert 0.0 <= sresid < 1.0, sresid
    del A, X
    torch.cuda.empty_cache()


def _host_array(a0, lld, pinned):
    """(lld, n) column-major host copy of a0 with PADVAL guard rows, page-locked or pageable."""
    import torch
    m, n = a0.shape
    host = torch.full((n, lld), PADVAL, dtype=torch.float64)
    if pinned:
        host = host.pin_memory()
    host[:, :m] = torch.from_numpy(np.ascontiguousarray(a0.T))
    return host


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("m,n,nb,slab_mb,save_mb", [(3072, 3072, 256, 0, 16384), (2000, 1500, 128, 0, 16384), (1500, 2000, 128, 0, 16384),
                                                    (4096, 4096, 512, 0, 40), (4096, 4096, 256, 4, 16384), (2048, 2048, 128, 0, 0),
                                                    (4096, 4096, 128, 16, 16384), (8192, 8192, 256, 64, 16384)])
def test_host_resident_streaming_is_bit_identical(S, O, ctx11, m, n, nb, slab_mb, save_mb, pinned):
    """A host-resident caller (pinned or pageable memory): A is uploaded in column slabs that join the sweep as they arrive
    (replay of the steps they missed) and the factors go back in block rows during the factorisation.  The host array must
    end up bit-identical to the device-resident factorisation, guard rows untouched.  slab_mb=0: one slab per block column
    (a join in almost every step); small save_mb: the kept-panel budget runs out and the remaining slabs are waited for."""
    a0 = O.matgen64_tile(max(m, n), 4321, 0, m, 0, n)
    lld = m + 3
    lu_dev, ipiv_dev, info_dev, _ = run_getrf(S, O, ctx11, a0, nb, pad=0, device=True)
    S.set_option("e2e_overlap_min_mb", 0); S.set_option("e2e_slab_mb", slab_mb); S.set_option("e2e_save_mb", save_mb)
    S.set_option("la_split_min", 512); S.set_option("lookahead_min_us", 0)
    try:
        host = _host_array(a0, lld, pinned)
        desc, info = S.descinit(m, n, nb, nb, 0, 0, ctx11, lld)
        ipiv = np.zeros(m + nb, np.int32)
        S.reset_counters()
        assert S.pdgetrf(m, n, host.numpy(), 1, 1, desc, ipiv) == info_dev == 0
        assert S.get_counter("e2e_upload_overlapped") == 1
        assert S.get_counter("h2d_bytes") == m * n * 8 and S.get_counter("d2h_bytes") == m * n * 8
    finally:
        S.set_option("e2e_overlap_min_mb", 256); S.set_option("e2e_slab_mb", 1024); S.set_option("e2e_save_mb", 16384)
        S.set_option("la_split_min", 6144); S.set_option("lookahead_min_us", 4000)
    out = host.numpy().T                                        # (lld, n)
    assert np.array_equal(ipiv[:min(m, n)], ipiv_dev)
    assert np.array_equal(out[:m, :], lu_dev)
    assert np.all(out[m:, :] == PADVAL)


def test_host_resident_small_matrix_takes_the_one_copy_path(S, O, ctx11):
    n, nb = 600, 64
    a0 = O.pdmatgen(n, n, 100)
    S.reset_counters()
    lu, ipiv, info, _ = run_getrf(S, O, ctx11, a0, nb, pad=2)
    assert S.get_counter("e2e_upload_overlapped") == 0
    check_against_oracle(O, a0, lu, ipiv, info, nb)


@pytest.mark.xfail(reason="written after the round's GPU budget was spent: first hardware run (passes on the CPU emulation)", strict=False)
def test_product_against_the_executed_reference_fortran(S, ctx11):
    """PDGETRF / PDGETRS through the C-ABI against tests/golden/lu_reference.npz = what the reference's OWN Fortran (pdgetrf.f, pdgetf2.f,
    pdlaswp.f, pdgetrs.f, executed by tests/fortran77_mini.py with numpy PBLAS leaves; tests/golden/make_lu_golden.py) produces on a
    1 x 1 grid: INFO and IPIV exactly (sub-matrix operands, partial blocks, M != N, NB > N, zero pivot columns), factors and solutions to
    rounding, nothing outside sub(A) touched.  No oracle in between."""
    import oracle as O                                        # only for the PDMATGEN input matrices (closed-form generator)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "lu_reference.npz"))
    ncases = sum(1 for k in g.files if k.startswith("case"))
    for i in range(ncases):
        m, n, nb, mg, ng, ia, ja, zero_col, info_ref = [int(v) for v in g[f"case{i}"]]
        a0 = O.pdmatgen(mg, ng, 100).copy(order="F")
        if zero_col >= 0:
            a0[:, zero_col] = 0.0
        lld = mg + 1
        al = np.full((lld, ng), PADVAL, order="F"); al[:mg, :] = a0
        desc, info = S.descinit(mg, ng, nb, nb, 0, 0, ctx11, lld)
        ipiv = np.full(mg + nb, -77, np.int32)
        info = S.pdgetrf(m, n, al, ia, ja, desc, ipiv)
        lu_ref, ipiv_ref = g[f"lu{i}"], g[f"ipiv{i}"]
        mn = min(m, n)
        assert info == info_ref, (i, info, info_ref)
        assert np.array_equal(ipiv[ia - 1:ia - 1 + mn], ipiv_ref[ia - 1:ia - 1 + mn]), i
        anorm = max(np.abs(a0[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n]).sum(axis=1).max(), 1e-300)
        assert np.abs(al[:mg, :] - lu_ref).max() / (anorm * max(m, n) * 2.0 ** -53) < 1.0, i
        outside = np.ones((mg, ng), bool); outside[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n] = False
        assert np.array_equal(al[:mg, :][outside], a0[outside]) and np.all(al[mg:, :] == PADVAL), i
        for trans in "NT":
            if f"x{trans}{i}" not in g.files:
                continue
            b = np.asfortranarray(O.pdmatgen(n, 3, 200).copy(order="F"))
            descb, _ = S.descinit(n, 3, nb, 1, 0, 0, ctx11, n)
            assert S.pdgetrs(trans, n, 3, al, 1, 1, desc, ipiv, b, 1, 1, descb) == 0
            xref = g[f"x{trans}{i}"]
            assert np.abs(b - xref).max() <= 1e-9 * np.abs(xref).max(), (i, trans)
