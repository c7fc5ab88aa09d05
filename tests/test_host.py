"""CPU tests of the product's host side: the C-ABI library loads, exports every declared symbol, its TOOLS
restatement equals the oracle's, argument checks return the reference's INFO codes, and the multi-process
control plane (BLACS grid setup, scoped barriers, PCHK1MAT consistency) works at world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(S):
    hdr = open(os.path.join(ROOT, "include", "scalapack_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", hdr))
    names -= {"defined", "if", "sizeof"}
    names = {n for n in names if n.endswith("_") or n.startswith("Cblacs") or n.startswith("slb200_")}
    assert len(names) > 50
    L = S.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing


def test_tools_match_oracle(S, O, ctx11):
    rng = np.random.default_rng(0)
    for _ in range(200):
        n, nb, P = int(rng.integers(0, 200)), int(rng.integers(1, 17)), int(rng.integers(1, 6))
        p, src = int(rng.integers(0, P)), int(rng.integers(0, P))
        assert S.numroc(n, nb, p, src, P) == O.numroc(n, nb, p, src, P)
        ig = int(rng.integers(1, 300))
        assert S.indxg2p(ig, nb, 0, src, P) == O.indxg2p(ig, nb, 0, src, P)
        assert S.indxg2l(ig, nb, 0, 0, P) == O.indxg2l(ig, nb, 0, 0, P)
        assert S.indxl2g(ig, nb, p, src, P) == O.indxl2g(ig, nb, p, src, P)
        Q, q = int(rng.integers(1, 5)), 0
        desc = [1, 0, 500, 400, nb, nb, src, 0, 500]
        gr, gc = int(rng.integers(1, 500)), int(rng.integers(1, 400))
        assert S.infog2l(gr, gc, desc, P, Q, p, q) == O.infog2l(gr, gc, desc, P, Q, p, q)
    assert S.iceil(7, 3) == 3 and S.ilcm(4, 6) == 12


def test_index_tools_against_the_block_cyclic_definition(S, O):
    """NUMROC / INDXG2P / INDXG2L / INDXL2G / INFOG2L of the product AND of the oracle against a brute-force enumeration of the
    2D block-cyclic map itself (global index g, 0-based, lives on process (isrc + g // nb) % P; its local index is its rank among the
    indices that process owns) -- independent of both restatements of TOOLS/*.f, plus literal values worked by hand from the
    Fortran (TOOLS/numroc.f, indxg2l.f examples: N=10, NB=2, P=3)."""
    for impl in (S, O):
        assert [impl.numroc(10, 2, p, 0, 3) for p in range(3)] == [4, 4, 2]
        assert [impl.numroc(10, 2, p, 1, 3) for p in range(3)] == [2, 4, 4]
        assert impl.numroc(0, 4, 0, 0, 2) == 0 and impl.numroc(7, 8, 0, 0, 2) == 7 and impl.numroc(7, 8, 1, 0, 2) == 0
        assert [impl.indxg2p(g, 2, 0, 0, 3) for g in range(1, 11)] == [0, 0, 1, 1, 2, 2, 0, 0, 1, 1]
        assert [impl.indxg2l(g, 2, 0, 0, 3) for g in range(1, 11)] == [1, 2, 1, 2, 1, 2, 3, 4, 3, 4]
    rng = np.random.default_rng(7)
    for _ in range(60):
        n, nb, P = int(rng.integers(1, 150)), int(rng.integers(1, 12)), int(rng.integers(1, 6))
        src = int(rng.integers(0, P))
        owner = [(src + g // nb) % P for g in range(n)]
        local = []                                               # 1-based local index of global index g+1
        cnt = [0] * P
        for g in range(n):
            cnt[owner[g]] += 1; local.append(cnt[owner[g]])
        for impl in (S, O):
            for p in range(P):
                assert impl.numroc(n, nb, p, src, P) == cnt[p]
            for g in range(n):
                assert impl.indxg2p(g + 1, nb, 0, src, P) == owner[g]
                assert impl.indxg2l(g + 1, nb, 0, 0, P) == local[g]
                assert impl.indxl2g(local[g], nb, owner[g], src, P) == g + 1
        # INFOG2L: (local row, local column) of the first element at or after (gr, gc) on every process, and its owner
        Q = int(rng.integers(1, 5)); csrc = int(rng.integers(0, Q)); m = int(rng.integers(1, 150))
        cown = [(csrc + g // nb) % Q for g in range(m)]
        desc = [1, 0, n, m, nb, nb, src, csrc, max(1, n)]
        gr, gc = int(rng.integers(1, n + 1)), int(rng.integers(1, m + 1))
        for pr in range(P):
            for pc in range(Q):
                want_lr = 1 + sum(1 for g in range(gr - 1) if owner[g] == pr)
                want_lc = 1 + sum(1 for g in range(gc - 1) if cown[g] == pc)
                for impl in (S, O):
                    lr, lc, rs, cs = impl.infog2l(gr, gc, desc, P, Q, pr, pc)
                    assert (lr, lc, rs, cs) == (want_lr, want_lc, owner[gr - 1], cown[gc - 1]), (impl.__name__, n, m, nb, P, Q, src, csrc, gr, gc, pr, pc)


def test_descinit_and_chk1mat(S, O, ctx11):
    d, info = S.descinit(10, 12, 4, 4, 0, 0, ctx11, 10)
    assert info == 0 and d == [1, ctx11, 10, 12, 4, 4, 0, 0, 10]
    do, io = O.descinit(10, 12, 4, 4, 0, 0, ctx11, 10, 1, 1, 0)
    assert d == do and io == 0
    # illegal LLD -> -9 and clamped (descinit.f:172-186)
    d2, info2 = S.descinit(10, 12, 4, 4, 0, 0, ctx11, 3)
    assert info2 == -9 and d2[8] == 10
    assert S.descinit(-1, 12, 4, 4, 0, 0, ctx11, 10)[1] == -2
    assert S.descinit(10, 12, 4, 4, 1, 0, ctx11, 10)[1] == -6
    # CHK1MAT codes (chk1mat.f:92-171) against the oracle restatement
    for (ma, na, ia, ja, desc) in [(10, 12, 1, 1, d), (11, 12, 1, 1, d), (10, 12, 0, 1, d), (10, 12, 1, 13, d),
                                   (10, 12, 1, 1, [2] + d[1:]), (10, 12, 1, 1, d[:4] + [0] + d[5:]), (10, 12, 1, 1, d[:8] + [4])]:
        assert S.chk1mat(ma, 1, na, 2, ia, ja, desc, 6) == O.chk1mat(ma, 1, na, 2, ia, ja, desc, 6, 1, 1, 0, 0)


def test_pdgetrf_error_exits(S, ctx11, capfd):
    """INFO codes and PXERBLA text of SRC/pdgetrf.f:169-197 (no GPU is touched on these paths)."""
    d, _ = S.descinit(10, 12, 4, 4, 0, 0, ctx11, 10)
    a = np.zeros((10, 12), order="F")
    ipiv = np.zeros(16, np.int32)
    assert S.pdgetrf(-1, 12, a, 1, 1, d, ipiv) == -1
    assert S.pdgetrf(10, -3, a, 1, 1, d, ipiv) == -2
    assert S.pdgetrf(8, 12, a, 2, 1, d, ipiv) == -4
    assert S.pdgetrf(10, 8, a, 1, 3, d, ipiv) == -5
    assert S.pdgetrf(10, 12, a, 1, 1, d[:5] + [3] + d[6:], ipiv) == -606
    assert S.pdgetrf(10, 12, a, 1, 1, [1, 77] + d[2:], ipiv) == -602
    assert S.pdgetrf(10, 12, a, 1, 1, d[:8] + [5], ipiv) == -609
    err = capfd.readouterr().err
    assert "On entry to PDGETRF parameter number 606 had an illegal value" in err
    # quick returns (pdgetrf.f:201-206)
    assert S.pdgetrf(0, 12, a, 1, 1, d, ipiv) == 0
    d1, _ = S.descinit(1, 1, 4, 4, 0, 0, ctx11, 1)
    ipiv[:] = 0
    assert S.pdgetrf(1, 1, np.ones((1, 1), order="F"), 1, 1, d1, ipiv) == 0 and ipiv[0] == 1


def test_pdgetrs_pdgesv_error_exits(S, ctx11):
    da, _ = S.descinit(8, 8, 4, 4, 0, 0, ctx11, 8)
    db, _ = S.descinit(8, 2, 4, 1, 0, 0, ctx11, 8)
    a = np.zeros((8, 8), order="F"); b = np.zeros((8, 2), order="F"); ipiv = np.zeros(12, np.int32)
    assert S.pdgetrs("X", 8, 2, a, 1, 1, da, ipiv, b, 1, 1, db) == -1
    assert S.pdgetrs("N", -1, 2, a, 1, 1, da, ipiv, b, 1, 1, db) == -2
    assert S.pdgetrs("N", 8, -1, a, 1, 1, da, ipiv, b, 1, 1, db) == -3
    assert S.pdgetrs("N", 8, 2, a, 1, 1, da, ipiv, b, 1, 1, db[:4] + [2] + db[5:]) == -1206      # the reference reports NB_ here (pdgetrs.f:223-224)
    assert S.pdgetrs("N", 8, 2, a, 1, 1, da[:5] + [2] + da[6:], ipiv, b, 1, 1, db) == -706
    assert S.pdgesv(-1, 2, a, 1, 1, da, ipiv, b, 1, 1, db) == -1
    assert S.pdgesv(8, -1, a, 1, 1, da, ipiv, b, 1, 1, db) == -2
    assert S.pdgesv(8, 2, a, 1, 1, da, ipiv, b, 1, 1, db[:4] + [2] + db[5:]) == -1106
    assert S.pdgetrs("N", 0, 2, a, 1, 1, da, ipiv, b, 1, 1, db) == 0          # quick return


def test_no_cpu_fallback(S, ctx11):
    """Without a GPU the compute entry point must die loudly, never compute on the host."""
    if S.has_cuda():
        pytest.skip("GPU present")
    code = ("import numpy as np, scalapack_b200 as S; c=S.blacs_gridinit(0,'R',1,1); d,_=S.descinit(8,8,4,4,0,0,c,8);"
            "a=np.eye(8,order='F'); S.pdgetrf(8,8,a,1,1,d,np.zeros(12,np.int32)); print('COMPUTED')")
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=120)
    assert p.returncode != 0 and "COMPUTED" not in p.stdout
    assert "no CPU fallback" in p.stderr


WORKER = r"""
import sys, json
sys.path.insert(0, %r)
import numpy as np, scalapack_b200 as S
me, n = S.blacs_pinfo()
out = {"me": me, "n": n}
ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 2, 1)
out["grid"] = S.blacs_gridinfo(ctx)
ctx2 = S.blacs_gridinit(S.blacs_get(-1, 0), "Col-major", 1, 2)
out["grid2"] = S.blacs_gridinfo(ctx2)
out["pnum"] = [S.blacs_pnum(ctx, 0, 0), S.blacs_pnum(ctx, 1, 0), S.blacs_pnum(ctx, 2, 0)]
out["pcoord"] = S.blacs_pcoord(ctx, 1)
S.blacs_barrier(ctx, "All"); S.blacs_barrier(ctx, "Column"); S.blacs_barrier(ctx, "Row")
d, info = S.descinit(10, 10, 2, 2, 0, 0, ctx, 6)
out["numroc"] = S.numroc(10, 2, out["grid"][2], 0, 2)
a = np.zeros((6, 10), order="F"); ipiv = np.zeros(10, np.int32)
# PCHK1MAT: N differs between the two processes -> both must return -2 (pchkxmat.f:404-490)
out["incons"] = S.pdgetrf(10, 10 if me == 0 else 9, a, 1, 1, d, ipiv)
# one-process grid inside a two-process job: rank 1 is outside -> context -1, gridinfo all -1
c1 = S.blacs_gridinit(S.blacs_get(-1, 0), "R", 1, 1)
out["solo_ctx_valid"] = c1 >= 0
out["solo_info"] = S.blacs_gridinfo(c1)
# a grid created AFTER one that only rank 0 belongs to: the context handles now differ between the ranks (as in the reference,
# where a handle is an index into the process's own table, blacs_map_.c:72-77,127) -- collectives must not depend on them
ctx3 = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 2, 1)
out["ctx3"] = ctx3; out["grid3"] = S.blacs_gridinfo(ctx3)
S.blacs_barrier(ctx3, "All")
d3, _ = S.descinit(10, 10, 2, 2, 0, 0, ctx3, 6)
out["incons3"] = S.pdgetrf(10 if me == 0 else 8, 10, np.zeros((6, 10), order="F"), 1, 1, d3, np.zeros(10, np.int32))
import ctypes as C
v = (C.c_int * 2)(me + 5, 10 - me)
S.lib().igamn2d_(C.byref(C.c_int(ctx)), b"All", b" ", C.byref(C.c_int(2)), C.byref(C.c_int(1)), v, C.byref(C.c_int(2)),
                 None, None, C.byref(C.c_int(-1)), C.byref(C.c_int(-1)), C.byref(C.c_int(-1)))
out["igamn"] = list(v)
S.blacs_gridexit(ctx); S.blacs_gridexit(ctx2); S.blacs_exit(0)
print("RESULT" + json.dumps(out), flush=True)
"""


def test_two_process_control_plane(S):
    """world_size-2 run on CPU (the reference's analogue: mpiexec -n 2, TESTING/traditional/CMakeLists.txt:22-34)."""
    import json
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SLB200_PORT_OFFSET="0")
        procs.append(subprocess.Popen([sys.executable, "-c", WORKER % ROOT], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    res = []
    for p in procs:
        o, e = p.communicate(timeout=120)
        assert p.returncode == 0, e
        res.append(json.loads([l for l in o.splitlines() if l.startswith("RESULT")][0][6:]))
    res.sort(key=lambda d: d["me"])
    assert [r["n"] for r in res] == [2, 2]
    assert res[0]["grid"] == [2, 1, 0, 0] and res[1]["grid"] == [2, 1, 1, 0]           # row-major map (blacs_init_.c)
    assert res[0]["grid2"] == [1, 2, 0, 0] and res[1]["grid2"] == [1, 2, 0, 1]
    assert res[0]["pnum"] == [0, 1, -1] and res[0]["pcoord"] == [1, 0]
    assert [r["numroc"] for r in res] == [6, 4]
    assert [r["incons"] for r in res] == [-2, -2]
    assert res[0]["solo_ctx_valid"] and not res[1]["solo_ctx_valid"] and res[1]["solo_info"] == [-1, -1, -1, -1]
    assert res[0]["igamn"] == [5, 9] and res[1]["igamn"] == [5, 9]
    assert res[0]["ctx3"] != res[1]["ctx3"] and res[0]["grid3"] == [2, 1, 0, 0] and res[1]["grid3"] == [2, 1, 1, 0]
    assert [r["incons3"] for r in res] == [-1, -1]                      # PCHK1MAT on the later grid: M differs -> -1 on both


def test_bench_reference_arm_rank_contract():
    """bench.py --impl reference: under torchrun only rank 0 works and prints the JSON line (with impl / cpu_baseline / e2e keys);
    the other ranks exit 0 without output.  Runs on the host cores (no GPU)."""
    import json
    env1 = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        cwd=ROOT, env=env1, capture_output=True, text=True, timeout=120)
    assert p1.returncode == 0 and p1.stdout.strip() == ""
    env0 = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0", SLB200_BENCH_CPU_SAMPLE_N="4096")
    p0 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        cwd=ROOT, env=env0, capture_output=True, text=True, timeout=300)
    assert p0.returncode == 0, p0.stderr[-2000:]
    line = json.loads([l for l in p0.stdout.splitlines() if l.startswith("{")][0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["unit"] == "TFLOP/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]


def test_bench_config_table():
    """bench.py's BASELINE config table: sizes, grids, flop models (pdludriver.f:913-918) and the shared workload string."""
    import argparse
    import bench as B
    def cfg(gpus, config="", n=0, nb=0):
        return B.pick_config(argparse.Namespace(gpus=gpus, config=config, n=n, nb=nb))
    assert cfg(1)["name"] == "c2" and cfg(1)["n"] == 65536 and cfg(1)["nb"] == 512
    assert cfg(8)["name"] == "weak" and cfg(8)["n"] == int(65536 * 8 ** 0.5) // 512 * 512 == 185344
    assert cfg(4)["n"] == 131072 and cfg(2)["n"] == 92672
    c3, c4, c5 = cfg(8, "c3"), cfg(8, "c4"), cfg(8, "c5")
    assert (c3["routine"], c3["n"], c3["scaling"]) == ("pdgesv", 131072, "strong")
    assert (c4["routine"], c4["n"]) == ("pdgetrf", 262144) and (c5["routine"], c5["n"], c5["nb"], c5["cplx"]) == ("pzgetrf", 65536, 256, True)
    n = 131072.0
    head, ref = B.flop_counts(c3)
    assert head == 2.0 / 3.0 * n ** 3 + 2 * n * n and ref == 2.0 / 3.0 * n ** 3 - 0.5 * n * n + 2 * n * n
    assert B.flop_counts(c5)[0] == 4 * 2.0 / 3.0 * 65536.0 ** 3
    assert "grid 2x4" in B.workload_string(c4, 8) and "64.0 GiB" in B.workload_string(c4, 8)
    assert cfg(1, "c5", n=8192)["n"] == 8192 and set(B.METRIC) == {"pdgetrf", "pdgesv", "pzgetrf"}


def test_compiled_c_caller_links_against_the_boundary(S, tmp_path):
    """examples/pdgesv_example.c (the flow of EXAMPLE/pdscaex.f written against include/scalapack_b200.h) compiles with plain gcc
    and links against the shared library -- the drop-in boundary needs nothing but the header and the .so.  Without a GPU the program
    must stop loudly (no CPU fallback); with one it solves the reference's 6 x 6 tutorial system below the example's threshold."""
    exe = str(tmp_path / "pdgesv_example")
    libdir = os.path.join(ROOT, "scalapack_b200", "lib")
    cc = subprocess.run(["gcc", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "examples", "pdgesv_example.c"), "-I" + os.path.join(ROOT, "include"),
                         "-L" + libdir, "-lscalapack_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], capture_output=True, text=True, timeout=120)
    assert cc.returncode == 0, cc.stderr
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    if S.has_cuda():
        assert run.returncode == 0 and "scaled residual" in run.stdout, run.stderr
    else:
        assert run.returncode != 0 and "no CPU fallback" in run.stderr


def test_compiled_expert_caller_links_against_the_boundary(S, tmp_path):
    """examples/expert_example.c: PDGESVX / PDLANGE / PDGECON / PDGEMR2D / PDGETRI / PDGEMM / PDPOSV called from plain C with the
    header's prototypes; without a GPU it must stop loudly, with one every step checks its own residual."""
    exe = str(tmp_path / "expert_example")
    libdir = os.path.join(ROOT, "scalapack_b200", "lib")
    cc = subprocess.run(["gcc", "-O1", "-Wall", "-Werror", os.path.join(ROOT, "examples", "expert_example.c"), "-I" + os.path.join(ROOT, "include"),
                         "-L" + libdir, "-lscalapack_b200", "-Wl,-rpath," + libdir, "-lm", "-o", exe], capture_output=True, text=True, timeout=120)
    assert cc.returncode == 0, cc.stderr
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    if S.has_cuda():
        assert run.returncode == 0 and "expert example ok" in run.stdout, (run.stdout, run.stderr)
    else:
        assert run.returncode != 0 and "no CPU fallback" in run.stderr


def test_submatrix_window_against_infog2l_and_enumeration(S):
    """The local window of a block-aligned sub(A) = A(IA:IA+M-1, JA:JA+N-1) that PDGETRF / PDGETRS work on (api.cu `window`): offsets
    equal INFOG2L's local indices (TOOLS/infog2l.f, which pdgetrf.f:202 uses), the window's source process is the owner of (IA, JA),
    and its extent equals a brute-force count of the owned rows / columns -- on every process of random P x Q grids."""
    import ctypes as C
    L = S.lib()
    rng = np.random.default_rng(11)
    out = (C.c_int64 * 6)()
    for _ in range(150):
        P, Q, nb = int(rng.integers(1, 5)), int(rng.integers(1, 5)), int(rng.integers(1, 9))
        M, N = int(rng.integers(nb, 12 * nb)), int(rng.integers(nb, 12 * nb))
        rs, cs = int(rng.integers(0, P)), int(rng.integers(0, Q))
        ia = 1 + nb * int(rng.integers(0, (M - 1) // nb + 1)); ja = 1 + nb * int(rng.integers(0, (N - 1) // nb + 1))
        m = int(rng.integers(1, M - ia + 2)); n = int(rng.integers(1, N - ja + 2))
        desc = [1, 0, M, N, nb, nb, rs, cs, M]
        for pr in range(P):
            for pc in range(Q):
                L.slb200_test_window(m, n, ia, ja, (C.c_int * 9)(*desc), P, Q, pr, pc, out)
                lr, lc, orow, ocol = S.infog2l(ia, ja, desc, P, Q, pr, pc)
                rows = [g for g in range(ia - 1, ia - 1 + m) if (rs + g // nb) % P == pr]
                cols = [g for g in range(ja - 1, ja - 1 + n) if (cs + g // nb) % Q == pc]
                assert list(out) == [lr - 1, lc - 1, len(rows), len(cols), orow, ocol], (P, Q, nb, M, N, rs, cs, ia, ja, m, n, pr, pc, list(out))
                # the window's rows are CONTIGUOUS in the local array: local index of the k-th owned row = offset + k
                if rows:
                    before = sum(1 for g in range(rows[0]) if (rs + g // nb) % P == pr)
                    assert before == out[0] and [S.indxg2l(g + 1, nb, 0, 0, P) - 1 for g in rows] == list(range(before, before + len(rows)))


@pytest.mark.parametrize("threads", [1, 2, 5])
def test_worker_pool_2d_copy(S, threads):
    """The pageable-memory path of the host <-> HBM link moves the caller's array with a pool of memcpy threads (stage.cu): 2-D copies
    of every shape it cuts (many short columns grouped, long columns split along their length, contiguous blocks) equal numpy's."""
    import ctypes as C
    L = S.lib()
    rng = np.random.default_rng(threads)
    for width, n, spitch, dpitch in [(8, 1000, 24, 8), (4096, 300, 8192, 4096), (1 << 20, 5, (1 << 20) + 64, 1 << 20), (700000, 3, 700000, 700008),
                                     (1 << 16, 64, 1 << 16, 1 << 16), (1, 1, 1, 1), (333, 0, 400, 333)]:
        src = rng.integers(0, 255, size=max(1, spitch * max(n, 1)), dtype=np.uint8)
        dst = np.full(max(1, dpitch * max(n, 1)), 7, np.uint8)
        L.slb200_test_copy2d(dst.ctypes.data_as(C.c_void_p), C.c_size_t(dpitch), src.ctypes.data_as(C.c_void_p), C.c_size_t(spitch),
                             C.c_size_t(width), C.c_int64(n), threads)
        want = np.full_like(dst, 7)
        for j in range(n):
            want[j * dpitch:j * dpitch + width] = src[j * spitch:j * spitch + width]
        assert np.array_equal(dst, want), (width, n, spitch, dpitch)


def test_fortran_interface_module_matches_the_c_header(S):
    """No Fortran compiler exists in the build image or on the GPU boxes, so fortran/scalapack_b200_iface.f90 cannot be compiled.  It is
    parsed with a real Fortran parser instead (numpy.f2py's crackfortran): every interface must bind a symbol the library exports, with
    as many arguments as the C prototype in include/scalapack_b200.h, each argument declared with a C-interoperable type of the right
    class (integer -> int*, real(c_double) -> double*, complex -> slb200_z*, character -> char*)."""
    import ctypes as C
    import re
    import numpy.f2py.crackfortran as cf
    cf.verbose = 0
    cf.quiet = 1
    blocks = cf.crackfortran([os.path.join(ROOT, "fortran", "scalapack_b200_iface.f90")])
    procs = []

    def walk(bs):
        for b in bs:
            if b.get("block") in ("subroutine", "function"):
                procs.append(b)
            walk(b.get("body", []))
    walk(blocks)
    assert len(procs) >= 24
    header = open(os.path.join(ROOT, "include", "scalapack_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    L = S.lib()
    for p in procs:
        name = p["name"]
        cname = p["bindlang"][name]["name"]
        assert p["bindlang"][name]["lang"] == "c"
        assert hasattr(L, cname), f"{cname} is not exported"
        m = re.search(r"\b%s\s*\(([^;]*?)\)\s*;" % re.escape(cname), header, flags=re.S)
        assert m, f"{cname} is not declared in the header"
        cargs = [a.strip() for a in m.group(1).split(",")] if m.group(1).strip() not in ("", "void") else []
        assert len(cargs) == len(p["args"]), f"{cname}: {len(p['args'])} Fortran arguments, {len(cargs)} in the C prototype"
        for fa, ca in zip(p["args"], cargs):
            spec = p["vars"][fa]
            ts = spec["typespec"]
            if ts == "integer":
                assert re.search(r"\b(int|int64_t|uint64_t)\b", ca) and "*" in ca, (cname, fa, ca)
            elif ts == "real":
                assert "double" in ca and "*" in ca, (cname, fa, ca)
            elif ts == "complex":
                assert "slb200_z" in ca and "*" in ca, (cname, fa, ca)
            elif ts == "character":
                assert "char" in ca and "*" in ca, (cname, fa, ca)
            else:
                raise AssertionError((cname, fa, ts))


def _tools_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "tools_reference.npz"))


def test_index_tools_against_the_reference_source(S, O):
    """NUMROC / INDXG2P / INDXG2L / INDXL2G / ICEIL / ILCM / INFOG2L of the product (tools.cpp) and of the oracle against values
    produced by EXECUTING the reference's own Fortran (tests/golden/tools_reference.npz, made by tests/golden/make_tools_golden.py with
    the Fortran-77 mini interpreter of tests/fortran77_mini.py): neither restatement is checked against the other any more."""
    g = _tools_golden()
    for n, nb, ip, isrc, np_, ig, il, a, b, r_numroc, r_g2p, r_g2l, r_l2g, r_ceil, r_lcm in g["index"].tolist():
        assert S.numroc(n, nb, ip, isrc, np_) == r_numroc == O.numroc(n, nb, ip, isrc, np_)
        assert S.indxg2p(ig, nb, ip, isrc, np_) == r_g2p == O.indxg2p(ig, nb, ip, isrc, np_)
        assert S.indxg2l(ig, nb, ip, isrc, np_) == r_g2l == O.indxg2l(ig, nb, ip, isrc, np_)
        assert S.indxl2g(il, nb, ip, isrc, np_) == r_l2g == O.indxl2g(il, nb, ip, isrc, np_)
        assert S.iceil(a, b) == r_ceil and S.ilcm(a, b) == r_lcm
    for row in g["infog2l"].tolist():
        P, Q, r, c = row[:4]; d = row[4:13]; gi, gj = row[13:15]; want = tuple(row[15:19])
        assert S.infog2l(gi, gj, d, P, Q, r, c) == want == O.infog2l(gi, gj, d, P, Q, r, c)


def test_descinit_chk1mat_against_the_reference_source(S, O, ctx11):
    """DESCINIT / CHK1MAT (INFO codes and the descriptor DESCINIT leaves behind on illegal input) against the executed reference source:
    the oracle on every P x Q grid of the golden set, the product on the cases whose grid is the 1 x 1 grid a single process can make."""
    g = _tools_golden()
    n11 = 0
    for row in g["descinit"].tolist():
        P, Q, r, c, m, n, mb, nb, rs, cs, ictxt, lld = row[:12]; want_desc = row[12:21]; want_info = row[21]
        d, info = O.descinit(m, n, mb, nb, rs, cs, ictxt, lld, P, Q, r)
        assert (d, info) == (want_desc, want_info), row
        if (P, Q) == (1, 1):
            d, info = S.descinit(m, n, mb, nb, rs, cs, ctx11, lld)
            assert info == want_info and d[:1] + d[2:] == want_desc[:1] + want_desc[2:], row      # all but the context handle
            n11 += 1
    for row in g["chk1mat"].tolist():
        P, Q, r, c = row[:4]; d = row[4:13]; ma, na, ia, ja, info_in, want = row[13:19]
        assert O.chk1mat(ma, 1, na, 2, ia, ja, d, 6, P, Q, r, c, info=info_in) == want, row
        if (P, Q) == (1, 1):
            d1 = list(d); d1[1] = ctx11
            assert S.chk1mat(ma, 1, na, 2, ia, ja, d1, 6, info=info_in) == want, row
            n11 += 1
    assert n11 >= 20


def test_reference_fortran_executed_live(O):
    """When the reference tree is on this machine: fresh random inputs through the interpreted Fortran (the golden file is not stale)."""
    if not os.path.exists("/root/reference/TOOLS/numroc.f"):
        pytest.skip("no reference tree here")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import fortran77_mini as F
    it = F.load_tools("/root/reference")
    rng = np.random.default_rng(7)
    for _ in range(300):
        np_ = int(rng.integers(1, 9)); nb = int(rng.integers(1, 33)); n = int(rng.integers(0, 5000)); ip = int(rng.integers(0, np_)); isrc = int(rng.integers(0, np_))
        assert it.call("NUMROC", n, nb, ip, isrc, np_)["__result__"] == O.numroc(n, nb, ip, isrc, np_)
        ig = int(rng.integers(1, 5000))
        assert it.call("INDXG2P", ig, nb, ip, isrc, np_)["__result__"] == O.indxg2p(ig, nb, ip, isrc, np_)
        assert it.call("INDXL2G", ig, nb, ip, isrc, np_)["__result__"] == O.indxl2g(ig, nb, ip, isrc, np_)


def test_bench_supplementary_leg_cannot_hold_the_bench_line(monkeypatch):
    """bench.py's supplementary leg (the 8(f) measurements of bench_next.py): skipped when the run has used its time, reported when it
    returns, ABANDONED when it does not come back within its limit -- the bench line is printed either way."""
    import importlib
    import time as _time
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    import bench_next
    rep, th = bench.next_rows_leg(1, 760.0)
    assert th is None and "skipped" in rep["error"]
    rep, th = bench.next_rows_leg(4, 570.0)
    assert th is None and "skipped" in rep["error"]
    calls = {}
    monkeypatch.setattr(bench_next, "run_all", lambda per_row_timeout, total_timeout: calls.update(t=total_timeout) or {"summary": {"ok": 3}})
    rep, th = bench.next_rows_leg(1, 300.0)
    assert rep == {"summary": {"ok": 3}} and calls["t"] == 240.0 and not th.is_alive()
    rep, th = bench.next_rows_leg(2, 500.0)
    assert calls["t"] == 100.0
    monkeypatch.setattr(bench_next, "run_all", lambda per_row_timeout, total_timeout: (_ for _ in ()).throw(RuntimeError("boom")))
    rep, th = bench.next_rows_leg(1, 0.0)
    assert "boom" in rep["error"]
    monkeypatch.setattr(bench_next, "run_all", lambda per_row_timeout, total_timeout: _time.sleep(3600))
    t0 = _time.time()
    rep, th = bench.next_rows_leg(1, 739.0, slack=-40.0)          # limit 41 s, waits 1 s
    assert "abandoned" in rep["error"] and th.is_alive() and _time.time() - t0 < 10.0
