"""Host logic of the SURVEY 8(f) entry points on the CPU: the product's own sources (refine.cu, ...) built against the stub CUDA
runtime of tests/emul, run on 1x1, 1x2, 2x1, 2x2 and 2x3 process grids over the TCP control plane, checked against the oracle.
This covers what a GPU-less machine can: index arithmetic of the simple kernels, argument checks, block loops, the order of
collectives.  The GPU tests (tests/test_gpu_next.py) run the same cases through the real library."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="session")
def emul_lib():
    subprocess.check_call(["make", "-C", EMUL, "-s", "-j8"])
    return os.path.join(EMUL, "libslb_emul.so")


def spawn(P, Q, cases, timeout=150):
    world = P * Q
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   SLB200_PORT_OFFSET="0", OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1", SLB200_EMUL="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "next_worker.py"), json.dumps(dict(P=P, Q=Q, cases=cases))],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            o, e = p.communicate(timeout=timeout)
            assert p.returncode == 0, e[-3000:]
            outs.append(json.loads([ln for ln in o.splitlines() if ln.startswith("RESULT")][0][6:]))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    bad = [(o["rank"], r["case"], r["msgs"]) for o in outs for r in o["results"] if not r["ok"]]
    assert not bad, bad


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (3, 2)])
def test_refinement_family(emul_lib, P, Q):
    spawn(P, Q, "F1_CASES")


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 2), (2, 3)])
def test_redistribution(emul_lib, P, Q):
    spawn(P, Q, "F2_CASES")


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (3, 2)])
def test_cholesky(emul_lib, P, Q):
    spawn(P, Q, "F3_CASES")


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 3), (3, 2)])
def test_inverse(emul_lib, P, Q):
    spawn(P, Q, "F4_CASES")


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 3)])
def test_pblas_entry_points(emul_lib, P, Q):
    spawn(P, Q, "F4B_CASES")


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 1), (2, 2), (2, 3)])
def test_level3_solve(emul_lib, P, Q):
    spawn(P, Q, "F5_CASES")


def test_compiled_c_caller_on_the_emulation(emul_lib, tmp_path):
    """examples/expert_example.c (PDGESVX -> PDGECON -> PDGEMR2D -> PDGETRI -> PDGEMM -> PDPOSV from plain C) linked against the
    emulation library runs its whole flow on the CPU and checks its own residuals."""
    exe = str(tmp_path / "expert_emul")
    cc = subprocess.run(["gcc", "-O1", "-Wall", os.path.join(ROOT, "examples", "expert_example.c"), "-I" + os.path.join(ROOT, "include"), "-L" + EMUL,
                         "-lslb_emul", "-Wl,-rpath," + EMUL, "-lm", "-lstdc++", "-o", exe], capture_output=True, text=True, timeout=120)
    assert cc.returncode == 0, cc.stderr
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    run = subprocess.run([exe], capture_output=True, text=True, timeout=120, env=env)
    assert run.returncode == 0 and "expert example ok" in run.stdout, (run.stdout, run.stderr)


@pytest.mark.parametrize("P,Q", [(1, 1), (2, 2), (1, 4), (4, 1)])
def test_reference_lu_driver_with_est(emul_lib, P, Q):
    """TESTING/traditional/LU.dat (EST = T) on its own process grids: the flow of pdludriver.f through the real entry points."""
    spawn(P, Q, "LUDAT_CASES")


def test_argument_errors_of_the_8f_entry_points(emul_lib):
    """INFO codes of the reference for illegal arguments (SRC/pdgesvx.f:499-573, pdgerfs.f:340-400, pdgecon.f:241-254, pdpotrf.f:181-190,
    pdpotrs.f:193-206, pdgetri.f:231-240): the checks run before any device work, so the emulation exercises exactly the product's code."""
    code = r'''
import sys
sys.path.insert(0, "%(root)s"); sys.path.insert(0, "%(root)s/tests")
import numpy as np
import scalapack_b200.api as api
api._SO = "%(root)s/tests/emul/libslb_emul.so"
import scalapack_b200 as S, oracle as O
ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 1, 1)
n, nb, nrhs = 12, 4, 2
a = np.asfortranarray(O.pdmatgen(n, n, 100)); b = np.asfortranarray(O.pdmatgen(n, nrhs, 200))
da, _ = S.descinit(n, n, nb, nb, 0, 0, ctx, n); db, _ = S.descinit(n, nrhs, nb, 1, 0, 0, ctx, n)
af = np.zeros_like(a); x = np.zeros_like(b); ipiv = np.zeros(n + nb, np.int32); r = np.ones(n); c = np.ones(n); fe = np.zeros(nrhs); be = np.zeros(nrhs)
def svx(fact, trans, equed, r=r, c=c):
    return S.pdgesvx(fact, trans, n, nrhs, a.copy(order="F"), 1, 1, da, af, 1, 1, da, ipiv, equed, r, c, b.copy(order="F"), 1, 1, db, x, 1, 1, db, fe, be)[2]
rb = r.copy(); rb[3] = 0.0
cb = c.copy(); cb[5] = -1.0
out = [svx("X", "N", "N"), svx("N", "X", "N"), svx("N", "N", "N"), svx("F", "N", "Q"), svx("F", "N", "R", r=rb), svx("F", "N", "C", c=cb), svx("F", "N", "B"), svx("F", "T", "B"),
       S.pdgerfs("X", n, nrhs, a, 1, 1, da, af, 1, 1, da, ipiv, b, 1, 1, db, x, 1, 1, db, fe, be), S.pdgecon("1", n, af, 1, 1, da, -1.0)[1],
       S.pdpotrf("X", n, a.copy(order="F"), 1, 1, da), S.pdpotrf("L", 4, a.copy(order="F"), 2, 1, da), S.pdpotrs("Q", n, nrhs, a, 1, 1, da, b.copy(order="F"), 1, 1, db),
       S.pdgetri(4, a.copy(order="F"), 2, 2, da, ipiv), S.pdgetri(n, a.copy(order="F"), 1, 1, da, ipiv, lwork=1), S.pdgetri(n, a.copy(order="F"), 1, 1, da, ipiv, liwork=0)]
print("CODES", out)
''' % dict(root=ROOT)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=env)
    assert run.returncode == 0, run.stderr[-2000:]
    line = [ln for ln in run.stdout.splitlines() if ln.startswith("CODES")][0]
    assert line == "CODES [-1, -2, 0, -13, -14, -15, 0, 0, -1, -7, -1, -4, -1, -4, -8, -10]", line
    assert "On entry to PDGESVX parameter number  13 had an illegal value" in run.stderr       # PXERBLA's text


def test_supplementary_bench_rows_on_the_emulation(emul_lib):
    """bench_next.py (the 8(f) measurements bench.py embeds as `next_rows`) checks every result it times with a size-independent property.
    Its own logic -- layouts, calls, checks -- is exercised here at a small size with host operands on the emulation; on a B200 the same
    functions run on device-resident operands."""
    code = r'''
import sys, json
sys.path.insert(0, "%(root)s")
import scalapack_b200.api as api
api._SO = "%(root)s/tests/emul/libslb_emul.so"
import bench_next as BN
out = {}
for row in BN.ROWS:
    out[row] = BN.run_row(row, n=72, nb=32, device="cpu")
print("ROWS" + json.dumps(out))
''' % dict(root=ROOT)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert run.returncode == 0, run.stderr[-2000:]
    out = json.loads([ln for ln in run.stdout.splitlines() if ln.startswith("ROWS")][0][4:])
    entries = {f"{row}.{k}": v for row, r in out.items() for k, v in r.items() if isinstance(v, dict)}
    assert len(entries) >= 25, sorted(entries)
    bad = {k: v for k, v in entries.items() if not v["ok"]}
    assert not bad, bad
    assert all(r["_kernel_launches"] > 0 for r in out.values())


def test_supplementary_bench_reports_a_failing_row_without_raising(monkeypatch):
    """bench.py must get a result object even when a row's process dies or hangs: run_all reports the error per row."""
    sys.path.insert(0, ROOT)
    import bench_next as BN
    monkeypatch.setattr(BN, "ROWS", ["potrf", "getri"])
    monkeypatch.setattr(BN.sys, "executable", "/bin/false")
    out = BN.run_all(per_row_timeout=5.0, total_timeout=20.0)
    assert out["summary"]["rows_failed"] == ["potrf", "getri"] and out["summary"]["entries"] == 0
    assert all("error" in out[r] for r in ("potrf", "getri"))


@pytest.mark.parametrize("world", [2, 4])
def test_supplementary_bench_rows_on_emulated_grids(emul_lib, world):
    """The same rows on 1 x 2 and 2 x 2 grids (gloo for the checker's own collectives, the emulation for the library): block-cyclic
    operands, results assembled on every rank, every size-independent check green, identical reports on all ranks."""
    code = r'''
import sys, json
sys.path.insert(0, "%(root)s")
import scalapack_b200.api as api
api._SO = "%(root)s/tests/emul/libslb_emul.so"
import bench_next as BN
print("ROWS" + json.dumps(BN.run_row(sys.argv[1], n=70, nb=16, device="cpu")))
''' % dict(root=ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    for k, row in enumerate(["potrf", "refine"] if world == 4 else ["getri", "pblas", "gemr2d", "getrs_l3"]):
        procs = []
        for r in range(world):
            env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port + 50 * k),
                       OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")       # torch's store on MASTER_PORT, the library's control plane on + 23
            env.pop("SLB200_PORT_OFFSET", None)
            procs.append(subprocess.Popen([sys.executable, "-c", code, row], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
        outs = []
        try:
            for p in procs:
                o, e = p.communicate(timeout=300)
                assert p.returncode == 0, e[-2000:]
                outs.append(json.loads([ln for ln in o.splitlines() if ln.startswith("ROWS")][0][4:]))
        finally:
            for p in procs:
                if p.poll() is None:
                    p.kill()
        entries = {k_: v for k_, v in outs[0].items() if isinstance(v, dict)}
        bad = {k_: v for k_, v in entries.items() if not v["ok"]}
        assert entries and not bad, (row, bad)
        assert outs[0]["_grid"] == ("1x2" if world == 2 else "2x2")
        for o in outs[1:]:                                       # every rank assembled the same results and drew the same conclusions
            assert {k_: v["ok"] for k_, v in o.items() if isinstance(v, dict)} == {k_: v["ok"] for k_, v in entries.items()}


def test_info_codes_against_the_executed_reference_source(emul_lib):
    """tests/golden/errors_reference.json: what the reference's OWN source returns as INFO (executed, tests/golden/make_errors_golden.py) for
    894 illegal or unusual argument combinations (single mutations and seeded pairs) of PDGETRF / PDGETRS / PDGESV / PDPOTRF / PDPOTRS / PDPOSV / PDGECON / PDGERFS / PDGESVX /
    PDGETRI / PDGEEQU (tests/error_cases.py: bad scalars, characters, offsets, every descriptor field, workspace sizes, foreign contexts).
    The product returns the same number for every one of them, and crashes on none."""
    code = r'''
import sys, json
sys.path.insert(0, "%(root)s"); sys.path.insert(0, "%(root)s/tests")
import scalapack_b200.api as api
api._SO = "%(root)s/tests/emul/libslb_emul.so"
import scalapack_b200 as S
import error_cases as E
ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 1, 1)
ctx2 = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 1, 1)
g = json.load(open("%(root)s/tests/golden/errors_reference.json"))
bad = []
for i, c in enumerate(g["cases"]):
    ch = {k: ([tuple(x) for x in v] if v and isinstance(v[0], list) else tuple(v)) if isinstance(v, list) else v for k, v in c["changes"].items()}
    print("AT", c["routine"], c["label"], flush=True)
    info = E.product_info(S, ctx, c["routine"], E.apply(c["routine"], ch), ctx2)
    if info != c["info"]:
        bad.append([c["routine"], c["label"], c["info"], info])
    if c["routine"] in ("PDGETRF", "PDGETRS", "PDGESV"):                  # the complex twins: same source up to the type names, same INFO
        z = c["routine"].replace("PD", "PZ")
        print("AT", z, c["label"], flush=True)
        info = E.product_info(S, ctx, z, E.apply(z, ch), ctx2)
        want = c["info"] if c["info"] <= 0 else None                    # a zero pivot (> 0) depends on the data, which differs
        if want is not None and info != want and not (want == 0 and info > 0):
            bad.append([z, c["label"], c["info"], info])
print("REPLAY" + json.dumps([len(g["cases"]), bad]))
''' % dict(root=ROOT)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    last = [ln for ln in run.stdout.splitlines() if ln.startswith("AT")][-1:]
    assert run.returncode == 0, (last, run.stderr[-1500:])                # a crash names the case it happened in
    ncases, bad = json.loads([ln for ln in run.stdout.splitlines() if ln.startswith("REPLAY")][0][6:])
    assert ncases >= 890 and not bad, bad


def test_supplementary_bench_whole_leg_under_two_ranks(emul_lib):
    """bench_next.run_all as bench.py calls it under torchrun: every rank starts its own rank of each row's process group on shifted
    ports (the caller's control plane and store keep theirs), collects the row's report, and both ranks end with the same summary."""
    code = r'''
import sys, json
sys.path.insert(0, "%(root)s")
import bench_next as BN
out = BN.run_all(per_row_timeout=200.0, total_timeout=400.0, n=48, nb=16, rows=["potrf", "gemr2d"],
                 extra_args=["--device", "cpu", "--lib", "%(root)s/tests/emul/libslb_emul.so"])
print("ALL" + json.dumps(out))
''' % dict(root=ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   TORCHELASTIC_RUN_ID="x", OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        env.pop("SLB200_PORT_OFFSET", None)
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            o, e = p.communicate(timeout=500)
            assert p.returncode == 0, e[-2000:]
            outs.append(json.loads([ln for ln in o.splitlines() if ln.startswith("ALL")][0][3:]))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for o in outs:
        assert o["summary"]["rows_failed"] == [] and o["summary"]["grid"] == "1x2", o
        assert o["summary"]["entries"] == o["summary"]["ok"] == 8, o["summary"]
        assert o["potrf"]["_grid"] == "1x2"


def test_info_codes_on_a_process_grid_are_collective(emul_lib):
    """The same 894 argument combinations on a 2 x 2 grid (four processes over the TCP control plane): every call comes back on every
    rank -- nobody returns early and leaves the others in a collective -- with the SAME INFO on all four, as PCHK1MAT / PCHK2MAT guarantee
    in the reference; the grid-independent majority equals the executed reference's value."""
    code = r'''
import sys, json
sys.path.insert(0, "%(root)s"); sys.path.insert(0, "%(root)s/tests")
import scalapack_b200.api as api
api._SO = "%(root)s/tests/emul/libslb_emul.so"
import scalapack_b200 as S
import error_cases as E
ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 2, 2)
ctx2 = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 2, 2)
g = json.load(open("%(root)s/tests/golden/errors_reference.json"))
out = []
for c in g["cases"]:
    ch = {k: ([tuple(x) for x in v] if v and isinstance(v[0], list) else tuple(v)) if isinstance(v, list) else v for k, v in c["changes"].items()}
    print("AT", c["routine"], c["label"], flush=True)
    out.append(E.product_info(S, ctx, c["routine"], E.apply(c["routine"], ch), ctx2))
print("INFOS" + json.dumps(out))
''' % dict(root=ROOT)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(4):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="4", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SLB200_PORT_OFFSET="0",
                   OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            o, e = p.communicate(timeout=300)
            last = [ln for ln in o.splitlines() if ln.startswith("AT")][-1:]
            assert p.returncode == 0, (last, e[-1500:])
            outs.append(json.loads([ln for ln in o.splitlines() if ln.startswith("INFOS")][0][5:]))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "errors_reference.json")))["cases"]
    differ = [(c["routine"], c["label"], [o[i] for o in outs]) for i, c in enumerate(ref) if len({o[i] for o in outs}) != 1]
    assert len(outs[0]) == len(ref) >= 890 and not differ, differ[:5]
    assert sum(1 for i, c in enumerate(ref) if outs[0][i] == c["info"]) >= 0.8 * len(ref)


def test_invalid_pivots_stop_the_run_with_a_message(emul_lib):
    """An IPIV that is not the pivot sequence of PDGETRF for this sub-matrix (here: the pivots of sub-matrix (1, 1) passed with IA = 5)
    indexes outside the row permutation -- silently in the reference; the product stops with a message instead of corrupting memory."""
    code = r'''
import sys
sys.path.insert(0, "%(root)s")
import numpy as np
import scalapack_b200.api as api
api._SO = "%(root)s/tests/emul/libslb_emul.so"
import scalapack_b200 as S
ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 1, 1)
a = np.asfortranarray(np.eye(12) * 3.0); b = np.asfortranarray(np.ones((12, 2)))
da, _ = S.descinit(12, 12, 4, 4, 0, 0, ctx, 12); db, _ = S.descinit(12, 2, 4, 2, 0, 0, ctx, 12)
ipiv = np.arange(1, 17, dtype=np.int32)
assert S.pdgetrs("N", 8, 2, a, 1, 1, da, ipiv, b, 1, 1, db) == 0          # pivots of sub-matrix (1, 1): fine
ipiv[4:12] = np.arange(1, 9)                                                 # rows 5 .. 12 "exchanged" with rows 1 .. 8: not PDGETRF's output for IA = 5
S.pdgetrs("N", 8, 2, a, 5, 5, da, ipiv, b, 5, 1, db)
print("NOT REACHED")
''' % dict(root=ROOT)
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, env=env)
    assert run.returncode != 0 and "NOT REACHED" not in run.stdout
    assert "IPIV: the entry for row 1 of sub(A) is -3" in run.stderr, run.stderr[-800:]
