"""GPU parity of the SURVEY 8(f) rows through the C-ABI: PDLANGE / PDGEEQU / PDLAQGE / PDGECON / PDGERFS / PDGESVX (row 1), ...
against the oracle (tests/next_cases.py).  Each group runs in its own process (one rank per GPU), like tests/test_gpu_multi.py."""
import json
import os
import socket
import subprocess
import sys

import pytest

# These rows were written after the round's GPU budget was spent: the cases below have passed on the CPU under the host-logic
# emulation (tests/test_emul_next.py) but this file has not run on hardware yet.  Until it has, a failure here must not mask the
# hardware-validated suite that runs before it (pytest -x): non-strict xfail reports a pass as XPASS and a failure as XFAIL.
pytestmark = [pytest.mark.gpu, pytest.mark.xfail(reason="SURVEY 8(f) rows: first hardware run", strict=False)]
_HUNG = []          # a case that hung once is likely to hang again: later groups give up immediately instead of burning GPU time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import next_cases  # noqa: E402


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def spawn(P, Q, cases, timeout=400):
    if _HUNG:
        pytest.fail(f"skipped after the hang of {_HUNG[0]}")
    world = P * Q
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   SLB200_PORT_OFFSET="0", OPENBLAS_NUM_THREADS="2", OMP_NUM_THREADS="2")
        env.pop("SLB200_EMUL", None)
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "next_worker.py"), json.dumps(dict(P=P, Q=Q, cases=cases))],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs, tails = [], []
    try:
        for p in procs:
            try:
                o, e = p.communicate(timeout=timeout)
            except subprocess.TimeoutExpired:
                _HUNG.append(f"{P}x{Q} {str(cases[-1])[:80]}")
                raise
            tails.append(e[-3000:])
            assert p.returncode == 0, e[-3000:]
            outs.append(json.loads([ln for ln in o.splitlines() if ln.startswith("RESULT")][0][6:]))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    bad = [(o["rank"], r["case"], r["msgs"]) for o in outs for r in o["results"] if not r["ok"]]
    assert not bad, bad


# larger cases than the emulation's: the 1x1 solves take the two-stream fast path (nb <= 512), the matrix passes fill the GPU
F1_GPU = [
    dict(kind="lange", mg=3000, ng=2500, nb=128, ia=130, ja=7, m=2000, n=2200),
    dict(kind="equ", n=2000, m=2300, nb=128, cond=6),
    dict(kind="gecon", n=2048, nb=256), dict(kind="gecon", n=1500, nb=64, cond=2), dict(kind="gecon", n=1000, nb=1024), dict(kind="gecon", n=1536, nb=128, dev=True),
    dict(kind="gerfs", n=2048, nb=256, nrhs=2), dict(kind="gerfs", n=1500, nb=64, nrhs=2, trans="T", cond=1),
    dict(kind="gesvx", n=2048, nb=256, fact="N"), dict(kind="gesvx", n=1500, nb=128, fact="E", cond=4),
    dict(kind="gesvx", n=1000, nb=64, fact="E", cond=4, trans="T"),
    dict(kind="gecon", n=2048, nb=256, lapack_estimator=True), dict(kind="gerfs", n=1500, nb=64, nrhs=2, lapack_estimator=True),
    dict(kind="gesvx", n=1500, nb=128, fact="E", cond=4, lapack_estimator=True),
]


@pytest.mark.parametrize("P,Q", [(1, 2), (2, 1)])
def test_refinement_family_2gpus(P, Q):
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    spawn(P, Q, next_cases.F1_CASES + F1_GPU)


def test_refinement_family_2x2():
    if ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    spawn(2, 2, next_cases.F1_CASES + F1_GPU)


# ---- row 2: PDGEMR2D ----
F2_GPU = [
    dict(kind="gemr2d", m=4096, n=4096, shape_a=(4096, 4096), shape_b=(4096, 4096), blk_a=(64, 64), blk_b=(512, 512)),     # NB 64 -> 512
    dict(kind="gemr2d", m=3000, n=2000, ia=65, ja=130, ib=7, jb=3, shape_a=(3100, 2200), shape_b=(3010, 2005), blk_a=(64, 32), blk_b=(100, 256),
         src_a=(1, 1), src_b=(0, 1)),
    dict(kind="gemr2d", m=1500, n=1500, shape_a=(1500, 1500), shape_b=(1500, 1500), blk_a=(128, 128), blk_b=(32, 32), z=True),
    dict(kind="gemr2d", m=2000, n=1800, ia=3, ja=2, shape_a=(2100, 1900), shape_b=(2000, 1800), blk_a=(64, 64), blk_b=(256, 256), dev=True),
    dict(kind="gemr2d", m=2048, n=2048, shape_a=(2048, 2048), shape_b=(2048, 2048), blk_a=(64, 64), blk_b=(512, 512), ga=(1, 2), gb=(2, 1)),
    dict(kind="gemr2d", m=2048, n=2048, shape_a=(2048, 2048), shape_b=(2048, 2048), blk_a=(64, 64), blk_b=(512, 512), ga=(2, 2), gb=(1, 1)),
]


@pytest.mark.parametrize("P,Q", [(1, 2), (2, 2)])
def test_redistribution_multi(P, Q):
    if ngpus() < P * Q:
        pytest.skip(f"needs {P * Q} GPUs")
    spawn(P, Q, next_cases.F2_CASES + F2_GPU)


# ---- row 3: PDPOTRF / PDPOTRS / PDPOSV ----
F3_GPU = [
    dict(kind="potrf", n=2048, nb=256, uplo="L"), dict(kind="potrf", n=2048, nb=256, uplo="U"),
    dict(kind="potrf", n=2500, nb=512, uplo="L", nrhs=1), dict(kind="potrf", n=1500, nb=64, uplo="U", nrhs=3),
    dict(kind="potrf", n=1000, nb=128, uplo="L", notpd=700), dict(kind="potrf", n=1536, nb=256, uplo="L", nrhs=200), dict(kind="potrf", n=1024, nb=128, uplo="U", nrhs=100), dict(kind="potrf", n=1536, nb=256, uplo="L", dev=True), dict(kind="potrf", n=1100, nb=128, uplo="U", dev=True), dict(kind="potrf", n=1200, nb=128, uplo="U", off=2, rsrc=1, csrc=1),
]


@pytest.mark.parametrize("P,Q", [(1, 2), (2, 1), (2, 2)])
def test_cholesky_multi(P, Q):
    if ngpus() < P * Q:
        pytest.skip(f"needs {P * Q} GPUs")
    spawn(P, Q, next_cases.F3_CASES + F3_GPU)


# ---- row 4: PDGETRI ----
F4_GPU = [
    dict(kind="getri", n=2048, nb=256), dict(kind="getri", n=2048, nb=128, dominant=True), dict(kind="getri", n=1500, nb=64, cond=1), dict(kind="getri", n=2200, nb=512),
    dict(kind="getri", n=1200, nb=128, off=2, rsrc=1, csrc=1), dict(kind="getri", n=1000, nb=128, singular=900), dict(kind="getri", n=1536, nb=256, dev=True),
]


@pytest.mark.parametrize("P,Q", [(1, 2), (2, 1), (2, 2)])
def test_inverse_multi(P, Q):
    if ngpus() < P * Q:
        pytest.skip(f"needs {P * Q} GPUs")
    spawn(P, Q, next_cases.F4_CASES + F4_GPU)


# ---- row 4, second half: PDGEMM / PDTRSM / PDTRAN entry points ----
F4B_GPU = [
    dict(kind="pdgemm", m=2000, n=1500, k=1800, ta="N", tb="N", alpha=1.5, beta=-0.5, ija=(3, 2), ijb=(2, 6), ijc=(4, 3), blk_a=(64, 48), blk_b=(128, 128), blk_c=(256, 256)),
    dict(kind="pdgemm", m=1024, n=1024, k=1024, ta="T", tb="N", blk_a=(128, 128), blk_b=(128, 128), blk_c=(128, 128)),
    dict(kind="pdgemm", m=1000, n=1100, k=900, ta="N", tb="T", beta=0.0, blk_a=(100, 100), blk_b=(64, 64), blk_c=(512, 512)),
    dict(kind="pdtrsm", m=2048, n=512, side="L", uplo="L", ta="N", diag="U", blk_a=(256, 256), blk_b=(256, 256)),
    dict(kind="pdtrsm", m=1500, n=700, side="L", uplo="U", ta="N", diag="N", alpha=0.5, ija=(2, 3), ijb=(3, 2), blk_a=(128, 128), blk_b=(64, 64)),
    dict(kind="pdtrsm", m=600, n=1500, side="R", uplo="L", ta="T", diag="N", blk_a=(128, 128), blk_b=(128, 128)),
    dict(kind="pdtrsm", m=900, n=1000, side="R", uplo="U", ta="N", diag="U", blk_a=(200, 200), blk_b=(512, 512)),
    dict(kind="pdtran", m=1500, n=2000, alpha=2.0, beta=0.5, blk_a=(64, 64), blk_c=(256, 128)),
    dict(kind="pdgemm", m=1024, n=768, k=512, ta="N", tb="N", alpha=-1.0, beta=1.0, blk_a=(128, 128), blk_b=(128, 128), blk_c=(128, 128), dev=True),
]


@pytest.mark.parametrize("P,Q", [(1, 2), (2, 2)])
def test_pblas_entry_points_multi(P, Q):
    if ngpus() < P * Q:
        pytest.skip(f"needs {P * Q} GPUs")
    spawn(P, Q, next_cases.F4B_CASES + F4B_GPU)


# ---- PDGETRS with many right-hand sides (level-3 path; `entry`: through pdgetrs_ itself, which switches above 64) ----
F5_GPU = [
    dict(kind="getrs_l3", n=2048, nb=256, nrhs=300), dict(kind="getrs_l3", n=2048, nb=256, nrhs=300, trans="T"),
    dict(kind="getrs_l3", n=1500, nb=128, nrhs=100, nbb=32, entry=True), dict(kind="getrs_l3", n=1500, nb=128, nrhs=100, nbb=32, entry=True, trans="T"),
    dict(kind="getrs_l3", n=1024, nb=128, nrhs=65, off=2, rsrc=1, csrc=1, entry=True),
]


@pytest.mark.parametrize("P,Q", [(1, 2), (2, 2)])
def test_level3_solve_multi(P, Q):
    if ngpus() < P * Q:
        pytest.skip(f"needs {P * Q} GPUs")
    spawn(P, Q, next_cases.F5_CASES + F5_GPU)


# ---- the reference's own LU test driver with EST = T on the LU.dat grid (PDGETRF -> PDGECON -> PDGETRS -> PDGERFS, guard zones) ----
LUD_GPU = [dict(kind="ludriver", n=1000, nb=64, nrhs=3, nbrhs=2), dict(kind="ludriver", n=2048, nb=256, nrhs=1, nbrhs=1),
           dict(kind="ludriver", n=1000, nb=64, nrhs=3, nbrhs=2, lapack_estimator=True)]


# ---- one GPU: one process per ENTRY POINT, so that a fault in one routine (a sticky CUDA error ends every later case of its process)
# cannot hide what the others do; the report then says which routines passed on hardware (XPASS) and which did not (XFAIL) ----
ONE_GPU = {}
for _cs in (next_cases.F1_CASES + F1_GPU + next_cases.F2_CASES + F2_GPU + next_cases.F3_CASES + F3_GPU + next_cases.F4_CASES + F4_GPU
            + next_cases.F4B_CASES + F4B_GPU + next_cases.F5_CASES + F5_GPU + next_cases.LUDAT_CASES + LUD_GPU):
    _key = _cs["kind"] + ("_" + _cs["uplo"] if _cs["kind"] == "potrf" else "") + ("_lapack_estimator" if _cs.get("lapack_estimator") else "")
    ONE_GPU.setdefault(_key, []).append(_cs)


@pytest.mark.parametrize("entry", sorted(ONE_GPU))
def test_one_gpu(entry):
    spawn(1, 1, ONE_GPU[entry])


@pytest.mark.parametrize("P,Q", [(2, 2), (1, 4), (4, 1)])
def test_reference_lu_driver_with_est_multi(P, Q):
    if ngpus() < P * Q:
        pytest.skip(f"needs {P * Q} GPUs")
    spawn(P, Q, next_cases.LUDAT_CASES)
