"""Randomised differential run of the SURVEY 8(f) entry points on emulated process grids: seeded random VALID argument sets (orders, block
sizes, offsets, source processes, transpositions, numbers of right-hand sides) for the case functions of tests/next_cases.py, product
(host-logic emulation, tests/emul) against the oracle.  The hand-written case lists of the test suite pick the shapes their author thought
of; this walks the rest of the argument space.  Not part of the pytest suite.
    python scripts/fuzz_next_rows.py [--grids 1x1,2x2,2x3] [--count 40] [--seed 1] [--kinds potrf,getri]"""
import argparse
import json
import os
import random
import socket
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gen(kind, rng):
    nb = rng.choice([1, 2, 3, 4, 5, 8, 16])
    n = rng.randint(1, 60)
    if kind == "lange":
        mg, ng = rng.randint(1, 60), rng.randint(1, 60)
        ia, ja = rng.randint(1, mg), rng.randint(1, ng)
        return dict(kind=kind, mg=mg, ng=ng, nb=nb, ia=ia, ja=ja, m=rng.randint(1, mg - ia + 1), n=rng.randint(1, ng - ja + 1), rsrc=rng.randint(0, 3), csrc=rng.randint(0, 3))
    if kind == "equ":
        return dict(kind=kind, n=n, m=rng.randint(1, 60), nb=nb, cond=rng.choice([None, 1, 6]))
    if kind == "gecon":
        return dict(kind=kind, n=max(n, 2), nb=nb, cond=rng.choice([None, 1, 2]), lapack_estimator=rng.random() < 0.5)
    if kind == "gerfs":
        return dict(kind=kind, n=n, nb=nb, nrhs=rng.randint(1, 6), trans=rng.choice("NT"), cond=rng.choice([None, 1]), nbr=rng.randint(1, 3), lapack_estimator=rng.random() < 0.3)
    if kind == "gesvx":
        return dict(kind=kind, n=n, nb=nb, nrhs=rng.randint(1, 4), fact=rng.choice("NE"), trans=rng.choice("NT"), cond=rng.choice([None, 2, 5]), lapack_estimator=rng.random() < 0.3)
    if kind == "gemr2d":
        m, n2 = rng.randint(1, 50), rng.randint(1, 50)
        ia, ja, ib, jb = (rng.randint(1, 9) for _ in range(4))
        return dict(kind=kind, m=m, n=n2, ia=ia, ja=ja, ib=ib, jb=jb, shape_a=(m + ia + rng.randint(0, 5), n2 + ja + rng.randint(0, 5)),
                    shape_b=(m + ib + rng.randint(0, 5), n2 + jb + rng.randint(0, 5)), blk_a=(rng.randint(1, 9), rng.randint(1, 9)), blk_b=(rng.randint(1, 20), rng.randint(1, 20)),
                    src_a=(rng.randint(0, 3), rng.randint(0, 3)), src_b=(rng.randint(0, 3), rng.randint(0, 3)), z=rng.random() < 0.2,
                    **({"ga": rng.choice([(1, 1), (1, 2), (2, 1), (2, 2), (1, 3), (3, 1)])} if rng.random() < 0.3 else {}),
                    **({"gb": rng.choice([(1, 1), (1, 2), (2, 1), (2, 2), (1, 3), (3, 1)])} if rng.random() < 0.3 else {}))
    if kind == "potrf":
        return dict(kind=kind, n=n, nb=nb, uplo=rng.choice("LU"), nrhs=rng.choice([1, 2, 5, 70]), off=rng.randint(0, 2), rsrc=rng.randint(0, 3), csrc=rng.randint(0, 3),
                    notpd=(rng.randint(0, n - 1) if rng.random() < 0.15 else None))
    if kind == "getri":
        return dict(kind=kind, n=n, nb=nb, dominant=rng.random() < 0.5, off=rng.randint(0, 2), rsrc=rng.randint(0, 3), csrc=rng.randint(0, 3),
                    singular=(rng.randint(0, n - 1) if rng.random() < 0.15 else None))
    blk = lambda: (rng.randint(1, 9), rng.randint(1, 9))  # noqa: E731
    ij = lambda: (rng.randint(1, 6), rng.randint(1, 6))   # noqa: E731
    src = lambda: (rng.randint(0, 3), rng.randint(0, 3))  # noqa: E731
    if kind == "pdgemm":
        return dict(kind=kind, m=rng.randint(1, 40), n=rng.randint(1, 40), k=rng.randint(0, 40), ta=rng.choice("NT"), tb=rng.choice("NT"), alpha=rng.choice([1.0, -0.5, 0.0]),
                    beta=rng.choice([0.0, 1.0, 2.0]), ija=ij(), ijb=ij(), ijc=ij(), blk_a=blk(), blk_b=blk(), blk_c=blk(), src_a=src(), src_b=src(), src_c=src())
    if kind == "pdtrsm":
        return dict(kind=kind, m=rng.randint(1, 40), n=rng.randint(1, 40), side=rng.choice("LR"), uplo=rng.choice("LU"), ta=rng.choice("NT"), diag=rng.choice("NU"),
                    alpha=rng.choice([1.0, 0.5, 0.0]), ija=ij(), ijb=ij(), blk_a=blk(), blk_b=blk())
    if kind == "pdtran":
        return dict(kind=kind, m=rng.randint(1, 40), n=rng.randint(1, 40), alpha=rng.choice([1.0, 2.0, 0.0]), beta=rng.choice([0.0, 0.5, 1.0]), blk_a=blk(), blk_c=blk())
    if kind == "getrs_l3":
        return dict(kind=kind, n=n, nb=nb, nrhs=rng.choice([1, 3, 17, 70]), trans=rng.choice("NT"), nbb=rng.randint(1, 9), off=rng.randint(0, 2), rsrc=rng.randint(0, 3), csrc=rng.randint(0, 3),
                    entry=rng.random() < 0.5)
    if kind == "ludriver":
        return dict(kind=kind, n=n, nb=nb, nrhs=rng.randint(1, 4), nbrhs=rng.randint(1, 3))
    raise KeyError(kind)


KINDS = ["lange", "equ", "gecon", "gerfs", "gesvx", "gemr2d", "potrf", "getri", "pdgemm", "pdtrsm", "pdtran", "getrs_l3", "ludriver"]


def run_grid(P, Q, cases, timeout=1200):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(P * Q):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(P * Q), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SLB200_PORT_OFFSET="0",
                   SLB200_EMUL="1", OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "next_worker.py"), json.dumps(dict(P=P, Q=Q, cases=cases))], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    bad = []
    try:
        for r, p in enumerate(procs):
            o, e = p.communicate(timeout=timeout)
            line = [ln for ln in o.splitlines() if ln.startswith("RESULT")]
            if p.returncode != 0 or not line:
                bad.append((r, "process", f"rc {p.returncode}: " + e[-600:]))
                continue
            bad += [(r, x["case"], x["msgs"]) for x in json.loads(line[0][6:])["results"] if not x["ok"]]
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    return bad


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="1x1,2x2,2x3,3x2")
    ap.add_argument("--count", type=int, default=30)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--kinds", default=",".join(KINDS))
    args = ap.parse_args()
    total = 0
    for kind in args.kinds.split(","):
        rng = random.Random(args.seed * 1000 + sum(map(ord, kind)))
        cases = [gen(kind, rng) for _ in range(args.count)]
        for grid in args.grids.split(","):
            P, Q = (int(v) for v in grid.split("x"))
            t0 = time.time()
            bad = run_grid(P, Q, cases)
            total += len(bad)
            print(f"{kind} {grid}: {len(cases)} cases, {len(bad)} failures, {time.time() - t0:.0f} s", flush=True)
            for b in bad[:6]:
                print("    ", str(b)[:700])
    sys.exit(1 if total else 0)
