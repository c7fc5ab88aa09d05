"""Randomised differential run of the LU path itself (PDGETRF / PDGETRS, real and complex, sub-matrix operands, non-zero source
processes, the pipelined and host-streamed schedules) on emulated process grids: seeded random valid argument sets for tests/mp_worker.py,
the product's orchestration (lu.cu, solve.cu, api.cu compiled unchanged into tests/emul) against the oracle.  Not part of the pytest suite.
    python scripts/fuzz_lu_path.py [--grids 1x1,2x2,2x3] [--count 40] [--seed 1]"""
import argparse
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)


def gen(rng, P, Q):
    nb = rng.choice([1, 2, 3, 4, 8, 16, 32])
    if rng.random() < 0.5:                                       # sub-matrix operand, source process anywhere
        k, l = rng.randint(0, 3), rng.randint(0, 3)
        m, n = rng.randint(1, 70), rng.randint(1, 70)
        if rng.random() < 0.5:
            n = m
        return dict(P=P, Q=Q, mg=k * nb + m + rng.randint(0, 9), ng=l * nb + n + rng.randint(0, 9), nb=nb, ia=k * nb + 1, ja=l * nb + 1, m=m, n=n,
                    rsrc=rng.randint(0, P - 1), csrc=rng.randint(0, Q - 1), nrhs=rng.randint(0, 4))
    m = rng.randint(1, 120)
    n = m if rng.random() < 0.6 else rng.randint(1, 120)
    cs = dict(P=P, Q=Q, m=m, n=n, nb=nb, nrhs=rng.randint(0, 3), z=rng.random() < 0.25)
    if rng.random() < 0.4:
        cs["split"] = nb * rng.randint(1, 4)
        cs["hoststream"] = rng.random() < 0.5
    return cs


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--grids", default="1x1,1x2,2x1,2x2,2x3,3x2")
    ap.add_argument("--count", type=int, default=40)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    import test_emul_lu as T
    T.emul_lib.__wrapped__() if hasattr(T.emul_lib, "__wrapped__") else None
    total = 0
    for grid in args.grids.split(","):
        P, Q = (int(v) for v in grid.split("x"))
        rng = random.Random(args.seed * 100 + P * 10 + Q)
        cases = [gen(rng, P, Q) for _ in range(args.count)]
        try:
            T.spawn(P * Q, cases, timeout=1500)
            print(f"{grid}: {len(cases)} cases ok", flush=True)
        except AssertionError as ex:
            total += 1
            print(f"{grid}: FAILED", str(ex)[:2500], flush=True)
    # host-resident callers on one process: the column slabs of A arrive late by a varying number of polls, with and without the budget
    # that keeps factored panels for the replay (tests/test_emul_lu.py::test_host_streaming_with_late_slabs, random shapes here)
    for delay in (0, 1, 3, 7):
        for save_mb in (16384, 0):
            rng = random.Random(args.seed * 1000 + delay * 10 + (save_mb > 0))
            cases = []
            for _ in range(args.count // 4):
                nb = rng.choice([4, 8, 16, 32])
                m = rng.randint(1, 300)
                n = m if rng.random() < 0.5 else rng.randint(1, 300)
                cs = dict(P=1, Q=1, m=m, n=n, nb=nb, nrhs=rng.randint(0, 2), z=rng.random() < 0.25, hoststream=True)
                if rng.random() < 0.6:
                    cs["split"] = nb * rng.randint(2, 4)
                cases.append(cs)
            try:
                T.spawn(1, cases, timeout=1500, extra_env={"SLB200_EMUL_SLAB_DELAY": str(delay), "SLB200_E2E_SAVE_MB": str(save_mb), "SLB200_E2E_SLAB_MB": "0"})
                print(f"late slabs delay={delay} save_mb={save_mb}: {len(cases)} cases ok", flush=True)
            except AssertionError as ex:
                total += 1
                print(f"late slabs delay={delay} save_mb={save_mb}: FAILED", str(ex)[:2000], flush=True)
    sys.exit(1 if total else 0)
