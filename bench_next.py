"""Supplementary measurement of the SURVEY 8(f) rows (the callers and data formats either side of the LU) on the GPUs of one box.

    python bench_next.py                  # every row, each in its own process; one JSON object on stdout
    python bench_next.py --row potrf      # one row in this process (under torchrun: one rank of the P x Q grid)

bench.py runs this after its own timed region and embeds the result as `next_rows` in its JSON line; a row that fails, crashes or
exceeds its time limit is reported as such and cannot touch the headline number (each row is a separate process with a hard limit).

What a row reports: the wall-clock time (max over the ranks) of ONE call through the reference-facing entry point (PDPOTRF, PDGETRI,
PDGEMR2D, ...) with its operands already resident in HBM in 2D block-cyclic layout (the calls are synchronous: they return after their
last kernel), after one warm-up call of the same size, the rate in the reference's own flop / byte model, and a SIZE-INDEPENDENT CHECK
of the result: the distributed result is assembled on every rank (torch.distributed, not the library) and tested with torch's FP64
(cuBLAS as the checker, never on the measured path): ||A - L L'||, ||A inv(A) - I||, a bit-exact round trip, ||A X - B||, the true
condition number.  Inputs are synthetic (torch's generator, the same seed on every rank).  These rows were written after the round's
GPU budget was spent: this script is how their first hardware numbers get recorded."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
EPS = 2.0 ** -53
ROWS = ["potrf", "getri", "gemr2d", "refine", "pblas", "getrs_l3"]
GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}


class Env:
    """The process grid of this run, seen from one rank."""

    def __init__(self, S, torch, dist, ctx, P, Q, device):
        self.S, self.torch, self.dist, self.ctx, self.P, self.Q, self.device = S, torch, dist, ctx, P, Q, device
        _, _, self.myrow, self.mycol = S.blacs_gridinfo(ctx)

    def sync(self):
        if self.device != "cpu":
            self.torch.cuda.synchronize()

    def barrier(self):
        self.sync()
        if self.dist is not None:
            self.dist.barrier()

    def maxr(self, v):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def timed(self, f):
        """wall clock of one collective call: barrier, call, device idle; the slowest rank counts"""
        self.barrier()
        t0 = time.perf_counter()
        r = f()
        self.sync()
        return self.maxr(time.perf_counter() - t0), r

    def rand(self, m, n, seed):
        g = self.torch.Generator(device=self.device); g.manual_seed(seed)
        return self.torch.rand(m, n, dtype=self.torch.float64, device=self.device, generator=g) * 2.0 - 1.0

    def eye(self, n):
        return self.torch.eye(n, dtype=self.torch.float64, device=self.device)

    def mat(self, m, n, mb, nb, full=None):
        return DMat(self, m, n, mb, nb, full)


class DMat:
    """An m x n matrix in 2D block-cyclic layout (mb x nb blocks, first block on process (0, 0)): .flat is this rank's column-major
    local array (what the library is given), .desc its descriptor, .full() the assembled global matrix on every rank."""

    def __init__(self, env, m, n, mb, nb, full=None):
        torch, S = env.torch, env.S
        self.env, self.m, self.n = env, m, n
        ri = torch.arange(m, device=env.device); ci = torch.arange(n, device=env.device)
        self.rows = ri[(ri // mb) % env.P == env.myrow]                 # my global rows / columns, in local order
        self.cols = ci[(ci // nb) % env.Q == env.mycol]
        self.mloc, self.nloc = int(self.rows.numel()), int(self.cols.numel())
        assert self.mloc == S.numroc(m, mb, env.myrow, 0, env.P) and self.nloc == S.numroc(n, nb, env.mycol, 0, env.Q)
        self.lld = max(1, self.mloc)
        self.desc, info = S.descinit(m, n, mb, nb, 0, 0, env.ctx, self.lld)
        assert info == 0
        self.flat = torch.zeros(max(1, self.lld * self.nloc), dtype=torch.float64, device=env.device)
        if full is not None:
            self.set(full)
        env.sync()     # the library runs on its own non-blocking streams: torch's fill must have finished before a call reads it

    def local(self):
        """(mloc, nloc) view of the local array"""
        return self.flat[:self.lld * self.nloc].view(self.nloc, self.lld).t()[:self.mloc]

    def set(self, full):
        if self.mloc and self.nloc:
            self.local().copy_(full[self.rows][:, self.cols])
        self.env.sync()

    def full(self):
        env, torch = self.env, self.env.torch
        if env.dist is None:
            return self.local().clone()
        g = torch.zeros(self.m, self.n, dtype=torch.float64, device=env.device)
        if self.mloc and self.nloc:
            g[self.rows.unsqueeze(1), self.cols.unsqueeze(0)] = self.local()
        env.dist.all_reduce(g)                                          # every element has one owner: the sum adds zeros to it
        return g


def _norm1(x):
    return float(x.abs().sum(dim=0).max())


def row_potrf(E, n, nb):
    S, torch = E.S, E.torch
    out = {}
    g = E.rand(n, n, 1)
    a0 = g + g.t() + 2.0 * n * E.eye(n)
    del g
    anorm = _norm1(a0)
    for uplo in "LU":
        A = E.mat(n, n, nb, nb, a0)
        assert S.pdpotrf(uplo, n, A.flat, 1, 1, A.desc) == 0                                 # warm-up (workspaces, communicators)
        A.set(a0)
        sec, info = E.timed(lambda: S.pdpotrf(uplo, n, A.flat, 1, 1, A.desc))
        af = A.full()
        f = torch.tril(af) if uplo == "L" else torch.triu(af).t()
        resid = _norm1(f @ f.t() - a0) / (anorm * n * EPS)                                   # pdlltdriver's check: ||A - L L'|| / (||A|| N eps)
        other_ok = bool(torch.equal(torch.triu(af, 1), torch.triu(a0, 1)) if uplo == "L" else torch.equal(torch.tril(af, -1), torch.tril(a0, -1)))
        del af, f
        out["pdpotrf_" + uplo] = {"n": n, "nb": nb, "seconds": sec, "tflops": (n ** 3 / 3.0) / sec / 1e12, "flops_model": "N^3/3", "info": info,
                                  "resid": resid, "other_triangle_untouched": other_ok, "ok": info == 0 and resid < 10.0 and other_ok}
        # PDPOTRS, one right-hand side (HBM-bound sweeps: the triangle is read twice) and 256 (level 3)
        for nrhs in (1, 256):
            b0 = E.rand(n, nrhs, 2)
            B = E.mat(n, nrhs, nb, nb, b0)
            S.pdpotrs(uplo, n, nrhs, A.flat, 1, 1, A.desc, B.flat, 1, 1, B.desc)
            B.set(b0)
            sec, info = E.timed(lambda: S.pdpotrs(uplo, n, nrhs, A.flat, 1, 1, A.desc, B.flat, 1, 1, B.desc))
            x = B.full()
            resid = _norm1(a0 @ x - b0) / (anorm * _norm1(x) * n * EPS)
            r = {"n": n, "nrhs": nrhs, "seconds": sec, "info": info, "resid": resid, "ok": info == 0 and resid < 10.0}
            if nrhs == 1:
                r["gbs"] = 8.0 * n * n / sec / 1e9; r["bytes_model"] = "8 N^2 (the triangle read twice)"
            else:
                r["tflops"] = 2.0 * n * n * nrhs / sec / 1e12; r["flops_model"] = "2 N^2 NRHS"
            out[f"pdpotrs_{uplo}_nrhs{nrhs}"] = r
    return out


def row_getri(E, n, nb):
    import numpy as np
    S = E.S
    a0 = E.rand(n, n, 3)
    A = E.mat(n, n, nb, nb)
    sec = info = None
    for _ in range(2):                                                                       # first pass = warm-up
        A.set(a0)
        ipiv = np.zeros(A.mloc + nb, np.int32)
        assert S.pdgetrf(n, n, A.flat, 1, 1, A.desc, ipiv) == 0
        sec, info = E.timed(lambda: S.pdgetri(n, A.flat, 1, 1, A.desc, ipiv))
    x = A.full()
    an, xn = _norm1(a0), _norm1(x)
    resid = _norm1(a0 @ x - E.eye(n)) / (n * an * xn * EPS)                                   # pdinvdriver's check
    return {"pdgetri": {"n": n, "nb": nb, "seconds": sec, "tflops": (4.0 * n ** 3 / 3.0) / sec / 1e12,
                        "flops_model": "4/3 N^3 (pdinvdriver.f); the level-3 solve of L U X = P executes 2 N^3", "info": info, "resid": resid,
                        "cond1": an * xn, "ok": info == 0 and resid < 10.0}}


def row_gemr2d(E, n, nb):
    # NB = 64 -> nb and back, sub-matrices that start inside blocks on both sides
    S, torch = E.S, E.torch
    big = n + 100
    a0 = E.rand(big, big, 4)
    A = E.mat(big, big, 64, 64, a0)
    B = E.mat(big, big, nb, nb)
    C = E.mat(big, big, 64, 64)
    ia, ja, ib, jb = 3, 70, 41, 5
    S.pdgemr2d(n, n, A.flat, ia, ja, A.desc, B.flat, ib, jb, B.desc, E.ctx)
    B.flat.zero_(); E.sync()
    sec1, _ = E.timed(lambda: S.pdgemr2d(n, n, A.flat, ia, ja, A.desc, B.flat, ib, jb, B.desc, E.ctx))
    sec2, _ = E.timed(lambda: S.pdgemr2d(n, n, B.flat, ib, jb, B.desc, C.flat, ia, ja, A.desc, E.ctx))
    bf = B.full()
    sub = a0[ia - 1:ia - 1 + n, ja - 1:ja - 1 + n]
    moved = bool(torch.equal(bf[ib - 1:ib - 1 + n, jb - 1:jb - 1 + n], sub))
    bf[ib - 1:ib - 1 + n, jb - 1:jb - 1 + n] = 0.0
    outside = not bool(bf.any())                                                             # nothing written outside sub(B)
    del bf
    back = bool(torch.equal(C.full()[ia - 1:ia - 1 + n, ja - 1:ja - 1 + n], sub))
    return {f"pdgemr2d_nb64_to_nb{nb}": {"n": n, "seconds": sec1, "gbs": 16.0 * n * n / sec1 / 1e9,
                                          "bytes_model": "16 N^2 over the whole grid (every element read and written once)",
                                          "bit_exact": moved, "nothing_outside_subB": outside, "ok": moved and outside},
            f"pdgemr2d_nb{nb}_to_nb64": {"n": n, "seconds": sec2, "gbs": 16.0 * n * n / sec2 / 1e9, "round_trip_bit_exact": back, "ok": back}}


def row_refine(E, n, nb):
    import numpy as np
    S, torch = E.S, E.torch
    out = {}
    a0 = E.rand(n, n, 5)
    A = E.mat(n, n, nb, nb, a0)
    # PDLANGE: one pass over A (two for 'F')
    for norm, ref in (("1", _norm1(a0)), ("I", _norm1(a0.t())), ("M", float(a0.abs().max())), ("F", float(torch.linalg.norm(a0)))):
        S.pdlange(norm, n, n, A.flat, 1, 1, A.desc)
        sec, v = E.timed(lambda: S.pdlange(norm, n, n, A.flat, 1, 1, A.desc))
        out["pdlange_" + norm] = {"n": n, "seconds": sec, "gbs": (2 if norm == "F" else 1) * 8.0 * n * n / sec / 1e9,
                                  "passes_over_A": 2 if norm == "F" else 1, "value": v, "rel_err": abs(v - ref) / ref, "ok": abs(v - ref) <= 1e-12 * ref}
    anorm = out["pdlange_1"]["value"]
    AF = E.mat(n, n, nb, nb, a0)
    ipiv = np.zeros(AF.mloc + nb, np.int32)
    assert S.pdgetrf(n, n, AF.flat, 1, 1, AF.desc, ipiv) == 0
    # PDGECON against the true condition number.  Any estimate bounds ||inv(A)|| from below: rcond_est >= rcond_true.  The default
    # returns what the reference's source returns -- the alternating-sign value only, pdlacon.f:188-189, one pair of solves; it drifts
    # away from the true RCOND as N grows (20x at N = 72, 1500x at N = 1000), so only the bound is checked for it; option
    # lacon_keep_estimate = 1 is LAPACK's full estimator (usually within 3x, ~5 pairs of solves)
    inv = torch.linalg.inv(a0)
    true = 1.0 / (anorm * _norm1(inv))
    del inv
    for key, keep, slack in (("pdgecon_1", 0, float("inf")), ("pdgecon_1_lapack_estimator", 1, 10.0)):
        S.set_option("lacon_keep_estimate", keep)
        S.pdgecon("1", n, AF.flat, 1, 1, AF.desc, anorm)
        sec, (rcond, info) = E.timed(lambda: S.pdgecon("1", n, AF.flat, 1, 1, AF.desc, anorm))
        out[key] = {"n": n, "seconds": sec, "rcond": rcond, "rcond_true": true, "over_true": rcond / true, "info": info,
                    "ok": info == 0 and true <= rcond * (1 + 1e-6) and rcond <= slack * true}
    S.set_option("lacon_keep_estimate", 0)
    # PDGERFS: x from PDGETRS, perturbed, refined; BERR must reach rounding level and the true error must respect FERR
    nrhs = 2
    x_true = E.rand(n, nrhs, 6)
    b0 = a0 @ x_true
    B = E.mat(n, nrhs, nb, nb, b0)
    X = E.mat(n, nrhs, nb, nb, b0)
    assert S.pdgetrs("N", n, nrhs, AF.flat, 1, 1, AF.desc, ipiv, X.flat, 1, 1, X.desc) == 0
    X.flat.mul_(1.0 + 1e-7); E.sync()                                                        # something to refine
    ferr, berr = np.zeros(max(1, B.nloc)), np.zeros(max(1, B.nloc))

    def errs(Xm, fe, be):
        """true error per column; FERR / BERR live on the process column that owns the column of B: bring them to every rank"""
        x = Xm.full()
        err = [float((x[:, k] - x_true[:, k]).abs().max() / x[:, k].abs().max()) for k in range(nrhs)]
        fe_, be_ = [E.maxr(float(fe[k]) if B.nloc > k else 0.0) for k in range(nrhs)], [E.maxr(float(be[k]) if B.nloc > k else 0.0) for k in range(nrhs)]
        return err, fe_, be_
    # the reliable error bound is LAPACK's estimator's: it is taken first, and the true error of BOTH runs is held against it (a FERR
    # built on the reference's alternating-sign value can fall short of the true error)
    xstart = X.flat.clone()
    fbound = None
    for key, keep in (("pdgerfs_lapack_estimator", 1), ("pdgerfs", 0)):
        S.set_option("lacon_keep_estimate", keep)
        X.flat.copy_(xstart); E.sync()
        sec, info = E.timed(lambda: S.pdgerfs("N", n, nrhs, A.flat, 1, 1, A.desc, AF.flat, 1, 1, AF.desc, ipiv, B.flat, 1, 1, B.desc,
                                              X.flat, 1, 1, X.desc, ferr, berr))
        err, fe_, be_ = errs(X, ferr, berr)
        fbound = fe_ if keep else fbound
        out[key] = {"n": n, "nrhs": nrhs, "seconds": sec, "info": info, "berr": be_, "ferr": fe_, "true_err": err,
                    "ok": info == 0 and max(be_) <= 4 * (n + 1) * EPS and all(e <= 4 * f + 1e-15 for e, f in zip(err, fbound))
                    and all(f <= fb * (1 + 1e-9) for f, fb in zip(fe_, fbound))}
    S.set_option("lacon_keep_estimate", 0)
    # PDGESVX, FACT = 'E' (equilibrate, factor, estimate, solve, refine in one call)
    A2 = E.mat(n, n, nb, nb, a0); AF2 = E.mat(n, n, nb, nb)
    B2 = E.mat(n, nrhs, nb, nb, b0); X2 = E.mat(n, nrhs, nb, nb)
    r, c = np.zeros(max(1, A2.mloc)), np.zeros(max(1, A2.nloc))
    ip2 = np.zeros(A2.mloc + nb, np.int32)
    ferr[:] = 0.0; berr[:] = 0.0
    sec, (equed, rc2, info) = E.timed(lambda: S.pdgesvx("E", "N", n, nrhs, A2.flat, 1, 1, A2.desc, AF2.flat, 1, 1, AF2.desc, ip2, "N", r, c,
                                                        B2.flat, 1, 1, B2.desc, X2.flat, 1, 1, X2.desc, ferr, berr))
    err, fe_, be_ = errs(X2, ferr, berr)
    out["pdgesvx_E"] = {"n": n, "nrhs": nrhs, "seconds": sec, "info": info, "equed": equed, "rcond": rc2, "true_err": err, "ferr": fe_,
                        "ok": info == 0 and all(e <= 4 * f + 1e-15 for e, f in zip(err, fbound)) and true <= rc2 * (1 + 1e-6)}
    return out


def row_pblas(E, n, nb):
    S, torch = E.S, E.torch
    out = {}
    a0, b0, c0 = E.rand(n, n, 7), E.rand(n, n, 8), E.rand(n, n, 9)
    A, B, C = E.mat(n, n, nb, nb, a0), E.mat(n, n, nb, nb, b0), E.mat(n, n, nb, nb)
    for ta, tb in (("N", "N"), ("T", "N")):
        C.set(c0)
        S.pdgemm(ta, tb, n, n, n, 0.5, A.flat, 1, 1, A.desc, B.flat, 1, 1, B.desc, -1.5, C.flat, 1, 1, C.desc)
        C.set(c0)
        sec, _ = E.timed(lambda: S.pdgemm(ta, tb, n, n, n, 0.5, A.flat, 1, 1, A.desc, B.flat, 1, 1, B.desc, -1.5, C.flat, 1, 1, C.desc))
        ref = 0.5 * ((a0.t() if ta == "T" else a0) @ b0) - 1.5 * c0
        err = float((C.full() - ref).abs().max()) / (n * EPS * float(a0.abs().max()) * float(b0.abs().max()) + EPS)
        del ref
        out[f"pdgemm_{ta}{tb}"] = {"n": n, "seconds": sec, "tflops": 2.0 * n ** 3 / sec / 1e12, "flops_model": "2 M N K, operand redistribution inside the call",
                                   "err_over_n_eps": err, "ok": err < 10.0}
    # PDTRSM, left / lower / non-unit and right / upper / unit, well-conditioned triangles
    t0 = torch.tril(E.rand(n, n, 10)) / n + E.eye(n)
    for side, uplo, diag in (("L", "L", "N"), ("R", "U", "U")):
        tri = t0 if uplo == "L" else t0.t().contiguous()
        A.set(tri); B.set(b0)
        S.pdtrsm(side, uplo, "N", diag, n, n, 2.0, A.flat, 1, 1, A.desc, B.flat, 1, 1, B.desc)
        B.set(b0)
        sec, _ = E.timed(lambda: S.pdtrsm(side, uplo, "N", diag, n, n, 2.0, A.flat, 1, 1, A.desc, B.flat, 1, 1, B.desc))
        eff = tri if diag == "N" else (tri - torch.diag(torch.diag(tri)) + E.eye(n))
        x = B.full()
        back = (eff @ x) if side == "L" else (x @ eff)
        resid = _norm1(back - 2.0 * b0) / (_norm1(eff) * _norm1(x) * n * EPS)
        del back, x
        out[f"pdtrsm_{side}{uplo}N{diag}"] = {"n": n, "seconds": sec, "tflops": float(n) ** 3 / sec / 1e12, "flops_model": "M N^2", "resid": resid, "ok": resid < 10.0}
    A.set(a0); C.set(c0)
    sec, _ = E.timed(lambda: S.pdtran(n, n, 2.0, A.flat, 1, 1, A.desc, 0.5, C.flat, 1, 1, C.desc))
    ok = bool(torch.equal(C.full(), 0.5 * c0 + 2.0 * a0.t()))
    out["pdtran"] = {"n": n, "seconds": sec, "gbs": 24.0 * n * n / sec / 1e9, "bytes_model": "24 N^2 (A read, C read and written)", "bit_exact": ok, "ok": ok}
    return out


def row_getrs_l3(E, n, nb):
    import numpy as np
    S = E.S
    out = {}
    a0 = E.rand(n, n, 11)
    AF = E.mat(n, n, nb, nb, a0)
    ipiv = np.zeros(AF.mloc + nb, np.int32)
    assert S.pdgetrf(n, n, AF.flat, 1, 1, AF.desc, ipiv) == 0
    nrhs = min(1024, n)
    b0 = E.rand(n, nrhs, 12)
    B = E.mat(n, nrhs, nb, nb, b0)
    anorm = _norm1(a0)
    for trans in "NT":
        S.pdgetrs(trans, n, nrhs, AF.flat, 1, 1, AF.desc, ipiv, B.flat, 1, 1, B.desc)       # > solve_l3_min_nrhs: the level-3 path
        B.set(b0)
        sec, info = E.timed(lambda: S.pdgetrs(trans, n, nrhs, AF.flat, 1, 1, AF.desc, ipiv, B.flat, 1, 1, B.desc))
        x = B.full()
        resid = _norm1((a0 if trans == "N" else a0.t()) @ x - b0) / (anorm * _norm1(x) * n * EPS)
        out[f"pdgetrs_{trans}_nrhs{nrhs}"] = {"n": n, "nrhs": nrhs, "seconds": sec, "tflops": 2.0 * n * n * nrhs / sec / 1e12, "flops_model": "2 N^2 NRHS",
                                              "info": info, "resid": resid, "ok": info == 0 and resid < 10.0}
        B.set(b0)
    return out


RUN = {"potrf": row_potrf, "getri": row_getri, "gemr2d": row_gemr2d, "refine": row_refine, "pblas": row_pblas, "getrs_l3": row_getrs_l3}
DEFAULT_N = {"potrf": 16384, "getri": 16384, "gemr2d": 16384, "refine": 8192, "pblas": 8192, "getrs_l3": 16384}


def run_row(row, n=None, nb=512, device="cuda", backend=None):
    """One row on the grid of this run (WORLD_SIZE ranks: 1 x 1, 1 x 2, 2 x 2 or 2 x 4); returns {entry: {...}} (the same on every rank).
    device='cpu' (with backend='gloo' when WORLD_SIZE > 1) exists for the CPU test of this script's own logic."""
    import torch
    import scalapack_b200 as S
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    if device != "cpu":
        assert S.has_cuda(), "the product library needs a B200 (there is no CPU fallback)"
        torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        kw = {"device_id": torch.device("cuda", local)} if device != "cpu" else {}
        dist_.init_process_group(backend or ("gloo" if device == "cpu" else "nccl"), **kw)
        dist = dist_
    P, Q = GRIDS[world]
    ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", P, Q)
    E = Env(S, torch, dist, ctx, P, Q, device)
    S.reset_counters()
    res = RUN[row](E, n or DEFAULT_N[row], nb)
    res["_kernel_launches"] = int(S.get_counter("kernel_launches"))
    res["_grid"] = f"{P}x{Q}"
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    return res


def _finite(o):
    """non-finite numbers as strings: the report is embedded in bench.py's JSON line, which must stay strict JSON"""
    if isinstance(o, float) and (o != o or o in (float("inf"), float("-inf"))):
        return repr(o)
    if isinstance(o, dict):
        return {k: _finite(v) for k, v in o.items()}
    if isinstance(o, list):
        return [_finite(v) for v in o]
    return o


def run_all(per_row_timeout=60.0, total_timeout=240.0, n=None, nb=512, port_shift=300, rows=None, extra_args=()):
    """Every row in its own process (a fault in one cannot poison the CUDA context of the next, nor of the caller).  Under torchrun
    every rank calls this: rank r starts rank r of each row's process group, which meets on MASTER_PORT + port_shift."""
    out, t_start = {}, time.time()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rows = list(rows or ROWS)
    for k, row in enumerate(rows):
        left = total_timeout - (time.time() - t_start)
        if left < 10.0:
            out[row] = {"error": "skipped: the time limit of the supplementary rows was used up"}
            continue
        cmd = [sys.executable, os.path.abspath(__file__), "--row", row, "--nb", str(nb)] + (["--n", str(n)] if n else []) + list(extra_args)
        env = dict(os.environ)
        if world > 1:                                                                         # ports of its own, new ones for every row
            env["MASTER_PORT"] = str(int(os.environ.get("MASTER_PORT", "29500")) + port_shift + 40 * k)
            env["MASTER_ADDR"] = os.environ.get("MASTER_ADDR", "127.0.0.1")
        else:
            env.update(WORLD_SIZE="1", RANK="0")
        for v in ("TORCHELASTIC_RUN_ID", "TORCHELASTIC_USE_AGENT_STORE", "TORCHELASTIC_RESTART_COUNT", "TORCHELASTIC_MAX_RESTARTS"):
            env.pop(v, None)                                                                  # the children rendezvous on their own TCP store
        try:
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=min(per_row_timeout, left), env=env)
            lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            out[row] = json.loads(lines[-1]) if (p.returncode == 0 and lines) else {"error": f"rc {p.returncode}: " + (p.stderr or p.stdout)[-400:]}
        except subprocess.TimeoutExpired:
            out[row] = {"error": f"no result within {min(per_row_timeout, left):.0f} s"}
        except Exception as ex:  # noqa: BLE001
            out[row] = {"error": repr(ex)[:300]}
    flat = [v for r in out.values() if "error" not in r for k, v in r.items() if isinstance(v, dict)]
    out["summary"] = {"entries": len(flat), "ok": sum(1 for v in flat if v.get("ok")), "rows_failed": [r for r in rows if "error" in out[r]],
                      "timing": "wall clock (max over ranks) of one synchronous call, operands resident in HBM, after one warm-up call",
                      "grid": "%dx%d" % GRIDS.get(world, (0, 0))}
    return _finite(out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--row", default="", choices=[""] + ROWS)
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--nb", type=int, default=512)
    ap.add_argument("--device", default="cuda", help=argparse.SUPPRESS)       # 'cpu' + --lib: the CPU test of this script's own logic
    ap.add_argument("--lib", default="", help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.lib:
        import scalapack_b200.api as api
        api._SO = args.lib
    if args.row:
        print(json.dumps(run_row(args.row, args.n or None, args.nb, device=args.device)), flush=True)
    else:
        print(json.dumps(run_all(n=args.n or None, nb=args.nb)), flush=True)


if __name__ == "__main__":
    main()
