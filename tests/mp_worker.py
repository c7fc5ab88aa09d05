"""One rank of a multi-GPU parity run (spawned by tests/test_gpu_multi.py or scripts/run_mp.py).
Every rank rebuilds the global test matrix with the oracle, factors its block-cyclic piece through the
C-ABI (PDGETRF / PDGETRS / PDGESV on a P x Q grid, one GPU per rank) and compares its local result with the
oracle's serial factorisation scattered the same way.  Prints one RESULT json line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import oracle as O  # noqa: E402
if os.environ.get("SLB200_EMUL") == "1":      # host-logic emulation (tests/emul, CPU): the real lu.cu / api.cu over contract kernels
    import scalapack_b200.api as _api  # noqa: E402
    _api._SO = os.path.join(ROOT, "tests", "emul", "libslb_emul.so")
import scalapack_b200 as S  # noqa: E402

EPS = 2.0 ** -53


def run_case(ctx, P, Q, m, n, nb, nrhs, cplx=False, device=False, split=0, hoststream=False):
    # split > 0: force the near | far column pipeline (and look-ahead overlap) at this small size
    S.set_option("la_split_min", split if split else 6144)
    S.set_option("lookahead_min_us", 0 if split else 4000)
    # hoststream: a host-resident caller takes the streaming path (block rows written back during the factorisation) at this size
    S.set_option("e2e_overlap_min_mb", 0 if hoststream else 256)
    _, _, r, c = S.blacs_gridinfo(ctx)
    res = {"case": f"{P}x{Q} m={m} n={n} nb={nb} nrhs={nrhs} z={int(cplx)} dev={int(device)} split={split} hoststream={int(hoststream)}", "ok": True, "msgs": []}
    if r < 0:
        return res
    gen = O.pzmatgen if cplx else O.pdmatgen
    a0 = gen(m, n, 100)
    ref = a0.copy(order="F")
    ipr, infr = O.getrf(ref, nb)
    mloc, nloc = S.numroc(m, nb, r, 0, P), S.numroc(n, nb, c, 0, Q)
    lld = max(1, mloc) + 1                                     # one guard row
    al = O.scatter(a0, nb, nb, P, Q, r, c, lld=lld)
    al[mloc:, :] = -9923.0
    desca, info = S.descinit(m, n, nb, nb, 0, 0, ctx, lld)
    ipiv = np.full(mloc + nb, -77, np.int32)
    f = S.pzgetrf if cplx else S.pdgetrf
    if device:
        import torch
        torch.cuda.set_device(S.device())                       # the GPU this BLACS process drives (LOCAL_RANK), not cuda:0
        t = torch.from_numpy(np.ascontiguousarray(al.T)).cuda()
        info = f(m, n, t, 1, 1, desca, ipiv)
        al = np.asfortranarray(t.cpu().numpy().T)
    else:
        info = f(m, n, al, 1, 1, desca, ipiv)
    print(f"[rank {r},{c}] factor done info={info}", file=sys.stderr, flush=True)
    refl = O.scatter(ref, nb, nb, P, Q, r, c, lld=lld)
    mn = min(m, n)
    ipl = O.ipiv_local(m, mn, nb, P, r, ipr, mloc + nb, fill=-77)
    own = ipl != -77
    if info != infr:
        res["ok"] = False; res["msgs"].append(f"info {info} != {infr}")
    if not np.array_equal(ipiv[own], ipl[own]):
        bad = np.nonzero(ipiv[own] != ipl[own])[0]
        res["ok"] = False; res["msgs"].append(f"ipiv mismatch at local idx {bad[:5].tolist()} got {ipiv[own][bad[:5]].tolist()} want {ipl[own][bad[:5]].tolist()}")
    anorm = np.abs(a0).sum(axis=1).max()
    err = float(np.abs(al[:mloc, :nloc] - refl[:mloc, :nloc]).max() / (anorm * max(m, n) * EPS)) if mloc and nloc else 0.0
    res["lu_err"] = err
    if not err < 1.0:
        res["ok"] = False; res["msgs"].append(f"lu_err {err}")
    if not np.all(al[mloc:, :].real == -9923.0):
        res["ok"] = False; res["msgs"].append("guard row overwritten")
    if m == n and nrhs > 0 and res["ok"]:
        b0 = gen(n, nrhs, 200)
        nbr = 2
        nlocb = S.numroc(nrhs, nbr, c, 0, Q)
        bl = O.scatter(b0, nb, nbr, P, Q, r, c, lld=max(1, mloc))
        descb, _ = S.descinit(n, nrhs, nb, nbr, 0, 0, ctx, max(1, mloc))
        g = S.pzgetrs if cplx else S.pdgetrs
        inf2 = g("N", n, nrhs, al, 1, 1, desca, ipiv, bl, 1, 1, descb)
        xr = b0.copy(order="F"); O.getrs(ref, ipr, xr)
        xl = O.scatter(xr, nb, nbr, P, Q, r, c, lld=max(1, mloc))
        if inf2 != 0:
            res["ok"] = False; res["msgs"].append(f"getrs info {inf2}")
        if mloc and nlocb:
            e2 = float(np.abs(bl[:mloc, :nlocb] - xl[:mloc, :nlocb]).max() / max(1e-300, np.abs(xr).max()))
            res["x_err"] = e2
            if not e2 < 1e-8:
                res["ok"] = False; res["msgs"].append(f"x_err {e2}")
        for trans in (("T", "C") if cplx else ("T",)):           # transposed solves (pdgetrs.f:268-284)
            blt = O.scatter(b0, nb, nbr, P, Q, r, c, lld=max(1, mloc))
            inft = g(trans, n, nrhs, al, 1, 1, desca, ipiv, blt, 1, 1, descb)
            xt = b0.copy(order="F"); O.getrs(ref, ipr, xt, trans)
            xtl = O.scatter(xt, nb, nbr, P, Q, r, c, lld=max(1, mloc))
            if inft != 0:
                res["ok"] = False; res["msgs"].append(f"getrs {trans} info {inft}")
            if mloc and nlocb:
                e3 = float(np.abs(blt[:mloc, :nlocb] - xtl[:mloc, :nlocb]).max() / max(1e-300, np.abs(xt).max()))
                if not e3 < 1e-8:
                    res["ok"] = False; res["msgs"].append(f"x_err[{trans}] {e3}")
        # PDGESV in one call on fresh copies
        al2 = O.scatter(a0, nb, nb, P, Q, r, c, lld=lld); bl2 = O.scatter(b0, nb, nbr, P, Q, r, c, lld=max(1, mloc))
        ip2 = np.zeros(mloc + nb, np.int32)
        h = S.pzgesv if cplx else S.pdgesv
        inf3 = h(n, nrhs, al2, 1, 1, desca, ip2, bl2, 1, 1, descb)
        if inf3 != 0 or (mloc and nlocb and not np.allclose(bl2[:mloc, :nlocb], bl[:mloc, :nlocb], rtol=0, atol=1e-300)):
            res["ok"] = False; res["msgs"].append(f"pdgesv differs from pdgetrf+pdgetrs (info {inf3})")
    return res


def run_case_general(ctx, P, Q, cs):
    """Sub-matrix operands on a grid with non-zero source processes: sub(A) = A(ia:ia+m-1, ja:ja+n-1) of an mg x ng matrix
    distributed from (rsrc, csrc); PDGETRF, then PDGETRS with TRANS in N/T on sub(B) rows aligned with sub(A)."""
    mg, ng, nb, ia, ja, m, n = cs["mg"], cs["ng"], cs["nb"], cs["ia"], cs["ja"], cs["m"], cs["n"]
    rsrc, csrc, nrhs = cs.get("rsrc", 0), cs.get("csrc", 0), cs.get("nrhs", 2)
    _, _, r, c = S.blacs_gridinfo(ctx)
    res = {"case": f"{P}x{Q} general {cs}", "ok": True, "msgs": []}
    if r < 0:
        return res
    def fail(msg):
        res["ok"] = False; res["msgs"].append(msg)
    ag = O.pdmatgen(mg, ng, 100)
    sub0 = np.asfortranarray(ag[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n])
    ref = sub0.copy(order="F"); ipr, infr = O.getrf(ref, nb)
    G = ag.copy(order="F"); G[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n] = ref
    mloc, nloc = S.numroc(mg, nb, r, rsrc, P), S.numroc(ng, nb, c, csrc, Q)
    lld = max(1, mloc) + 1
    al = O.scatter(ag, nb, nb, P, Q, r, c, rsrc=rsrc, csrc=csrc, lld=lld); al[mloc:, :] = -9923.0
    exp = O.scatter(G, nb, nb, P, Q, r, c, rsrc=rsrc, csrc=csrc, lld=lld); exp[mloc:, :] = -9923.0
    desca, info = S.descinit(mg, ng, nb, nb, rsrc, csrc, ctx, lld)
    ipiv = np.full(mloc + nb, -77, np.int32)
    info = S.pdgetrf(m, n, al, ia, ja, desca, ipiv)
    if info != infr:
        fail(f"info {info} != {infr}")
    ipe = np.full(mloc + nb, -77, np.int32)
    for i in range(min(m, n)):
        gi = ia + i                                              # 1-based global row of A
        if S.indxg2p(gi, nb, r, rsrc, P) == r:
            ipe[S.indxg2l(gi, nb, r, rsrc, P) - 1] = ipr[i] + ia - 1
    if not np.array_equal(ipiv, ipe):
        bad = np.nonzero(ipiv != ipe)[0]
        fail(f"ipiv mismatch at local idx {bad[:5].tolist()} got {ipiv[bad[:5]].tolist()} want {ipe[bad[:5]].tolist()}")
    anorm = np.abs(sub0).sum(axis=1).max()
    err = float(np.abs(al - exp).max() / (anorm * max(m, n) * EPS))
    res["lu_err"] = err
    if not err < 1.0:
        fail(f"lu_err {err}")
    # outside sub(A): untouched bits.  Mask = positions whose global index lies inside the window.
    Mk = np.zeros((mg, ng), order="F"); Mk[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n] = 1.0
    mk = O.scatter(Mk, nb, nb, P, Q, r, c, rsrc=rsrc, csrc=csrc, lld=lld)
    if not np.array_equal(al[mk == 0.0], exp[mk == 0.0]):
        fail("something outside sub(A) was written")
    if m == n and nrhs > 0 and res["ok"]:
        nbb, csrcb = 2, (csrc + 1) % Q
        bg = O.pdmatgen(mg, nrhs, 200)
        b0 = np.asfortranarray(bg[ia - 1:ia - 1 + n, :])
        descb, _ = S.descinit(mg, nrhs, nb, nbb, rsrc, csrcb, ctx, max(1, mloc))
        for trans in ("N", "T"):
            bl = O.scatter(bg, nb, nbb, P, Q, r, c, rsrc=rsrc, csrc=csrcb, lld=max(1, mloc))
            inf2 = S.pdgetrs(trans, n, nrhs, al, ia, ja, desca, ipiv, bl, ia, 1, descb)
            xr = b0.copy(order="F"); O.getrs(ref, ipr, xr, trans)
            XG = bg.copy(order="F"); XG[ia - 1:ia - 1 + n, :] = xr
            xl = O.scatter(XG, nb, nbb, P, Q, r, c, rsrc=rsrc, csrc=csrcb, lld=max(1, mloc))
            nlocb = S.numroc(nrhs, nbb, c, csrcb, Q)
            if inf2 != 0:
                fail(f"getrs {trans} info {inf2}")
            if mloc and nlocb:
                e2 = float(np.abs(bl[:mloc, :nlocb] - xl[:mloc, :nlocb]).max() / max(1e-300, np.abs(xr).max()))
                if not e2 < 1e-8:
                    fail(f"x_err[{trans}] {e2}")
    return res


def main():
    me, np_ = S.blacs_pinfo()
    # every rank runs the oracle's BLAS: share the host cores (8 ranks x all-core OpenBLAS pools spin against each other)
    O.set_threads(max(1, (os.cpu_count() or 1) // max(1, np_)))
    cases = json.loads(sys.argv[1])
    out = []
    grids = {}                      # one BLACS grid (and one set of NCCL communicators) per (P, Q)
    for cs in cases:
        if cs["P"] * cs["Q"] > np_:
            continue
        key = (cs["P"], cs["Q"])
        if key not in grids:
            grids[key] = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", cs["P"], cs["Q"])
        print(f"[rank {me}] case {cs}", file=sys.stderr, flush=True)
        if "ia" in cs:
            out.append(run_case_general(grids[key], cs["P"], cs["Q"], cs))
        else:
            out.append(run_case(grids[key], cs["P"], cs["Q"], cs["m"], cs["n"], cs["nb"], cs.get("nrhs", 1), cs.get("z", False), cs.get("dev", False), cs.get("split", 0), cs.get("hoststream", False)))
    S.blacs_exit(0)
    print("RESULT" + json.dumps({"rank": me, "results": out}), flush=True)


if __name__ == "__main__":
    main()
