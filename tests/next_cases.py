"""Parity cases of the SURVEY 8(f) rows, written once against the reference's interface and run three ways:
on the GPU through libscalapack_b200.so (tests/test_gpu_next.py, multi-GPU via tests/mp_worker.py) and, for the host logic
only, on the CPU through the emulation library (tests/emul; tests/test_emul_next.py).  Every rank of a P x Q grid builds the
global test matrix, takes its block-cyclic piece, calls the routine and compares its local result with the oracle's."""
import numpy as np

import oracle as O

EPS = 2.0 ** -53


class Grid:
    def __init__(self, S, ctx):
        self.S, self.ctx = S, ctx
        self.P, self.Q, self.r, self.c = S.blacs_gridinfo(ctx)

    def dist(self, ag, nb, rsrc=0, csrc=0, extra=0, nbc=None):
        """(local array with `extra` guard rows, descriptor) of the global matrix ag"""
        S = self.S
        nbc = nb if nbc is None else nbc
        m, n = ag.shape
        mloc = S.numroc(m, nb, self.r, rsrc, self.P)
        lld = max(1, mloc) + extra
        al = O.scatter(ag, nb, nbc, self.P, self.Q, self.r, self.c, rsrc=rsrc, csrc=csrc, lld=lld)
        if extra:
            al[mloc:, :] = -9923.0
        desc, info = S.descinit(m, n, nb, nbc, rsrc, csrc, self.ctx, lld)
        assert info == 0
        return al, desc

    def local_of(self, ag, nb, rsrc=0, csrc=0, lld=None, nbc=None):
        return O.scatter(ag, nb, nb if nbc is None else nbc, self.P, self.Q, self.r, self.c, rsrc=rsrc, csrc=csrc, lld=lld)

    def rows_of(self, v, nb, rsrc=0):
        """local entries (LOCr) of a vector aligned with the rows of a distributed matrix"""
        return self.local_of(np.asfortranarray(v.reshape(-1, 1)), nb, rsrc=rsrc, nbc=1)[:, 0] if self.c == 0 or True else None


class OnDevice:
    """Device-resident operands (GPU runs only, case key dev=True): the local arrays go to the process's GPU as torch tensors with the
    same column-major memory, the routine works on them in place, back() copies them into the numpy arrays the checks look at."""

    def __init__(self, S, enabled):
        self.S, self.enabled, self.pairs = S, bool(enabled), []

    def __call__(self, al):
        if not self.enabled or al is None:
            return al
        import torch
        torch.cuda.set_device(self.S.device())
        t = torch.from_numpy(np.ascontiguousarray(al.T)).cuda()
        self.pairs.append((al, t))
        return t

    def back(self):
        for al, t in self.pairs:
            al[...] = np.asfortranarray(t.cpu().numpy().T)


def matrix(n, m=None, seed=100, cond=None):
    a = O.pdmatgen(n, m or n, seed)
    if cond:
        rng = np.random.default_rng(seed)
        a = (10.0 ** rng.uniform(-cond, cond, (n, 1))) * a * (10.0 ** rng.uniform(-cond, cond, (1, m or n)))
    return np.asfortranarray(a)


def _close(msgs, what, got, want, rtol, atol=0.0):
    got, want = np.asarray(got, dtype=float), np.asarray(want, dtype=float)
    if got.shape != want.shape or not np.allclose(got, want, rtol=rtol, atol=atol):
        err = np.abs(got - want).max() if got.shape == want.shape and got.size else None
        msgs.append(f"{what}: max abs diff {err} (rtol {rtol})")


def case_lange(G, cs):
    """PDLANGE on a general (not block-aligned) sub-matrix"""
    S, msgs = G.S, []
    mg, ng, nb, ia, ja, m, n = cs["mg"], cs["ng"], cs["nb"], cs["ia"], cs["ja"], cs["m"], cs["n"]
    rsrc, csrc = cs.get("rsrc", 0) % G.P, cs.get("csrc", 0) % G.Q
    ag = matrix(mg, ng)
    al, desc = G.dist(ag, nb, rsrc, csrc, extra=1)
    sub = np.asfortranarray(ag[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n])
    for nm in ("M", "1", "O", "I", "F", "E"):
        got, want = S.pdlange(nm, m, n, al, ia, ja, desc), O.dlange(nm, sub)
        if not abs(got - want) <= 1e-12 * max(1.0, want):       # sums in another order than the serial oracle
            msgs.append(f"pdlange {nm}: {got} != {want}")
    return msgs


def case_equ(G, cs):
    """PDGEEQU + PDLAQGE"""
    S, msgs = G.S, []
    n, m, nb, cond = cs["n"], cs.get("m", cs["n"]), cs["nb"], cs.get("cond")
    ag = matrix(m, n, cond=cond)
    if cs.get("zero_row") is not None:
        ag[cs["zero_row"], :] = 0.0
    al, desc = G.dist(ag, nb, extra=1)
    mloc, nloc = S.numroc(m, nb, G.r, 0, G.P), S.numroc(n, nb, G.c, 0, G.Q)
    r, c = np.full(max(1, mloc), -1.0), np.full(max(1, nloc), -1.0)
    rowcnd, colcnd, amax, info = S.pdgeequ(m, n, al, 1, 1, desc, r, c)
    r0, c0, rowcnd0, colcnd0, amax0, info0 = O.dgeequ(ag)
    if cs.get("zero_row") is not None:
        if info != info0:
            msgs.append(f"pdgeequ info {info} != {info0}")
        return msgs
    if info != info0 or not np.allclose([rowcnd, colcnd, amax], [rowcnd0, colcnd0, amax0], rtol=1e-14):
        msgs.append(f"pdgeequ scalars {(rowcnd, colcnd, amax, info)} != {(rowcnd0, colcnd0, amax0, info0)}")
    rl = G.local_of(np.asfortranarray(r0.reshape(-1, 1)), nb, nbc=1)[:mloc, 0] if G.c == 0 else None
    rl = O.scatter(np.asfortranarray(r0.reshape(-1, 1)), nb, 1, G.P, 1, G.r, 0)[:mloc, 0]
    cl = O.scatter(np.asfortranarray(c0.reshape(1, -1)), 1, nb, 1, G.Q, 0, G.c)[0, :nloc]
    _close(msgs, "R", r[:mloc], rl, 1e-15); _close(msgs, "C", c[:nloc], cl, 1e-15)
    eq = S.pdlaqge(m, n, al, 1, 1, desc, r, c, rowcnd, colcnd, amax)
    a2 = ag.copy(order="F")
    eq0 = O.dlaqge(a2, r0, c0, rowcnd0, colcnd0, amax0)
    if eq != eq0:
        msgs.append(f"pdlaqge equed {eq} != {eq0}")
    exp = G.local_of(a2, nb, lld=al.shape[0])
    _close(msgs, "scaled A", al[:mloc, :nloc], exp[:mloc, :nloc], 1e-15)
    if not np.all(al[mloc:, :] == -9923.0):
        msgs.append("guard row overwritten")
    return msgs


def _factored(G, ag, nb, rsrc=0, csrc=0):
    """oracle factors of ag distributed over the grid: (local LU, descriptor, local IPIV, global LU, global ipiv)"""
    S = G.S
    lu = ag.copy(order="F"); ipg, info = O.getrf(lu, nb)
    assert info == 0
    n = ag.shape[0]
    ll, desc = G.dist(lu, nb, rsrc, csrc)
    mloc = S.numroc(n, nb, G.r, rsrc, G.P)
    ipl = O.ipiv_local(n, n, nb, G.P, G.r, ipg, mloc + nb, rsrc=rsrc, fill=-77)
    return ll, desc, ipl, lu, ipg


def case_gecon(G, cs):
    """PDGECON on the oracle's factors"""
    S, msgs = G.S, []
    n, nb = cs["n"], cs["nb"]
    ag = matrix(n, cond=cs.get("cond"))
    ll, desc, ipl, lu, ipg = _factored(G, ag, nb)
    dev = OnDevice(S, cs.get("dev")); ll_host = ll; ll = dev(ll)
    for nm in ("1", "I"):
        anorm = O.dlange(nm, ag)
        rc, info = S.pdgecon(nm, n, ll, 1, 1, desc, anorm)
        want = O.dgecon(nm, lu, anorm)
        # the solves inside the estimator round differently on the GPU (blocked order): the estimate moves by ~cond * eps
        if info != 0 or not abs(rc - want) <= 1e-6 * want:
            msgs.append(f"pdgecon {nm}: rcond {rc} info {info}, oracle {want}")
    if cs.get("unaligned"):
        # the factors as a sub-matrix that starts inside a block (the reference's PDTRSV accepts that): same estimate
        oi, oj = cs["unaligned"]
        big = matrix(n + oi + 2, n + oj + 1, seed=9); big[oi:oi + n, oj:oj + n] = lu
        bl_, descb_ = G.dist(np.asfortranarray(big), nb, 1 % G.P, 1 % G.Q)
        anorm = O.dlange("1", ag)
        rc, info = S.pdgecon("1", n, bl_, oi + 1, oj + 1, descb_, anorm)
        want = O.dgecon("1", lu, anorm)
        if info != 0 or not abs(rc - want) <= 1e-6 * want:
            msgs.append(f"pdgecon on an unaligned sub-matrix: rcond {rc} info {info}, oracle {want}")
    rc, info = S.pdgecon("X", n, ll, 1, 1, desc, 1.0)
    if info != -1:
        msgs.append(f"pdgecon bad NORM: info {info}")
    rc, info = S.pdgecon("1", n, ll, 1, 1, desc, 1.0, lwork=1)
    if info != -10:
        msgs.append(f"pdgecon short LWORK: info {info}")
    return msgs


def case_gerfs(G, cs):
    """PDGERFS: a perturbed solution is refined; X, FERR, BERR against the oracle"""
    S, msgs = G.S, []
    n, nb, nrhs, trans = cs["n"], cs["nb"], cs.get("nrhs", 2), cs.get("trans", "N")
    nbr = cs.get("nbr", 1)
    ag = matrix(n, cond=cs.get("cond")); bg = matrix(n, nrhs, seed=200)
    ll, desc, ipl, lu, ipg = _factored(G, ag, nb)
    al, desca = G.dist(ag, nb)
    xg = bg.copy(order="F"); O.getrs(lu, ipg, xg, trans)
    xg *= 1.0 + 1e-8 * np.sin(np.arange(n))[:, None]
    xg = np.asfortranarray(xg)
    bl, descb = G.dist(bg, nb, nbc=nbr); xl, descx = G.dist(xg, nb, extra=1, nbc=nbr)
    nlocb = S.numroc(nrhs, nbr, G.c, 0, G.Q); mloc = S.numroc(n, nb, G.r, 0, G.P)
    ferr, berr = np.full(max(1, nlocb), -1.0), np.full(max(1, nlocb), -1.0)
    info = S.pdgerfs(trans, n, nrhs, al, 1, 1, desca, ll, 1, 1, desc, ipl, bl, 1, 1, descb, xl, 1, 1, descx, ferr, berr)
    xstart = xg.copy(order="F")
    ferr0, berr0 = O.dgerfs(trans, ag, lu, ipg, bg, xg)
    # the RELIABLE error bound is LAPACK's: the reference's PDLACON returns its alternating-sign value only (pdlacon.f:188-189), and a
    # FERR built on it can fall an order of magnitude short of the true error
    O.lacon_keep_est(True); fbound, _ = O.dgerfs(trans, ag, lu, ipg, bg, xstart); O.lacon_keep_est(bool(cs.get("lapack_estimator")))
    if info != 0:
        msgs.append(f"pdgerfs info {info}")
    xe = G.local_of(xg, nb, lld=xl.shape[0], nbc=nbr)
    # both refined solutions are within the bound of the truth, so they are within twice the bound of each other
    _close(msgs, "X", xl[:mloc, :nlocb], xe[:mloc, :nlocb], 0.0, atol=max(1e-12, 4.0 * fbound.max()) * np.abs(xg).max())
    fl = O.scatter(np.asfortranarray(ferr0.reshape(1, -1)), 1, nbr, 1, G.Q, 0, G.c)[0, :nlocb]
    blc = O.scatter(np.asfortranarray(berr0.reshape(1, -1)), 1, nbr, 1, G.Q, 0, G.c)[0, :nlocb]
    # FERR is built on |r| + (n + 1) eps (|A||x| + |b|): for tiny n the rounding noise of the residual r is not negligible against the second term
    _close(msgs, "FERR", ferr[:nlocb], fl, max(0.05, min(0.6, 4.0 / max(n, 1))), atol=1e-14)
    if n > 1 and nlocb and not (np.all(berr[:nlocb] >= 0) and np.all(berr[:nlocb] < 1e-14) and np.all(blc < 1e-14)):
        msgs.append(f"BERR {berr[:nlocb]} vs {blc}")
    if not np.all(xl[mloc:, :] == -9923.0):
        msgs.append("guard row of X overwritten")
    xt = np.linalg.solve(ag if trans == "N" else ag.T, bg)
    for k in range(nrhs if n > 1 else 0):                       # FERR bounds the true error (N <= 1: quick return, pdgerfs.f:457-463)
        if not np.abs(xg[:, k] - xt[:, k]).max() / np.abs(xt[:, k]).max() <= fbound[k] * 1.001:
            msgs.append(f"oracle FERR[{k}] (LAPACK's estimator) is not a bound")
        if not ferr0[k] <= fbound[k] * (1 + 1e-9):
            msgs.append(f"FERR[{k}] {ferr0[k]} above LAPACK's {fbound[k]}")
    return msgs


def case_gesvx(G, cs):
    """PDGESVX: FACT = N / E (then F on the result), TRANS = N / T"""
    S, msgs = G.S, []
    n, nb, nrhs, fact, trans = cs["n"], cs["nb"], cs.get("nrhs", 2), cs.get("fact", "E"), cs.get("trans", "N")
    ag = matrix(n, cond=cs.get("cond")); bg = matrix(n, nrhs, seed=200)
    if cs.get("singular"):
        ag[:, 3] = ag[:, 2]
    al, desca = G.dist(ag, nb, extra=1); bl, descb = G.dist(bg, nb, nbc=1)
    mloc, nloc, nlocb = S.numroc(n, nb, G.r, 0, G.P), S.numroc(n, nb, G.c, 0, G.Q), S.numroc(nrhs, 1, G.c, 0, G.Q)
    afl = np.zeros_like(al); xl = np.zeros_like(bl)
    ipiv = np.full(mloc + nb, -77, np.int32)
    r, c = np.zeros(max(1, mloc)), np.zeros(max(1, nloc))
    ferr, berr = np.full(max(1, nlocb), -1.0), np.full(max(1, nlocb), -1.0)
    eq, rcond, info = S.pdgesvx(fact, trans, n, nrhs, al, 1, 1, desca, afl, 1, 1, desca, ipiv, "N", r, c, bl, 1, 1, descb, xl, 1, 1,
                                descb, ferr, berr)
    a1, b1 = ag.copy(order="F"), bg.copy(order="F")
    af0, x0 = np.zeros((n, n), order="F"), np.zeros((n, nrhs), order="F")
    ip0, r0, c0 = np.zeros(n, np.int32), np.zeros(n), np.zeros(n)
    eq0, rcond0, ferr0, berr0, info0 = O.dgesvx(fact, trans, a1, af0, ip0, "N", r0, c0, b1, x0, nb=nb)
    O.lacon_keep_est(True)                                          # the reliable error bound: LAPACK's estimator (see case_gerfs)
    fbound = O.dgesvx(fact, trans, ag.copy(order="F"), np.zeros((n, n), order="F"), np.zeros(n, np.int32), "N", np.zeros(n), np.zeros(n),
                      bg.copy(order="F"), np.zeros((n, nrhs), order="F"), nb=nb)[2]
    O.lacon_keep_est(bool(cs.get("lapack_estimator")))
    if cs.get("singular"):
        # two equal columns: the last pivot is rounding noise, so is RCOND (a few 1e-17 with the reference's alternating-sign estimate) and
        # with it the side of eps it falls on.  Each implementation must be consistent with its OWN estimate (pdgesvx.f:738-741).
        for who, rc, inf in (("product", rcond, info), ("oracle", rcond0, info0)):
            if not (rc < 1e-13 and (0 < inf <= n and rc == 0.0 or inf == (n + 1 if rc < EPS else 0))):
                msgs.append(f"singular matrix, {who}: rcond {rc}, info {inf}")
        return msgs
    if (eq, info) != (eq0, info0):
        msgs.append(f"pdgesvx equed/info {(eq, info)} != {(eq0, info0)}")
        return msgs
    if info0 != 0:
        if info0 <= n and rcond != 0.0:
            msgs.append(f"singular: rcond {rcond}")
        return msgs
    if not abs(rcond - rcond0) <= 1e-6 * rcond0:
        msgs.append(f"rcond {rcond} != {rcond0}")
    ipl = O.ipiv_local(n, n, nb, G.P, G.r, ip0, mloc + nb, fill=-77)
    own = ipl != -77
    if not np.array_equal(ipiv[own], ipl[own]):
        msgs.append("IPIV differs")
    _close(msgs, "A (equilibrated)", al[:mloc, :nloc], G.local_of(a1, nb)[:mloc, :nloc], 1e-15)
    _close(msgs, "B (scaled)", bl[:mloc, :nlocb], G.local_of(b1, nb, nbc=1)[:mloc, :nlocb], 1e-15)
    anorm = np.abs(a1).sum(axis=1).max()
    lerr = np.abs(afl[:mloc, :nloc] - G.local_of(af0, nb)[:mloc, :nloc]).max() / (anorm * n * EPS) if mloc and nloc else 0.0
    if not lerr < 1.0:
        msgs.append(f"AF lu_err {lerr}")
    _close(msgs, "X", xl[:mloc, :nlocb], G.local_of(x0, nb, nbc=1)[:mloc, :nlocb], 0.0, atol=max(1e-12, 4.0 * fbound.max()) * np.abs(x0).max())   # both within the bound of the truth
    _close(msgs, "FERR", ferr[:nlocb], O.scatter(np.asfortranarray(ferr0.reshape(1, -1)), 1, 1, 1, G.Q, 0, G.c)[0, :nlocb], max(0.1, min(0.6, 4.0 / max(n, 1))), atol=1e-14)   # see case_gerfs
    if not np.all(al[mloc:, :] == -9923.0):
        msgs.append("guard row of A overwritten")
    # FACT = 'F': the factors and scalings just returned reproduce X
    bl2 = G.local_of(bg, nb, nbc=1); xl2 = np.zeros_like(bl2)
    eq2, rcond2, info2 = S.pdgesvx("F", trans, n, nrhs, al, 1, 1, desca, afl, 1, 1, desca, ipiv, eq, r, c, bl2, 1, 1, descb, xl2, 1, 1,
                                   descb, ferr, berr)
    if info2 != 0 or eq2 != eq or not abs(rcond2 - rcond) <= 1e-10 * rcond:
        msgs.append(f"FACT=F: {(eq2, rcond2, info2)}")
    _close(msgs, "X (FACT=F)", xl2[:mloc, :nlocb], xl[:mloc, :nlocb], 0.0, atol=max(1e-12, 4.0 * fbound.max()) * np.abs(x0).max())
    return msgs


def _subgrid(G, Pn, Qn, cache={}):
    """a context on the first Pn*Qn processes (every process of the main grid calls; outsiders get -1)"""
    key = (id(G.S), G.ctx, Pn, Qn)
    if key not in cache:
        cache[key] = G.S.blacs_gridinit(G.S.blacs_get(-1, 0), "Row-major", Pn, Qn)
    return cache[key]


def case_gemr2d(G, cs):
    """PDGEMR2D / PZGEMR2D: sub(A) on grid ga (block mba x nba, source (rsa, csa)) -> sub(B) on grid gb (any other layout)"""
    S, msgs = G.S, []
    z = cs.get("z", False)
    (Pa, Qa), (Pb, Qb) = cs.get("ga", (G.P, G.Q)), cs.get("gb", (G.P, G.Q))
    if Pa * Qa > G.P * G.Q or Pb * Qb > G.P * G.Q:
        return msgs
    ca, cb = _subgrid(G, Pa, Qa), _subgrid(G, Pb, Qb)
    m, n, ia, ja, ib, jb = cs["m"], cs["n"], cs.get("ia", 1), cs.get("ja", 1), cs.get("ib", 1), cs.get("jb", 1)
    (mga, nga), (mgb, ngb) = cs["shape_a"], cs["shape_b"]
    (mba, nba), (mbb, nbb) = cs["blk_a"], cs["blk_b"]
    (rsa, csa), (rsb, csb) = cs.get("src_a", (0, 0)), cs.get("src_b", (0, 0))
    rsa, csa, rsb, csb = rsa % Pa, csa % Qa, rsb % Pb, csb % Qb
    gen = O.pzmatgen if z else O.pdmatgen
    ag = np.asfortranarray(gen(mga, nga, 100))
    bg = np.full((mgb, ngb), -9923.0, dtype=ag.dtype, order="F")
    want = bg.copy(order="F"); want[ib - 1:ib - 1 + m, jb - 1:jb - 1 + n] = ag[ia - 1:ia - 1 + m, ja - 1:ja - 1 + n]

    def local(ctx, Pn, Qn, glob, mb, nb, rs, cs_):
        _, _, r, c = S.blacs_gridinfo(ctx) if ctx >= 0 else (0, 0, -1, -1)
        if r < 0:
            return None, [1, -1, glob.shape[0], glob.shape[1], mb, nb, rs, cs_, 1], None
        lld = max(1, S.numroc(glob.shape[0], mb, r, rs, Pn)) + 1
        al = O.scatter(glob, mb, nb, Pn, Qn, r, c, rsrc=rs, csrc=cs_, lld=lld)
        al[lld - 1:, :] = -555.0
        desc, info = S.descinit(glob.shape[0], glob.shape[1], mb, nb, rs, cs_, ctx, lld)
        assert info == 0
        return al, desc, (r, c)
    al, desca, _ = local(ca, Pa, Qa, ag, mba, nba, rsa, csa)
    bl, descb, rc = local(cb, Pb, Qb, bg, mbb, nbb, rsb, csb)
    a_before = None if al is None else al.copy()
    f = S.pzgemr2d if z else S.pdgemr2d
    S.set_option("redist_chunk_mb", cs.get("chunk_mb", 1024))      # 0: one column per exchange (the chunked walk of the public entry)
    dev = OnDevice(S, cs.get("dev") and not z)
    f(m, n, dev(al) if al is not None else np.zeros(1, dtype=ag.dtype), ia, ja, desca, dev(bl) if bl is not None else np.zeros(1, dtype=ag.dtype), ib, jb,
      descb, G.ctx)
    dev.back()
    if al is not None and not np.array_equal(al, a_before):
        msgs.append("A was modified")
    if bl is not None:
        exp = O.scatter(want, mbb, nbb, Pb, Qb, rc[0], rc[1], rsrc=rsb, csrc=csb, lld=bl.shape[0])
        exp[bl.shape[0] - 1:, :] = -555.0
        if not np.array_equal(bl, exp):
            bad = np.argwhere(bl != exp)
            msgs.append(f"B differs at {len(bad)} local entries, first {bad[0].tolist()}: got {bl[tuple(bad[0])]} want {exp[tuple(bad[0])]}")
    return msgs


def spd(n, seed=100):
    """symmetric positive definite test matrix (the reference's PDMATGEN 'S' + diagonal dominance idea, pdlltdriver.f): A0 + A0' + 2 n I"""
    a = O.pdmatgen(n, n, seed)
    return np.asfortranarray(a + a.T + 2.0 * n * np.eye(n))


def case_potrf(G, cs):
    """PDPOTRF ('L' / 'U') on a sub-matrix, then PDPOTRS, then PDPOSV; the other triangle and the guard rows stay untouched"""
    S, msgs = G.S, []
    n, nb, uplo, nrhs = cs["n"], cs["nb"], cs.get("uplo", "L"), cs.get("nrhs", 2)
    off = cs.get("off", 0)                                       # sub(A) = A(off*nb+1 : , off*nb+1 : ) of a larger matrix
    rsrc, csrc = cs.get("rsrc", 0) % G.P, cs.get("csrc", 0) % G.Q
    ng = n + off * nb
    ag = O.pdmatgen(ng, ng, 77); a0 = spd(n)
    if cs.get("notpd") is not None:
        a0[cs["notpd"], cs["notpd"]] = -1.0
    ag[off * nb:, off * nb:] = a0
    ag = np.asfortranarray(ag)
    al, desca = G.dist(ag, nb, rsrc, csrc, extra=1)
    ia = off * nb + 1
    dev = OnDevice(S, cs.get("dev"))
    info = S.pdpotrf(uplo, n, dev(al), ia, ia, desca)
    dev.back()
    ref = a0.copy(order="F"); info0 = O.dpotrf(uplo, ref, nb)
    if info != info0:
        msgs.append(f"pdpotrf info {info} != {info0}")
        return msgs
    if info0 != 0:
        return msgs                                              # same INFO; the partially factored matrix is not compared
    want = ag.copy(order="F"); want[off * nb:, off * nb:] = ref
    mloc, nloc = S.numroc(ng, nb, G.r, rsrc, G.P), S.numroc(ng, nb, G.c, csrc, G.Q)
    exp = G.local_of(want, nb, rsrc, csrc, lld=al.shape[0])
    anorm = np.abs(a0).sum(axis=1).max()
    if mloc and nloc:
        err = np.abs(al[:mloc, :nloc] - exp[:mloc, :nloc]).max() / (anorm * n * EPS)
        if not err < 1.0:
            msgs.append(f"factor error {err} (x ||A|| n eps)")
        # everything outside the UPLO triangle of sub(A) is bit-identical to the input
        gi = np.array([S.indxl2g(i + 1, nb, G.r, rsrc, G.P) - 1 for i in range(mloc)]); gj = np.array([S.indxl2g(j + 1, nb, G.c, csrc, G.Q) - 1 for j in range(nloc)])
        I, J = np.meshgrid(gi, gj, indexing="ij")
        insub = (I >= off * nb) & (J >= off * nb)
        tri = (I >= J) if uplo == "L" else (I <= J)
        untouched = ~(insub & tri)
        orig = G.local_of(ag, nb, rsrc, csrc, lld=al.shape[0])
        if not np.array_equal(al[:mloc, :nloc][untouched], orig[:mloc, :nloc][untouched]):
            msgs.append("elements outside the triangle were modified")
    if not np.all(al[mloc:, :] == -9923.0):
        msgs.append("guard row overwritten")
    # the reference's own check (pdlltdriver.f: PDPOTRRV + PDLAFCHK): || L L' - A || / (||A|| N eps) <= 3.0 (LLT.dat)
    fac = _to_root(G, n, n, al, ia, ia, desca)
    if fac is not None:
        t = np.tril(fac) if uplo == "L" else np.triu(fac)
        rec = t @ t.T if uplo == "L" else t.T @ t
        fres = np.abs(rec - a0).sum(axis=1).max() / (np.abs(a0).sum(axis=1).max() * n * EPS)
        if not fres <= 3.0:
            msgs.append(f"Cholesky factor residual {fres} > 3.0")
    if off:
        return msgs
    # PDPOTRS on the factor, PDPOSV from scratch (l3: force the many-right-hand-sides path whatever NRHS is)
    S.set_option("potrs_l3_min_nrhs", 1 if cs.get("l3") else 64)
    bg = matrix(n, nrhs, seed=200)
    bl, descb = G.dist(bg, nb, rsrc, 0, nbc=1); nlocb = S.numroc(nrhs, 1, G.c, 0, G.Q)
    info = S.pdpotrs(uplo, n, nrhs, al, 1, 1, desca, bl, 1, 1, descb)
    xg = bg.copy(order="F"); O.dpotrs(uplo, ref, xg)
    xe = G.local_of(xg, nb, rsrc, 0, nbc=1)
    if info != 0:
        msgs.append(f"pdpotrs info {info}")
    _close(msgs, "X (PDPOTRS)", bl[:mloc, :nlocb], xe[:mloc, :nlocb], 1e-10, atol=1e-14 * np.abs(xg).max())
    al2, _ = G.dist(np.asfortranarray(a0), nb, rsrc, csrc); bl2, _ = G.dist(bg, nb, rsrc, 0, nbc=1)
    info = S.pdposv(uplo, n, nrhs, al2, 1, 1, desca[:8] + [al2.shape[0]], bl2, 1, 1, descb)
    if info != 0:
        msgs.append(f"pdposv info {info}")
    _close(msgs, "X (PDPOSV)", bl2[:mloc, :nlocb], xe[:mloc, :nlocb], 1e-10, atol=1e-14 * np.abs(xg).max())
    return msgs


def _to_root(G, m, n, al, ia, ja, desc, cache={}):
    """sub(A) gathered on process (0, 0) through PDGEMR2D onto a 1 x 1 grid (None elsewhere); every process of the grid calls"""
    S = G.S
    key = (id(S), G.ctx)
    if key not in cache:
        cache[key] = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 1, 1)
    root = (G.r, G.c) == (0, 0)
    full = np.zeros((m + 1, n), order="F") if root else np.zeros(1)
    descf = S.descinit(m, n, m, n, 0, 0, cache[key], m + 1)[0] if root else [1, -1, m, n, m, n, 0, 0, 1]
    S.pdgemr2d(m, n, al, ia, ja, desc, full, 1, 1, descf, G.ctx)
    return np.asfortranarray(full[:m, :]) if root else None


def case_getri(G, cs):
    """PDGETRI on the oracle's factors (optionally of a sub-matrix with shifted source processes)"""
    S, msgs = G.S, []
    n, nb, off = cs["n"], cs["nb"], cs.get("off", 0)
    rsrc, csrc = cs.get("rsrc", 0) % G.P, cs.get("csrc", 0) % G.Q
    ng = n + off * nb
    a0 = matrix(n, cond=cs.get("cond"))
    if cs.get("dominant"):
        a0 = np.asfortranarray(a0 + n * np.eye(n))               # PDMATGEN( ..., 'N', 'D', ... ): the matrices of pdinvdriver.f
    lu = a0.copy(order="F"); ipg, info = O.getrf(lu, nb)
    if cs.get("singular") is not None:
        lu[cs["singular"], cs["singular"]] = 0.0
    big = O.pdmatgen(ng, ng, 55); big[off * nb:, off * nb:] = lu; big = np.asfortranarray(big)
    al, desca = G.dist(big, nb, rsrc, csrc, extra=1)
    mloc, nloc = S.numroc(ng, nb, G.r, rsrc, G.P), S.numroc(ng, nb, G.c, csrc, G.Q)
    ipfull = np.concatenate([np.arange(1, off * nb + 1, dtype=np.int32), ipg + off * nb]).astype(np.int32)
    ipl = O.ipiv_local(ng, ng, nb, G.P, G.r, ipfull, mloc + nb, rsrc=rsrc, fill=-77)
    ia = off * nb + 1
    dev = OnDevice(S, cs.get("dev"))
    info = S.pdgetri(n, dev(al), ia, ia, desca, ipl)
    dev.back()
    inv = lu.copy(order="F"); info0 = O.dgetri(inv, ipg, nb)
    if info != info0:
        msgs.append(f"pdgetri info {info} != {info0}")
        return msgs
    want = big.copy(order="F")
    if info0 == 0:
        want[off * nb:, off * nb:] = inv
    exp = G.local_of(want, nb, rsrc, csrc, lld=al.shape[0])
    if mloc and nloc:
        scale = np.abs(inv).max() if info0 == 0 else 1.0
        cond = np.linalg.cond(a0)
        err = np.abs(al[:mloc, :nloc] - exp[:mloc, :nloc]).max() / scale
        if not err < 50 * n * EPS * cond:
            msgs.append(f"inverse differs: {err} (cond {cond:.3g})")
    if not np.all(al[mloc:, :] == -9923.0):
        msgs.append("guard row overwritten")
    if cs.get("dominant") and info0 == 0:
        # the reference's own check (TESTING/traditional/LIN/pdinvchk.f:378): || inv(A) A - I ||_1 / (N eps ||A||_1) <= 1.0 (INV.dat)
        inv_root = _to_root(G, n, n, al, ia, ia, desca)
        if inv_root is not None:
            fres = np.abs(inv_root @ a0 - np.eye(n)).sum(axis=0).max() / (n * EPS * np.abs(a0).sum(axis=0).max())
            if not fres <= 1.0:
                msgs.append(f"pdinvchk residual {fres} > 1.0")
    if S.pdgetri(n, al, ia, ia, desca, ipl, lwork=0) != -8:
        msgs.append("short LWORK not reported")
    return msgs


def _place(G, S, glob, blk, src=(0, 0), guard=-9923.0):
    """local array (+ one guard row) and descriptor of a global matrix with mb x nb blocks from process src"""
    mb, nb = blk
    rs, cs_ = src[0] % G.P, src[1] % G.Q
    mloc = S.numroc(glob.shape[0], mb, G.r, rs, G.P)
    lld = max(1, mloc) + 1
    al = O.scatter(np.asfortranarray(glob), mb, nb, G.P, G.Q, G.r, G.c, rsrc=rs, csrc=cs_, lld=lld)
    al[mloc:, :] = guard
    desc, info = S.descinit(glob.shape[0], glob.shape[1], mb, nb, rs, cs_, G.ctx, lld)
    assert info == 0
    return al, desc, (mb, nb, rs, cs_, mloc)


def _expect(G, glob, lay, lld, guard=-9923.0):
    mb, nb, rs, cs_, mloc = lay
    e = O.scatter(np.asfortranarray(glob), mb, nb, G.P, G.Q, G.r, G.c, rsrc=rs, csrc=cs_, lld=lld)
    e[mloc:, :] = guard
    return e


def case_pdgemm(G, cs):
    """PDGEMM on sub-matrices with unrelated alignments and blockings, all four transposition pairs"""
    S, msgs = G.S, []
    m, n, k = cs["m"], cs["n"], cs["k"]
    ta, tb, alpha, beta = cs.get("ta", "N"), cs.get("tb", "N"), cs.get("alpha", 1.0), cs.get("beta", 1.0)
    (ia, ja), (ib, jb), (ic, jc) = cs.get("ija", (1, 1)), cs.get("ijb", (1, 1)), cs.get("ijc", (1, 1))
    sa = (k, m) if ta in "TC" else (m, k); sb = (n, k) if tb in "TC" else (k, n)
    ag = matrix(ia - 1 + sa[0] + 2, ja - 1 + sa[1] + 1, seed=11); bg = matrix(ib - 1 + sb[0] + 1, jb - 1 + sb[1] + 3, seed=12)
    cg = matrix(ic - 1 + m + 2, jc - 1 + n + 2, seed=13)
    al, desca, _ = _place(G, S, ag, cs.get("blk_a", (4, 4)), cs.get("src_a", (0, 0)))
    bl, descb, _ = _place(G, S, bg, cs.get("blk_b", (4, 4)), cs.get("src_b", (0, 0)))
    cl, descc, layc = _place(G, S, cg, cs.get("blk_c", (4, 4)), cs.get("src_c", (0, 0)))
    a_before, b_before = al.copy(), bl.copy()
    dev = OnDevice(S, cs.get("dev"))
    S.pdgemm(ta, tb, m, n, k, alpha, dev(al), ia, ja, desca, dev(bl), ib, jb, descb, beta, dev(cl), ic, jc, descc)
    dev.back()
    want = cg.copy(order="F")
    want[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = O.dgemm(ta, tb, alpha, ag[ia - 1:ia - 1 + sa[0], ja - 1:ja - 1 + sa[1]],
                                                         bg[ib - 1:ib - 1 + sb[0], jb - 1:jb - 1 + sb[1]], beta, cg[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n])
    exp = _expect(G, want, layc, cl.shape[0])
    scale = max(1.0, np.abs(want).max())
    if not np.allclose(cl, exp, rtol=0, atol=1e-13 * scale * max(1, k)):
        msgs.append(f"C differs by {np.abs(cl - exp).max()}")
    outside = np.ones_like(want, dtype=bool); outside[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = False
    mask = _expect(G, outside.astype(float), layc, cl.shape[0], guard=1.0) == 1.0
    if not np.array_equal(cl[mask], exp[mask]):
        msgs.append("elements outside sub(C) were modified")
    if not (np.array_equal(al, a_before) and np.array_equal(bl, b_before)):
        msgs.append("an input operand was modified")
    return msgs


def case_pdtrsm(G, cs):
    """PDTRSM: every SIDE / UPLO / TRANS / DIAG combination on non-aligned sub-matrices; the other triangle holds NaNs"""
    S, msgs = G.S, []
    m, n, alpha = cs["m"], cs["n"], cs.get("alpha", 1.0)
    side, uplo, ta, diag = cs.get("side", "L"), cs.get("uplo", "L"), cs.get("ta", "N"), cs.get("diag", "N")
    (ia, ja), (ib, jb) = cs.get("ija", (1, 1)), cs.get("ijb", (1, 1))
    na = m if side == "L" else n
    # well conditioned triangles at any size (random triangles with O(1) entries have exponentially growing inverses): the
    # off-diagonal row sums stay below the diagonal, unit or not
    base = matrix(na, seed=21)
    tri = base * (2.0 / max(na, 4))
    tri[np.diag_indices(na)] = 1.0 + np.abs(np.diag(base))
    tri = np.tril(tri) if uplo == "L" else np.triu(tri)
    stored = tri.copy()
    stored[np.triu_indices(na, 1) if uplo == "L" else np.tril_indices(na, -1)] = np.nan     # must never be read
    if diag == "U":
        stored[np.diag_indices(na)] = np.nan
    ag = matrix(ia - 1 + na + 1, ja - 1 + na + 2, seed=22); ag[ia - 1:ia - 1 + na, ja - 1:ja - 1 + na] = stored
    bg = matrix(ib - 1 + m + 2, jb - 1 + n + 1, seed=23)
    al, desca, _ = _place(G, S, ag, cs.get("blk_a", (4, 4)), cs.get("src_a", (0, 0)))
    bl, descb, layb = _place(G, S, bg, cs.get("blk_b", (4, 4)), cs.get("src_b", (0, 0)))
    S.pdtrsm(side, uplo, ta, diag, m, n, alpha, al, ia, ja, desca, bl, ib, jb, descb)
    want = bg.copy(order="F")
    want[ib - 1:ib - 1 + m, jb - 1:jb - 1 + n] = O.dtrsm(side, uplo, ta, diag, alpha, tri, bg[ib - 1:ib - 1 + m, jb - 1:jb - 1 + n])
    exp = _expect(G, want, layb, bl.shape[0])
    if not np.allclose(bl, exp, rtol=1e-10, atol=1e-12 * max(1.0, np.abs(want).max())):
        msgs.append(f"X differs by {np.nanmax(np.abs(bl - exp))}")
    return msgs


def case_pdtran(G, cs):
    S, msgs = G.S, []
    m, n, alpha, beta = cs["m"], cs["n"], cs.get("alpha", 1.0), cs.get("beta", 0.0)
    (ia, ja), (ic, jc) = cs.get("ija", (1, 1)), cs.get("ijc", (1, 1))
    ag = matrix(ia - 1 + n + 1, ja - 1 + m + 2, seed=31); cg = matrix(ic - 1 + m + 2, jc - 1 + n + 1, seed=32)
    al, desca, _ = _place(G, S, ag, cs.get("blk_a", (4, 4)), cs.get("src_a", (0, 0)))
    cl, descc, layc = _place(G, S, cg, cs.get("blk_c", (4, 4)), cs.get("src_c", (0, 0)))
    S.pdtran(m, n, alpha, al, ia, ja, desca, beta, cl, ic, jc, descc)
    want = cg.copy(order="F")
    want[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] = beta * cg[ic - 1:ic - 1 + m, jc - 1:jc - 1 + n] + alpha * ag[ia - 1:ia - 1 + n, ja - 1:ja - 1 + m].T
    exp = _expect(G, want, layc, cl.shape[0])
    if not np.allclose(cl, exp, rtol=1e-14, atol=1e-14):
        msgs.append(f"C differs by {np.abs(cl - exp).max()}")
    return msgs


def case_getrs_l3(G, cs):
    """PDGETRS through the level-3 (many right-hand sides) path: TRANS = N / T, sub-matrix factors, B with its own column blocking"""
    S, msgs = G.S, []
    n, nb, nrhs, trans, off = cs["n"], cs["nb"], cs["nrhs"], cs.get("trans", "N"), cs.get("off", 0)
    rsrc, csrc = cs.get("rsrc", 0) % G.P, cs.get("csrc", 0) % G.Q
    nbb = cs.get("nbb", nb)
    ng = n + off * nb
    a0 = matrix(n, cond=cs.get("cond")); lu = a0.copy(order="F"); ipg, info = O.getrf(lu, nb)
    big = O.pdmatgen(ng, ng, 55); big[off * nb:, off * nb:] = lu; big = np.asfortranarray(big)
    al, desca = G.dist(big, nb, rsrc, csrc)
    mloc = S.numroc(ng, nb, G.r, rsrc, G.P)
    ipfull = np.concatenate([np.arange(1, off * nb + 1, dtype=np.int32), ipg + off * nb]).astype(np.int32)
    ipl = O.ipiv_local(ng, ng, nb, G.P, G.r, ipfull, mloc + nb, rsrc=rsrc, fill=-77)
    bg = matrix(ng, nrhs + 3, seed=200)                          # sub(B) = B(off*nb+1 :, 3 : 3+nrhs)
    bl, descb, layb = _place(G, S, bg, (nb, nbb), (rsrc, 1))
    ia = off * nb + 1
    f = S.pdgetrs if cs.get("entry") else S.pdgetrs_l3
    info = f(trans, n, nrhs, al, ia, ia, desca, ipl, bl, ia, 3, descb)
    x = np.asfortranarray(bg[off * nb:, 2:2 + nrhs].copy()); O.getrs(lu, ipg, x, trans)
    want = bg.copy(order="F"); want[off * nb:, 2:2 + nrhs] = x
    exp = _expect(G, want, layb, bl.shape[0])
    if info != 0:
        msgs.append(f"info {info}")
    if not np.allclose(bl, exp, rtol=0, atol=1e-8 * np.abs(x).max()):
        msgs.append(f"X differs by {np.abs(bl - exp).max()} (|x| max {np.abs(x).max()})")
    return msgs


def case_ludriver(G, cs):
    """The reference's own LU test driver with EST = T (TESTING/traditional/LIN/pdludriver.f:380-900 on the LU.dat grid): PDLANGE ->
    PDGETRF -> PDGECON -> PDGETRS -> solve residual (pdlaschk.f, threshold 1.0) -> PDGERFS -> solve residual again, with the driver's
    guard zones (PADVAL) around A0, A, IPIV, B0, B, FERR, BERR checked after every call (PDCHEKPAD)."""
    S, msgs = G.S, []
    n, nb, nrhs, nbrhs = cs["n"], cs["nb"], cs["nrhs"], cs["nbrhs"]
    PAD = -9923.0
    a0g = O.pdmatgen(n, n, 100); b0g = O.pdmatgen(n, nrhs, 200)
    mloc, nloc, nlocb = S.numroc(n, nb, G.r, 0, G.P), S.numroc(n, nb, G.c, 0, G.Q), S.numroc(nrhs, nbrhs, G.c, 0, G.Q)
    lld = max(1, mloc) + 2
    def padded(glob, nbc):
        al = O.scatter(np.asfortranarray(glob), nb, nbc, G.P, G.Q, G.r, G.c, lld=lld)
        al[mloc:, :] = PAD
        return al
    a0l, al = padded(a0g, nb), padded(a0g, nb)
    b0l, bl = padded(b0g, nbrhs), padded(b0g, nbrhs)
    desca, _ = S.descinit(n, n, nb, nb, 0, 0, G.ctx, lld); descb, _ = S.descinit(n, nrhs, nb, nbrhs, 0, 0, G.ctx, lld)
    ipiv = np.full(mloc + nb + 2, -77, np.int32)             # LIPIV + a guard zone
    ferr, berr = np.full(max(1, nlocb) + 2, PAD), np.full(max(1, nlocb) + 2, PAD)
    def pads(where):
        ok = (np.all(a0l[mloc:, :] == PAD) and np.all(al[mloc:, :] == PAD) and np.all(b0l[mloc:, :] == PAD) and np.all(bl[mloc:, :] == PAD)
              and np.all(ipiv[mloc + nb:] == -77) and np.all(ferr[max(1, nlocb):] == PAD) and np.all(berr[max(1, nlocb):] == PAD)
              and np.array_equal(a0l[:mloc, :nloc], G.local_of(a0g, nb, lld=lld)[:mloc, :nloc]))
        if not ok:
            msgs.append(f"a guard zone (or A0) was overwritten by {where}")
    anorm1 = S.pdlange("1", n, n, al, 1, 1, desca)
    info = S.pdgetrf(n, n, al, 1, 1, desca, ipiv[:mloc + nb]); pads("PDGETRF")
    if info != 0:
        msgs.append(f"PDGETRF info {info}")
        return msgs
    rcond, info = S.pdgecon("1", n, al, 1, 1, desca, anorm1); pads("PDGECON")
    true_rc = 1.0 / (np.abs(a0g).sum(axis=0).max() * np.abs(np.linalg.inv(a0g)).sum(axis=0).max())
    # any estimate bounds ||inv(A)|| from below.  LAPACK's stays within ~10x of the truth; the reference's PDLACON returns its
    # alternating-sign value only (pdlacon.f:188-189), which drifts away with N (1500x at N = 1000): only the bound is checked then
    slack = 10.0 if cs.get("lapack_estimator") else float("inf")
    if info != 0 or not (true_rc * (1 - 1e-8) <= rcond <= slack * true_rc):
        msgs.append(f"PDGECON rcond {rcond} (true {true_rc}) info {info}")
    lu_o = a0g.copy(order="F"); O.getrf(lu_o, nb)
    want = O.dgecon("1", lu_o, np.abs(a0g).sum(axis=0).max())
    if not abs(rcond - want) <= 1e-6 * want:
        msgs.append(f"PDGECON rcond {rcond}, oracle {want}")
    info = S.pdgetrs("N", n, nrhs, al, 1, 1, desca, ipiv[:mloc + nb], bl, 1, 1, descb); pads("PDGETRS")
    # the residual check needs the global X: collect the pieces over the control plane (tiny) via PDGEMR2D onto one process
    xall = np.zeros((n + 1, nrhs), order="F") if (G.r, G.c) == (0, 0) else None
    one = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", 1, 1) if cs.get("_ctx1") is None else cs["_ctx1"]
    cs["_ctx1"] = one
    descx = S.descinit(n, nrhs, n, nrhs, 0, 0, one, n + 1)[0] if xall is not None else [1, -1, n, nrhs, n, nrhs, 0, 0, 1]
    def check(tag):
        S.pdgemr2d(n, nrhs, bl, 1, 1, descb, xall if xall is not None else np.zeros(1), 1, 1, descx, G.ctx)
        if xall is not None:
            res = O.sresid(a0g, np.asfortranarray(xall[:n, :]), b0g)
            if not res < 1.0:
                msgs.append(f"solve residual after {tag}: {res}")
    check("PDGETRS")
    info = S.pdgerfs("N", n, nrhs, a0l, 1, 1, desca, al, 1, 1, desca, ipiv[:mloc + nb], b0l, 1, 1, descb, bl, 1, 1, descb, ferr[:max(1, nlocb)], berr[:max(1, nlocb)])
    pads("PDGERFS")
    if info != 0:
        msgs.append(f"PDGERFS info {info}")
    check("PDGERFS")
    if nlocb and n > 1 and not (np.all(berr[:nlocb] >= 0) and np.all(berr[:nlocb] < 1e-13) and np.all(ferr[:nlocb] >= 0)):
        msgs.append(f"FERR {ferr[:nlocb]} BERR {berr[:nlocb]}")
    return msgs


CASES = {"lange": case_lange, "equ": case_equ, "gecon": case_gecon, "gerfs": case_gerfs, "gesvx": case_gesvx, "gemr2d": case_gemr2d, "potrf": case_potrf, "getri": case_getri, "pdgemm": case_pdgemm, "pdtrsm": case_pdtrsm, "pdtran": case_pdtran, "getrs_l3": case_getrs_l3, "ludriver": case_ludriver}


def run(S, ctx, cases):
    G = Grid(S, ctx)
    out = []
    for cs in cases:
        res = {"case": str(cs), "ok": True, "msgs": []}
        if G.r >= 0:
            # lapack_estimator: the full Higham iteration with EST carried between its stages (LAPACK's DLACON) in product AND oracle,
            # instead of what the reference's source returns (pdlacon.f:188-189, the default of both)
            keep = bool(cs.get("lapack_estimator"))
            S.set_option("lacon_keep_estimate", 1 if keep else 0); O.lacon_keep_est(keep)
            # exact ties in the pivot search (equilibrated matrices are full of entries that are exactly 1) are broken the way the
            # reference's PDAMAX breaks them on THIS grid: the serial oracle is told the grid's process rows and A's source row
            O.tie_grid(G.P, cs.get("rsrc", 0) % G.P)
            try:
                res["msgs"] = CASES[cs["kind"]](G, cs)
            except Exception as e:          # noqa: BLE001 - a rank must report, not vanish
                import traceback
                res["msgs"] = [f"exception {e!r}", traceback.format_exc()[-1500:]]
            S.set_option("lacon_keep_estimate", 0); O.lacon_keep_est(False); O.tie_grid(1, 0)
            res["ok"] = not res["msgs"]
        out.append(res)
    return out


# the default case lists: small enough for the serial emulation, shaped to hit partial blocks, P != Q ownership and non-aligned windows
F1_CASES = [
    dict(kind="lange", mg=45, ng=37, nb=4, ia=6, ja=3, m=30, n=29, rsrc=1, csrc=1),
    dict(kind="lange", mg=64, ng=64, nb=8, ia=1, ja=1, m=64, n=64),
    dict(kind="lange", mg=20, ng=20, nb=32, ia=2, ja=2, m=1, n=7),
    dict(kind="equ", n=37, m=45, nb=4, cond=6), dict(kind="equ", n=64, nb=8), dict(kind="equ", n=30, nb=4, cond=1, zero_row=7),
    dict(kind="gecon", n=64, nb=8), dict(kind="gecon", n=45, nb=4, cond=2), dict(kind="gecon", n=2, nb=2), dict(kind="gecon", n=37, nb=8, unaligned=(3, 5)),
    dict(kind="gerfs", n=64, nb=8, nrhs=3), dict(kind="gerfs", n=45, nb=4, nrhs=2, trans="T", cond=2), dict(kind="gerfs", n=30, nb=4, nrhs=5, nbr=2),
    dict(kind="gesvx", n=64, nb=8, fact="N"), dict(kind="gesvx", n=45, nb=4, fact="E", cond=5), dict(kind="gesvx", n=45, nb=4, fact="E", cond=5, trans="T"),
    dict(kind="gesvx", n=40, nb=8, fact="E"), dict(kind="gesvx", n=24, nb=4, fact="N", singular=True),
    dict(kind="gesvx", n=1, nb=4, fact="N", nrhs=1), dict(kind="gesvx", n=2, nb=4, fact="E", nrhs=1), dict(kind="gerfs", n=1, nb=2, nrhs=2), dict(kind="gecon", n=1, nb=2),
    dict(kind="gerfs", n=3, nb=2, nrhs=1, trans="T"),
    dict(kind="gesvx", n=40, nb=1, nrhs=1, fact="E", cond=2),          # equilibrated: exact ties in the pivot search, broken by process row (found by scripts/fuzz_next_rows.py on 3x2)
    dict(kind="gecon", n=64, nb=8, lapack_estimator=True), dict(kind="gecon", n=45, nb=4, cond=2, lapack_estimator=True),
    dict(kind="gerfs", n=45, nb=4, nrhs=2, trans="T", cond=2, lapack_estimator=True), dict(kind="gesvx", n=45, nb=4, fact="E", cond=5, lapack_estimator=True),
    dict(kind="gesvx", n=24, nb=4, fact="N", singular=True, lapack_estimator=True),
]

# PDGEMR2D: other block sizes (the NB=64 -> NB=512 use), rectangular blocks, shifted source processes, non-aligned sub-matrices,
# other grids (1 x np line, a single process, a transposed grid); ga / gb cases are skipped when the run has too few processes
F2_CASES = [
    dict(kind="gemr2d", m=50, n=40, shape_a=(50, 40), shape_b=(50, 40), blk_a=(4, 4), blk_b=(16, 16)),
    dict(kind="gemr2d", m=33, n=29, ia=5, ja=8, ib=2, jb=11, shape_a=(45, 41), shape_b=(37, 50), blk_a=(4, 3), blk_b=(7, 5), src_a=(1, 0), src_b=(0, 1)),
    dict(kind="gemr2d", m=64, n=64, shape_a=(64, 64), shape_b=(64, 64), blk_a=(8, 8), blk_b=(8, 8), src_b=(1, 1)),
    dict(kind="gemr2d", m=40, n=30, shape_a=(40, 30), shape_b=(40, 30), blk_a=(4, 4), blk_b=(5, 5), gb=(1, 1)),
    dict(kind="gemr2d", m=40, n=30, shape_a=(40, 30), shape_b=(40, 30), blk_a=(40, 30), blk_b=(3, 3), ga=(1, 1)),
    dict(kind="gemr2d", m=37, n=41, shape_a=(37, 41), shape_b=(37, 41), blk_a=(4, 4), blk_b=(6, 2), ga=(1, 2), gb=(2, 1)),
    dict(kind="gemr2d", m=37, n=41, ia=2, shape_a=(40, 41), shape_b=(37, 41), blk_a=(4, 4), blk_b=(6, 2), ga=(2, 2), gb=(1, 4)),
    dict(kind="gemr2d", m=37, n=41, shape_a=(37, 41), shape_b=(37, 41), blk_a=(4, 4), blk_b=(6, 2), ga=(1, 3), gb=(3, 2)),
    dict(kind="gemr2d", m=20, n=25, ja=3, shape_a=(20, 30), shape_b=(25, 25), ib=6, blk_a=(2, 2), blk_b=(4, 4), z=True),
    dict(kind="gemr2d", m=1, n=1, ia=7, ja=9, ib=3, jb=2, shape_a=(10, 10), shape_b=(5, 5), blk_a=(2, 2), blk_b=(3, 3)),
    dict(kind="gemr2d", m=33, n=29, ia=5, ja=8, ib=2, jb=11, shape_a=(45, 41), shape_b=(37, 50), blk_a=(4, 3), blk_b=(7, 5), src_a=(1, 0), src_b=(0, 1), chunk_mb=0),
]

# Cholesky: both triangles, partial last blocks, blocks wider than 32 (the diagonal block's inner loop), sub-matrices, shifted sources,
# a matrix that is not positive definite
F3_CASES = [
    dict(kind="potrf", n=64, nb=8, uplo="L"), dict(kind="potrf", n=64, nb=8, uplo="U"),
    dict(kind="potrf", n=45, nb=4, uplo="L", nrhs=3), dict(kind="potrf", n=45, nb=4, uplo="U", nrhs=3),
    dict(kind="potrf", n=45, nb=4, uplo="L", nrhs=9, l3=True), dict(kind="potrf", n=45, nb=4, uplo="U", nrhs=9, l3=True), dict(kind="potrf", n=64, nb=16, uplo="L", nrhs=2, l3=True), dict(kind="potrf", n=50, nb=8, uplo="L", nrhs=5, l3=True, rsrc=1, csrc=1),
    dict(kind="potrf", n=50, nb=8, uplo="U", nrhs=5, l3=True, rsrc=1, csrc=1), dict(kind="potrf", n=50, nb=8, uplo="L", nrhs=2, rsrc=1, csrc=1),
    dict(kind="potrf", n=150, nb=40, uplo="L"), dict(kind="potrf", n=150, nb=40, uplo="U"),
    dict(kind="potrf", n=100, nb=100, uplo="L"), dict(kind="potrf", n=7, nb=16, uplo="U"), dict(kind="potrf", n=1, nb=4, uplo="L", nrhs=1), dict(kind="potrf", n=2, nb=1, uplo="U", nrhs=1),
    dict(kind="potrf", n=40, nb=8, uplo="L", off=2, rsrc=1, csrc=1), dict(kind="potrf", n=40, nb=8, uplo="U", off=1, csrc=1),
    dict(kind="potrf", n=64, nb=8, uplo="L", notpd=37), dict(kind="potrf", n=64, nb=8, uplo="U", notpd=0), dict(kind="potrf", n=90, nb=40, uplo="L", notpd=75),
]

F4_CASES = [
    dict(kind="getri", n=64, nb=8), dict(kind="getri", n=1, nb=3),
    dict(kind="getri", n=50, nb=6, dominant=True), dict(kind="getri", n=15, nb=4, dominant=True), dict(kind="getri", n=30, nb=20, dominant=True), dict(kind="getri", n=2, nb=1), dict(kind="getri", n=45, nb=4, cond=2), dict(kind="getri", n=150, nb=40), dict(kind="getri", n=7, nb=16),
    dict(kind="getri", n=100, nb=100), dict(kind="getri", n=40, nb=8, off=2, rsrc=1, csrc=1), dict(kind="getri", n=64, nb=8, singular=37),
]

F4B_CASES = (
    [dict(kind="pdgemm", m=30, n=25, k=20, ta=ta, tb=tb, alpha=1.5, beta=-0.5, ija=(3, 2), ijb=(2, 6), ijc=(4, 3), blk_a=(4, 3), blk_b=(5, 5), blk_c=(6, 6),
          src_a=(1, 0), src_b=(0, 1), src_c=(1, 1)) for ta in "NT" for tb in "NT"]
    + [dict(kind="pdgemm", m=32, n=32, k=32, blk_a=(8, 8), blk_b=(8, 8), blk_c=(8, 8)), dict(kind="pdgemm", m=17, n=9, k=0, beta=2.0),
       dict(kind="pdgemm", m=17, n=9, k=5, alpha=0.0, beta=0.0), dict(kind="pdgemm", m=40, n=33, k=45, beta=0.0, blk_c=(16, 16), tb="T")]
    + [dict(kind="pdtrsm", m=26, n=19, side=sd, uplo=ul, ta=ta, diag=dg, alpha=0.75, ija=(2, 3), ijb=(3, 2), blk_a=(4, 4), blk_b=(5, 5), src_a=(0, 1), src_b=(1, 0))
       for sd in "LR" for ul in "LU" for ta in "NT" for dg in "NU"]
    + [dict(kind="pdtrsm", m=64, n=8, blk_a=(8, 8), blk_b=(8, 8)), dict(kind="pdtrsm", m=12, n=50, side="R", uplo="U", alpha=0.0),
       dict(kind="pdtrsm", m=45, n=45, side="R", uplo="L", ta="T", blk_b=(16, 16), blk_a=(16, 16))]
    + [dict(kind="pdtran", m=23, n=31, alpha=2.0, beta=0.5, ija=(2, 2), ijc=(3, 1), blk_a=(4, 4), blk_c=(7, 3), src_a=(1, 1)), dict(kind="pdtran", m=16, n=16, blk_a=(8, 8), blk_c=(8, 8))]
)

F5_CASES = [
    dict(kind="getrs_l3", n=64, nb=8, nrhs=20), dict(kind="getrs_l3", n=64, nb=8, nrhs=20, trans="T"),
    dict(kind="getrs_l3", n=45, nb=4, nrhs=7, nbb=3, cond=1), dict(kind="getrs_l3", n=45, nb=4, nrhs=7, nbb=3, trans="T"),
    dict(kind="getrs_l3", n=40, nb=8, nrhs=50, off=2, rsrc=1, csrc=1), dict(kind="getrs_l3", n=40, nb=8, nrhs=50, off=2, rsrc=1, csrc=1, trans="T"),
    dict(kind="getrs_l3", n=100, nb=40, nrhs=1),
    # through pdgetrs_ itself, which takes this path above 64 right-hand sides (and the replicated path below)
    dict(kind="getrs_l3", n=64, nb=8, nrhs=70, entry=True), dict(kind="getrs_l3", n=64, nb=8, nrhs=70, entry=True, trans="T"),
    dict(kind="getrs_l3", n=40, nb=8, nrhs=66, off=1, rsrc=1, csrc=1, entry=True), dict(kind="getrs_l3", n=64, nb=8, nrhs=5, entry=True),
]

# TESTING/traditional/LU.dat with EST = T: the square problem sizes x NB x NRHS x NBRHS of the reference's own input file
LUDAT_CASES = [dict(kind="ludriver", n=n, nb=nb, nrhs=nrhs, nbrhs=nbrhs) for n in (4, 13) for nb in (2, 3, 4) for nrhs in (1, 3, 9) for nbrhs in (1, 3, 5)]
