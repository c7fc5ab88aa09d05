"""Illegal (and a few unusual but legal) argument combinations for every entry point of the path and of the SURVEY 8(f) rows, in ONE
description that can be replayed through the reference's own source (executed by tests/fortran77_mini.py, 1 x 1 grid) and through the
product's C-ABI: the INFO a caller gets must be the same number.  TEST INFRASTRUCTURE (tests/golden/make_errors_golden.py writes what the
executed reference returns, tests/test_emul_next.py replays the list through the product)."""
import numpy as np

MG, NB, NBRHS = 12, 4, 4        # A, AF: 12 x 12 in 4 x 4 blocks; B, X: 12 x 4; sub-matrices of order 8 with 2 right-hand sides by default

ROUTINES = {
    "PDGETRF": ["m", "n", "ia", "ja", "desca"],
    "PDGETRS": ["trans", "n", "nrhs", "ia", "ja", "desca", "ib", "jb", "descb"],
    "PDGESV": ["n", "nrhs", "ia", "ja", "desca", "ib", "jb", "descb"],
    "PDPOTRF": ["uplo", "n", "ia", "ja", "desca"],
    "PDPOTRS": ["uplo", "n", "nrhs", "ia", "ja", "desca", "ib", "jb", "descb"],
    "PDPOSV": ["uplo", "n", "nrhs", "ia", "ja", "desca", "ib", "jb", "descb"],
    "PDGECON": ["norm", "n", "ia", "ja", "desca", "anorm", "lwork", "liwork"],
    "PDGERFS": ["trans", "n", "nrhs", "ia", "ja", "desca", "iaf", "jaf", "descaf", "ib", "jb", "descb", "ix", "jx", "descx"],
    "PDGESVX": ["fact", "trans", "n", "nrhs", "ia", "ja", "desca", "iaf", "jaf", "descaf", "equed", "ib", "jb", "descb", "ix", "jx", "descx"],
    "PDGETRI": ["n", "ia", "ja", "desca", "lwork", "liwork"],
    "PDGEEQU": ["m", "n", "ia", "ja", "desca"],
}


def base():
    d = [1, 0, MG, MG, NB, NB, 0, 0, MG]
    db = [1, 0, MG, NBRHS, NB, NBRHS, 0, 0, MG]
    return dict(m=8, n=8, nrhs=2, ia=1, ja=1, iaf=1, jaf=1, ib=1, jb=1, ix=1, jx=1, desca=list(d), descaf=list(d), descb=list(db), descx=list(db),
                trans="N", uplo="L", norm="1", fact="N", equed="N", anorm=1.0, lwork=4096, liwork=4096)


def mutations():
    """[(label, {argument: value}), ...]; a descriptor entry is given as (index, value)"""
    out = [("valid", {})]
    for k, vals in (("m", (-1, 0, 3)), ("n", (-1, 0, 1, 4)), ("nrhs", (-1, 0, 1)), ("ia", (0, 2, 5, 9)), ("ja", (0, 2, 5, 9)), ("iaf", (2, 5)), ("jaf", (2, 5)),
                    ("ib", (0, 2, 5)), ("jb", (0, 2, 4)), ("ix", (2, 5)), ("jx", (2, 4)), ("trans", ("T", "C", "X", "n")), ("uplo", ("U", "X", "u")),
                    ("norm", ("I", "O", "X")), ("fact", ("E", "X")), ("anorm", (-1.0, 0.0)), ("lwork", (1, -1)), ("liwork", (0, -1))):
        out += [(f"{k}={v!r}", {k: v}) for v in vals]
    for dname in ("desca", "descaf", "descb", "descx"):
        for idx, vals in ((0, (2,)), (1, (1,)), (2, (-1, 7)), (3, (-1, 1)), (4, (0, 3)), (5, (0, 3)), (6, (1, -1)), (7, (1,)), (8, (3, 0))):
            if dname == "desca" and idx == 1:
                continue                                     # A's context is THE context: an invalid one is a BLACS matter, not an INFO code
            out += [(f"{dname}[{idx}]={v}", {dname: (idx, v)}) for v in vals]
    out += [("fact=F equed=Q", {"fact": "F", "equed": "Q"}), ("fact=F equed=R", {"fact": "F", "equed": "R"}),
            ("mb!=nb both", {"desca": (4, 3), "descaf": (4, 3)}), ("ia=5 ja=5", {"ia": 5, "ja": 5}), ("ia=5 ja=5 ib=5", {"ia": 5, "ja": 5, "ib": 5, "ix": 5, "iaf": 5, "jaf": 5}),
            ("n=0 nrhs=0", {"n": 0, "nrhs": 0})]
    return out


def pair_mutations(routine, count=40, seed=7):
    """two mutations at once (which error is reported first is part of the behaviour): a seeded sample per routine"""
    import random
    rng = random.Random(seed + sum(map(ord, routine)))
    singles = [(lab, ch) for lab, ch in mutations()[1:] if apply(routine, ch) is not None and len(ch) == 1]
    out = []
    while len(out) < count:
        (l1, c1), (l2, c2) = rng.sample(singles, 2)
        k1, k2 = next(iter(c1)), next(iter(c2))
        if k1 == k2 and not k1.startswith("desc"):
            continue
        if k1 == k2:                                           # two fields of the same descriptor: a list of (index, value) pairs
            if c1[k1][0] == c2[k2][0]:
                continue
            out.append((f"{l1} & {l2}", {k1: [c1[k1], c2[k2]]}))
        else:
            out.append((f"{l1} & {l2}", {**c1, **c2}))
    return out


def apply(routine, changes):
    """the argument set of one call, or None when the mutation does not concern this routine"""
    names = ROUTINES[routine.replace("PZ", "PD")]
    if any(k not in names for k in changes):
        return None
    a = base()
    for k, v in changes.items():
        if k.startswith("desc"):
            for idx, val in (v if isinstance(v, list) else [v]):
                a[k][idx] = val
        else:
            a[k] = v
    return a


def matrices():
    """A (12 x 12, symmetric and strongly diagonally dominant: every principal sub-matrix is positive definite and its triangles are
    usable as "factors"), B (12 x 4)"""
    rng = np.random.default_rng(42)
    g = rng.uniform(-1, 1, (MG, MG))
    return np.asfortranarray(g + g.T + 2.0 * MG * np.eye(MG)), np.asfortranarray(rng.uniform(-1, 1, (MG, NBRHS)))


def product_info(S, ctx, routine, a, ctx_other=None):
    """the same call through the product's C-ABI (1 x 1 grid, host operands); context 0 of the description is ctx, 1 is ctx_other: ANOTHER
    VALID 1 x 1 context (the executed source sees a valid grid behind every handle; an invalid handle is a different error, on RSRC)"""
    A, B = matrices()
    d = {k: [(ctx_other if v == 1 else ctx) if i == 1 else v for i, v in enumerate(a[k])] for k in ("desca", "descaf", "descb", "descx")}
    # "no interchange" pivots: the entry of every local row is that row's own global index (the layout PDGETRF leaves, pdgetrf.f:118-121)
    P, _, myrow, _ = S.blacs_gridinfo(ctx)
    mb, rs = (a["desca"][4] if a["desca"][4] > 0 else 1), (a["desca"][6] if 0 <= a["desca"][6] < P else 0)
    nloc = S.numroc(MG, mb, myrow, rs, P)
    ip = np.zeros(MG + NB + 4, np.int32)
    ip[:nloc] = [S.indxl2g(l + 1, mb, myrow, rs, P) for l in range(nloc)]
    if routine in ("PZGETRF", "PZGETRS", "PZGESV"):                   # the complex entry points: the same checks (their source is the real one, type-swapped)
        Az, Bz = (A + 0.5j * A.T).copy(order="F"), (B * (1 + 0.25j)).copy(order="F")
        if routine == "PZGETRF":
            return S.pzgetrf(a["m"], a["n"], Az, a["ia"], a["ja"], d["desca"], ip)
        if routine == "PZGETRS":
            return S.pzgetrs(a["trans"], a["n"], a["nrhs"], Az, a["ia"], a["ja"], d["desca"], ip, Bz, a["ib"], a["jb"], d["descb"])
        return S.pzgesv(a["n"], a["nrhs"], Az, a["ia"], a["ja"], d["desca"], ip, Bz, a["ib"], a["jb"], d["descb"])
    if routine == "PDGETRF":
        return S.pdgetrf(a["m"], a["n"], A, a["ia"], a["ja"], d["desca"], ip)
    if routine == "PDGETRS":
        return S.pdgetrs(a["trans"], a["n"], a["nrhs"], A, a["ia"], a["ja"], d["desca"], ip, B, a["ib"], a["jb"], d["descb"])
    if routine == "PDGESV":
        return S.pdgesv(a["n"], a["nrhs"], A, a["ia"], a["ja"], d["desca"], ip, B, a["ib"], a["jb"], d["descb"])
    if routine == "PDPOTRF":
        return S.pdpotrf(a["uplo"], a["n"], A, a["ia"], a["ja"], d["desca"])
    if routine == "PDPOTRS":
        return S.pdpotrs(a["uplo"], a["n"], a["nrhs"], A, a["ia"], a["ja"], d["desca"], B, a["ib"], a["jb"], d["descb"])
    if routine == "PDPOSV":
        return S.pdposv(a["uplo"], a["n"], a["nrhs"], A, a["ia"], a["ja"], d["desca"], B, a["ib"], a["jb"], d["descb"])
    if routine == "PDGECON":
        return S.pdgecon(a["norm"], a["n"], A, a["ia"], a["ja"], d["desca"], a["anorm"], lwork=a["lwork"], liwork=a["liwork"])[1]
    fe, be = np.zeros(16), np.zeros(16)
    if routine == "PDGERFS":
        return S.pdgerfs(a["trans"], a["n"], a["nrhs"], A, a["ia"], a["ja"], d["desca"], A.copy(order="F"), a["iaf"], a["jaf"], d["descaf"], ip, B, a["ib"],
                         a["jb"], d["descb"], B.copy(order="F"), a["ix"], a["jx"], d["descx"], fe, be)
    if routine == "PDGESVX":
        return S.pdgesvx(a["fact"], a["trans"], a["n"], a["nrhs"], A, a["ia"], a["ja"], d["desca"], A.copy(order="F"), a["iaf"], a["jaf"], d["descaf"], ip,
                         a["equed"], np.ones(32), np.ones(32), B, a["ib"], a["jb"], d["descb"], B.copy(order="F"), a["ix"], a["jx"], d["descx"], fe, be)[2]
    if routine == "PDGETRI":
        return S.pdgetri(a["n"], A, a["ia"], a["ja"], d["desca"], ip, lwork=a["lwork"], liwork=a["liwork"])
    if routine == "PDGEEQU":
        return S.pdgeequ(a["m"], a["n"], A, a["ia"], a["ja"], d["desca"], np.zeros(32), np.zeros(32))[3]
    raise KeyError(routine)
