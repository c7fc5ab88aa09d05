"""Golden vectors of the index / descriptor tools, produced by EXECUTING the reference's own Fortran source
(TOOLS/numroc.f, indxg2p.f, indxg2l.f, indxl2g.f, iceil.f, ilcm.f, infog2l.f, descinit.f, chk1mat.f under /root/reference) with the
mini interpreter of tests/fortran77_mini.py -- there is no Fortran compiler in this image.  Writes tests/golden/tools_reference.npz.
Run here (the reference tree is not on the GPU boxes):  python tests/golden/make_tools_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fortran77_mini as F  # noqa: E402


def cases(rng, count):
    it = F.load_tools("/root/reference")
    idx, inf, dsc, chk = [], [], [], []
    for _ in range(count):
        np_ = int(rng.integers(1, 9)); nb = int(rng.integers(1, 9)); n = int(rng.integers(0, 200))
        ip = int(rng.integers(0, np_)); isrc = int(rng.integers(0, np_)); ig = int(rng.integers(1, 200)); il = int(rng.integers(1, 60))
        a, b = int(rng.integers(1, 60)), int(rng.integers(1, 60))
        idx.append([n, nb, ip, isrc, np_, ig, il, a, b,
                    it.call("NUMROC", n, nb, ip, isrc, np_)["__result__"], it.call("INDXG2P", ig, nb, ip, isrc, np_)["__result__"],
                    it.call("INDXG2L", ig, nb, ip, isrc, np_)["__result__"], it.call("INDXL2G", il, nb, ip, isrc, np_)["__result__"],
                    it.call("ICEIL", a, b)["__result__"], it.call("ILCM", a, b)["__result__"]])
        P, Q = int(rng.integers(1, 5)), int(rng.integers(1, 5)); r, c = int(rng.integers(0, P)), int(rng.integers(0, Q))
        it.state["grid"] = (P, Q, r, c)
        # DESCINIT with mostly legal, sometimes illegal arguments
        m, nn, mb, nbb = int(rng.integers(-1, 80)), int(rng.integers(-1, 80)), int(rng.integers(0, 9)), int(rng.integers(0, 9))
        rs, cs, lld = int(rng.integers(-1, P + 1)), int(rng.integers(-1, Q + 1)), int(rng.integers(0, 60))
        desc = [0] * 9
        out = it.call("DESCINIT", desc, m, nn, mb, nbb, rs, cs, 3, lld, 0)
        dsc.append([P, Q, r, c, m, nn, mb, nbb, rs, cs, 3, lld] + desc + [out["INFO"]])
        # INFOG2L / CHK1MAT on a legal descriptor
        M, N = int(rng.integers(1, 120)), int(rng.integers(1, 120)); mb, nbb = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        rs, cs = int(rng.integers(0, P)), int(rng.integers(0, Q))
        d = [1, 3, M, N, mb, nbb, rs, cs, max(1, it.call("NUMROC", M, mb, r, rs, P)["__result__"])]
        gi, gj = int(rng.integers(1, M + 1)), int(rng.integers(1, N + 1))
        o = it.call("INFOG2L", gi, gj, list(d), P, Q, r, c, 0, 0, 0, 0)
        inf.append([P, Q, r, c] + d + [gi, gj, o["LRINDX"], o["LCINDX"], o["RSRC"], o["CSRC"]])
        d2 = list(d)
        if rng.integers(0, 3) == 0:
            d2[int(rng.integers(0, 9))] = int(rng.integers(-2, 3))          # sometimes an illegal entry
        ma, na, ia, ja = int(rng.integers(-1, M + 3)), int(rng.integers(-1, N + 3)), int(rng.integers(0, M + 2)), int(rng.integers(0, N + 2))
        info_in = int(rng.choice([0, 0, 0, -3, -702]))
        o = it.call("CHK1MAT", ma, 1, na, 2, ia, ja, list(d2), 6, info_in)
        chk.append([P, Q, r, c] + d2 + [ma, na, ia, ja, info_in, o["INFO"]])
    return np.array(idx, np.int64), np.array(inf, np.int64), np.array(dsc, np.int64), np.array(chk, np.int64)


if __name__ == "__main__":
    idx, inf, dsc, chk = cases(np.random.default_rng(20261017), 400)
    np.savez_compressed(os.path.join(HERE, "tools_reference.npz"), index=idx, infog2l=inf, descinit=dsc, chk1mat=chk)
    print("wrote", idx.shape, inf.shape, dsc.shape, chk.shape)
