"""What the reference's OWN source returns as INFO for the argument combinations of tests/error_cases.py: every entry point's Fortran
(SRC/pdgetrf.f, pdgetrs.f, pdgesv.f, pdpotrf.f, pdpotrs.f, pdposv.f, pdgecon.f, pdgerfs.f, pdgesvx.f, pdgetri.f, pdgeequ.f under
/root/reference) is executed on a 1 x 1 grid.  Writes tests/golden/errors_reference.json.  python tests/golden/make_errors_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.dirname(HERE))
import error_cases as E  # noqa: E402
import fortran_chol_runner as RC  # noqa: E402
import fortran_lu_runner as RL  # noqa: E402
import fortran_refine_runner as RR  # noqa: E402


def reference_info(its, routine, a):
    itl, itc, itr = its
    A, B = E.matrices()
    flat = lambda m_: m_.reshape(-1, order="F").copy()  # noqa: E731
    ident = list(range(1, E.MG + 1)) + [0] * E.NB
    if routine == "PDGETRF":
        return itl.call("PDGETRF", a["m"], a["n"], A, a["ia"], a["ja"], a["desca"], [0] * (E.MG + E.NB), 0)["INFO"]
    if routine == "PDGETRS":
        return itl.call("PDGETRS", a["trans"], a["n"], a["nrhs"], A, a["ia"], a["ja"], a["desca"], ident, B, a["ib"], a["jb"], a["descb"], 0)["INFO"]
    if routine == "PDGESV":
        return itl.call("PDGESV", a["n"], a["nrhs"], A, a["ia"], a["ja"], a["desca"], [0] * (E.MG + E.NB), B, a["ib"], a["jb"], a["descb"], 0)["INFO"]
    if routine == "PDPOTRF":
        return itc.call("PDPOTRF", a["uplo"], a["n"], flat(A), a["ia"], a["ja"], a["desca"], 0)["INFO"]
    if routine in ("PDPOTRS", "PDPOSV"):
        return itc.call(routine, a["uplo"], a["n"], a["nrhs"], flat(A), a["ia"], a["ja"], a["desca"], flat(B), a["ib"], a["jb"], a["descb"], 0)["INFO"]
    W, IW = np.zeros(8192), np.zeros(8192, np.int64)
    ip = np.array(ident, np.int64)
    if routine == "PDGECON":
        return itr.call("PDGECON", a["norm"], a["n"], flat(A), a["ia"], a["ja"], a["desca"], a["anorm"], 0.0, W, a["lwork"], IW, a["liwork"], 0)["INFO"]
    if routine == "PDGERFS":
        return itr.call("PDGERFS", a["trans"], a["n"], a["nrhs"], flat(A), a["ia"], a["ja"], a["desca"], flat(A), a["iaf"], a["jaf"], a["descaf"], ip,
                        flat(B), a["ib"], a["jb"], a["descb"], flat(B), a["ix"], a["jx"], a["descx"], np.zeros(16), np.zeros(16), W, 4096, IW, 4096, 0)["INFO"]
    if routine == "PDGESVX":
        return itr.call("PDGESVX", a["fact"], a["trans"], a["n"], a["nrhs"], flat(A), a["ia"], a["ja"], a["desca"], flat(A), a["iaf"], a["jaf"], a["descaf"],
                        ip, a["equed"], np.ones(32), np.ones(32), flat(B), a["ib"], a["jb"], a["descb"], flat(B), a["ix"], a["jx"], a["descx"], 0.0,
                        np.zeros(16), np.zeros(16), W, 4096, IW, 4096, 0)["INFO"]
    if routine == "PDGETRI":
        return itr.call("PDGETRI", a["n"], flat(A), a["ia"], a["ja"], a["desca"], ip, W, a["lwork"], IW, a["liwork"], 0)["INFO"]
    if routine == "PDGEEQU":
        return itr.call("PDGEEQU", a["m"], a["n"], flat(A), a["ia"], a["ja"], a["desca"], np.zeros(32), np.zeros(32), 0.0, 0.0, 0.0, 0)["INFO"]
    raise KeyError(routine)


def _alarm(signum, frame):
    raise TimeoutError("no result within the limit")


if __name__ == "__main__":
    import signal
    signal.signal(signal.SIGALRM, _alarm)
    its = (RL.make(extra=(("SRC", "pdgesv"),)), RC.make(extra=(("SRC", "pdposv"),)), RR.make(extra=RR.SVX_UNITS + RR.TRI_UNITS + RR.LU_UNITS))
    out, undefined = [], []
    for routine in E.ROUTINES:
        for label, changes in E.mutations() + E.pair_mutations(routine):
            a = E.apply(routine, changes)
            if a is None:
                continue
            try:
                signal.alarm(4)                                           # a call that does not come back has no defined answer either
                info = int(reference_info(its, routine, a))
                signal.alarm(0)
            except (ZeroDivisionError, IndexError, ValueError, AssertionError, TypeError, KeyError, TimeoutError, OverflowError) as ex:
                signal.alarm(0)
                undefined.append((routine, label, type(ex).__name__))     # the source divides by a zero block size, indexes out of range ...: no defined answer
                continue
            out.append(dict(routine=routine, label=label, changes={k: ([list(x) for x in v] if isinstance(v, list) else list(v)) if isinstance(v, (tuple, list)) else v for k, v in changes.items()}, info=info))
    json.dump(dict(cases=out, undefined=undefined), open(os.path.join(HERE, "errors_reference.json"), "w"), indent=0)
    print(len(out), "cases;", len(undefined), "without a defined answer:", undefined[:12])
    for r in E.ROUTINES:
        print(r, sorted({c["info"] for c in out if c["routine"] == r}))
