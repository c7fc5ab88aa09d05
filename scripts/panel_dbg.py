import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, oracle as O, scalapack_b200 as S
S.set_option("panel_debug", 1)
for (m, jb) in [(33, 32), (5000, 512), (40000, 512), (65536, 512)]:
    a = O.matgen64_tile(m, 5, 0, m, 0, jb)
    ipiv = np.zeros(jb, np.int32); info = C.c_int(0)
    ms = S.lib().slb200_test_panel(m, jb, S.api._ptr(a), C.c_int64(m), ipiv.ctypes.data_as(C.c_void_p), C.byref(info), 0)
    dbg = np.zeros((2, 4096, 8), np.uint64)
    S.lib().slb200_test_panel_dbg(dbg.ctypes.data_as(C.c_void_p))
    for who in (0, 1):
        t = dbg[who, :jb, :6].astype(np.int64)
        d = np.diff(t, axis=1)                       # ns between stamps within a column
        nxt = t[1:, 0] - t[:-1, 5]                   # end of column -> start of next (incl. sub-panel phases every 32)
        inner = np.array([nxt[i] for i in range(len(nxt)) if (i + 1) % 32 != 0])
        outer = np.array([nxt[i] for i in range(len(nxt)) if (i + 1) % 32 == 0])
        print(f"m={m} jb={jb} cta={'first' if who == 0 else 'last'} total={ms:.3f} ms  per-column ns: reduce={d[:,0].mean():.0f} publish={d[:,1].mean():.0f} "
              f"poll={d[:,2].mean():.0f} fetch={d[:,3].mean():.0f} update={d[:,4].mean():.0f} gap={inner.mean() if len(inner) else 0:.0f} "
              f"subpanel_phases={outer.mean() if len(outer) else 0:.0f}")
