#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
SLB200_SOLVE_GRAPH=0 timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,launch__grid_size -k regex:"diag_solve|gemv_rows|inv32" -s 0 -c 40 --csv --log-file gpurun_out/r2h_solve_kernels.csv python scripts/ncu_driver.py solve 32768 512 > gpurun_out/r2h.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2h_solve_kernels.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"]
if hdr:
    h = rows[hdr[0]]; ki = h.index("Kernel Name"); mi = h.index("Metric Name"); vi = h.index("Metric Value"); ii = h.index("ID")
    cur = {}
    for r in rows[hdr[0] + 1:]:
        if len(r) <= vi: continue
        cur.setdefault(r[ii], {"k": r[ki][:60]})[r[mi]] = r[vi]
    for i, d in list(cur.items())[:40]:
        print(i, d)
PY
