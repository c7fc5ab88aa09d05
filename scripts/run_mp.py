"""Light multi-GPU parity run: python scripts/run_mp.py WORLD 'json list of cases' (see tests/mp_worker.py)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_gpu_multi import spawn  # noqa: E402

world = int(sys.argv[1])
cases = json.loads(sys.argv[2])
t = time.time()
outs = spawn(world, cases, timeout=int(sys.argv[3]) if len(sys.argv) > 3 else 300)
for r in outs[0]["results"]:
    print(r["case"], "ok" if r["ok"] else r["msgs"], "lu_err=%.3g" % r.get("lu_err", -1))
print("all ranks ok in %.1f s" % (time.time() - t))
