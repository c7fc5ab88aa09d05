"""Runs the FULL-SIZE case lists of tests/test_gpu_next.py (orders up to 2500) through the host-logic emulation on the CPU (1 x 1 grid,
device-resident variants stripped): checks that the GPU-only case definitions are well formed and that their tolerances hold for an
arithmetic ordered differently from the oracle's.  A few minutes; not part of the pytest suite.  python scripts/run_gpu_cases_on_emulation.py"""
import json, os, subprocess, sys, time, socket
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_next as t
lists = dict(t.ONE_GPU)                       # one list per entry point, exactly what tests/test_gpu_next.py::test_one_gpu runs
if len(sys.argv) > 1:
    lists = {k: v for k, v in lists.items() if k in sys.argv[1:]}
for name, cases in lists.items():
    cases = [{k: v for k, v in c.items() if k not in ("dev", "entry")} for c in cases]
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, RANK="0", LOCAL_RANK="0", WORLD_SIZE="1", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), SLB200_EMUL="1", OPENBLAS_NUM_THREADS="4")
    t0 = time.time()
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "next_worker.py"), json.dumps(dict(P=1, Q=1, cases=cases))], env=env, capture_output=True, text=True, timeout=3000)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
    if not line:
        print(name, "NO RESULT rc", p.returncode, p.stderr[-1500:]); continue
    res = json.loads(line[0][6:])["results"]
    bad = [(r["case"], r["msgs"]) for r in res if not r["ok"]]
    print(f"{name}: {len(res)} cases, {len(bad)} bad, {time.time() - t0:.0f} s", flush=True)
    for b in bad: print("   ", str(b)[:600])
