"""Multi-GPU parity: PDGETRF / PDGETRS / PDGESV on P x Q grids (one GPU per BLACS process, NCCL), against the
oracle.  Includes BASELINE config 1 (N=2000 NB=64 on a 2x2 grid) and the LU.dat grids 2x2, 1x4, 4x1."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def spawn(world, cases, timeout=600):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), SLB200_PORT_OFFSET="0")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_worker.py"), json.dumps(cases)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    import time
    deadline = time.time() + timeout
    try:
        for p in procs:
            o, e = p.communicate(timeout=max(1.0, deadline - time.time()))
            assert p.returncode == 0, e[-3000:]
            outs.append(json.loads([l for l in o.splitlines() if l.startswith("RESULT")][0][6:]))
    except BaseException:
        # a rank that died (or hangs) leaves its peers waiting in a collective: never leave them behind on the GPUs
        tails = []
        for p in procs:
            if p.poll() is None:
                p.kill()
            try:
                o, e = p.communicate(timeout=10)
                tails.append((e or "")[-1500:])
            except Exception:
                pass
        print("\n==== rank stderr tails ====\n" + "\n----\n".join(tails), file=sys.stderr)
        raise
    bad = [(o["rank"], r["case"], r["msgs"]) for o in outs for r in o["results"] if not r["ok"]]
    assert not bad, bad
    return outs


SMALL = [dict(m=m, n=n, nb=nb, nrhs=3) for (m, n) in [(4, 4), (10, 12), (17, 13), (13, 13)] for nb in (2, 3, 4)]
# sub-matrix operands (IA, JA > 1, M, N < descriptor) on grids whose first block does not live on process (0, 0)
GENERAL = [dict(mg=40, ng=40, nb=4, ia=9, ja=5, m=20, n=20, rsrc=0, csrc=0), dict(mg=300, ng=260, nb=32, ia=65, ja=33, m=200, n=200, rsrc=1, csrc=1),
           dict(mg=500, ng=500, nb=64, ia=1, ja=1, m=500, n=500, rsrc=1, csrc=0), dict(mg=640, ng=640, nb=64, ia=129, ja=65, m=300, n=400, rsrc=0, csrc=1)]


def general(P, Q):
    return [dict(c, P=P, Q=Q, rsrc=c["rsrc"] % P, csrc=c["csrc"] % Q) for c in GENERAL]


@pytest.mark.parametrize("P,Q", [(1, 2), (2, 1)])
def test_two_gpus(P, Q):
    if ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    cases = [dict(c, P=P, Q=Q) for c in SMALL] + [dict(P=P, Q=Q, m=200, n=200, nb=32, nrhs=2), dict(P=P, Q=Q, m=1000, n=1000, nb=64, nrhs=1, dev=True),
                                                   dict(P=P, Q=Q, m=300, n=200, nb=64, nrhs=0), dict(P=P, Q=Q, m=1536, n=1536, nb=512, nrhs=1),
                                                   dict(P=P, Q=Q, m=120, n=120, nb=16, nrhs=2, z=True),
                                                   dict(P=P, Q=Q, m=3072, n=3072, nb=128, nrhs=1, dev=True, split=256),   # pipelined halves
                                                   dict(P=P, Q=Q, m=3072, n=3072, nb=128, nrhs=1, split=256, hoststream=True),
                                                   dict(P=P, Q=Q, m=2000, n=1500, nb=64, nrhs=0, split=128, hoststream=True)]
    spawn(2, cases + general(P, Q))


@pytest.mark.parametrize("P,Q", [(2, 2), (1, 4), (4, 1)])
def test_four_gpus(P, Q):
    if ngpus() < 4:
        pytest.skip("needs 4 GPUs")
    cases = [dict(c, P=P, Q=Q) for c in SMALL] + [dict(P=P, Q=Q, m=2000, n=2000, nb=64, nrhs=1),           # BASELINE config 1
                                                   dict(P=P, Q=Q, m=777, n=513, nb=100, nrhs=0), dict(P=P, Q=Q, m=2048, n=2048, nb=512, nrhs=2, dev=True),
                                                   dict(P=P, Q=Q, m=200, n=200, nb=32, nrhs=2, z=True),
                                                   dict(P=P, Q=Q, m=4096, n=4096, nb=128, nrhs=1, dev=True, split=256),  # pipelined halves
                                                   dict(P=P, Q=Q, m=4096, n=4096, nb=128, nrhs=1, split=256, hoststream=True),
                                                   dict(P=P, Q=Q, m=1500, n=2000, nb=64, nrhs=0, split=128, hoststream=True)]
    spawn(4, cases + general(P, Q))


def test_eight_gpus():
    if ngpus() < 8:
        pytest.skip("needs 8 GPUs")
    cases = [dict(P=2, Q=4, m=2000, n=2000, nb=64, nrhs=1), dict(P=2, Q=4, m=4096, n=4096, nb=512, nrhs=1, dev=True), dict(P=2, Q=4, m=13, n=13, nb=2, nrhs=3),
             dict(P=4, Q=2, m=1000, n=1000, nb=64, nrhs=2), dict(P=2, Q=4, m=512, n=512, nb=64, nrhs=1, z=True),
             dict(P=2, Q=4, m=8192, n=8192, nb=256, nrhs=1, dev=True, split=512), dict(P=2, Q=4, m=2048, n=2048, nb=128, nrhs=2, z=True),
             dict(P=2, Q=4, m=8192, n=8192, nb=256, nrhs=1, split=512, hoststream=True)]
    spawn(8, cases + general(2, 4))
