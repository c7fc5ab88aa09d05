"""The LU path's own host orchestration on the CPU: lu.cu (panel gather / Pbuf scatter / row broadcast / column all-gather / look-ahead
and the two-half pipeline) and api.cu (argument checks, windows, IPIV distribution) are compiled UNCHANGED against the stub CUDA runtime
of tests/emul, their kernels replaced by contract stand-ins (tests/emul/backend.cpp), NCCL by the TCP control plane with the same call
order (point-to-point included).  tests/mp_worker.py -- the same worker the multi-GPU tests use -- then checks IPIV bit-exactly and
the factors / solutions against the oracle on grids up to 2 x 4 and 4 x 2 (eight processes).  This is the CPU stand-in for
tests/test_gpu_multi.py::test_eight_gpus, which needs an 8-GPU lease: it cannot see kernel bugs or stream races, it does see every
index, message size, root and ordering mistake of the distributed schedule."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMUL = os.path.join(ROOT, "tests", "emul")


@pytest.fixture(scope="session")
def emul_lib():
    subprocess.check_call(["make", "-C", EMUL, "-s", "-j8"])
    return os.path.join(EMUL, "libslb_emul.so")


def spawn(world, cases, timeout=240, extra_env=None):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   SLB200_PORT_OFFSET="0", SLB200_EMUL="1", OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1", **(extra_env or {}))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mp_worker.py"), json.dumps(cases)], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = []
    try:
        for p in procs:
            o, e = p.communicate(timeout=timeout)
            assert p.returncode == 0, e[-3000:]
            outs.append(json.loads([ln for ln in o.splitlines() if ln.startswith("RESULT")][0][6:]))
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    bad = [(o["rank"], r["case"], r["msgs"]) for o in outs for r in o["results"] if not r["ok"]]
    assert not bad, bad
    return outs


SMALL = [dict(m=m, n=n, nb=nb, nrhs=3) for (m, n) in [(4, 4), (10, 12), (17, 13), (13, 13)] for nb in (2, 3, 4)]
GENERAL = [dict(mg=40, ng=40, nb=4, ia=9, ja=5, m=20, n=20, rsrc=0, csrc=0), dict(mg=150, ng=130, nb=16, ia=33, ja=17, m=100, n=100, rsrc=1, csrc=1),
           dict(mg=200, ng=200, nb=32, ia=1, ja=1, m=200, n=200, rsrc=1, csrc=0), dict(mg=256, ng=256, nb=32, ia=65, ja=33, m=120, n=160, rsrc=0, csrc=1)]


def cases_for(P, Q):
    mid = [dict(m=200, n=200, nb=32, nrhs=2), dict(m=300, n=200, nb=64, nrhs=0), dict(m=250, n=250, nb=16, nrhs=1),
           dict(m=120, n=120, nb=16, nrhs=2, z=True),
           dict(m=384, n=384, nb=32, nrhs=1, split=64),                      # two-half pipeline + look-ahead forced at this size
           dict(m=384, n=384, nb=32, nrhs=1, split=64, hoststream=True),    # block rows go back to the host caller during the sweep
           dict(m=300, n=420, nb=32, nrhs=0, split=64, hoststream=True), dict(m=200, n=200, nb=32, nrhs=1, z=True, split=64)]
    gen = [dict(c, rsrc=c["rsrc"] % P, csrc=c["csrc"] % Q) for c in GENERAL]
    return [dict(c, P=P, Q=Q) for c in SMALL + mid + gen]


@pytest.mark.parametrize("P,Q", [(1, 1), (1, 2), (2, 1), (2, 2), (1, 4), (4, 1), (2, 3), (2, 4), (4, 2)])
def test_lu_orchestration_on_the_cpu(emul_lib, P, Q):
    spawn(P * Q, cases_for(P, Q))


@pytest.mark.parametrize("delay", [0, 1, 2, 5, 11])
@pytest.mark.parametrize("save_mb", [16384, 0])
def test_host_streaming_with_late_slabs(emul_lib, delay, save_mb):
    """1 x 1 grid, host-resident caller: the column slabs of A land in the staging copy after a varying number of polls (the emulated
    HostLink poisons what has not arrived), so they join the sweep at different block steps and are replayed through the steps they
    missed; with no panel-keep budget (save_mb = 0) every slab is forced in at once.  The factors must not depend on any of that."""
    cases = [dict(P=1, Q=1, m=384, n=384, nb=32, nrhs=1, split=64, hoststream=True), dict(P=1, Q=1, m=300, n=420, nb=32, nrhs=0, split=64, hoststream=True),
             dict(P=1, Q=1, m=420, n=300, nb=32, nrhs=0, hoststream=True), dict(P=1, Q=1, m=256, n=256, nb=16, nrhs=2, z=True, split=32, hoststream=True)]
    spawn(1, cases, extra_env={"SLB200_EMUL_SLAB_DELAY": str(delay), "SLB200_E2E_SAVE_MB": str(save_mb), "SLB200_E2E_SLAB_MB": "0"})


def test_gpu_interface_tests_against_the_emulation(emul_lib):
    """The host-resident cases of tests/test_gpu_iface.py and tests/test_gpu_lu.py (argument checks and INFO codes, sub-matrix operands,
    RSRC / CSRC, TRANS = N / T / C, the reference's 6 x 6 example and LU.dat grid, zero pivots, guard rows, complex) run UNCHANGED
    against the emulation library: the entry points' host side is checked on every CPU run, not only when a GPU is at hand.
    Deselected: device-resident operands, kernel-vs-kernel bit-identity tests, the on-device generators, the full-size runs."""
    env = dict(os.environ, SLB200_EMUL="1", OPENBLAS_NUM_THREADS="2", OMP_NUM_THREADS="2")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    run = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_iface.py"), os.path.join(ROOT, "tests", "test_gpu_lu.py"),
                          "-q", "-m", "gpu", "-p", "no:cacheprovider",
                          "-k", "not True and not bit_identical and not device_generators and not full_size and not large_properties"],
                         env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = run.stdout[-1500:]
    assert run.returncode == 0, tail + run.stderr[-1500:]
    assert " passed" in tail and "failed" not in tail, tail


@pytest.mark.parametrize("P,Q", [(1, 1), (2, 2), (3, 2)])
def test_degenerate_split_settings_are_clamped(emul_lib, P, Q):
    """`la_split_min` below 2 NB (or below 4 columns) used to split steps whose near half could not hold the next panel: with 1-3 trailing
    columns the near half ran empty and the next panel was factored before its update (wrong factors, silently; found by
    scripts/fuzz_lu_path.py).  The option is now taken no lower than max(2 NB, 4); these are the reduced failing configurations."""
    spawn(P * Q, [dict(P=P, Q=Q, m=6, n=6, nb=1, nrhs=1, split=3), dict(P=P, Q=Q, m=11, n=11, nb=1, nrhs=2, split=1, hoststream=True),
                  dict(P=P, Q=Q, m=24, n=24, nb=3, nrhs=1, split=3), dict(P=P, Q=Q, m=63, n=53, nb=2, nrhs=1, split=2, hoststream=True),
                  dict(P=P, Q=Q, m=47, n=44, nb=3, nrhs=1, z=True, split=3), dict(P=P, Q=Q, m=117, n=117, nb=1, nrhs=2, split=1, hoststream=True)])
