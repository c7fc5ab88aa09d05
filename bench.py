#!/usr/bin/env python
"""bench.py -- PDGETRF FP64 TFLOP/s (2/3 N^3) through the C-ABI drop-in, on N GPUs of one node.

  python bench.py --gpus 1 --steps K --warmup W            # BASELINE config 2: N=65536 NB=512, 1x1 grid
  torchrun ... bench.py --gpus N                           # 1x2 / 2x2 / 2x4 grids, 32 GiB of A per GPU (weak scaling)
  torchrun ... bench.py --gpus N --config c3|c4|c5         # BASELINE configs 3 (PDGESV N=131072, strong scaling),
                                                           # 4 (PDGETRF N=262144 on 2x4), 5 (PZGETRF N=65536 NB=256 on 2x4)
  python bench.py --impl reference                         # the reference algorithm on the host cores

A step = one factorisation of the freshly regenerated matrix (64-bit LCG generator, on the device).  `value` is the
device-timed factorisation with A resident in HBM; `e2e` is the same call with A in HOST memory (H2D + factor + D2H
inside the timed region; pinned and pageable callers are both measured).  The matrix (>= 8 GiB per GPU) is far larger than
the 126 MB L2, so no explicit L2 flush is needed between steps.

Before anything is timed, a PARITY PRE-FLIGHT factors two small reference-generator matrices (PDMATGEN N=4096 NB=512 and
BASELINE config 1's N=2000 NB=64) on the SAME P x Q grid through the same entry points and compares IPIV (bit-exact), the
LU factors, FRESID and the PDGETRS solution with the CPU oracle (the oracle is the checker here, never the thing measured);
the run aborts if the pre-flight fails, so every line this script prints carries parity evidence for its grid
(TESTING/traditional/LIN/pdludriver.f:738,961-969 prints the same PASSED/FAILED verdict next to every timing line).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GRIDS = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}
A_SEED, B_SEED = 20261017, 777
EPS = 2.0 ** -53
CPU_SAMPLE_N = int(os.environ.get("SLB200_BENCH_CPU_SAMPLE_N", 16384))   # the reference arm's bounded sample: a full factorisation at this
                                                                          # fixed N (see cpu_sample; the variable exists for the CPU test suite)

# BASELINE.json configs.  "weak" is the driver's default scaling run: 32 GiB of A per GPU at every grid.
CONFIGS = {
    "c2":   dict(routine="pdgetrf", n=65536,  nb=512, gpus=(1,),        scaling="weak",   cplx=False),
    "weak": dict(routine="pdgetrf", n=None,   nb=512, gpus=(1, 2, 4, 8), scaling="weak",  cplx=False),
    "c3":   dict(routine="pdgesv",  n=131072, nb=512, gpus=(2, 4, 8),   scaling="strong", cplx=False),
    "c4":   dict(routine="pdgetrf", n=262144, nb=512, gpus=(8,),        scaling="weak",   cplx=False),
    "c5":   dict(routine="pzgetrf", n=65536,  nb=256, gpus=(8,),        scaling="weak",   cplx=True),
}
METRIC = {"pdgetrf": "pdgetrf_fp64_tflops", "pdgesv": "pdgesv_fp64_tflops", "pzgetrf": "pzgetrf_fp64_tflops"}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.rows, self.p = dev, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.dev), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._rd, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _rd(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 9 and r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 9:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0.5 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(self.rows), "reasons": sorted(reasons)}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------- workload description (shared by both arms)
def pick_config(args):
    name = args.config or ("c2" if args.gpus == 1 else "weak")
    cfg = dict(CONFIGS[name]); cfg["name"] = name
    if args.nb:
        cfg["nb"] = args.nb
    if args.n:
        cfg["n"] = args.n
    elif cfg["n"] is None:
        cfg["n"] = int(65536 * math.sqrt(args.gpus)) // cfg["nb"] * cfg["nb"]      # 32 GiB of A per GPU at every grid
    return cfg


def workload_string(cfg, gpus):
    P, Q = GRIDS[gpus]
    per_gpu = cfg["n"] ** 2 * (16 if cfg["cplx"] else 8) / gpus / 2 ** 30
    what = {"pdgetrf": "PDGETRF", "pdgesv": "PDGESV NRHS=1", "pzgetrf": "PZGETRF"}[cfg["routine"]]
    return f"{what} N={cfg['n']} NB={cfg['nb']} grid {P}x{Q} ({per_gpu:.1f} GiB of A per GPU), 64-bit LCG matrix"


def flop_counts(cfg):
    """(headline flops, flops by the reference driver's own model pdludriver.f:913-918)."""
    n = float(cfg["n"]); mul = 4.0 if cfg["cplx"] else 1.0
    head = mul * (2.0 / 3.0) * n ** 3
    ref = mul * ((2.0 / 3.0) * n ** 3 - 0.5 * n ** 2)
    if cfg["routine"] == "pdgesv":
        head += 2.0 * n * n; ref += 2.0 * n * n
    return head, ref


# --------------------------------------------------------------------------- reference arm (CPU)
def cpu_sample(nb, threads=None, n=CPU_SAMPLE_N):
    """The oracle port of the reference algorithm (unblocked panel + BLAS-3 update in the reference's order,
    oracle/oracle.c, bundled OpenBLAS as the host BLAS) on the host cores, on a bounded sample of the workload: one FULL
    factorisation of the same generator's matrix at the fixed size N = CPU_SAMPLE_N (the full-size matrix would take
    minutes per step on a CPU).  The thread count is set explicitly (torchrun exports OMP_NUM_THREADS=1)."""
    import oracle as O
    cores = os.cpu_count() or 1
    threads = threads or cores
    O.set_threads(threads)
    n = max(2 * nb, n // nb * nb)
    a = O.matgen64_tile(n, A_SEED, 0, n, 0, n)
    t = time.perf_counter()
    ipiv, info = O.getrf(a, nb)
    dt = time.perf_counter() - t
    tf = (2.0 / 3.0) * n ** 3 / dt / 1e12
    return {"value": tf, "unit": "TFLOP/s", "cores": threads, "kind": "port",
            "sample": f"one full PDGETRF (oracle/oracle.c restatement + OpenBLAS, {threads} threads) of the same generator's "
                      f"matrix at the fixed sample size N={n} NB={nb}: {dt:.2f} s; the reference (Fortran 77 + MPI) cannot be "
                      f"built in this image",
            "n": n, "seconds": dt, "info": int(info)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = pick_config(args)
    nb = min(cfg["nb"], 512)
    for _ in range(args.warmup and 1):                         # one untimed pass pages the BLAS in
        cpu_sample(nb, n=4096)
    res = [cpu_sample(nb) for _ in range(max(1, args.steps))]
    mean_s = sum(r["seconds"] for r in res) / len(res)
    best = dict(res[0]); best["seconds"] = mean_s
    best["value"] = (2.0 / 3.0) * best["n"] ** 3 / mean_s / 1e12
    line = {"impl": "reference", "metric": METRIC[cfg["routine"]], "value": best["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_string(cfg, args.gpus), "flops_model": "2/3 N^3",
                       "timed_sample": best["sample"], "sample_n": best["n"]},
            "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": best["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- parity pre-flight (oracle = checker)
def parity_preflight(S, ctx, P, Q, dist, cplx, cases=None):
    """PDGETRF + PDGETRS on the run's own P x Q grid against the CPU oracle.  Returns a dict for the JSON line;
    `ok` False aborts the run.  Every rank checks its block-cyclic piece; the verdicts are combined over the grid."""
    import numpy as np
    import torch
    import oracle as O
    _, _, r, c = S.blacs_gridinfo(ctx)
    world = P * Q
    O.set_threads(max(1, (os.cpu_count() or 1) // world))
    out = {"ok": True, "grid": f"{P}x{Q}", "cases": []}
    if cases is None:
        cases = [(4096, 512, False), (2000, 64, False)] + ([(1024, 256, True)] if cplx else [])
    for n, nb, z in cases:
        gen = O.pzmatgen if z else O.pdmatgen
        a0 = gen(n, n, 100); b0 = gen(n, 1, 200)
        ref = a0.copy(order="F")
        ipr, infr = O.getrf(ref, nb)
        xr = b0.copy(order="F"); O.getrs(ref, ipr, xr)
        mloc, nloc = S.numroc(n, nb, r, 0, P), S.numroc(n, nb, c, 0, Q)
        lld = max(1, mloc)
        al = O.scatter(a0, nb, nb, P, Q, r, c, lld=lld)
        desca, _ = S.descinit(n, n, nb, nb, 0, 0, ctx, lld)
        ipiv = np.full(mloc + nb, -77, np.int32)
        info = (S.pzgetrf if z else S.pdgetrf)(n, n, al, 1, 1, desca, ipiv)
        refl = O.scatter(ref, nb, nb, P, Q, r, c, lld=lld)
        ipl = O.ipiv_local(n, n, nb, P, r, ipr, mloc + nb, fill=-77)
        own = ipl != -77
        ipiv_exact = bool(info == infr == 0 and np.array_equal(ipiv[own], ipl[own]))
        anorm = np.abs(a0).sum(axis=1).max()
        lu_err = float(np.abs(al[:mloc, :nloc] - refl[:mloc, :nloc]).max() / (anorm * n * EPS)) if mloc and nloc else 0.0
        # solve with the factors: x against the oracle's x
        bl = O.scatter(b0, nb, 1, P, Q, r, c, lld=lld)
        descb, _ = S.descinit(n, 1, nb, 1, 0, 0, ctx, lld)
        inf2 = (S.pzgetrs if z else S.pdgetrs)("N", n, 1, al, 1, 1, desca, ipiv, bl, 1, 1, descb)
        xl = O.scatter(xr, nb, 1, P, Q, r, c, lld=lld)
        x_err = float(np.abs(bl[:mloc, :1] - xl[:mloc, :1]).max() / np.abs(xr).max()) if (mloc and c == 0) else 0.0
        # FRESID of the assembled factors (pdlafchk.f:225-226): pieces summed over the grid, checked on rank 0
        lug = np.zeros((n, n), dtype=a0.dtype, order="F")
        O.gather_into(lug, np.asfortranarray(al), nb, nb, P, Q, r, c)
        if dist is not None:
            t = torch.from_numpy(lug.ravel(order="K").view(np.float64)).cuda()      # memory (column-major) order
            dist.all_reduce(t); lug = t.cpu().numpy().view(a0.dtype).reshape((n, n), order="F")
        fres = float(O.fresid(np.asfortranarray(lug), ipr, a0)) if ipiv_exact else float("nan")
        v = torch.tensor([0.0 if ipiv_exact and inf2 == 0 else 1.0, lu_err, x_err], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
        bad, lu_err, x_err = [float(x) for x in v.tolist()]
        ok = bad == 0.0 and lu_err < 1.0 and fres < 1.0 and x_err < 1e-8
        out["cases"].append({"matrix": f"{'PZMATGEN' if z else 'PDMATGEN'} N={n} NB={nb} seed 100", "ipiv_exact": bad == 0.0,
                             "lu_err": lu_err, "fresid": fres, "x_err": x_err, "ok": ok})
        out["ok"] = out["ok"] and ok
    out["ipiv_exact"] = all(cs["ipiv_exact"] for cs in out["cases"])
    out["lu_err"] = max(cs["lu_err"] for cs in out["cases"])
    out["fresid"] = max(cs["fresid"] for cs in out["cases"])
    out["checker"] = "oracle/oracle.c (CPU restatement) on every rank's block-cyclic piece; tolerances: IPIV bit-exact, lu_err < 1, FRESID < 1, x_err < 1e-8"
    return out


# --------------------------------------------------------------------------- our arm (GPU)
def zresid_check(torch, dist, S, ctx, n, nb, P, Q, myrow, mycol, X, lld, W):
    """Solve residual of pdlaschk.f:187,296 for the complex workload: A (regenerated into the scratch array W by the device
    generator) times the distributed solution with torch (a checker, not the product path)."""
    mloc, nloc = S.numroc(n, nb, myrow, 0, P), S.numroc(n, nb, mycol, 0, Q)
    dev = X.device
    gi = torch.arange(mloc, device=dev); gi = ((gi // nb) * P + myrow) * nb + gi % nb
    gj = torch.arange(nloc, device=dev); gj = ((gj // nb) * Q + mycol) * nb + gj % nb
    xg = torch.zeros(n, dtype=torch.complex128, device=dev)
    if mycol == 0:
        xg[gi] = X[:mloc]
    if dist is not None:
        t = torch.view_as_real(xg); dist.all_reduce(t)
    S.zmatgen64(ctx, n, n, nb, nb, W, lld, A_SEED)
    Aloc = W.view(nloc, lld)[:, :mloc]                       # row j = local column j
    part = torch.zeros(n, dtype=torch.complex128, device=dev); rowabs = torch.zeros(n, dtype=torch.float64, device=dev)
    step = max(1, (1 << 27) // max(1, mloc))
    acc = torch.zeros(mloc, dtype=torch.complex128, device=dev); ra = torch.zeros(mloc, dtype=torch.float64, device=dev)
    for j0 in range(0, nloc, step):
        blk = Aloc[j0:j0 + step]
        acc += xg[gj[j0:j0 + step]] @ blk
        ra += blk.abs().sum(dim=0)
    part[gi] = acc; rowabs[gi] = ra
    if dist is not None:
        dist.all_reduce(torch.view_as_real(part)); dist.all_reduce(rowabs)
    S.zmatgen64(ctx, n, 1, nb, 1, W, lld, B_SEED)
    bg = torch.zeros(n, dtype=torch.complex128, device=dev)
    if mycol == 0:
        bg[gi] = W[:mloc]
    if dist is not None:
        dist.all_reduce(torch.view_as_real(bg))
    return float(((bg - part).abs().max() / (xg.abs().max() * rowabs.max() * EPS * n)).item())


def run_ours(args):
    # ONE JSON line on stdout: everything native libraries print to file descriptor 1 (NCCL's version banner) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import numpy as np
    import torch
    import scalapack_b200 as S
    t_start = time.time()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} needs WORLD_SIZE={args.gpus} (torchrun), got {world}"
    cfg = pick_config(args)
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    P, Q = GRIDS[args.gpus]
    n, nb, cplx, routine = cfg["n"], cfg["nb"], cfg["cplx"], cfg["routine"]
    ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", P, Q)
    _, _, myrow, mycol = S.blacs_gridinfo(ctx)

    # ---- parity pre-flight on this grid (aborts the run when it fails) ----
    pre = None
    if not args.no_preflight:
        pre = parity_preflight(S, ctx, P, Q, dist, cplx)
        if rank == 0:
            log("parity pre-flight:", json.dumps(pre))
        if not pre["ok"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC[routine], "value": None, "n_gpus": args.gpus, "parity_preflight": pre,
                                  "error": "parity pre-flight failed: nothing was timed"}), file=json_out, flush=True)
            sys.exit(3)

    mloc, nloc = S.numroc(n, nb, myrow, 0, P), S.numroc(n, nb, mycol, 0, Q)
    lld = max(1, mloc)
    desca, info = S.descinit(n, n, nb, nb, 0, 0, ctx, lld)
    assert info == 0
    dt = torch.complex128 if cplx else torch.float64
    esz = 16 if cplx else 8
    A = torch.empty(nloc * lld, dtype=dt, device="cuda")                  # column-major local array in HBM
    ipiv = np.zeros(mloc + nb, np.int32)
    flops, flops_ref = flop_counts(cfg)
    matgen = S.zmatgen64 if cplx else S.matgen64
    getrf = S.pzgetrf if cplx else S.pdgetrf
    getrs = S.pzgetrs if cplx else S.pdgetrs
    descb, _ = S.descinit(n, 1, nb, 1, 0, 0, ctx, lld)
    X = torch.zeros(max(1, lld), dtype=dt, device="cuda")

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step():
        matgen(ctx, n, n, nb, nb, A, lld, A_SEED)               # untimed: regenerate the matrix in HBM
        if routine == "pdgesv":
            matgen(ctx, n, 1, nb, 1, X, lld, B_SEED)
            barrier()
            inf = S.pdgesv(n, 1, A, 1, 1, desca, ipiv, X, 1, 1, descb)
            barrier()
            assert inf == 0, inf
            return maxr(S.last_factor_ms()) + maxr(S.last_solve_ms())    # device-timed factor + solve (pdludriver.f:926-956)
        barrier()
        inf = getrf(n, n, A, 1, 1, desca, ipiv)
        barrier()
        assert inf == 0, inf
        return maxr(S.last_factor_ms())

    if args.profile:
        S.set_option("profile", 1)
    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local); sampler.start()
    S.reset_counters()
    times, upd = [], []
    for _ in range(args.steps):
        times.append(step())
        upd.append(S.last_update())
    clocks = sampler.stop()
    launches = S.get_counter("kernel_launches")
    prof = {k: S.get_counter(k) for k in ("prof_panel_us", "prof_swap_us", "prof_trsm_us", "prof_gemm_us")} if args.profile else None
    ms = sum(times) / len(times)
    value = flops / (ms * 1e-3) / 1e12
    factor_ms = maxr(S.last_factor_ms())

    # ---- correctness of the timed workload: solve with the factors, reference residual on regenerated A, b ----
    if routine != "pdgesv":
        matgen(ctx, n, 1, nb, 1, X, lld, B_SEED)                # b = column 0 of the generator with B_SEED
        inf = getrs("N", n, 1, A, 1, 1, desca, ipiv, X, 1, 1, descb)
        assert inf == 0
    solve_ms = solve_first_ms = maxr(S.last_solve_ms())
    if cplx:
        W = torch.empty(nloc * lld, dtype=dt, device="cuda")
        sresid = zresid_check(torch, dist, S, ctx, n, nb, P, Q, myrow, mycol, X, lld, W)
        del W
    else:
        sresid = S.pdlaschk(ctx, n, 1, X, descb, desca, A_SEED, B_SEED, gen=64)
    if world > 1 and routine != "pdgesv":   # the first multi-GPU solve pays NCCL's lazy connection set-up of the world communicator
        matgen(ctx, n, 1, nb, 1, X, lld, B_SEED)
        inf = getrs("N", n, 1, A, 1, 1, desca, ipiv, X, 1, 1, descb)
        assert inf == 0
        solve_ms = maxr(S.last_solve_ms())
    peaks = measured_peaks()
    hbm_peak = peaks.get("hbm_gbs") or 6546.9
    solve_bytes = esz * float(n) * n / world                   # L and U read once, this GPU's share
    solve_gbs = solve_bytes / (solve_ms * 1e-3) / 1e9 if solve_ms > 0 else None
    solve_kernel = ("PDGETRS 'N' two-stream sweep: diag_solve_kernel -> gemv_rows_kernel (top) beside gemv_bulk_kernel (solve_fast.cu)"
                    if (world == 1 and not cplx and nb <= 512) else
                    "PDGETRS block substitution with NCCL reductions / broadcasts per block (solve.cu, solve_kernels.cu)")
    roof_solve = {"bound": "hbm", "kernel": solve_kernel, "achieved": solve_gbs, "peak": hbm_peak,
                  "unit": "GB/s", "frac": solve_gbs / hbm_peak if solve_gbs else None, "traffic": None,
                  "algorithmic_bytes_per_gpu": solve_bytes, "solve_ms": solve_ms,
                  "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks.get("hbm_gbs") else "fallback 6546.9 GB/s (B200_PROFILING.md)"}

    # ---- roofline of the dominant kernel (trailing update, FP64 DMMA): live CUDA-event time over the timed steps ----
    u_ms = sum(u[0] for u in upd); u_fl = sum(u[1] for u in upd); u_n = sum(u[2] for u in upd)
    dmma_peak = S.lib().slb200_bench_dmma_tflops(20000)
    dfma_peak = S.lib().slb200_bench_dfma_tflops(20000)
    achieved = u_fl / (u_ms * 1e-3) / 1e12 if u_ms > 0 else None
    # DRAM traffic of the update kernel: one `ncu --set full` capture (profiles/r02_gemm_ncu.md, M=N=32768, K=512) read
    # 11.48 GB and wrote 8.54 GB for 17.45 GB of algorithmic bytes (16*m*n for C + the operands) = 1.147x; scaled here to
    # the average launch of this run (algorithmic C bytes of a launch = its flops * 8 / NB).  Not a live counter.
    traffic = (u_fl / u_n) * 8.0 / nb * 1.147 if (u_n and not cplx) else None
    kern = ("dgemm_minus_packed<CMODE=1> (complex update on the real FP64 DMMA kernel: K doubled, interleaved C; gemm_packed.cu)" if cplx else
            "dgemm_minus_packed (FP64 DMMA trailing update, gemm_packed.cu; dgemm_minus_p8b for m < 3072)")
    roof = {"bound": "tensor", "kernel": kern, "achieved": achieved, "peak": dmma_peak,
            "unit": "TFLOP/s", "frac": (achieved / dmma_peak) if achieved else None, "traffic": traffic,
            "traffic_unit": "bytes per average launch = algorithmic C bytes x 1.147 (dram read+write / algorithmic bytes of one ncu --set full capture at M=N=32768, not a live counter)",
            "peak_source": "FP64 DMMA peak measured live by slb200_bench_dmma_tflops (MEASURED_PEAKS.json has no FP64 entry)",
            "fp64_fma_peak_tflops": dfma_peak, "share_of_step": u_ms / (factor_ms * len(upd)) if (upd and factor_ms) else None,
            "launches": u_n, "avg_launch_ms": u_ms / u_n if u_n else None}

    # ---- e2e: the same call with A in HOST memory (H2D + factor + D2H inside the timed region), pinned and pageable ----
    e2e, e2e_pageable = None, None
    if not args.no_e2e:
        def e2e_leg(pinned):
            # every decision here is COLLECTIVE (min / max over the ranks): a rank that skipped a leg on its own would leave the
            # others waiting in the next barrier
            need = nloc * lld * esz * world
            try:
                import psutil
                avail = float(psutil.virtual_memory().available)
            except Exception:
                avail = float("inf")
            avail = -maxr(-avail)
            if need > 0.6 * avail:
                raise MemoryError(f"host copies of A need {need / 2**30:.0f} GiB, {avail / 2**30:.0f} GiB of host memory available")
            elapsed = maxr(time.time() - t_start)
            if not pinned and elapsed > args.time_budget:
                raise TimeoutError(f"skipped: {elapsed:.0f} s of the run's {args.time_budget} s budget were used before this leg")
            Ah = torch.empty(nloc * lld, dtype=dt, pin_memory=pinned)
            Xh = torch.empty(max(1, lld), dtype=dt, pin_memory=pinned)
            e_times = []
            # the driver allows 870 s per run: at 8 GPUs 25 device-resident steps alone take ~490 s (round 1: 662 s in total), so
            # the end-to-end legs shrink with the grid: one timed pass and no warm-up pass at 8 GPUs
            n_timed = max(1, min(args.steps, (args.e2e_steps if world < 8 else 1) if pinned else 1))
            # the first pass is a warm-up (staging buffers, copy threads); the pageable leg of a large grid reuses the pinned leg's
            n_warm = 1 if ((pinned and world < 8) or world < 4) else 0
            check_bits = pinned or world < 4                    # the extra device-resident factorisation is not repeated at scale
            for it in range(n_timed + n_warm):
                matgen(ctx, n, n, nb, nb, A, lld, A_SEED)
                Ah.copy_(A)
                if routine == "pdgesv":
                    matgen(ctx, n, 1, nb, 1, X, lld, B_SEED); Xh.copy_(X)
                barrier()
                t0 = time.perf_counter()
                if routine == "pdgesv":
                    inf = S.pdgesv(n, 1, Ah.numpy(), 1, 1, desca, ipiv, Xh.numpy(), 1, 1, descb)
                else:
                    inf = getrf(n, n, Ah.numpy(), 1, 1, desca, ipiv)
                barrier()
                dt_ = maxr(time.perf_counter() - t0)
                assert inf == 0
                if it >= n_warm:
                    e_times.append(dt_)
            # the factors in the caller's host array against a device-resident factorisation of the same matrix: same bits
            same = None
            if check_bits:
                ip_host = ipiv.copy()
                if routine == "pdgesv":
                    matgen(ctx, n, 1, nb, 1, X, lld, B_SEED)
                    assert S.pdgesv(n, 1, A, 1, 1, desca, ipiv, X, 1, 1, descb) == 0
                else:
                    assert getrf(n, n, A, 1, 1, desca, ipiv) == 0
                same = bool(np.array_equal(ip_host, ipiv))
                chunk = 1 << 27
                for o in range(0, nloc * lld, chunk):
                    same = same and bool(torch.equal(Ah[o:o + chunk].cuda(), A[o:o + chunk]))
                same = maxr(0.0 if same else 1.0) == 0.0
            e_s = sum(e_times) / len(e_times)
            nbytes = nloc * lld * esz
            res = {"value": flops / e_s / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": nbytes * world,
                   "d2h_bytes_per_step": nbytes * world + 4 * n, "seconds": e_s, "steps": len(e_times),
                   "host_memory": "pinned" if pinned else "pageable", "h2d_overlapped": bool(S.get_counter("e2e_upload_overlapped")),
                   "d2h_overlapped": bool(S.get_counter("e2e_download_overlapped")), "bit_identical_to_device_resident": same}
            del Ah, Xh
            return res
        for pinned in (True, False):
            try:
                r_ = e2e_leg(pinned)
            except Exception as ex:  # host RAM for a copy of A may be missing
                r_ = {"value": None, "unit": "TFLOP/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None,
                      "host_memory": "pinned" if pinned else "pageable", "error": repr(ex)[:200]}
            if pinned:
                e2e = r_
            else:
                e2e_pageable = r_
            if args.no_pageable:
                break

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_sample(min(nb, 512))
    # ---- supplementary: first hardware numbers of the SURVEY 8(f) rows (bench_next.py), one process per row, hard time limits; never
    # part of `value` / `e2e`, and unable to delay or break this line: the leg is abandoned when it does not return in time ----
    next_rows, next_thread = None, None
    if not args.no_next:
        # every rank takes the same decision (max over the ranks): each one starts its own rank of every row's process group
        next_rows, next_thread = next_rows_leg(world, maxr(time.time() - t_start))
    if rank == 0:
        line = {"metric": METRIC[routine], "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_string(cfg, args.gpus), "baseline_config": cfg["name"],
                           "flops_model": "2/3 N^3" + (" x 4 (complex)" if cplx else "") + (" + 2 N^2 (solve)" if routine == "pdgesv" else ""),
                           "value_by_reference_flop_model": flops_ref / (ms * 1e-3) / 1e12,
                           "reference_flop_model": "2/3 N^3 - 1/2 N^2 (+ 2 N^2 NRHS), pdludriver.f:913-918",
                           "l2": "inputs (>= 8 GiB per GPU) far exceed the 126 MB L2; no flush needed",
                           "pct_of_fp64_tensor_peak": 100.0 * value / (dmma_peak * args.gpus) if dmma_peak else None,
                           "fp64_dmma_peak_tflops_per_gpu": dmma_peak, "fp64_fma_peak_tflops_per_gpu": dfma_peak,
                           "sresid": sresid, "factor_ms": factor_ms, "solve_ms": solve_ms, "solve_first_call_ms": solve_first_ms,
                           "hbm_gbs_measured": peaks.get("hbm_gbs")},
                "parity_preflight": pre, "roofline": roof, "roofline_solve": roof_solve, "cpu_baseline": cpu, "e2e": e2e,
                "e2e_pageable": e2e_pageable, "gpu_launches": launches, "clocks": clocks}
        if next_rows is not None:
            line["next_rows"] = next_rows
        if prof:
            line["phase_profile_us"] = prof
        print(json.dumps(line), file=json_out, flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()
    assert sresid < 1.0, f"solve residual {sresid} of the timed workload exceeds the reference threshold"
    if next_thread is not None and next_thread.is_alive():      # a child that cannot be reaped must not keep this process from exiting
        sys.stdout.flush(); sys.stderr.flush()
        os._exit(0)


def next_rows_leg(world, elapsed, slack=20.0):
    """Runs bench_next.run_all in a daemon thread and waits for it at most its time limit + slack: (report, thread).  The driver allows
    870 s per run; a multi-GPU run leaves less room, and at 8 GPUs this leg normally does not fit at all.  A leg that does not come back
    is abandoned (the caller prints its line and leaves with os._exit while the thread is still alive)."""
    limit = min(240.0, 780.0 - elapsed) if world == 1 else min(150.0, 600.0 - elapsed)
    if limit < 40.0:
        return {"error": f"skipped: {elapsed:.0f} s of the run's time were used before this leg"}, None
    import threading
    box = {}

    def _work():
        try:
            import bench_next
            box["r"] = bench_next.run_all(per_row_timeout=60.0, total_timeout=limit)
        except Exception as ex:  # noqa: BLE001
            box["r"] = {"error": repr(ex)[:300]}
    th = threading.Thread(target=_work, daemon=True)
    th.start(); th.join(limit + slack)
    return box.get("r", {"error": "abandoned: the supplementary rows did not return within their limit"}), th


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="", choices=[""] + sorted(CONFIGS), help="BASELINE config (default: c2 at 1 GPU, weak = 32 GiB per GPU otherwise)")
    ap.add_argument("--size", dest="n", type=int, default=0, help="override N of the config")
    ap.add_argument("--nb", type=int, default=0, help="override NB of the config")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-pageable", action="store_true", help="e2e with a pinned host array only")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-preflight", action="store_true")
    ap.add_argument("--no-next", action="store_true", help="skip the supplementary SURVEY 8(f) measurements (bench_next.py, 1 GPU only)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--time-budget", type=int, default=640, help="seconds after which the optional pageable e2e leg is skipped")
    ap.add_argument("--profile", action="store_true")
    args = ap.parse_args()
    if args.gpus not in GRIDS:
        sys.exit("--gpus must be 1, 2, 4 or 8")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
