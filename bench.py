#!/usr/bin/env python
"""bench.py -- PDGETRF FP64 TFLOP/s (2/3 N^3) through the C-ABI drop-in, on N GPUs of one node.

  python bench.py --gpus 1 --steps K --warmup W            # N=65536 NB=512, 1x1 grid (BASELINE config 2)
  torchrun ... bench.py --gpus N                           # 1x2 / 2x2 / 2x4 grids, 32 GiB of A per GPU
  python bench.py --impl reference                         # the reference algorithm on the host cores

A step = one PDGETRF of the freshly regenerated matrix (64-bit LCG generator, on the device).  `value` is the
device-timed factorisation with A resident in HBM; `e2e` is the same call with A in pinned HOST memory
(H2D + factor + D2H inside the timed region).  The matrix (>= 32 GiB per GPU) is far larger than the 126 MB
L2, so no explicit L2 flush is needed between steps.
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


# --------------------------------------------------------------------------- reference arm (CPU)
def cpu_sample(nb, target_s=15.0, threads=None):
    """The oracle port of the reference algorithm (serial panel + BLAS-3 update, oracle/oracle.c) on the host
    cores, on a bounded sample: a full factorisation of the same generator's matrix at a smaller N."""
    import numpy as np
    import oracle as O
    cores = os.cpu_count() or 1
    threads = threads or cores
    O.set_threads(threads)
    a = O.matgen64_tile(2048, A_SEED, 0, 2048, 0, 2048)
    t = time.perf_counter(); _ = a @ a; dt = time.perf_counter() - t
    rate = 2 * 2048 ** 3 / dt                                   # host dgemm flop/s probe (numpy shares the BLAS)
    n = int((1.5 * target_s * rate * 0.6) ** (1.0 / 3.0)) // nb * nb
    n = max(2 * nb, min(n, 16384))
    a = O.matgen64_tile(n, A_SEED, 0, n, 0, n)
    t = time.perf_counter()
    ipiv, info = O.getrf(a, nb)
    dt = time.perf_counter() - t
    tf = (2.0 / 3.0) * n ** 3 / dt / 1e12
    return {"value": tf, "unit": "TFLOP/s", "cores": threads, "kind": "port",
            "sample": f"full PDGETRF restatement (oracle/oracle.c + OpenBLAS {threads} threads) of the same generator's "
                      f"matrix at N={n} NB={nb}: {dt:.2f} s; the reference (Fortran+MPI) cannot be built in this image",
            "n": n, "seconds": dt, "info": int(info)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    res = []
    for _ in range(args.steps):
        res.append(cpu_sample(args.nb, target_s=float(os.environ.get("SLB200_BENCH_CPU_TARGET_S", max(3.0, 12.0 / max(1, args.steps))))))
    best = max(res, key=lambda r: r["value"])
    P, Q = GRIDS[args.gpus]
    n_full = workload_n(args)
    line = {"impl": "reference", "metric": "pdgetrf_fp64_tflops", "value": best["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": best["seconds"] * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"PDGETRF N={n_full} NB={args.nb} grid {P}x{Q}", "timed_sample": best["sample"]},
            "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": best["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- our arm (GPU)
def workload_n(args):
    if args.n:
        return args.n
    base = 65536                                               # 32 GiB of A per GPU at every grid (weak scaling)
    n = int(base * math.sqrt(args.gpus)) // args.nb * args.nb
    return n


def run_ours(args):
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"                     # keep NCCL's version banner off stdout (one JSON line only)
    import numpy as np
    import torch
    import scalapack_b200 as S
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} needs WORLD_SIZE={args.gpus} (torchrun), got {world}"
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_
    P, Q = GRIDS[args.gpus]
    n, nb = workload_n(args), args.nb
    ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", P, Q)
    _, _, myrow, mycol = S.blacs_gridinfo(ctx)
    mloc, nloc = S.numroc(n, nb, myrow, 0, P), S.numroc(n, nb, mycol, 0, Q)
    lld = max(1, mloc)
    desca, info = S.descinit(n, n, nb, nb, 0, 0, ctx, lld)
    assert info == 0
    A = torch.empty(nloc * lld, dtype=torch.float64, device="cuda")       # column-major local array in HBM
    ipiv = np.zeros(mloc + nb, np.int32)
    flops = (2.0 / 3.0) * float(n) ** 3

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
        S.matgen64(ctx, n, n, nb, nb, A, lld, A_SEED)            # untimed: regenerate the matrix in HBM
        barrier()
        inf = S.pdgetrf(n, n, A, 1, 1, desca, ipiv)
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

    # ---- correctness of the timed workload: solve with the factors, reference residual on regenerated A, b ----
    descb, _ = S.descinit(n, 1, nb, 1, 0, 0, ctx, lld)
    X = torch.zeros(max(1, lld), dtype=torch.float64, device="cuda")
    S.matgen64(ctx, n, 1, nb, 1, X, lld, B_SEED)                 # b = column 0 of the generator with B_SEED
    inf = S.pdgetrs("N", n, 1, A, 1, 1, desca, ipiv, X, 1, 1, descb)
    assert inf == 0
    solve_ms = solve_first_ms = maxr(S.last_solve_ms())
    if world > 1:            # the first multi-GPU solve pays NCCL's lazy connection set-up of the world communicator
        S.matgen64(ctx, n, 1, nb, 1, X, lld, B_SEED)
        inf = S.pdgetrs("N", n, 1, A, 1, 1, desca, ipiv, X, 1, 1, descb)
        assert inf == 0
        solve_ms = maxr(S.last_solve_ms())
    sresid = S.pdlaschk(ctx, n, 1, X, descb, desca, A_SEED, B_SEED, gen=64)

    # ---- roofline of the dominant kernel (trailing update, FP64 DMMA): live CUDA-event time over the timed steps ----
    u_ms = sum(u[0] for u in upd); u_fl = sum(u[1] for u in upd); u_n = sum(u[2] for u in upd)
    dmma_peak = S.lib().slb200_bench_dmma_tflops(20000)
    dfma_peak = S.lib().slb200_bench_dfma_tflops(20000)
    achieved = u_fl / (u_ms * 1e-3) / 1e12 if u_ms > 0 else None
    # DRAM traffic of the update kernel: one `ncu --set full` capture (profiles/r01_gemm_v9_ncu.md, M=N=32768, K=512) read
    # 11.48 GB and wrote 8.54 GB for 17.45 GB of algorithmic bytes (16*m*n for C + the operands) = 1.147x; scaled here to
    # the average launch of this run (algorithmic C bytes of a launch = its flops * 8 / NB).
    traffic = (u_fl / u_n) * 8.0 / nb * 1.147 if u_n else None
    roof = {"bound": "tensor", "kernel": "dgemm_minus_packed (FP64 DMMA trailing update, gemm_packed.cu; dgemm_minus_p8b for m < 3072)",
            "achieved": achieved, "peak": dmma_peak,
            "unit": "TFLOP/s", "frac": (achieved / dmma_peak) if achieved else None, "traffic": traffic,
            "traffic_unit": "bytes per average launch = algorithmic C bytes x 1.147 (measured dram read+write / algorithmic bytes, ncu capture at M=N=32768)",
            "peak_source": "FP64 DMMA peak measured live by slb200_bench_dmma_tflops (MEASURED_PEAKS.json has no FP64 entry)",
            "fp64_fma_peak_tflops": dfma_peak, "share_of_step": u_ms / sum(times) if times else None,
            "launches": u_n, "avg_launch_ms": u_ms / u_n if u_n else None}

    # ---- e2e: the same call with A in pinned HOST memory (H2D + factor + D2H inside the timed region) ----
    e2e = None
    if not args.no_e2e:
        try:
            # every rank of this node pins a host copy of its local array: refuse rather than drive the box out of memory
            need = nloc * lld * 8 * world
            try:
                import psutil
                avail = psutil.virtual_memory().available
            except Exception:
                avail = None
            if avail is not None and need > 0.7 * avail:
                raise MemoryError(f"pinned host copies of A need {need / 2**30:.0f} GiB, {avail / 2**30:.0f} GiB of host memory available")
            Ah = torch.empty(nloc * lld, dtype=torch.float64, pin_memory=True)
            e_times = []
            for _ in range(max(1, min(args.steps, args.e2e_steps))):
                S.matgen64(ctx, n, n, nb, nb, A, lld, A_SEED)
                Ah.copy_(A); barrier()
                t0 = time.perf_counter()
                inf = S.pdgetrf(n, n, Ah.numpy(), 1, 1, desca, ipiv)
                barrier()
                e_times.append(maxr(time.perf_counter() - t0))
                assert inf == 0
            e_s = sum(e_times) / len(e_times)
            nbytes = nloc * lld * 8
            e2e = {"value": flops / e_s / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": nbytes * world, "d2h_bytes_per_step": nbytes * world + 4 * n,
                   "seconds": e_s, "steps": len(e_times), "host_memory": "pinned"}
            del Ah
        except Exception as ex:  # host RAM for a pinned copy of A may be missing
            e2e = {"value": None, "unit": "TFLOP/s", "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "error": repr(ex)[:200]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_sample(nb)
    if rank == 0:
        peaks = measured_peaks()
        line = {"metric": "pdgetrf_fp64_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"PDGETRF N={n} NB={nb} grid {P}x{Q} ({nloc * lld * 8 / 2**30:.1f} GiB of A per GPU), 64-bit LCG matrix",
                           "flops_model": "2/3 N^3", "l2": "inputs (>= 32 GiB) far exceed the 126 MB L2; no flush needed",
                           "pct_of_fp64_tensor_peak": 100.0 * value / (dmma_peak * args.gpus) if dmma_peak else None,
                           "fp64_dmma_peak_tflops_per_gpu": dmma_peak, "fp64_fma_peak_tflops_per_gpu": dfma_peak,
                           "sresid": sresid, "solve_ms": solve_ms, "solve_first_call_ms": solve_first_ms, "hbm_gbs_measured": peaks.get("hbm_gbs")},
                "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if prof:
            line["phase_profile_us"] = prof
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier(); dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=0, help="override N (default: 65536*sqrt(gpus), BASELINE config 2 at 1 GPU)")
    ap.add_argument("--nb", type=int, default=512)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
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
