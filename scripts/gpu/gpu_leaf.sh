#!/bin/bash
# Call G: packet-mailbox leaf kernel: parity + panel timings + bench timeline.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_lu.py -m gpu -q -x -k "not variants and not packed" --timeout 300 --timeout-method=thread -p no:cacheprovider 2>&1 | tail -n 8 | tee gpurun_out/leaf_parity.log
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/panel_times.txt
import ctypes as C, numpy as np, torch, scalapack_b200 as S
I64 = C.c_int64
for m in (512, 2048, 4096, 8192, 16384, 32768, 65536):
    jb = 512
    a = torch.rand(m * jb, dtype=torch.float64, device="cuda") - 0.5
    ipiv = np.zeros(jb, np.int32); info = C.c_int(0)
    best = 1e9
    for rep in range(3):
        w = a.clone()
        ms = S.lib().slb200_test_panel(m, jb, S.api._ptr(w), I64(m), ipiv.ctypes.data_as(C.c_void_p), C.byref(info), 0)
        best = min(best, ms)
    print(f"panel {m}x{jb}: {best:.3f} ms  ({best * 1e3 / jb:.2f} us/column)")
PY
SLB200_LA_TRACE=1 timeout 300 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/trace_leaf.json 2> gpurun_out/trace_leaf.err
grep la_trace gpurun_out/trace_leaf.err | tail -n 24
python -c "
import json; d=json.loads([l for l in open('gpurun_out/trace_leaf.json') if l.startswith('{')][0]); print('bench', round(d['value'],3), round(d['ms_per_step'],1), round(d['roofline']['achieved'],2), round(d['roofline']['share_of_step'],3), d['config']['sresid'])"
