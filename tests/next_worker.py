"""One rank of a parity run of the SURVEY 8(f) entry points (tests/next_cases.py) on a P x Q grid.
Default: the product library on the GPU(s) (spawned by tests/test_gpu_next.py).  With SLB200_EMUL=1: the HOST-LOGIC emulation
(tests/emul: the same sources compiled against a stub CUDA runtime) on the CPU -- test infrastructure, never the product path."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import scalapack_b200.api as api  # noqa: E402

EMUL = os.environ.get("SLB200_EMUL") == "1"
if EMUL:
    api._SO = os.path.join(ROOT, "tests", "emul", "libslb_emul.so")
import scalapack_b200 as S  # noqa: E402
import next_cases  # noqa: E402


def main():
    spec = json.loads(sys.argv[1])
    if EMUL:
        assert S.lib().slb200_is_emulation() == 1
    else:
        assert S.has_cuda(), "the product library needs a B200 (no CPU fallback)"
    me, np_ = S.blacs_pinfo()
    ctx = S.blacs_gridinit(S.blacs_get(-1, 0), "Row-major", spec["P"], spec["Q"])
    cases = getattr(next_cases, spec["cases"]) if isinstance(spec["cases"], str) else spec["cases"]
    results = next_cases.run(S, ctx, cases)
    print("RESULT" + json.dumps({"rank": me, "results": results}), flush=True)
    S.blacs_exit(0)


if __name__ == "__main__":
    main()
