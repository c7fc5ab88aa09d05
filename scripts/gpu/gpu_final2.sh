#!/bin/bash
# Last check of the round: parity suite on the pruned build (cp.async kernel + packed kernel only).
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -q -x -k "not packed or (packed and opts0 and shape0)" --timeout 90 --timeout-method=thread -p no:cacheprovider > gpurun_out/final2_tests.log 2>&1
echo "tests exit=$? $(tail -n 1 gpurun_out/final2_tests.log)"
