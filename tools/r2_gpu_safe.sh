#!/bin/sh
# developer helper (runs ON the GPU box): a small parity subset under a short timeout first (a kernel that hangs must not eat
# the call's limit), then the whole suite, then the A/B timing
tag=${1:-r2x}
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest_small.log 2>&1
rc=$?
tail -4 gpurun_out/${tag}_pytest_small.log
[ $rc -ne 0 ] && { echo "small parity run failed (rc $rc)"; grep -m5 "Error\|error\|assert" gpurun_out/${tag}_pytest_small.log; exit 1; }
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log
shift
sh tools/r2_gpu_ab.sh $tag 0 "$@"
