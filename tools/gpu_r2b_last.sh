#!/bin/bash
# last 1-GPU checks of the round: dependent-issue latencies, compute-sanitizer on the smoke problem, full -m gpu suite
mkdir -p gpurun_out
[ -x tools/micro/lat ] && tools/micro/lat | tee gpurun_out/micro_lat.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_r2b.log 2>&1; echo "sanitizer exit $?"; grep -E "ERROR SUMMARY|smoke:" gpurun_out/sanitizer_r2b.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -5
