#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "robustness or chunking or factor_cache or core_golden" > gpurun_out/pytest_gpu_w.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_w.log
grep -v "^$" gpurun_out/pytest_gpu_w.log | tail -8
echo "== fit"; timeout 900 python tools/fit_profile.py 2>&1 | tail -12 | tee gpurun_out/fit_profile_w.log
timeout 900 python tools/fit_cfg3.py 2>&1 | tail -1 | tee gpurun_out/fit_cfg3_w.json
