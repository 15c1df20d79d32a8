#!/bin/bash
# One gpurun call: parity tests, micro-measurements, a short bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
GPB_GEMM_BN=128 timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "gemm or potrf or oracle_medium" > gpurun_out/pytest_gpu_bn128.log 2>&1; echo "pytest bn128 exit $?"
tail -5 gpurun_out/pytest_gpu_bn128.log
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; echo "microbench exit $?"
cat gpurun_out/microbench.log | tail -50
BN=128 timeout 600 python tools/microbench.py > gpurun_out/microbench_bn128.log 2>&1; echo "microbench bn128 exit $?"
grep wl gpurun_out/microbench_bn128.log | tail -12
timeout 900 python bench.py --steps 2 --warmup 3 --batch 16 > gpurun_out/bench_b16.log 2>&1; echo "bench exit $?"
tail -5 gpurun_out/bench_b16.log
