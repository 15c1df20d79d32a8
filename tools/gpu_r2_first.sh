#!/bin/bash
# Round 2, first GPU call: the whole -m gpu suite (with durations), small-batch latency baseline, default bench.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
free -g > gpurun_out/host.txt; nproc >> gpurun_out/host.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=25 -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -60
timeout 300 python tools/b1_latency.py > gpurun_out/b1_latency.log 2>&1; cat gpurun_out/b1_latency.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -2 gpurun_out/bench_default.json
