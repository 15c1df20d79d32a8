#!/bin/bash
# parity tests, then the ncu launch list and one full capture of the tile GEMM
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -rA 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log | tail -45
CMD="python bench.py --steps 1 --warmup 3 --batch 8"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 780 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"; tail -2 gpurun_out/plain.log | cut -c1-400
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 2600 -c 3 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | head -20
