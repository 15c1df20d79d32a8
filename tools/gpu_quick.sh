#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -15
timeout 600 python tools/microbench.py 2>&1 | grep wl | cut -c1-330
CMD="python bench.py --steps 1 --warmup 3 --batch 8"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 780 -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 372 -c 8 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
