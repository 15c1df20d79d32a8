#!/bin/bash
# parity tests, bench, then the ncu launch list and one full capture of the tile GEMM
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg3.json 2>gpurun_out/bench_cfg3.err; cut -c1-400 gpurun_out/bench_cfg3.json
timeout 300 python tools/ob_sweep.py 2>&1 | grep "OB=4"
CMD="python bench.py --steps 1 --warmup 3 --batch 8"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 790 -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
python tools/launch_summary.py gpurun_out/launches.csv
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 380 -c 6 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -2 gpurun_out/ncu_full.log | cut -c1-200
