#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 8"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:diag_kernel -s 130 -c 2 -o gpurun_out/prof_diag $CMD > gpurun_out/ncu_diag.log 2>&1
echo "ncu diag exit $?"; tail -2 gpurun_out/ncu_diag.log | cut -c1-200
