#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --batch 8"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"build_kernel|grad_kernel" -s 6 -c 2 -o gpurun_out/prof_cov $CMD > gpurun_out/ncu_cov.log 2>&1
echo "ncu cov exit $?"; tail -2 gpurun_out/ncu_cov.log | cut -c1-200
CMD2="python bench.py --steps 1 --warmup 3 --workload cfg5"
$CMD2 > gpurun_out/plain5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ks_build_kernel|OpPred" -s 200 -c 2 -o gpurun_out/prof_pred $CMD2 > gpurun_out/ncu_pred.log 2>&1
echo "ncu pred exit $?"; tail -2 gpurun_out/ncu_pred.log | cut -c1-200
