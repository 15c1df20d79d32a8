#!/bin/bash
# Round 2, third GPU call: lanes / multi-CTA prep / tensor-map loader / predict_full fix.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -30
timeout 300 python tools/debug_predict_full.py > gpurun_out/debug_predict_full.log 2>&1; tail -20 gpurun_out/debug_predict_full.log
echo "== b1 latency (lanes on)"; timeout 300 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency.log
echo "== b1 latency (GPB_LANES=0)"; GPB_LANES=0 timeout 300 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency_nolanes.log
echo "== mid batch"; timeout 300 python tools/mid_batch.py 2>&1 | tee gpurun_out/mid_batch.log
for ld in auto auto_bulk tensor; do
  echo "== bench cfg3 GPB_LOADER=$ld"; GPB_LOADER=$ld timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_cfg3_$ld.json 2> gpurun_out/bench_cfg3_$ld.err; python -c "
import json,sys
d=json.load(open('gpurun_out/bench_cfg3_$ld.json'))
print(d['value'], d['roofline']['phase_ms_per_step'])"
done
echo "== bench cfg2"; timeout 600 python bench.py --workload cfg2 --no-cpu-baseline > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; python -c "
import json
d=json.load(open('gpurun_out/bench_cfg2.json'))
print(d['value'], d['roofline']['phase_ms_per_step'])"
echo "== fit cfg3"; timeout 900 python tools/fit_cfg3.py 2>&1 | tail -1 | tee gpurun_out/fit_cfg3.json
# ncu --set full evidence (B=16): inverse-phase ops + covariance kernels, then a window of the potrf's OpSyrk launches
timeout 120 python tools/one_step.py cfg3 16 > gpurun_out/one_step_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'OpSyrk2|OpRecX|OpRecW|grad_kernel|build_kernel' -c 15 -o gpurun_out/r02_full_inverse python tools/one_step.py cfg3 16 > gpurun_out/ncu_full_a.log 2>&1; echo "ncu a exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'OpSyrk,|OpPanel|diag_kernel' -s 60 -c 12 -o gpurun_out/r02_full_potrf python tools/one_step.py cfg3 16 > gpurun_out/ncu_full_b.log 2>&1; echo "ncu b exit $?"
timeout 120 python tools/one_step.py cfg2 64 > gpurun_out/one_step_cfg2_plain.log 2>&1 && \
timeout 600 ncu --profile-from-start off --set full --clock-control none -k regex:'grad_kernel|build_kernel' -c 2 -o gpurun_out/r02_full_cfg2_cov python tools/one_step.py cfg2 64 > gpurun_out/ncu_full_c.log 2>&1; echo "ncu c exit $?"
ls -la gpurun_out/*.ncu-rep
