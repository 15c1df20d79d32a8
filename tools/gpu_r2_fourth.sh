#!/bin/bash
# Round 2, fourth GPU call: PDL + quarter tiles + fused gradient distances + tensor maps in predict.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -12
echo "== b1 latency (default: PDL + quarter tiles)"; timeout 300 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency.log
echo "== b1 latency GPB_PDL=0"; GPB_PDL=0 timeout 300 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency_nopdl.log
echo "== b1 latency GPB_QUARTER=0"; GPB_QUARTER=0 timeout 300 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency_noquarter.log
echo "== mid batch"; timeout 300 python tools/mid_batch.py 2>&1 | tee gpurun_out/mid_batch.log
for wl in cfg3 cfg2 cfg5; do
  echo "== bench $wl"; timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], d['roofline']['frac'], d['roofline']['phase_ms_per_step'])"
done
echo "== bench cfg3 GPB_PDL=0"; GPB_PDL=0 timeout 600 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['phase_ms_per_step'])"
echo "== fit profile"; timeout 900 python tools/fit_profile.py 2>&1 | tail -12 | tee gpurun_out/fit_profile.log
# ncu --set full evidence (B=16)
timeout 120 python tools/one_step.py cfg3 16 > gpurun_out/one_step_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none --kernel-name-base mangled -k regex:'OpSyrk2|OpRecX|OpRecW' -c 13 -o gpurun_out/r02_full_inverse2 python tools/one_step.py cfg3 16 > gpurun_out/ncu_full_a.log 2>&1; echo "ncu a exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --kernel-name-base mangled -k regex:'OpSyrkE|OpPanel|diag_kernel' -s 66 -c 12 -o gpurun_out/r02_full_potrf python tools/one_step.py cfg3 16 > gpurun_out/ncu_full_b.log 2>&1; echo "ncu b exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'grad_kernel' -c 1 -o gpurun_out/r02_full_grad python tools/one_step.py cfg3 16 > gpurun_out/ncu_full_c.log 2>&1; echo "ncu c exit $?"
ls -la gpurun_out/*.ncu-rep | tail -5
