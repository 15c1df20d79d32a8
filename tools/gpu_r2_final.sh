#!/bin/bash
# Round 2 evidence run on one B200: full -m gpu suite, every bench workload (with the CPU reference arm), latency
# tables, fit profile, launch / DRAM-traffic tables and ncu --set full summaries (made on the box, reports deleted:
# gpurun_out must stay under 64 MiB).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,driver_version --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -16
timeout 120 python tools/fp64_peak.py > gpurun_out/fp64_peak.json 2>/dev/null
for wl in cfg3 cfg2 cfg4 cfg5; do
  timeout 900 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['phase_ms_per_step'], d['cpu_baseline'])"
done
timeout 600 python bench.py --impl reference > gpurun_out/bench_cfg3_reference_arm.json 2>/dev/null; tail -c 600 gpurun_out/bench_cfg3_reference_arm.json
timeout 600 python bench.py --scaling strong --batch 1024 --no-cpu-baseline > gpurun_out/bench_cfg3_strong_b1024_1.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg3_strong_b1024_1.json').read().strip().splitlines()[-1]); print('strong b1024 x1', d['value'], d['ms_per_step'])"
echo "== latency"; timeout 300 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency.log
timeout 300 python tools/mid_batch.py 2>&1 | tee gpurun_out/mid_batch.log
REPS=8 timeout 120 python tools/hit_once.py 2>&1 | tail -1 | tee gpurun_out/hit_latency.log
timeout 300 python tools/predict_latency.py 2>&1 | tail -3 | tee gpurun_out/predict_latency.log
timeout 300 python tools/append_latency.py 2>&1 | tail -5 | tee gpurun_out/append_latency.log
echo "== fit"; timeout 900 python tools/fit_profile.py 2>&1 | tail -12 | tee gpurun_out/fit_profile.log
timeout 900 python tools/fit_cfg3.py 2>&1 | tail -1 | tee gpurun_out/fit_cfg3.json
# ---- ncu
cap() {
  local name=$1; shift
  timeout 900 ncu --profile-from-start off --set full --clock-control none "$@" -o gpurun_out/$name python tools/one_step.py cfg3 16 > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page details 2>/dev/null | grep -E "^  [a-zA-Z_].*\(|Duration|Throughput|Pipe|Warp Cycles Per Issued|Stall|No Eligible|Eligible Warps|Issued Warp|Registers Per|Theoretical Occ|Achieved Occ|L2 Hit|Bank conflicts|One or More Eligible" > gpurun_out/${name}_details.txt
  rm -f gpurun_out/$name.ncu-rep
}
timeout 120 python tools/one_step.py cfg3 16 > gpurun_out/one_step_plain.log 2>&1 || exit 1
cap r02_full_inverse --kernel-name-base mangled -k regex:'OpSyrk2|OpRecX|OpRecW' -s 8 -c 5
cap r02_full_potrf --kernel-name-base mangled -k regex:'OpSyrkE|OpPanel|diag_kernel' -s 70 -c 6
cap r02_full_cov -k regex:'grad_kernel|build_kernel' -c 2
timeout 120 python tools/one_step.py cfg5 16 > /dev/null 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'gemm_nt_kernel|ks_build_kernel' -c 2 -o gpurun_out/r02_full_predict python tools/one_step.py cfg5 16 > gpurun_out/ncu_r02_full_predict.log 2>&1; python tools/ncu_summary.py gpurun_out/r02_full_predict.ncu-rep > gpurun_out/r02_full_predict_summary.txt 2>&1; rm -f gpurun_out/r02_full_predict.ncu-rep
timeout 120 python tools/one_step.py cfg2 64 > /dev/null 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'grad_kernel|build_kernel' -c 2 -o gpurun_out/r02_full_cfg2_cov python tools/one_step.py cfg2 64 > gpurun_out/ncu_r02_full_cfg2_cov.log 2>&1; python tools/ncu_summary.py gpurun_out/r02_full_cfg2_cov.ncu-rep > gpurun_out/r02_full_cfg2_cov_summary.txt 2>&1; rm -f gpurun_out/r02_full_cfg2_cov.ncu-rep
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_cfg3_b64.csv python tools/one_step.py cfg3 64 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
python tools/ncu_traffic.py gpurun_out/r02_traffic_cfg3_b64.csv cfg3 64 > gpurun_out/r02_traffic_cfg3_b64.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b1_nlz.csv python tools/one_step.py cfg3 1 0 > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/r02_launches_b1_nlz.csv > gpurun_out/r02_launches_b1_nlz.txt
du -sh gpurun_out
