#!/bin/bash
# Round 2, second session: evidence run on one B200 after the gradient-kernel / triangular-trimming / left-looking
# changes.  DRAM-traffic table first (bench.py reads it for roofline.traffic), full -m gpu suite, every bench workload
# with the CPU reference beside it, latency tables, fit, ncu --set full summaries (made on the box, reports deleted).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total,driver_version --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 python tools/one_step.py cfg3 64 > gpurun_out/one_step_plain.log 2>&1 || { tail -5 gpurun_out/one_step_plain.log; exit 1; }
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_cfg3_b64.csv python tools/one_step.py cfg3 64 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
python tools/ncu_traffic.py gpurun_out/r02_traffic_cfg3_b64.csv cfg3 64 --update > gpurun_out/r02_traffic_cfg3_b64.txt; cp profiles/r02_dram_bytes.json gpurun_out/r02_dram_bytes.json
head -8 gpurun_out/r02_traffic_cfg3_b64.txt
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -16
timeout 120 python tools/fp64_peak.py > gpurun_out/fp64_peak.json 2>/dev/null
for wl in cfg3 cfg2 cfg4 cfg5; do
  timeout 900 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_$wl.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['phase_ms_per_step'], d['cpu_baseline'])"
done
timeout 600 python bench.py --impl reference > gpurun_out/bench_cfg3_reference_arm.json 2>/dev/null; tail -c 400 gpurun_out/bench_cfg3_reference_arm.json
echo "== latency"; timeout 300 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency.log
timeout 300 python tools/mid_batch.py 2>&1 | tee gpurun_out/mid_batch.log
REPS=8 timeout 120 python tools/hit_once.py 2>&1 | tail -1 | tee gpurun_out/hit_latency.log
echo "== fit"; timeout 900 python tools/fit_profile.py 2>&1 | tail -12 | tee gpurun_out/fit_profile.log
timeout 900 python tools/fit_cfg3.py 2>&1 | tail -1 | tee gpurun_out/fit_cfg3.json
# ---- ncu --set full
cap() {
  local name=$1 wl=$2 bb=$3; shift 3
  timeout 900 ncu --profile-from-start off --set full --clock-control none "$@" -o gpurun_out/$name python tools/one_step.py $wl $bb > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page details 2>/dev/null | grep -E "^  [a-zA-Z_].*\(|Duration|Throughput|Pipe|Warp Cycles Per Issued|Stall|No Eligible|Eligible Warps|Issued Warp|Registers Per|Theoretical Occ|Achieved Occ|L2 Hit|Bank conflicts|One or More Eligible" > gpurun_out/${name}_details.txt
  rm -f gpurun_out/$name.ncu-rep
}
cap r02b_full_inverse cfg3 16 --kernel-name-base mangled -k regex:'OpSyrk2|OpRecX|OpRecW' -s 8 -c 5
cap r02b_full_potrf cfg3 16 --kernel-name-base mangled -k regex:'OpSyrkE|OpPanel|diag_kernel' -s 70 -c 6
cap r02b_full_cov cfg3 16 -k regex:'grad_rows_kernel|build_kernel' -c 2
cap r02b_full_cfg2_cov cfg2 64 -k regex:'grad_rows_kernel|build_kernel' -c 2
timeout 120 python tools/one_step.py cfg5 16 > /dev/null 2>&1 && cap r02b_full_predict cfg5 16 -k regex:'gemm_nt_kernel|ks_build_kernel' -c 2
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b1_nlz.csv python tools/one_step.py cfg3 1 0 > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/r02_launches_b1_nlz.csv > gpurun_out/r02_launches_b1_nlz.txt
grep -E "Kernel Name|time_duration|fp64.avg|dmma|dram_throughput" gpurun_out/r02b_full_*_summary.txt | head -80
du -sh gpurun_out
