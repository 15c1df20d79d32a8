#!/bin/bash
# parity files + the two batched benches + ncu --set full of the gradient kernel (cfg3 B=16, cfg2 B=64)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -m gpu -q -x --timeout 600 -k "not 8192" > gpurun_out/pytest_gpu_r2b2.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2b2.log
grep -v "^$" gpurun_out/pytest_gpu_r2b2.log | tail -6
for wl in cfg3 cfg2; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_r2b2_$wl.json 2> gpurun_out/bench_r2b2_$wl.err; echo "bench $wl exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_r2b2_$wl.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['phase_ms_per_step'])"
done
cap() {
  local name=$1 wl=$2 bb=$3; shift 3
  timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none "$@" -o gpurun_out/$name python tools/one_step.py $wl $bb > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page details 2>/dev/null | grep -E "^  [a-zA-Z_].*\(|Duration|Throughput|Pipe|Warp Cycles Per Issued|Stall|No Eligible|Eligible Warps|Issued Warp|Registers Per|Theoretical Occ|Achieved Occ|L2 Hit|Bank conflicts|One or More Eligible" > gpurun_out/${name}_details.txt
  ncu -i gpurun_out/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${name}_source.csv.gz
  rm -f gpurun_out/$name.ncu-rep
}
cap r02b_full_grad cfg3 16 -k regex:'grad_rows_kernel' -c 1
cap r02b_full_grad_cfg2 cfg2 64 -k regex:'grad_rows_kernel' -c 1
grep -E "Kernel Name|time_duration|fp64|dram_throughput|warps_active" gpurun_out/r02b_full_grad_summary.txt gpurun_out/r02b_full_grad_cfg2_summary.txt
