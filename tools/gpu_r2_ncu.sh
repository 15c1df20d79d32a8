#!/bin/bash
# ncu --set full evidence, summarised ON THE BOX (a kernel's report is 6-16 MB; gpurun_out must stay < 64 MiB).
mkdir -p gpurun_out
cap() {   # name, extra ncu args..., then -- command
  local name=$1; shift
  timeout 900 ncu --profile-from-start off --set full --clock-control none "$@" -o gpurun_out/$name python tools/one_step.py cfg3 16 > gpurun_out/ncu_$name.log 2>&1
  echo "ncu $name exit $?"
  python tools/ncu_summary.py gpurun_out/$name.ncu-rep > gpurun_out/${name}_summary.txt 2>&1
  ncu -i gpurun_out/$name.ncu-rep --page details 2>/dev/null | grep -E "^  [a-zA-Z_].*\(|Duration|Throughput|Pipe|Warp Cycles Per Issued|Stall|No Eligible|Eligible Warps|Issued Warp|Registers Per|Theoretical Occ|Achieved Occ|L2 Hit|DRAM Throughput|Bank conflicts|One or More Eligible" > gpurun_out/${name}_details.txt
  if [ "$name" == "r02_full_grad" ]; then ncu -i gpurun_out/$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/${name}_source.csv.gz; fi
  rm -f gpurun_out/$name.ncu-rep
}
timeout 120 python tools/one_step.py cfg3 16 > gpurun_out/one_step_plain.log 2>&1 || exit 1
cap r02_full_inverse --kernel-name-base mangled -k regex:'OpSyrk2|OpRecX|OpRecW' -s 8 -c 5
cap r02_full_potrf --kernel-name-base mangled -k regex:'OpSyrkE|OpPanel|diag_kernel' -s 70 -c 6
cap r02_full_grad -k regex:"grad_kernel|build_kernel" -c 2 --import-source on
# launch list + DRAM bytes of one cfg3 step at B=64 (tensor-map loader), and of one B=1 nlZ-only evaluation
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_cfg3_b64.csv python tools/one_step.py cfg3 64 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
python tools/ncu_traffic.py gpurun_out/r02_traffic_cfg3_b64.csv cfg3 64 > gpurun_out/r02_traffic_cfg3_b64.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b1_nlz.csv python tools/one_step.py cfg3 1 0 > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/r02_launches_b1_nlz.csv > gpurun_out/r02_launches_b1_nlz.txt
du -sh gpurun_out; ls -la gpurun_out | head -30
# outer-block sweep with the tensor-map loader (B=64, cfg3)
for ob in 3 4 6 8; do echo -n "GPB_OUTER_BLOCK=$ob: "; GPB_OUTER_BLOCK=$ob timeout 300 python bench.py --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['phase_ms_per_step']['factor'])"; done
