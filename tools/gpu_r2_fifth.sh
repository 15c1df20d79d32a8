#!/bin/bash
# Round 2, fifth GPU call: cached-solve latency, ncu evidence (kept small: gpurun_out must stay under 64 MiB).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "toggles or factor_cache or core_golden or potrf or gemm_nt or medium" > gpurun_out/pytest_gpu_subset.log 2>&1; tail -3 gpurun_out/pytest_gpu_subset.log
echo "== cached solve (3 rows, N=5000)"; REPS=8 timeout 120 python tools/hit_once.py 2>&1 | tail -1
echo "== cached solve GPB_PDL=0"; GPB_PDL=0 REPS=8 timeout 120 python tools/hit_once.py 2>&1 | tail -1
echo "== b1"; timeout 200 python tools/b1_latency.py 2>&1 | tail -1
# launch list + DRAM bytes of one cfg3 step at B=64 (tensor-map loader)
timeout 200 python tools/one_step.py cfg3 64 > gpurun_out/one_step_b64_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_traffic_cfg3_b64.csv python tools/one_step.py cfg3 64 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
python tools/ncu_traffic.py gpurun_out/r02_traffic_cfg3_b64.csv cfg3 64 > gpurun_out/r02_traffic_cfg3_b64.txt; tail -22 gpurun_out/r02_traffic_cfg3_b64.txt
# launch list of one B=1 nlZ-only evaluation
timeout 100 python tools/one_step.py cfg3 1 0 > /dev/null 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b1_nlz.csv python tools/one_step.py cfg3 1 0 > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/r02_launches_b1_nlz.csv | tee gpurun_out/r02_launches_b1_nlz.txt | head -14
# --set full, few kernels per report (each kernel is ~3.6 MB)
timeout 120 python tools/one_step.py cfg3 16 > gpurun_out/one_step_plain.log 2>&1 && \
timeout 900 ncu --profile-from-start off --set full --clock-control none --kernel-name-base mangled -k regex:'OpSyrk2|OpRecX|OpRecW' -s 8 -c 5 -o gpurun_out/r02_full_inverse python tools/one_step.py cfg3 16 > gpurun_out/ncu_full_a.log 2>&1; echo "ncu a exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --kernel-name-base mangled -k regex:'OpSyrkE|OpPanel|diag_kernel' -s 70 -c 5 -o gpurun_out/r02_full_potrf python tools/one_step.py cfg3 16 > gpurun_out/ncu_full_b.log 2>&1; echo "ncu b exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none -k regex:'grad_kernel' -c 1 -o gpurun_out/r02_full_grad python tools/one_step.py cfg3 16 > gpurun_out/ncu_full_c.log 2>&1; echo "ncu c exit $?"
rm -f gpurun_out/*.log.bak; du -sh gpurun_out; ls -la gpurun_out/*.ncu-rep
