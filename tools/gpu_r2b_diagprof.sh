#!/bin/bash
mkdir -p gpurun_out
GPB_DIAG_DBG=1 timeout 120 python tools/diag_dbg.py 2>&1 | tail -4 | tee gpurun_out/diag_dbg_r2b.txt
timeout 120 python tools/diag_bench.py 2>&1 | tee -a gpurun_out/diag_dbg_r2b.txt
timeout 600 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'diag_kernel' -s 5 -c 1 -o gpurun_out/r02b_diag python tools/one_step.py cfg3 1 0 > gpurun_out/ncu_r02b_diag.log 2>&1; echo "ncu exit $?"
ncu -i gpurun_out/r02b_diag.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02b_diag_source.csv.gz
python tools/ncu_summary.py gpurun_out/r02b_diag.ncu-rep > gpurun_out/r02b_diag_summary.txt 2>&1
rm -f gpurun_out/r02b_diag.ncu-rep
ls -la gpurun_out/r02b_diag*
