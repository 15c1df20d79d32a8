#!/bin/bash
# evidence for the round: gpu tests, benches, ncu launch list of the default bench command,
# one ncu --set full capture of the dominant kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_tail.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for wl in cfg3 cfg2 cfg5; do
  timeout 600 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"
  python -c "import json;d=json.load(open('gpurun_out/bench_$wl.json'));print(d['metric'], round(d['value'],1), d['unit'], '| e2e', round(d['e2e']['value'],1), '| roofline', round(d['roofline']['achieved'],2), round(d['roofline']['frac'],3), '| cpu', d['cpu_baseline']['value'], '| launches', d['gpu_launches'], '| clocks', d['clocks'])"
done
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/bench_ref.json
CMD="python bench.py --steps 1 --warmup 3"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
python tools/launch_summary.py gpurun_out/launches_default.csv
if [ -n "$NCU_FULL" ]; then
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_nt_kernel -s 371 -c 8 -o gpurun_out/prof_gemm $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; tail -1 gpurun_out/ncu_full.log | cut -c1-120
fi
timeout 300 python tools/b1_latency.py 2>&1 | tail -3 | tee gpurun_out/b1_latency.txt
timeout 100 python tools/diag_dbg.py 2>&1 | grep cycles | tail -1 | tee gpurun_out/diag_phases.txt
timeout 300 python tools/mid_batch.py 2>&1 | tail -1 | tee gpurun_out/mid_batch.txt
timeout 300 python tools/append_latency.py 2>&1 | tail -4 | tee gpurun_out/append_latency.txt
timeout 600 python tools/fit_profile.py 2>&1 | tail -9 | tee gpurun_out/fit_profile.txt
