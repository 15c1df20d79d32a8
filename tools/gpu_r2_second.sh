#!/bin/bash
# Round 2, second GPU call: full -m gpu suite again, quadrature debug, FP64 peak artefact, every bench workload,
# the per-launch DRAM-traffic table of one cfg3 step.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -25
timeout 600 python tools/debug_quadnoise.py > gpurun_out/debug_quadnoise.log 2>&1; tail -40 gpurun_out/debug_quadnoise.log
timeout 120 python tools/fp64_peak.py > gpurun_out/fp64_peak.json 2> gpurun_out/fp64_peak.err; cat gpurun_out/fp64_peak.json
for wl in cfg3 cfg2 cfg4 cfg5; do
  timeout 900 python bench.py --workload $wl > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?"; tail -c 3000 gpurun_out/bench_$wl.json; tail -3 gpurun_out/bench_$wl.err
done
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/traffic_cfg3_b64.csv python tools/one_step.py cfg3 64 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu exit $?"; tail -2 gpurun_out/ncu_traffic.log
python tools/ncu_traffic.py gpurun_out/traffic_cfg3_b64.csv cfg3 64 | tail -30
