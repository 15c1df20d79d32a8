#!/bin/bash
# final evidence for the round: gpu tests, default bench, ncu launch list of the default bench command
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; cut -c1-300 gpurun_out/bench_default.json
python -c "import json;d=json.load(open('gpurun_out/bench_default.json'));print('launches per run', d['gpu_launches'], 'steps', d['steps'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline'])"
CMD="python bench.py --steps 1 --warmup 3"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 648 -c 260 --csv --log-file gpurun_out/launches_default.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
python tools/launch_summary.py gpurun_out/launches_default.csv
