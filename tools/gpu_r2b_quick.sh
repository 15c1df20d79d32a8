#!/bin/bash
# quick check after a gradient-kernel change: gradient parity tests + the two batched benches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "core_golden or lownoise or oracle_medium or large_rq or tile_boundaries or fuzz or wide_inputs" > gpurun_out/pytest_gpu_quick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_quick.log
grep -v "^$" gpurun_out/pytest_gpu_quick.log | tail -4
for wl in cfg3 cfg2; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_quick_$wl.json 2> gpurun_out/bench_quick_$wl.err; echo "bench $wl exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_quick_$wl.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['phase_ms_per_step'])"
done
