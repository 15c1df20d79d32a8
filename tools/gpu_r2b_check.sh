#!/bin/bash
# Round 2 (second session) check: full -m gpu suite, then the two batched benches (kernel changes: row-per-thread
# gradient kernel, triangular K-step trimming in the tile GEMM)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_gpu_r2b.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_r2b.log
grep -v "^$" gpurun_out/pytest_gpu_r2b.log | tail -12
for wl in cfg3 cfg2; do
  timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_r2b_$wl.json 2> gpurun_out/bench_r2b_$wl.err; echo "bench $wl exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_r2b_$wl.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'], d['roofline']['phase_ms_per_step'])"
done
