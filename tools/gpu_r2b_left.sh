#!/bin/bash
# left-looking inner updates: bit-identity test, then cfg3 / cfg2 benches per setting
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "toggles and (env9 or env10 or env11)" > gpurun_out/pytest_gpu_left.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_left.log
grep -v "^$" gpurun_out/pytest_gpu_left.log | tail -4
run() {
  local tag=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_left_${tag}_$wl.json 2> gpurun_out/bench_left_${tag}_$wl.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_left_${tag}_$wl.json').read().strip().splitlines()[-1])
print('$tag $wl', round(d['value'],2), 'frac', round(d['roofline']['frac'],4), {k: round(v,2) for k,v in d['roofline']['phase_ms_per_step'].items() if v > 0.01})"
}
run base cfg3 GPB_LEFT=0
run left4 cfg3 GPB_LEFT=1
run left6 cfg3 GPB_LEFT=1 GPB_OUTER_BLOCK=6
run left8 cfg3 GPB_LEFT=1 GPB_OUTER_BLOCK=8
run base cfg2 GPB_LEFT=0
run left4 cfg2 GPB_LEFT=1
run left6 cfg2 GPB_LEFT=1 GPB_OUTER_BLOCK=6
