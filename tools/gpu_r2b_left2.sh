#!/bin/bash
mkdir -p gpurun_out
run() {
  local tag=$1 wl=$2; shift 2
  env "$@" timeout 600 python bench.py --workload $wl --no-cpu-baseline > gpurun_out/bench_left_${tag}_$wl.json 2> gpurun_out/bench_left_${tag}_$wl.err
  python -c "
import json
d=json.loads(open('gpurun_out/bench_left_${tag}_$wl.json').read().strip().splitlines()[-1])
print('$tag $wl', round(d['value'],2), 'frac', round(d['roofline']['frac'],4), {k: round(v,2) for k,v in d['roofline']['phase_ms_per_step'].items() if v > 0.01})"
}
run left12 cfg3 GPB_LEFT=1 GPB_OUTER_BLOCK=12
run left16 cfg3 GPB_LEFT=1 GPB_OUTER_BLOCK=16
run left40 cfg3 GPB_LEFT=1 GPB_OUTER_BLOCK=40
run left8 cfg2 GPB_LEFT=1 GPB_OUTER_BLOCK=8
run left16 cfg2 GPB_LEFT=1 GPB_OUTER_BLOCK=16
echo "== mid batch default"; timeout 300 python tools/mid_batch.py 2>&1 | tail -2
echo "== mid batch LEFT OB8"; GPB_LEFT=1 GPB_OUTER_BLOCK=8 timeout 300 python tools/mid_batch.py 2>&1 | tail -2
