#!/bin/bash
# quick validation: numerics subset + gradient kernel A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "core_golden or medium or tile_boundaries or fuzz or lownoise or toggles or design or cfg2 or rq_iso" 2>&1 | tail -3
for wl in cfg3 cfg2; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new', d['value'], d['roofline']['phase_ms_per_step']['gradient'])"
  GPYREG_B200_LIB=$PWD/build/lib_grad4.so timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old', d['value'], d['roofline']['phase_ms_per_step']['gradient'])"
done
