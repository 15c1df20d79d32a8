#!/bin/bash
for w in 6000 3000 1500 800 400; do for ob in 0 2; do echo -n "LA_WIDE=$w LA_OB=$ob: "; NS=5000 GPB_LA_WIDE=$w GPB_LA_OB=$ob timeout 100 python tools/b1_latency.py 2>&1 | tail -1; done; done
echo "== mid batch default"; timeout 200 python tools/mid_batch.py | tail -1
echo "== mid batch LA_WIDE=1500"; GPB_LA_WIDE=1500 timeout 200 python tools/mid_batch.py | tail -1
echo "== cfg2 outer-block sweep"
for ob in 2 4 8; do echo -n "OB=$ob: "; GPB_OUTER_BLOCK=$ob timeout 300 python bench.py --workload cfg2 --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['phase_ms_per_step']['factor'])"; done
