#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6
timeout 600 python tools/microbench.py 2>&1 | grep wl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['wl'], 'B=%d' % d['B'], 'grad' if d['grad'] else 'nlz ', 'evals/s %.1f' % d['evals_per_s'], 'TF %.2f' % d['tflops_alg'], {k: round(v, 2) for k, v in d['phases_ms'].items() if v > 0.01})
"
timeout 600 python bench.py --steps 3 --warmup 3 2>/dev/null | cut -c1-200
