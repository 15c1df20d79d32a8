#!/bin/bash
# quick validation: numerics subset + small-batch latency
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "potrf or gemm_nt or core_golden or medium or tile_boundaries or fuzz or lownoise or toggles or factor_cache or robustness or design or fit_examples" 2>&1 | tail -4
timeout 200 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency.log
echo "GPB_GRAPH=0:"; GPB_GRAPH=0 NS=1000,5000 timeout 200 python tools/b1_latency.py 2>&1
REPS=8 timeout 120 python tools/hit_once.py 2>&1 | tail -1
timeout 300 python tools/mid_batch.py | tail -1
