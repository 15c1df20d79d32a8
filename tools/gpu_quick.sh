#!/bin/bash
# quick validation of a numerics subset + small/mid batch latency on one B200
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "potrf or core_golden or medium or tile_boundaries or fuzz or toggles or design or factor_cache or large_rq" 2>&1 | tail -4
echo "banded (default):"; NS=2000,5000 timeout 200 python tools/b1_latency.py 2>&1
timeout 300 python tools/mid_batch.py | tail -1
echo "GPB_LA_BAND=0:"; GPB_LA_BAND=0 NS=2000,5000 timeout 200 python tools/b1_latency.py 2>&1
GPB_LA_BAND=0 timeout 300 python tools/mid_batch.py | tail -1
echo "GPB_LA_BAND=12:"; GPB_LA_BAND=12 NS=5000 timeout 200 python tools/b1_latency.py 2>&1
echo "GPB_LA_BAND=16:"; GPB_LA_BAND=16 NS=5000 timeout 200 python tools/b1_latency.py 2>&1
