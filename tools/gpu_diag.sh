#!/bin/bash
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "potrf or medium or tile_boundaries or fuzz or core_golden" 2>&1 | tail -2
timeout 100 python tools/diag_bench.py 2>&1 | tail -3
GPB_DIAG_DBG=1 timeout 100 python tools/diag_bench.py 2>&1 | grep -A1 "phase cycles" | head -2
NS=5000 timeout 100 python tools/b1_latency.py | tail -1
