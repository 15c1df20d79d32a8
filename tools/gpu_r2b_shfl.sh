#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -k "potrf or core_golden or cholesky_failure" 2>&1 | tail -2
timeout 600 python tools/ab_bits.py build/var/lib_noshfl.so
echo "== shuffle-ahead"; GPB_DIAG_DBG=1 timeout 120 python tools/diag_dbg.py 2>&1 | tail -3 | head -2
echo "== without"; GPYREG_B200_LIB=$PWD/build/var/lib_noshfl.so GPB_DIAG_DBG=1 timeout 120 python tools/diag_dbg.py 2>&1 | tail -3 | head -2
echo "== shuffle-ahead"; timeout 120 python tools/diag_bench.py 2>&1; NS=5000 timeout 300 python tools/b1_latency.py 2>&1 | tail -1
echo "== without"; NS=5000 LIB=build/var/lib_noshfl.so timeout 300 python tools/b1_latency.py 2>&1 | tail -1
