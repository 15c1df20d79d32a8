#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ab_bits.py build/var/lib_noshfl.so
GPB_DIAG_DBG=1 timeout 120 python tools/diag_dbg.py 2>&1 | tail -3 | head -2
timeout 120 python tools/diag_bench.py 2>&1
timeout 300 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency_last.log
echo prev; NS=5000 LIB=build/var/lib_noshfl.so timeout 300 python tools/b1_latency.py 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -v "^$" gpurun_out/pytest_gpu.log | tail -4
