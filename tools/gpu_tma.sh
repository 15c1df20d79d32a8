#!/bin/bash
mkdir -p gpurun_out
export GPB_LOADER=tma
timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 100 -k "gemm" 2>&1 | tail -4
echo "gemm-only exit $?"
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4
for L in cpasync tma; do
  export GPB_LOADER=$L
  echo "== loader $L"
  timeout 300 python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from gpyreg_b200 import Engine
e = Engine(0)
for (M, N, K) in [(8192, 8192, 2048), (8192, 8192, 512), (8192, 8192, 128)]:
    ms = e.debug_gemm_bench(M, N, K, 5)
    print(f"gemm {M}x{N}x{K}: {2.0*M*N*K/(ms*1e-3)/1e12:.2f} TFLOP/s")
PY
  timeout 300 python tools/ob_sweep.py 2>&1 | grep "OB=4"
done
