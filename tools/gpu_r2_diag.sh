#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x -k "potrf or gemm_nt or core_golden or medium or tile_boundaries or fuzz or lownoise or toggles" > gpurun_out/pytest_gpu_subset.log 2>&1; tail -4 gpurun_out/pytest_gpu_subset.log
GPB_DIAG_DBG=1 timeout 100 python tools/diag_dbg.py 2>&1 | tail -3
echo "== b1"; timeout 200 python tools/b1_latency.py 2>&1 | tee gpurun_out/b1_latency.log
echo "== grad kernel 4 CTAs/SM (default) vs 3 (build/lib_grad3.so), bench cfg3 + cfg2"
for wl in cfg3 cfg2; do
  timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('new', d['value'], d['roofline']['phase_ms_per_step'])"
  GPYREG_B200_LIB=$PWD/build/lib_grad3.so timeout 300 python bench.py --workload $wl --no-cpu-baseline --steps 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('old', d['value'], d['roofline']['phase_ms_per_step'])"
done
