#!/bin/bash
# Multi-GPU evidence (run with gpurun --gpus N): product sharding path, strong scaling, the cfg5 sweep.
# usage: bash tools/gpu_r2_multi.sh N [full]
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
echo "== multi_gpu_fit x$N"
timeout 600 $TR tools/multi_gpu_fit.py 3000 8 > gpurun_out/multi_gpu_fit_$N.json 2> gpurun_out/multi_gpu_fit_$N.err; echo "exit $?"; tail -1 gpurun_out/multi_gpu_fit_$N.json
for B in 64 1024; do
  echo "== strong scaling cfg3, global batch $B, x$N"
  timeout 600 $TR bench.py --gpus $N --scaling strong --batch $B --no-cpu-baseline > gpurun_out/bench_cfg3_strong_b${B}_$N.json 2> gpurun_out/bench_cfg3_strong_b${B}_$N.err; echo "exit $?"
  python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg3_strong_b${B}_$N.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
echo "== weak scaling cfg3 x$N"
timeout 600 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_cfg3_weak_$N.json 2> gpurun_out/bench_cfg3_weak_$N.err; echo "exit $?"; python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg3_weak_$N.json').read().strip().splitlines()[-1])
print(d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'])"
if [ "$2" == "full" ]; then
  echo "== cfg5 as stated: 256 samples, 2^24 test points over $N GPUs"
  timeout 1200 $TR bench.py --gpus $N --workload cfg5 --points-total 16777216 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5_full_$N.json 2> gpurun_out/bench_cfg5_full_$N.err; echo "exit $?"
  tail -c 1500 gpurun_out/bench_cfg5_full_$N.json
fi
