#!/bin/bash
# 2-GPU confirmation with the final kernels of the round: weak scaling and strong scaling (global batch 64)
N=2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --no-cpu-baseline > gpurun_out/bench_cfg3_weak_$N.json 2> gpurun_out/bench_cfg3_weak_$N.err; echo "exit $?"
timeout 600 $TR bench.py --gpus $N --scaling strong --batch 64 --no-cpu-baseline > gpurun_out/bench_cfg3_strong_b64_$N.json 2> gpurun_out/bench_cfg3_strong_b64_$N.err; echo "exit $?"
for f in weak strong_b64; do python -c "
import json
d=json.loads(open('gpurun_out/bench_cfg3_${f}_$N.json').read().strip().splitlines()[-1])
print('$f', d['value'], d['unit'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'n_gpus', d['n_gpus'])"; done
