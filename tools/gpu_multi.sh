#!/bin/bash
NG=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/bench_cfg3_${NG}gpu.json 2> gpurun_out/bench_cfg3_${NG}gpu.err; echo "exit $?"
cut -c1-900 gpurun_out/bench_cfg3_${NG}gpu.json; tail -5 gpurun_out/bench_cfg3_${NG}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 3 --warmup 3 --workload cfg5 > gpurun_out/bench_cfg5_${NG}gpu.json 2> gpurun_out/bench_cfg5_${NG}gpu.err; echo "exit $?"
cut -c1-700 gpurun_out/bench_cfg5_${NG}gpu.json; tail -3 gpurun_out/bench_cfg5_${NG}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus $NG --steps 1 --warmup 1 2>&1 | tail -2 | cut -c1-300
