#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "cfg3 exit $?"; cut -c1-1500 gpurun_out/bench_cfg3.json; tail -3 gpurun_out/bench_cfg3.err
timeout 900 python bench.py --steps 3 --warmup 3 --workload cfg2 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; echo "cfg2 exit $?"; cut -c1-1500 gpurun_out/bench_cfg2.json; tail -3 gpurun_out/bench_cfg2.err
timeout 900 python bench.py --steps 3 --warmup 3 --workload cfg5 > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "cfg5 exit $?"; cut -c1-1500 gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench_cfg5.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; echo "ref exit $?"; cut -c1-600 gpurun_out/bench_ref.json
python -c "import __graft_entry__ as g; g.smoke()"
