#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 600 -x -k "potrf or medium or tile_boundaries or fuzz" > gpurun_out/pytest_gpu_subset.log 2>&1; tail -2 gpurun_out/pytest_gpu_subset.log
timeout 100 python tools/one_step.py cfg3 1 0 > /dev/null 2>&1 && \
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_b1_nlz.csv python tools/one_step.py cfg3 1 0 > /dev/null 2>&1; python tools/launch_summary.py gpurun_out/r02_launches_b1_nlz.csv > gpurun_out/r02_launches_b1_nlz.txt; head -7 gpurun_out/r02_launches_b1_nlz.txt
timeout 200 python tools/b1_latency.py 2>&1 | tail -1
