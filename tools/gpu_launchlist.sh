#!/bin/bash
# ncu launch list of one forward (per-kernel durations); usage: gpu_launchlist.sh <tag>
mkdir -p gpurun_out
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 200 --csv --log-file gpurun_out/launches_$1.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch_$1.log 2>&1; echo "ncu launches exit $?"
