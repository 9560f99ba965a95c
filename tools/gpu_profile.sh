#!/bin/bash
# launch list of one forward (share of each kernel) + ncu full capture of the fused edge kernel
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_edge -s 27 -c 2 -o gpurun_out/edge_cur python tools/phase_profile.py bf16 C2 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
