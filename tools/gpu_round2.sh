#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout=600 -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "rel-L2|agreement|passed|failed|vs fp32" gpurun_out/pytest_gpu.log | tail -40
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_c2.log 2>&1; echo "bench exit $?"; tail -3 gpurun_out/bench_c2.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 260 -c 200 --csv --log-file gpurun_out/launches_c2.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_edge -s 30 -c 2 -o gpurun_out/edge_bf16 python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"; ls -la gpurun_out
