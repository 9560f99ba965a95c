#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout=600 -rP -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "rel-L2|agreement|passed|failed|vs fp32|Error|error" gpurun_out/pytest_gpu.log | grep -E "tf32|bf16|passed|failed|rror" | tail -14
python tools/phase_profile.py bf16 C2; python tools/phase_profile.py tf32 C2
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c2.log 2>&1; echo "bench exit $?"; tail -1 gpurun_out/bench_c2.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'edge_ms',d['roofline']['launch_ms'],'frac',d['roofline']['frac'],'step_frac',d['roofline']['step_frac_of_sustained'])"
