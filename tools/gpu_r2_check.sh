#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_t100.py tests/test_gpu_dropin.py tests/test_gpu_ifm.py -m gpu -q -rP -k "seer or bond or inpaint or generate or sdf or ifm" > gpurun_out/pytest_check.log 2>&1
echo "rc=$?"; grep -E "passed|failed" gpurun_out/pytest_check.log | tail -2; grep -E "bond orders|seer logits|inpaint T=100|bond-order" gpurun_out/pytest_check.log | grep -v "print("
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_check.json').read().strip().splitlines()[-1]); print(d['value'], d['breakdown']['gcn_ms'], d['breakdown']['gcn_share_of_step'])"
