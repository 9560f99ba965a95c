#!/bin/bash
# Round-2 full pass on one GPU: whole GPU suite, smoke, default bench, reference arms.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --impl reference --ref-device cuda --steps 3 --warmup 1 > gpurun_out/bench_reference_gpu.json 2> gpurun_out/bench_reference_gpu.err; echo "refgpu rc=$?"; tail -c 300 gpurun_out/bench_reference_gpu.err
timeout 600 python bench.py --precision bf16 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bf16 rc=$?"
