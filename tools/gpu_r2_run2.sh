#!/bin/bash
# fp16 mode bring-up: parity tests of every mode, then the bench in fp16 and bf16.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_t100.py tests/test_gpu_properties.py -m gpu -q -rP -k "not shape" > gpurun_out/pytest_fp16.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_fp16.log
grep -E "passed|failed" gpurun_out/pytest_fp16.log | tail -3
timeout 600 python bench.py --precision fp16 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err
tail -c 300 gpurun_out/bench_fp16.err
timeout 300 python tools/phase_profile.py fp16 C2 > gpurun_out/phase_fp16.json 2>&1
timeout 300 python tools/phase_profile.py bf16 C2 > gpurun_out/phase_bf16.json 2>&1
