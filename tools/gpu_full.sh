#!/bin/bash
# Full round-end style run: tests, smoke, default bench, reference arm, mixed-size workload, profiles.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout=900 -rP > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"; tail -1 gpurun_out/bench_default.json | cut -c1-400
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref exit $?"; tail -1 gpurun_out/bench_reference.json | cut -c1-300
timeout 900 python bench.py --workload C3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "c3 exit $?"; tail -1 gpurun_out/bench_c3.json | cut -c1-300
timeout 900 python bench.py --precision tf32 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_tf32.json 2> gpurun_out/bench_tf32.err; echo "tf32 exit $?"; tail -1 gpurun_out/bench_tf32.json | cut -c1-300
python tools/phase_profile.py bf16 C2 > gpurun_out/phase_bf16.json 2>&1; python tools/phase_profile.py tf32 C2 > gpurun_out/phase_tf32.json 2>&1
bash tools/gpu_profile.sh
