#!/bin/bash
# Round-2 final pass on one GPU with k_tc_edge3 as the default edge kernel: whole GPU suite, smoke, default bench + reference
# arm + bf16 line, ncu launch list and a full capture of the edge kernel, phase profiles, A/B against k_tc_edge, memcheck.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -rP > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed" gpurun_out/pytest_gpu.log | tail -2
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_default.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"
timeout 600 python bench.py --precision bf16 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bf16 rc=$?"; cut -c1-120 gpurun_out/bench_bf16.json
MLCG_EDGE_V3=0 timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_v1_edge.json 2> gpurun_out/bench_v1_edge.err; echo "v1 rc=$?"; cut -c1-120 gpurun_out/bench_v1_edge.json
for a in "fp16 C2" "fp16 C3" "bf16 C2" "bf16 C3"; do timeout 300 python tools/phase_profile.py $a 2>/dev/null | tail -2; done > gpurun_out/r2_edge3_phase_profiles.jsonl
for a in "fp16 C2" "fp16 C3 --no-fp32" "bf16 C3 --no-fp32" "fp16 C1"; do timeout 500 python tools/edge_v3_check.py $a 2>/dev/null | tail -1; done > gpurun_out/r2_edge3_vs_edge.jsonl; cat gpurun_out/r2_edge3_vs_edge.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 420 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_edge3 -s 60 -c 3 -o gpurun_out/r2_edge python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_edge.log 2>&1; echo "edge rc=$?"
bash tools/ncu_export.sh > gpurun_out/ncu_export.log 2>&1; echo "export rc=$?"; tail -2 gpurun_out/ncu_export.log
timeout 500 compute-sanitizer --tool memcheck python tools/sanitizer_case.py fp16,bf16 2 > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
