#!/bin/bash
# ncu evidence for round 2: launch list of the default bench command, full captures of the fused edge kernel, the node GEMMs,
# the GCN GEMM and the step kernel; compute-sanitizer memcheck + bounded racecheck.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 330 -c 420 --csv --log-file gpurun_out/launches_default.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_edge -s 60 -c 3 -o gpurun_out/r2_edge python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras > gpurun_out/ncu_edge.log 2>&1; echo "edge rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 30 -c 8 -o gpurun_out/r2_gemm python tools/phase_profile.py fp16 C3 > gpurun_out/ncu_gemm.log 2>&1; echo "gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_step|k_decode|k_noise_init|k_egnn_prepare|k_egnn_readout|k_seer|k_lmul|k_lnorm" -s 20 -c 16 -o gpurun_out/r2_small python tools/small_kernels_case.py > gpurun_out/ncu_small.log 2>&1; echo "small rc=$?"
timeout 900 ncu --set full --clock-control none -k "regex:k_tc_gemm<1, 256" -s 2 -c 3 -o gpurun_out/r2_gcn_gemm python tools/small_kernels_case.py > gpurun_out/ncu_gcn.log 2>&1; echo "gcn rc=$?"
bash tools/ncu_export.sh > gpurun_out/ncu_export.log 2>&1; echo "export rc=$?"; tail -2 gpurun_out/ncu_export.log
timeout 600 compute-sanitizer --tool memcheck python tools/sanitizer_case.py fp16,bf16,tf32 4 > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck.log
timeout 700 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitizer_case.py fp16,tf32 1 > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/racecheck.log
timeout 600 python bench.py --impl reference --ref-device cuda --steps 3 --warmup 1 > gpurun_out/bench_reference_gpu.json 2> gpurun_out/bench_reference_gpu.err; echo "refgpu rc=$?"; tail -c 300 gpurun_out/bench_reference_gpu.err
timeout 600 python bench.py --workload C1 --steps 3 --warmup 2 --no-extras > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; echo "c1 rc=$?"
timeout 900 python bench.py --impl reference --workload C1 --ref-full --ref-mols 20 --steps 1 --warmup 0 > gpurun_out/bench_reference_c1_full.json 2> gpurun_out/bench_reference_c1_full.err; echo "ref c1 rc=$?"
