#!/bin/bash
# Runs the GPU suite in separate processes (a trapped kernel poisons the CUDA context of its process only).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=600 -k "$@" > gpurun_out/t_$name.log 2>&1; echo "$name exit $?"; tail -5 gpurun_out/t_$name.log; }
run steps "step_kernels or seer_inputs or argument_errors"
run fp32 "fp32 and not seer"
run seer "seer_golden"
run gemm "tc_gemm"
run tf32 "tf32 and not gemm and not seer"
run bf16 "bf16 and not gemm"
timeout 900 python -m pytest tests/test_gpu_dropin.py -q -m gpu --timeout=600 > gpurun_out/t_dropin.log 2>&1; echo "dropin exit $?"; tail -5 gpurun_out/t_dropin.log
