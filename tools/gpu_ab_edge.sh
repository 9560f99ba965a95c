#!/bin/bash
# correctness of the current build on the edge-kernel tests, then same-box A/B (build_ab/A.so vs B.so) in fp16 and bf16
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -k "egnn_forward_golden or every_molecule or variants or bitwise or reproducible or full_width or padding" > gpurun_out/pytest_ab.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ab.log
bash tools/ab.sh 3 fp16 > gpurun_out/ab_fp16.txt 2>&1; cat gpurun_out/ab_fp16.txt
bash tools/ab.sh 2 bf16 > gpurun_out/ab_bf16.txt 2>&1; cat gpurun_out/ab_bf16.txt
