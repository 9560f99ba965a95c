#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_ifm.py tests/test_gpu_noise.py tests/test_gpu_dropin.py -m gpu -q -rP > gpurun_out/pytest_ifm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_ifm.log
grep -E "passed|failed" gpurun_out/pytest_ifm.log | tail -3
( timeout 900 python tools/parity_study.py fp16 fp16
  MLCG_EDGE_DIST_FP32=1 timeout 600 python tools/parity_study.py fp16 fp16+dist_fp32
  timeout 600 python tools/parity_study.py tf32 tf32
  MLCG_EDGE_DIST_FP32=1 timeout 600 python tools/parity_study.py bf16 bf16+dist_fp32
  timeout 600 python tools/parity_study.py bf16 bf16 ) > gpurun_out/parity_study.jsonl 2> gpurun_out/parity_study.err
tail -3 gpurun_out/parity_study.err
