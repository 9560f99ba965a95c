#!/bin/bash
# first-light check of k_tc_edge3 on the GPU box (bounded by timeouts; a protocol bug traps after ~4 s per wait)
mkdir -p gpurun_out
for args in "fp16 C1" "fp16 C2" "bf16 C2 --no-fp32" "fp16 C3 --no-fp32"; do
  timeout 900 python tools/edge_v3_check.py $args 2>&1 | tail -3
done | tee gpurun_out/edge_v3_check.txt
