#!/bin/bash
mkdir -p gpurun_out
bash tools/ab.sh 3 ${1:-fp16} > gpurun_out/ab_quick.txt 2>&1; cat gpurun_out/ab_quick.txt
