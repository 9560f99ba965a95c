#!/bin/bash
# A/B of the node GEMMs for two library builds (build_ab/A.so, build_ab/B.so): launch ms of the three EGNN GEMM variants
L=ml_conformer_generator_b200/libmlcg_b200.so
cp $L /tmp/orig.so
for i in $(seq 1 ${1:-2}); do
  for v in A B; do
    cp build_ab/$v.so $L
    python tools/gemm_profile.py ${2:-bf16} C2 2>/dev/null | python -c "
import sys, json
print('$v', ' '.join('%s %.1f us (epi %.0f, mma-wait-epi %.0f)' % (d['gemm'], d['launch_ms'] * 1e3, d['epilogue'] / d['tiles_per_cta'], d['mma_wait_epilogue'] / d['tiles_per_cta']) for d in map(json.loads, sys.stdin)))
"
  done
done
cp /tmp/orig.so $L
