#!/bin/bash
# A/B two builds of the library on the same GPU box: build_ab/A.so and build_ab/B.so, alternating, bf16 phase profile
# (launch_ms of the GCL and equivariant edge kernels).  usage: tools/ab.sh [rounds]
mkdir -p gpurun_out
L=ml_conformer_generator_b200/libmlcg_b200.so
cp $L /tmp/orig.so
for i in $(seq 1 ${1:-3}); do
  for v in A B; do
    cp build_ab/$v.so $L
    python tools/phase_profile.py ${2:-bf16} C2 2>/dev/null | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); c = d['cycles_per_tile']
    print('$v', d['kind'], 'ms %.4f' % d['launch_ms'], 'agen %.0f tail %.0f p1 %.0f p2 %.0f bp %.0f' % (c['a_gen'], c['mma_tail'], c['pass1'], c['pass2'], c['a_ring_backpressure']))
"
  done
done
cp /tmp/orig.so $L
