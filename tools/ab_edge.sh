#!/bin/bash
# A/B of several builds of the library (build_ab/<name>.so) on the same GPU box: live launch times of the GCL and the
# equivariant edge kernel.  usage: tools/ab_edge.sh "E1 E2 E3" [precision] [workload] [rounds]
L=ml_conformer_generator_b200/libmlcg_b200.so
cp $L /tmp/orig.so
for i in $(seq 1 ${4:-2}); do
  for v in $1; do
    cp build_ab/$v.so $L
    python - ${2:-fp16} ${3:-C2} $v <<'PY'
import sys, torch
sys.path.insert(0, ".")
from bench import workload, normed_ctx
from ml_conformer_generator_b200.engine import Engine
from ml_conformer_generator_b200.weights import random_state_dicts
prec, wl, tag = sys.argv[1], workload(sys.argv[2]), sys.argv[3]
e = Engine(torch.device("cuda:0"), prec)
e.load_edm_state_dict(random_state_dicts(0)[0])
e.set_batch(wl["global_n_nodes"], wl["N"])
B, N = wl["B"], wl["N"]
e.egnn_forward(torch.full((B,), 0.5), torch.randn(B, N, 11), torch.from_numpy(normed_ctx(wl["ctx"], B)))
torch.cuda.synchronize()
print(tag, prec, sys.argv[2], "gcl %.4f ms  equiv %.4f ms" % (e.time_edge_kernel(0, 20), e.time_edge_kernel(2, 20)))
PY
  done
done
cp /tmp/orig.so $L
