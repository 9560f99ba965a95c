"""Prints the cycle counters of the three node GEMMs (mlcg_gemm_phase_profile) for a workload (run on the GPU box)."""
import sys, json
import torch
sys.path.insert(0, ".")
from bench import workload, normed_ctx
from ml_conformer_generator_b200.engine import Engine
from ml_conformer_generator_b200.weights import random_state_dicts
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
wl = workload(sys.argv[2] if len(sys.argv) > 2 else "C2")
e = Engine(torch.device("cuda:0"), prec)
sd, _ = random_state_dicts(0)
e.load_edm_state_dict(sd)
e.set_batch(wl["global_n_nodes"], wl["N"])
B, N = wl["B"], wl["N"]
z = torch.randn(B, N, 11, device="cuda")
e.egnn_forward(torch.full((B,), 0.5), z, torch.from_numpy(normed_ctx(wl["ctx"], B)))
torch.cuda.synchronize()
for which, name in ((0, "pq_projection"), (1, "node_mlp_silu"), (2, "node_mlp_residual")):
    for rep in range(2):
        prof = e.gemm_phase_profile(which)
    print(json.dumps({"precision": prec, "gemm": name, **prof}))
