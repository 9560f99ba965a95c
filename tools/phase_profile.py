"""Prints the per-phase cycle profile of the fused edge kernel for a workload (run on the GPU box)."""
import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
from bench import workload, normed_ctx
from ml_conformer_generator_b200.engine import Engine
from ml_conformer_generator_b200.weights import random_state_dicts
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
wl = workload(sys.argv[2] if len(sys.argv) > 2 else "C2")
e = Engine(torch.device("cuda:0"), prec)
sd, ssd = random_state_dicts(0)
e.load_edm_state_dict(sd)
e.set_batch(wl["global_n_nodes"], wl["N"])
B, N = wl["B"], wl["N"]
z = torch.randn(B, N, 11, device="cuda")
e.egnn_forward(torch.full((B,), 0.5), z, torch.from_numpy(normed_ctx(wl["ctx"], B)))
torch.cuda.synchronize()
for layer in (0, 2):
    prof = e.edge_phase_profile(layer)
    ms = e.time_edge_kernel(layer, 10)
    print(json.dumps({"precision": prec, "layer": layer, "kind": "equiv" if layer % 3 == 2 else "gcl", "launch_ms": ms,
                      "tiles": e.num_edge_tiles(), "cycles_per_tile": prof}))
