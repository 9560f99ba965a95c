"""Eager (no CUDA graph) short case for ncu captures of the small kernels and the GCN GEMMs: C2 batch, T = 2 sampler + GCN."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from bench import normed_ctx, workload  # noqa: E402
from ml_conformer_generator_b200.engine import Engine  # noqa: E402
from ml_conformer_generator_b200.weights import random_state_dicts  # noqa: E402

e = Engine(torch.device("cuda:0"), sys.argv[1] if len(sys.argv) > 1 else "fp16")
sd, ssd = random_state_dicts(0)
e.load_edm_state_dict(sd)
e.load_seer_state_dict(ssd)
wl = workload("C2")
e.set_batch(wl["global_n_nodes"], wl["N"])
ctx = torch.from_numpy(normed_ctx(wl["ctx"], wl["B"])).cuda()
for rep in range(2):
    x, cls = e.sample(ctx, 2, "forward", 0, seed=rep)
    el, d, a = e.seer_inputs(x, cls)
    e.seer_forward(el, d, a, want_logits=False)
torch.cuda.synchronize()
print("ok")
