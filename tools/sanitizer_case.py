import sys, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from ml_conformer_generator_b200.engine import Engine
from ml_conformer_generator_b200.weights import random_state_dicts
sd, ssd = random_state_dicts(0)
for prec in ("bf16", "tf32"):
    e = Engine(torch.device("cuda:0"), prec); e.load_edm_state_dict(sd); e.load_seer_state_dict(ssd)
    n = np.array([39, 17, 23, 39, 15, 31, 2, 1, 13, 12, 27, 39, 39, 20, 36, 38] * 4, np.int32)
    e.set_batch(n, 39)
    B = len(n)
    g = torch.Generator().manual_seed(1)
    z = torch.randn(B, 39, 11, generator=g)
    ctx = torch.zeros(B, 3)
    eps = e.egnn_forward(torch.full((B,), 0.5), z, ctx)
    x, cls = e.sample(ctx, 2, "forward", 0, seed=3)
    el, d, a = e.seer_inputs(x, cls)
    lo, bo = e.seer_forward(el, d, a)
    torch.cuda.synchronize()
    print(prec, "ok", float(eps.abs().max()))
