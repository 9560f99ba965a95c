"""Small end-to-end case for compute-sanitizer (memcheck / racecheck):  python tools/sanitizer_case.py [precisions] [reps]
EGNN forward, 2-step sampler (device Philox noise), GCN, IFM kernels on a ragged batch that exercises whole-target tiles,
split targets carried in shared memory and the side-buffer path."""
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/repo")
from ml_conformer_generator_b200.config import CONTEXT_NORMS  # noqa: E402
from ml_conformer_generator_b200.engine import Engine  # noqa: E402
from ml_conformer_generator_b200.weights import random_state_dicts  # noqa: E402

precs = (sys.argv[1] if len(sys.argv) > 1 else "fp16,bf16,tf32").split(",")
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
sd, ssd = random_state_dicts(0)
for prec in precs:
    e = Engine(torch.device("cuda:0"), prec)
    e.load_edm_state_dict(sd)
    e.load_seer_state_dict(ssd)
    n = np.array([39, 17, 23, 39, 15, 31, 2, 1, 13, 12, 27, 39, 39, 20, 36, 38] * reps, np.int32)
    e.set_batch(n, 39)
    B = len(n)
    g = torch.Generator().manual_seed(1)
    z = torch.randn(B, 39, 11, generator=g)
    ctx = torch.zeros(B, 3)
    eps = e.egnn_forward(torch.full((B,), 0.5), z, ctx)
    x, cls = e.sample(ctx, 2, "forward", 0, seed=3, sample_ids=np.arange(B)[::-1].copy())
    el, d, a = e.seer_inputs(x, cls)
    lo, bo = e.seer_forward(el, d, a)
    ff = torch.randn(8, 3, generator=g)
    c2, sh, rot, ng = e.ifm_context(ff, torch.tensor([89.87, 210.78, 217.78]), CONTEXT_NORMS, torch.from_numpy(n.clip(10, 39)))
    zk, fm = e.ifm_merge_inputs(x[:, :31].contiguous(), cls[:, :31].contiguous(), sh, rot, ff, torch.eye(8), 39)
    torch.cuda.synchronize()
    print(prec, "ok", float(eps.abs().max()), float(zk.abs().max()))
    e.close()
