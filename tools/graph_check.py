import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from bench import workload, normed_ctx
from ml_conformer_generator_b200.engine import Engine
from ml_conformer_generator_b200.weights import random_state_dicts
e = Engine(torch.device("cuda:0"), "bf16")
sd, ssd = random_state_dicts(0); e.load_edm_state_dict(sd); e.load_seer_state_dict(ssd)
for name, T in (("C1", 100), ("C2", 10)):
    wl = workload(name); ctx = normed_ctx(wl["ctx"], wl["B"])
    outs = []
    for k in range(4):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        o = e.generate_host(wl["global_n_nodes"], wl["N"], ctx, T, 0, seed=5 + (k == 3))
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        outs.append([t.clone() for t in o]); print(name, "call", k, "%.1f ms" % (dt * 1e3), "launches", e.kernel_launches())
    same = all(torch.equal(a, b) for a, b in zip(outs[0], outs[1])) and all(torch.equal(a, b) for a, b in zip(outs[0], outs[2]))
    print(name, "eager == captured == replayed:", same, " new seed differs:", not torch.equal(outs[0][0], outs[3][0]))
