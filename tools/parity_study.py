"""Parity / speed study of one kernel variant (run on the GPU box):  python tools/parity_study.py <precision> [tag]
Environment switches (MLCG_EDGE_DIST_FP32, ...) are read once per process, so every variant is its own process; the
exact-fp32 free-running reference (T = 100, > 10 000 atoms) is computed by the first process and cached in /tmp.
Prints one JSON line per measurement."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_conformer_generator_b200.engine import Engine  # noqa: E402
from ml_conformer_generator_b200.weights import random_state_dicts  # noqa: E402
from tools import parity_check as PC  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
tag = sys.argv[2] if len(sys.argv) > 2 else prec
sd, ssd = random_state_dicts(0)


def engine(p):
    e = Engine(torch.device("cuda:0"), p)
    e.load_edm_state_dict(sd)
    return e


def workloads():
    rng = np.random.RandomState(3)
    return {"C2": np.full(264, 39, np.int32), "C3": rng.randint(15, 40, 384).astype(np.int32)}


def tape_for(B, N=39, T=100, seed=2024):
    g = torch.Generator().manual_seed(seed)
    out = torch.empty(T + 2, B, N, 11)
    for k in range(T + 2):
        out[k, :, :, :3] = torch.randn(B, N, 3, generator=g)
        out[k, :, :, 3:] = torch.randn(B, N, 8, generator=g)
    return out


cache = "/tmp/mlcg_fp32_free_running.pt"
if not os.path.exists(cache):
    e = engine("fp32")
    ref = {}
    for name, n_nodes in workloads().items():
        e.set_batch(n_nodes, 39)
        x, cls = e.sample(PC.normed_context([53.6424, 108.3042, 151.4399], len(n_nodes)), 100, "forward", 0,
                          noise_tape=tape_for(len(n_nodes)))
        ref[name] = (x.cpu(), cls.cpu())
    torch.save(ref, cache)
    e.close()
ref = torch.load(cache)

e = engine(prec)
for gname in ("edm_forward_T100_n39", "edm_forward_T100_mixed"):
    g = PC.load_golden(gname)
    errs = PC.teacher_forced(e, g, chunk=26 if prec == "fp32" else 101)
    xmax = np.abs(g["traj_z"][:, :, :, :3]).reshape(101, -1).max(1)
    print(json.dumps({"variant": tag, "golden": gname, "teacher_forced_worst": max(errs), "median": float(np.median(errs)),
                      "worst_call": int(np.argmax(errs)), "ordinary_scale_worst": max(x for x, m in zip(errs, xmax) if m <= 100),
                      "calls_above_1e-3": int(sum(x > 1e-3 for x in errs))}), flush=True)
for name, n_nodes in workloads().items():
    e.set_batch(n_nodes, 39)
    x, cls = e.sample(PC.normed_context([53.6424, 108.3042, 151.4399], len(n_nodes)), 100, "forward", 0,
                      noise_tape=tape_for(len(n_nodes)))
    real = ref[name][1] >= 0
    diff = int((cls.cpu()[real] != ref[name][1][real]).sum())
    print(json.dumps({"variant": tag, "free_running_vs_fp32_cuda": name, "atoms": int(real.sum()), "differ": diff,
                      "agreement": 1 - diff / int(real.sum()), "x_rel_l2": PC.rel_l2(x.cpu(), ref[name][0])}), flush=True)
if prec != "fp32":
    e.set_batch(np.full(1024, 39, np.int32), 39)
    z = torch.randn(1024, 39, 11, device="cuda")
    e.egnn_forward(torch.full((1024,), 0.5), z, PC.normed_context([53.6424, 108.3042, 151.4399], 1024))
    torch.cuda.synchronize()
    bd = e.egnn_forward_breakdown(torch.full((1024,), 0.5), z, PC.normed_context([53.6424, 108.3042, 151.4399], 1024))
    print(json.dumps({"variant": tag, "C2_edge_gcl_ms": e.time_edge_kernel(0, 20), "C2_edge_equiv_ms": e.time_edge_kernel(2, 20),
                      "C2_forward_ms": bd["total_ms"]}), flush=True)
