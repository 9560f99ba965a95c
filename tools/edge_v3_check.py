"""A/B of the two fused edge kernels on the GPU box: k_tc_edge3 (resident A, ping-pong accumulators; default) against
k_tc_edge (MLCG_EDGE_V3=0).  One subprocess per variant (the switch is read once per process): the same EGNN forward on
the same input, eps compared between the variants and against the exact fp32 CUDA path, plus the live launch times of a
GCL and an equivariant sub-layer.   usage: python tools/edge_v3_check.py [fp16|bf16] [C2|C3|C1] [--no-fp32]"""
import json, os, subprocess, sys
import numpy as np

def child(prec, wl_name, out):
    import torch
    sys.path.insert(0, ".")
    from bench import workload, normed_ctx
    from ml_conformer_generator_b200.engine import Engine
    from ml_conformer_generator_b200.weights import random_state_dicts
    wl = workload(wl_name)
    e = Engine(torch.device("cuda:0"), prec)
    sd, _ = random_state_dicts(0)
    e.load_edm_state_dict(sd)
    e.set_batch(wl["global_n_nodes"], wl["N"])
    B, N = wl["B"], wl["N"]
    g = torch.Generator().manual_seed(7)
    z = torch.randn(B, N, 11, generator=g).cuda()
    ctx = torch.from_numpy(normed_ctx(wl["ctx"], B))
    eps = e.egnn_forward(torch.full((B,), 0.5), z, ctx)
    torch.cuda.synchronize()
    eps2 = e.egnn_forward(torch.full((B,), 0.5), z, ctx)
    torch.cuda.synchronize()
    res = {"bitwise_repro": bool(torch.equal(eps, eps2)), "finite": bool(torch.isfinite(eps).all())}
    if prec != "fp32":
        res["gcl_ms"] = e.time_edge_kernel(0, 20)
        res["equiv_ms"] = e.time_edge_kernel(2, 20)
    np.save(out, eps.float().cpu().numpy())
    print(json.dumps(res))

def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(sys.argv[2], sys.argv[3], sys.argv[4])
        sys.exit(0)
    prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
    wl = sys.argv[2] if len(sys.argv) > 2 else "C2"
    os.makedirs("gpurun_out", exist_ok=True)
    runs = [("v3", prec, {"MLCG_EDGE_V3": "1"}), ("v1", prec, {"MLCG_EDGE_V3": "0"})]
    if "--no-fp32" not in sys.argv:
        runs.append(("fp32", "fp32", {}))
    info, eps = {}, {}
    for tag, pr, env in runs:
        out = f"/tmp/eps_{tag}.npy"
        try:
            r = subprocess.run([sys.executable, __file__, "--child", pr, wl, out], env={**os.environ, **env},
                               capture_output=True, text=True, timeout=600)
        except subprocess.TimeoutExpired:
            print(tag, "TIMEOUT"); continue
        if r.returncode != 0:
            print(tag, "FAILED", r.stdout[-2000:], r.stderr[-3000:]); continue
        info[tag] = json.loads(r.stdout.strip().splitlines()[-1])
        eps[tag] = np.load(out)
    line = {"precision": prec, "workload": wl, **{k: v for k, v in info.items()}}
    if "v3" in eps and "v1" in eps:
        line["v3_vs_v1_rel_l2"] = rel(eps["v3"], eps["v1"])
    if "fp32" in eps:
        for t in ("v3", "v1"):
            if t in eps:
                line[f"{t}_vs_fp32_rel_l2"] = rel(eps[t], eps["fp32"])
    print(json.dumps(line))
