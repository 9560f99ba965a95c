"""Throughput of the GPU shape-similarity path (mlcg_shape_moments + mlcg_shape_tanimoto) on synthetic conformers, with the
CPU oracle (the reference's algorithm) timed on a few samples beside it.  usage: python tools/bench_shape.py [B] [n_atoms]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_conformer_generator_b200.shape_similarity import ShapeScorer  # noqa: E402


def blob(n, g):
    """A compact random 'molecule': points on a jittered 1.5 A lattice walk, so neighbour counts resemble real conformers."""
    pts = [torch.zeros(3)]
    while len(pts) < n:
        base = pts[int(torch.randint(0, len(pts), (1,), generator=g))]
        step = torch.randn(3, generator=g)
        cand = base + 1.5 * step / step.norm()
        if min(float((cand - p).norm()) for p in pts) > 1.2:
            pts.append(cand)
    return torch.stack(pts)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 39
    g = torch.Generator().manual_seed(0)
    ref = blob(n, g)
    base = torch.stack([blob(n, g) for _ in range(32)])
    coords = base[torch.arange(B) % 32] + 0.05 * torch.randn(B, n, 3, generator=g)
    n_nodes = torch.full((B,), n)
    sc = ShapeScorer()
    dev = sc.device
    coords_d = coords.to(dev)
    for _ in range(2):
        sc.evaluate(ref, coords_d, n_nodes)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        out = sc.evaluate(ref, coords_d, n_nodes)
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / reps
    # kernels only
    fr = sc.principal_frames(coords_d, n_nodes)
    rp = sc.principal_frames(ref.unsqueeze(0), torch.tensor([n]))["points"][0]
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    out16 = torch.empty(B, 16, device=dev)
    from ml_conformer_generator_b200.engine import _ptr
    nn = n_nodes.to(torch.int32).to(dev)
    e = sc.engine
    e0.record()
    for _ in range(reps):
        e.lib.mlcg_shape_moments(e.h, _ptr(coords_d), _ptr(nn), B, n, sc.amplitude, sc.atom_radius, 6, _ptr(out16), e._stream())
    e1.record()
    for _ in range(reps):
        sc.tanimoto(rp, coords_d, n_nodes, fr["frames"])
    e2.record()
    torch.cuda.synchronize()
    res = {"workload": "shape similarity, B=%d samples x %d atoms, 4 orientations, 40^3 grid" % (B, n),
           "samples_per_s_end_to_end": B / t_all, "ms_end_to_end": t_all * 1e3,
           "ms_moments_kernel": e0.elapsed_time(e1) / reps, "ms_tanimoto_call": e1.elapsed_time(e2) / reps,
           "mean_best_score": float(out["shape_tanimoto"].mean())}
    if "--no-cpu" not in sys.argv:
        from oracle import shape_oracle as S
        t0 = time.perf_counter()
        k = 2
        for b in range(k):
            S.evaluate_shape(ref, coords[b])
        res["cpu_oracle_s_per_sample"] = (time.perf_counter() - t0) / k
        res["cpu_note"] = "oracle/shape_oracle.py (vectorised restatement; the reference's Python clique backtracker is ~10x slower)"
    print(json.dumps(res))


if __name__ == "__main__":
    main()
