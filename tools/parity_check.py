"""Parity measurements against the committed golden runs of the reference (tests/golden/*.npz, produced by
oracle/make_golden*.py from the real reference).  Shared by the GPU tests, `bench.py` (the `parity` block of the JSON
line) and tools/parity_study.py.  Reads fixtures only -- never /root/reference, never the oracle."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
# stated tolerances on the per-step eps (relative L2 against the reference's fp32 torch output)
TOL = {"fp32": 2e-5, "tf32": 1e-3, "fp16": 1e-3, "bf16": 2e-2}


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def normed_context(raw, B):
    from ml_conformer_generator_b200.config import CONTEXT_NORMS
    c = (torch.tensor(np.asarray(raw), dtype=torch.float32) - torch.tensor(CONTEXT_NORMS["mean"])) / torch.tensor(
        CONTEXT_NORMS["mad"])
    return c.view(1, 3).repeat(B, 1)


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def teacher_forced(engine, g, chunk=101):
    """Per-step eps of the CUDA path on the reference's own z_t trajectory: every recorded denoiser call of the golden run
    is replayed (the calls are independent given z_t, so `chunk` of them are stacked into one batch with per-sample t).
    Returns the list of per-call relative L2 errors."""
    n_calls, B, N, _ = g["traj_z"].shape
    n_nodes = np.asarray(g["n_nodes"]).astype(np.int32)
    ctx1 = normed_context(g["raw_context"], B)
    errs = []
    for c0 in range(0, n_calls, chunk):
        k = min(chunk, n_calls - c0)
        engine.set_batch(np.tile(n_nodes, k), N)
        z = torch.from_numpy(g["traj_z"][c0:c0 + k]).reshape(k * B, N, 11)
        t = torch.from_numpy(g["traj_t"][c0:c0 + k]).reshape(k * B)
        eps = engine.egnn_forward(t, z, ctx1.repeat(k, 1)).cpu().reshape(k, B, N, 11)
        for j in range(k):
            errs.append(rel_l2(eps[j], g["traj_eps"][c0 + j]))
    return errs


def noise_tape(g):
    """The reference's noise for a golden run: it draws from the global CPU generator seeded with g['seed'], x-part
    (B,N,3) then h-part (B,N,8) per draw (reference equivariant_diffusion.py:347-362)."""
    B, N = len(g["n_nodes"]), int(g["n_max"])
    gen = torch.Generator().manual_seed(int(g["seed"]))
    out = torch.empty(int(g["n_pairs"]), B, N, 11)
    for k in range(int(g["n_pairs"])):
        out[k, :, :, :3] = torch.randn(B, N, 3, generator=gen)
        out[k, :, :, 3:] = torch.randn(B, N, 8, generator=gen)
    return out


def free_running(engine, g):
    """Whole reverse loop with the reference's injected noise.  Returns (x rel-L2, atom-type agreement, atoms)."""
    B, N = len(g["n_nodes"]), int(g["n_max"])
    engine.set_batch(np.asarray(g["n_nodes"]).astype(np.int32), N)
    x, cls = engine.sample(normed_context(g["raw_context"], B), int(g["T"]), "forward", int(g["resample_steps"]),
                           noise_tape=noise_tape(g))
    h = torch.from_numpy(g["h"])
    real = h.sum(-1) > 0
    agree = float((cls.cpu().long()[real] == h.argmax(-1)[real]).float().mean())
    return rel_l2(x.cpu(), g["x"]), agree, int(real.sum())


def bond_agreement(bonds, ref_bonds, n_nodes, ref_logits=None):
    """Bond-order agreement on the entries the reference consumes (reference utils/mol_utils.py:210-211: strict lower
    triangle of the argmax).  Returns a dict: agreement over all 861 strict-lower-triangle entries per molecule, over the
    pairs of real atoms only, the number of flips and -- when the reference logits are given -- the largest top-2 margin
    of the reference among the flipped entries (a flip at a margin below the logit error is a tie, not an error)."""
    bonds = torch.as_tensor(bonds).long().cpu()
    ref_bonds = torch.as_tensor(ref_bonds).long().cpu()
    B, D = bonds.shape[0], bonds.shape[1]
    tri = torch.tril(torch.ones(D, D, dtype=torch.bool), diagonal=-1).unsqueeze(0).expand(B, D, D)
    idx = torch.arange(D)
    n = torch.as_tensor(np.asarray(n_nodes)).long().view(B, 1, 1)
    real = tri & (idx.view(1, D, 1) < n) & (idx.view(1, 1, D) < n)
    eq = bonds == ref_bonds
    out = {"lower_triangle": float(eq[tri].double().mean()), "lower_triangle_entries": int(tri.sum()),
           "real_pairs": float(eq[real].double().mean()) if int(real.sum()) else 1.0, "real_pair_entries": int(real.sum()),
           "flips": int((~eq & tri).sum()), "flips_real_pairs": int((~eq & real).sum())}
    if ref_logits is not None:
        top2 = torch.topk(torch.as_tensor(ref_logits).double().cpu(), 2, dim=-1).values
        margin = top2[..., 0] - top2[..., 1]
        flipped = ~eq & tri
        out["max_ref_margin_at_flips"] = float(margin[flipped].max()) if int(flipped.sum()) else 0.0
        out["median_ref_margin"] = float(margin[tri].median())
    return out
