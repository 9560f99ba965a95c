"""Generates tests/golden/*.npz by running the REAL reference (imported from /root/reference with rdkit stubbed,
oracle/reference_loader.py) on seeded inputs.  Run once in the build container:

    python -m oracle.make_golden

The weights are not stored (46 M parameters): they are regenerated on any box by
`ml_conformer_generator_b200.weights.random_state_dicts(0)`, which this script loads into the reference modules
with `load_state_dict(strict=True)`.  Noise: the reference draws from the global torch CPU generator; seeding it
with `torch.manual_seed(seed)` makes it reproduce `NoiseTape.draw(..., seed)` (checked below), so fixtures only
store the seed.
"""
import os

import numpy as np
import torch

from ml_conformer_generator_b200.config import CONTEXT_NORMS
from ml_conformer_generator_b200.weights import random_state_dicts
from oracle import edm_oracle as O
from oracle.reference_loader import load_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
CEYYAG_CONTEXT = [50.5897, 105.3132, 133.5223]  # SURVEY.md 8d (reference get_context_shape on ceyyag heavy atoms)
ONNX_CONTEXT = [53.6424, 108.3042, 151.4399]  # reference onnx_export/onnx_export_utils.py:16-18


def build_reference(T):
    load_reference()
    from mlconfgen.adj_mat_seer import AdjMatSeer
    from mlconfgen.egnn import EGNNDynamics
    from mlconfgen.equivariant_diffusion import EquivariantDiffusion, PredefinedNoiseSchedule
    sd, ssd = random_state_dicts(0)
    dyn = EGNNDynamics(in_node_nf=9, context_node_nf=3, hidden_nf=420)
    edm = EquivariantDiffusion(dynamics=dyn, in_node_nf=8, timesteps=1000, noise_precision=1e-5)
    edm.load_state_dict(sd, strict=True)
    seer = AdjMatSeer(dimension=42, n_hidden=2048, embedding_dim=64, num_embeddings=36, num_bond_types=5)
    seer.load_state_dict(ssd, strict=True)
    # conformer_generator.py:104-113 schedule override
    edm.gamma = PredefinedNoiseSchedule(timesteps=T, precision=1e-5)
    edm.time_steps = torch.flip(torch.arange(0, T), dims=[0])
    edm.T = T
    return edm.eval(), seer.eval()


def inputs(sizes, n_max, context):
    from mlconfgen.utils.mol_utils import prepare_masks
    n_nodes = torch.tensor(sizes)
    node_mask, edge_mask = prepare_masks(n_nodes, n_max, torch.device("cpu"))
    norms = {k: torch.tensor(v) for k, v in CONTEXT_NORMS.items()}
    normed = (torch.tensor(context) - norms["mean"]) / norms["mad"]
    ctx = normed.unsqueeze(0).repeat(len(sizes), 1).unsqueeze(1).repeat(1, n_max, 1) * node_mask
    return n_nodes, node_mask, edge_mask, ctx


class Recorder:
    def __init__(self, dyn):
        self.z, self.t, self.eps = [], [], []
        dyn.register_forward_hook(self)

    def __call__(self, mod, args, out):
        self.t.append(args[0].clone())
        self.z.append(args[1].clone())
        self.eps.append(out.clone())

    def pack(self):
        return {"traj_z": torch.stack(self.z).numpy(), "traj_t": torch.stack(self.t).numpy(),
                "traj_eps": torch.stack(self.eps).numpy()}


def save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
                                 for k, v in arrs.items()})
    print("wrote", path, os.path.getsize(path))


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    load_reference()
    from mlconfgen.equivariant_diffusion import PredefinedNoiseSchedule
    from mlconfgen.utils.mol_utils import distance_matrix

    # --- schedule tables (a2)
    save("schedule", **{"gamma_%d" % T: PredefinedNoiseSchedule(T, 1e-5).gamma.detach() for T in (10, 100, 1000)})

    # --- masks (a1)
    n_nodes, nm, em, ctx = inputs([15, 17, 19, 16], 19, CEYYAG_CONTEXT)
    save("masks", n_nodes=n_nodes, n_max=19, node_mask=nm, edge_mask=em, context=ctx, raw_context=CEYYAG_CONTEXT)

    # --- single EGNN forwards (a6-a12)
    for tag, sizes, n_max, cvec, seed, tval in (("egnn_small", [15, 17, 19], 19, CEYYAG_CONTEXT, 101, 0.37),
                                                ("egnn_n39", [39, 23], 39, ONNX_CONTEXT, 102, 0.91)):
        edm, _ = build_reference(100)
        n_nodes, nm, em, ctx = inputs(sizes, n_max, cvec)
        g = torch.Generator().manual_seed(seed)
        xh = torch.randn(len(sizes), n_max, 11, generator=g) * nm
        xh[:, :, :3] *= 2.0
        t = torch.full((len(sizes), 1), tval)
        eps = edm.dynamics(t, xh, nm, em, ctx)
        save(tag, n_nodes=n_nodes, n_max=n_max, raw_context=cvec, t=t, xh=xh, eps=eps)

    # --- full sampler, plain forward (a3-a5, a13), with trajectory
    for tag, sizes, n_max, T, r, seed in (("edm_forward_T10", [15, 19, 17], 19, 10, 0, 11),
                                          ("edm_forward_T6_r1", [16, 18], 18, 6, 1, 12)):
        edm, _ = build_reference(T)
        rec = Recorder(edm.dynamics)
        n_nodes, nm, em, ctx = inputs(sizes, n_max, CEYYAG_CONTEXT)
        torch.manual_seed(seed)
        x, h = edm(nm, em, ctx, r)
        # check the tape/seed equivalence claimed in the header, and the port, right here
        n_pairs = 1 + T * (r + 1) + 1
        tape = O.NoiseTape.draw(n_pairs, len(sizes), n_max, seed)
        xo, ho = O.edm_forward(edm.state_dict(), edm.gamma.gamma, nm, em, ctx, tape, r)
        print(tag, "port vs reference: x", (x - xo).abs().max().item(), "h equal", torch.equal(h, ho),
              "|x|max", x.abs().max().item())
        save(tag, n_nodes=n_nodes, n_max=n_max, T=T, resample_steps=r, seed=seed, n_pairs=n_pairs,
             raw_context=CEYYAG_CONTEXT, x=x, h=h, **rec.pack())

    # --- inpaint (a14) and merge_fragments (a15)
    sizes, n_max, T = [21, 25, 23], 25, 6
    edm, _ = build_reference(T)
    n_nodes, nm, em, ctx = inputs(sizes, n_max, [89.8693, 210.7831, 217.7827])
    g = torch.Generator().manual_seed(77)
    n_frag = 8
    zk = torch.zeros(len(sizes), n_max, 11)
    frag_x = torch.randn(n_frag, 3, generator=g) * 1.5
    zk[:, :n_frag, :3] = frag_x
    cls = torch.tensor([6, 6, 0, 0, 0, 0, 0, 0])  # Cl, Cl, C x6 as frag_yibfeu.mol
    zk[:, :n_frag, 3:] = torch.nn.functional.one_hot(cls, 8).float()  # raw 0/1 (mol_utils.py:329-335)
    fm = torch.zeros(len(sizes), n_max, 1)
    fm[:, :n_frag] = 1.0
    rec = Recorder(edm.dynamics)
    torch.manual_seed(14)
    x, h = edm.inpaint(nm, em, ctx, zk, fm, 1, 3)
    n_pairs = 1 + T * 3 + 1
    xo, ho = O.edm_inpaint(edm.state_dict(), edm.gamma.gamma, nm, em, ctx, zk, fm,
                           O.NoiseTape.draw(n_pairs, len(sizes), n_max, 14), 1, 3)
    print("inpaint port vs reference", (x - xo).abs().max().item(), torch.equal(h, ho))
    save("edm_inpaint_T6", n_nodes=n_nodes, n_max=n_max, T=T, resample_steps=1, blend_power=3, seed=14,
         n_pairs=n_pairs, raw_context=[89.8693, 210.7831, 217.7827], z_known=zk, fixed_mask=fm, x=x, h=h,
         **rec.pack())

    edm, _ = build_reference(10)
    rec = Recorder(edm.dynamics)
    zk2 = zk.clone()
    zk2[:, n_frag:, :3] = torch.randn(len(sizes), n_max - n_frag, 3, generator=g) * nm[:, n_frag:]
    zk2[:, n_frag:, 3:] = torch.nn.functional.one_hot(torch.randint(0, 7, (len(sizes), n_max - n_frag), generator=g),
                                                      8).float() * nm[:, n_frag:]
    torch.manual_seed(15)
    x, h = edm.merge_fragments(nm, em, fm, ctx, zk2, diffusion_level=4, resample_steps=1, blend_power=3)
    n_pairs = 1 + 5 * 2 + 1
    xo, ho = O.edm_merge_fragments(edm.state_dict(), edm.gamma.gamma, nm, em, fm, ctx, zk2,
                                   O.NoiseTape.draw(n_pairs, len(sizes), n_max, 15), 4, 1, 3)
    print("merge port vs reference", (x - xo).abs().max().item(), torch.equal(h, ho))
    save("edm_merge_T10_L4", n_nodes=n_nodes, n_max=n_max, T=10, diffusion_level=4, resample_steps=1, blend_power=3,
         seed=15, n_pairs=n_pairs, raw_context=[89.8693, 210.7831, 217.7827], z_known=zk2, fixed_mask=fm, x=x, h=h,
         **rec.pack())

    # --- AdjMatSeer (a17-a19) on hand-built tensor inputs in the layout of prepare_adj_mat_seer_input (a16)
    _, seer = build_reference(10)
    g = torch.Generator().manual_seed(31)
    sizes = [17, 39, 15]
    el = torch.zeros(len(sizes), 42, dtype=torch.long)
    dm = torch.zeros(len(sizes), 42, 42)
    am = torch.zeros(len(sizes), 42, 42)
    ztab = torch.tensor(O.ATOMIC_NUMBERS)
    for k, m in enumerate(sizes):
        cls = torch.randint(0, 7, (m,), generator=g)
        el[k, :m] = ztab[cls]
        xyz = torch.randn(m, 3, generator=g) * 2.5
        d = distance_matrix(xyz)
        dm[k, :m, :m] = d
        a = (d < 1.9).float()
        am[k, :m, :m] = a
    dm = dm + torch.eye(42)
    am = am + torch.eye(42)
    am[am > 0] = 1
    logits = seer(el, dm, am)
    lo = O.seer_forward(seer.state_dict(), el, dm, am)
    bonds = torch.stack([torch.tril(torch.argmax(a, dim=2)) * (1 - torch.eye(42, dtype=torch.long)) for a in logits])
    print("seer port vs reference", (logits - lo).abs().max().item(), logits.abs().max().item(),
          torch.equal(bonds, O.bond_orders(lo)))
    save("seer", elements=el, dist_mat=dm, adj_mat=am, logits=logits, bonds=bonds, sizes=sizes)




@torch.no_grad()
def host_fixtures():
    """Golden vectors for the host-side tensor helpers (reference utils/mol_utils.py), incl. the demo molecules."""
    load_reference()
    from mlconfgen.utils import mol_utils as R
    from ml_conformer_generator_b200.mol_utils import read_mol_heavy_atoms
    out = {}
    demo = "/root/reference/assets/demo_files/"
    for name in ("ceyyag", "yibfeu", "frag_yibfeu"):
        sym, xyz = read_mol_heavy_atoms(demo + name + ".mol")
        out[name + "_xyz"] = xyz
        out[name + "_symbols"] = np.array(sym)
        centred = xyz - torch.mean(xyz, dim=0)
        ctx, rot = R.get_context_shape(centred)
        out[name + "_context"] = ctx
        out[name + "_rotated"] = rot
    # inertial fragment matching helpers (fragment expressed in the yibfeu centre-of-mass frame)
    yib = out["yibfeu_xyz"]
    ff_x = out["frag_yibfeu_xyz"] - yib.mean(dim=0)
    norms = {k: torch.tensor(v) for k, v in CONTEXT_NORMS.items()}
    n_nodes = torch.tensor([21, 25, 23, 22])
    nm, em, fctx, shift, rot = R.ifm_prepare_gen_fragment_context(
        fixed_fragment_x=ff_x, reference_context=out["yibfeu_context"], context_norms=norms, n_nodes=n_nodes,
        max_n_nodes=25, min_n_nodes=21, device=torch.device("cpu"))
    g = torch.Generator().manual_seed(5)
    xg = torch.randn(4, 17, 3, generator=g)
    hg = torch.nn.functional.one_hot(torch.randint(0, 7, (4, 17), generator=g), 8)
    inv = R.inverse_coord_transform(xg, shift, rot)
    ff_h = torch.nn.functional.one_hot(torch.tensor([6, 6, 0, 0, 0, 0, 0, 0]), 8)
    zk, fm = R.ifm_prepare_fragments_for_merge(ff_x, ff_h, inv, hg, torch.device("cpu"), 25)
    out.update(ifm_ff_x=ff_x, ifm_n_nodes=n_nodes, ifm_node_mask=nm, ifm_edge_mask=em, ifm_context=fctx, ifm_shift=shift,
               ifm_rotation=rot, ifm_xg=xg, ifm_hg=hg, ifm_inv=inv, ifm_ff_h=ff_h, ifm_z_known=zk, ifm_fixed_mask=fm)
    torch.manual_seed(123)
    nm2, em2, ctx2 = R.prepare_edm_input(6, out["ceyyag_context"], norms, 15, 19, torch.device("cpu"))
    out.update(edm_in_node_mask=nm2, edm_in_edge_mask=em2, edm_in_context=ctx2)
    save("host_utils", **out)


def shape_fixtures():
    """Golden vectors of the Gaussian shape-similarity path (reference cheminformatics/shape_similarity.py and the tensor
    part of cheminformatics/pipeline.py:37-86), produced by the reference's own functions."""
    load_reference()
    import importlib
    ss = importlib.import_module("mlconfgen.cheminformatics.shape_similarity")
    from ml_conformer_generator_b200.mol_utils import read_mol_heavy_atoms
    demo = "/root/reference/assets/demo_files/"
    _, ref_xyz = read_mol_heavy_atoms(demo + "ceyyag.mol")
    _, yib_xyz = read_mol_heavy_atoms(demo + "yibfeu.mol")
    g = torch.Generator().manual_seed(2024)
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    blob = torch.randn(30, 3, generator=g) * 2.0
    samples = [ref_xyz + 0.3 * torch.randn(ref_xyz.shape, generator=g),
               (ref_xyz + 0.15 * torch.randn(ref_xyz.shape, generator=g)) @ q + torch.tensor([3.0, -2.0, 1.0]),
               yib_xyz, blob, ref_xyz[:15].clone()]
    n_max = max(s.size(0) for s in samples)
    ref_c = ref_xyz - torch.mean(ref_xyz, dim=0)
    ref_mom, ref_pts = ss.get_shape_quadrupole_for_molecule(coordinates=ref_c)
    pi = torch.pi
    rots = [torch.tensor([pi, 0, 0]), torch.tensor([0, pi, 0]), torch.tensor([0, 0, pi])]
    B = len(samples)
    coords = torch.zeros(B, n_max, 3)
    pts = torch.zeros(B, n_max, 3)
    best_coord = torch.zeros(B, n_max, 3)
    moments = torch.zeros(B, 3)
    scores = torch.zeros(B, 4)
    best_idx = torch.zeros(B, dtype=torch.long)
    n_nodes = torch.tensor([s.size(0) for s in samples])
    for b, xyz in enumerate(samples):
        n = xyz.size(0)
        coords[b, :n] = xyz
        c = xyz - torch.mean(xyz, dim=0)
        mom, sq = ss.get_shape_quadrupole_for_molecule(coordinates=c)
        moments[b], pts[b, :n] = mom, sq
        best, bc = ss.tanimoto_score(ref_pts, sq), sq
        scores[b, 0] = best
        for k, ang in enumerate(rots):
            rc = ss.rotate_coord(coord=sq, angles=ang)
            sc = ss.tanimoto_score(ref_pts, rc)
            scores[b, k + 1] = sc
            if sc > best:
                best, bc, best_idx[b] = sc, rc, k + 1
        best_coord[b, :n] = bc
        print("shape sample", b, n, "scores", scores[b].tolist(), "best", int(best_idx[b]))
    save("shape", ref_xyz=ref_xyz, ref_moments=ref_mom, ref_pts=ref_pts, coords=coords, n_nodes=n_nodes, moments=moments,
         pts=pts, scores=scores, best_idx=best_idx, best_coord=best_coord)


if __name__ == "__main__":
    import sys
    if "--shape-only" in sys.argv:
        shape_fixtures()
        sys.exit(0)
    if "--host-only" not in sys.argv:
        main()
    host_fixtures()
    shape_fixtures()
