"""End-to-end use of the accelerated path without RDKit (run on a B200):

    python examples/generate_and_score.py reference.mol out.sdf [n_samples]

reads the heavy atoms of a V2000 mol file, generates `n_samples` conformers around its shape context (atom count +- 2,
the reference's `variance=2`), predicts bond orders with the AdjMatSeer GCN, writes the samples as an SDF file and prints
their Gaussian shape-Tanimoto similarity to the reference (the tensor part of the reference's `evaluate_samples`).

Weights: pass the reference's checkpoints with --edm / --seer (`edm_moi_chembl_15_39.pt`, `adj_mat_seer_chembl_15_39.pt`);
without them the architecture runs with random weights, which exercises every kernel but produces no chemistry.
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ml_conformer_generator_b200 import MLConformerGenerator, ShapeScorer  # noqa: E402
from ml_conformer_generator_b200.mol_utils import read_mol_heavy_atoms, get_context_shape, samples_to_sdf_blocks  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference")
    ap.add_argument("out_sdf")
    ap.add_argument("n_samples", nargs="?", type=int, default=64)
    ap.add_argument("--edm", default=None)
    ap.add_argument("--seer", default=None)
    ap.add_argument("--precision", default="fp16", choices=["fp16", "bf16", "tf32", "fp32"])
    args = ap.parse_args()

    _, xyz = read_mol_heavy_atoms(args.reference)
    context, _ = get_context_shape(xyz - xyz.mean(dim=0))
    kw = {}
    if args.edm and args.seer:
        kw = dict(edm_weights=args.edm, adj_mat_seer_weights=args.seer)
    else:
        from ml_conformer_generator_b200.weights import random_state_dicts
        edm_sd, seer_sd = random_state_dicts(0)
        kw = dict(edm_state_dict=edm_sd, adj_mat_seer_state_dict=seer_sd)
        print("no checkpoints given: running with random weights", file=sys.stderr)
    gen = MLConformerGenerator(diffusion_steps=100, device=torch.device("cuda:0"), precision=args.precision, **kw)
    out = gen.generate_tensors(context, n_atoms=xyz.size(0), n_samples=args.n_samples, variance=2)
    scores = ShapeScorer(gen.engine).evaluate(xyz, out["x"], out["n_nodes"])
    order = torch.argsort(scores["shape_tanimoto"], descending=True)
    blocks = samples_to_sdf_blocks(scores["aligned_coords"], out["atom_class"], out["bonds"], out["n_nodes"],
                                   names=["sample_%d shape_tanimoto=%.4f" % (i, float(scores["shape_tanimoto"][i]))
                                          for i in range(args.n_samples)])
    with open(args.out_sdf, "w") as fh:
        for i in order.tolist():
            fh.write(blocks[i])
    print("wrote %d samples to %s; best shape Tanimoto %.4f, median %.4f"
          % (args.n_samples, args.out_sdf, float(scores["shape_tanimoto"].max()), float(scores["shape_tanimoto"].median())))


if __name__ == "__main__":
    main()
