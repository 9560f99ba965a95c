"""ml_conformer_generator_b200 -- sm_100a implementation of ml_conformer_generator's hot path (EDM reverse diffusion with
the EGNN denoiser + AdjMatSeer bond GCN) behind the reference's `MLConformerGenerator` API."""
from .config import ATOM_DECODER, CONTEXT_NORMS, DIMENSION, MAX_N_NODES, MIN_N_NODES, NUM_BOND_TYPES  # noqa: F401
from .conformer_generator import MLConformerGenerator  # noqa: F401
from .engine import Engine, MlcgError  # noqa: F401
from .shape_similarity import ShapeScorer  # noqa: F401


def evaluate_samples(*args, **kwargs):
    """Re-export of the reference's scoring pipeline (cheminformatics/pipeline.py:17-96), which needs RDKit for the Morgan
    fingerprints and the mol blocks; this forwards to the reference package when it is installed.  The tensor part of
    it -- principal shape-quadrupole frames and the 4-orientation grid Tanimoto -- runs on the GPU through
    `ShapeScorer.evaluate(reference_coord, sample_coords, n_nodes)` without RDKit."""
    try:
        from mlconfgen import evaluate_samples as _ref
    except ImportError as exc:
        raise ImportError("evaluate_samples is the reference's RDKit/CPU scoring pipeline; install mlconfgen + rdkit") from exc
    return _ref(*args, **kwargs)


__all__ = ["MLConformerGenerator", "Engine", "MlcgError", "ShapeScorer", "evaluate_samples"]
