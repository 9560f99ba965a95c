"""ml_conformer_generator_b200 -- sm_100a implementation of ml_conformer_generator's hot path (EDM reverse diffusion with
the EGNN denoiser + AdjMatSeer bond GCN) behind the reference's `MLConformerGenerator` API."""
from .config import ATOM_DECODER, CONTEXT_NORMS, DIMENSION, MAX_N_NODES, MIN_N_NODES, NUM_BOND_TYPES  # noqa: F401
from .conformer_generator import MLConformerGenerator  # noqa: F401
from .engine import Engine, MlcgError  # noqa: F401


def evaluate_samples(*args, **kwargs):
    """Re-export of the reference's CPU scoring pipeline (cheminformatics/pipeline.py:17-96).  Shape / chemical
    Tanimoto scoring stays on CPU with RDKit (out of scope of the accelerated path); this forwards to the reference
    package when it is installed."""
    try:
        from mlconfgen import evaluate_samples as _ref
    except ImportError as exc:
        raise ImportError("evaluate_samples is the reference's RDKit/CPU scoring pipeline; install mlconfgen + rdkit") from exc
    return _ref(*args, **kwargs)


__all__ = ["MLConformerGenerator", "Engine", "MlcgError", "evaluate_samples"]
