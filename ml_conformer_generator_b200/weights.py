"""Seeded random-init weights with the reference's state_dict key layout.

The HuggingFace checkpoints (`edm_moi_chembl_15_39.pt`, `adj_mat_seer_chembl_15_39.pt`,
reference conformer_generator.py:35-36) are not available offline, so benches, smoke and parity tests use
random-init weights of the named architecture.  Key names and shapes are exactly those of
`EquivariantDiffusion.state_dict()` / `AdjMatSeer.state_dict()` (reference conformer_generator.py:90-102) so the real
checkpoints drop in unchanged.  Distributions follow torch's defaults for the reference's layers
(nn.Linear: U(+-1/sqrt(fan_in)) for weight and bias; nn.Embedding: N(0,1); coord_mlp.4: xavier_uniform gain 1e-3,
reference egnn.py:100-101).  Generated with an explicit torch.Generator so they are reproducible on any box."""
import math
from typing import Dict, Tuple

import torch

from .config import (DIMENSION, HIDDEN_NF, IN_NODE_NF, N_BLOCKS, NUM_BOND_TYPES, SEER_EMBEDDING_DIM, SEER_HIDDEN,
                     SEER_NUM_EMBEDDINGS)


def _linear(sd: Dict[str, torch.Tensor], key: str, out_f: int, in_f: int, g: torch.Generator, bias: bool = True):
    bound = 1.0 / math.sqrt(in_f)
    sd[key + ".weight"] = (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound
    if bias:
        sd[key + ".bias"] = (torch.rand(out_f, generator=g) * 2 - 1) * bound


def random_edm_state_dict(seed: int = 0, timesteps: int = 1000) -> Dict[str, torch.Tensor]:
    """230 tensors, 23,897,351 parameters."""
    from .schedule import gamma_table
    g = torch.Generator().manual_seed(seed)
    h = HIDDEN_NF
    sd: Dict[str, torch.Tensor] = {"gamma.gamma": gamma_table(timesteps)}
    p = "dynamics.egnn."
    _linear(sd, p + "embedding", h, IN_NODE_NF, g)
    _linear(sd, p + "embedding_out", IN_NODE_NF, h, g)
    for b in range(N_BLOCKS):
        for name in ("gcl_0", "gcl_1"):
            q = "%se_block_%d.%s." % (p, b, name)
            _linear(sd, q + "edge_mlp.0", h, 2 * h + 2, g)
            _linear(sd, q + "edge_mlp.2", h, h, g)
            _linear(sd, q + "node_mlp.0", h, 2 * h, g)
            _linear(sd, q + "node_mlp.2", h, h, g)
            _linear(sd, q + "att_mlp.0", 1, h, g)
        q = "%se_block_%d.gcl_equiv." % (p, b)
        _linear(sd, q + "coord_mlp.0", h, 2 * h + 2, g)
        _linear(sd, q + "coord_mlp.2", h, h, g)
        bound = 0.001 * math.sqrt(6.0 / (h + 1))
        sd[q + "coord_mlp.4.weight"] = (torch.rand(1, h, generator=g) * 2 - 1) * bound
    return sd


def random_seer_state_dict(seed: int = 1) -> Dict[str, torch.Tensor]:
    """22 tensors, 21,800,531 parameters."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    n, e, d = SEER_HIDDEN, SEER_EMBEDDING_DIM, DIMENSION
    _linear(sd, "gcn1.linear", n, e, g)
    for k in ("gcn2", "gcn3", "gcn4"):
        _linear(sd, k + ".linear", n, n, g)
    _linear(sd, "resize", d * NUM_BOND_TYPES, n, g)
    sd["nodes_embedding.weight"] = torch.randn(SEER_NUM_EMBEDDINGS, e, generator=g)
    _linear(sd, "nodes_coord_fc", d * e, d, g)
    _linear(sd, "gcn1_dm.linear", n, e, g)
    for k in ("gcn2_dm", "gcn3_dm"):
        _linear(sd, k + ".linear", n, n, g)
    _linear(sd, "dm_resize", 1, n, g)
    sd["dm_nodes_embedding.weight"] = torch.randn(SEER_NUM_EMBEDDINGS, e, generator=g)
    return sd


def random_state_dicts(seed: int = 0) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    return random_edm_state_dict(seed), random_seer_state_dict(seed + 1)
