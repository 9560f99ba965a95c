"""Gaussian shape similarity of generated conformers on the GPU -- host side.

Mirrors the tensor part of the reference's `evaluate_samples` (cheminformatics/pipeline.py:37-86) and the functions of
cheminformatics/shape_similarity.py it calls, with the same names where a counterpart exists.  The clique series and the
grid overlap run in `libmlcg_b200.so` (`mlcg_shape_moments`, `mlcg_shape_tanimoto`); the 3x3 eigen-decomposition stays on
the host in torch, exactly as the reference does it (`torch.linalg.eigh`, shape_similarity.py:139), so the principal-frame
sign / order conventions are the reference's.  There is no CPU fallback.
"""
from typing import Dict, Optional

import numpy as np
import torch

from .engine import Engine, _ptr

ATOM_RADIUS = 1.60   # reference shape_similarity.py:14
AMPLITUDE = 2.70     # reference shape_similarity.py:15
GRID_POINTS = 40     # tanimoto_score default, :470
MAX_ATOMS = 64


def get_alpha(atom_radius: float = ATOM_RADIUS, gaussian_amplitude: float = AMPLITUDE) -> float:
    """reference shape_similarity.py:322-329."""
    lyambda_ = 4 * np.pi / 3 / gaussian_amplitude
    k_a = np.pi / lyambda_ ** (2 / 3)
    return k_a / atom_radius ** 2


def orientation_matrices() -> torch.Tensor:
    """(4,3,3): identity and the three rotations by pi evaluate_samples tries (pipeline.py:46-51), built like
    rotate_coord (shape_similarity.py:422-436) from float32 cos/sin of float32 pi."""
    pi = torch.pi
    mats = [torch.eye(3)]
    for angles in (torch.tensor([pi, 0, 0]), torch.tensor([0, pi, 0]), torch.tensor([0, 0, pi])):
        c, s = torch.cos(angles), torch.sin(angles)
        rx = torch.tensor([[1, 0, 0], [0, c[0], -s[0]], [0, s[0], c[0]]])
        ry = torch.tensor([[c[1], 0, s[1]], [0, 1, 0], [-s[1], 0, c[1]]])
        rz = torch.tensor([[c[2], -s[2], 0], [s[2], c[2], 0], [0, 0, 1]])
        mats.append(rx @ ry @ rz)
    return torch.stack(mats).to(torch.float32)


def grid_axes(ref_pts: torch.Tensor, n: int = GRID_POINTS, bounds_scale: float = 6, max_sigma: float = ATOM_RADIUS) -> torch.Tensor:
    """(3,n) grid coordinates of the reference's Grid as tanimoto_score builds it (shape_similarity.py:476-480, 381-403).
    The reference reduces cat(ref, cand) over dim=1 -- the x/y/z components of each ATOM -- and reads entries 0, 1, 2, so
    the bounds of the x, y, z axes are the smallest / largest component of the first, second and third atom of the
    reference molecule (+- 6 * 1.6).  With at least three reference atoms the grid therefore does not depend on the
    candidate; that case is the one supported here."""
    if ref_pts.size(0) < 3:
        raise ValueError("the reference molecule needs at least 3 heavy atoms")
    p = ref_pts.detach().cpu().to(torch.float32)
    lo = p.min(dim=1).values - bounds_scale * max_sigma
    hi = p.max(dim=1).values + bounds_scale * max_sigma
    return torch.stack([torch.linspace(lo[k], hi[k], n) for k in range(3)])


class ShapeScorer:
    """Batched get_shape_quadrupole_for_molecule + tanimoto_score on one GPU."""

    def __init__(self, engine: Optional[Engine] = None, device: torch.device = torch.device("cuda:0"),
                 amplitude: float = AMPLITUDE, atom_radius: float = ATOM_RADIUS, n_terms: int = 6):
        self.engine = engine if engine is not None else Engine(device, "bf16")
        self.device = self.engine.device
        self.amplitude, self.atom_radius, self.n_terms = float(amplitude), float(atom_radius), int(n_terms)
        self._orient = orientation_matrices()

    # -- get_shape_quadrupole_for_molecule, batched ---------------------------------------------------------------
    def principal_frames(self, coords: torch.Tensor, n_nodes: torch.Tensor) -> Dict[str, torch.Tensor]:
        """coords (B,N,3), n_nodes (B).  Returns `moments` (B,3) descending principal shape-quadrupole moments, `frames`
        (B,12) = shift + rotation with principal = (x - shift) @ rotation, `points` (B,N,3) the principal-frame
        coordinates (zero beyond n_nodes), `volume` (B,), `tensor0` (B,3,3)."""
        e = self.engine
        coords = coords.to(self.device, torch.float32).contiguous()
        if coords.dim() != 3 or coords.size(2) != 3:
            raise ValueError("coords must be (B, N, 3)")
        B, N = coords.size(0), coords.size(1)
        n_nodes = torch.as_tensor(n_nodes).to(torch.int32).cpu()
        if n_nodes.numel() != B or int(n_nodes.min()) < 1 or int(n_nodes.max()) > min(N, MAX_ATOMS):
            raise ValueError("n_nodes must hold B values in [1, min(N, 64)]")
        nn_dev = n_nodes.to(self.device)
        out = torch.empty(B, 16, device=self.device)
        e._check(e.lib.mlcg_shape_moments(e.h, _ptr(coords), _ptr(nn_dev), B, N, self.amplitude, self.atom_radius,
                                          self.n_terms, _ptr(out), e._stream()), "shape_moments")
        o = out.cpu()
        t0 = o[:, 4:13].reshape(B, 3, 3).contiguous()
        # host: eigenvectors, moments in the rotated frame, descending order (reference :139-140, 193-201)
        _, vecs = torch.linalg.eigh(t0)
        rotated_t = vecs.transpose(1, 2) @ t0 @ vecs
        moments, order = torch.sort(torch.diagonal(rotated_t, dim1=1, dim2=2), descending=True, dim=1)
        rot = torch.gather(vecs, 2, order.unsqueeze(1).expand(B, 3, 3))
        shift = o[:, 13:16] + o[:, 1:4]
        frames = torch.cat([shift, rot.reshape(B, 9)], dim=1).contiguous()
        mask = (torch.arange(N).unsqueeze(0) < n_nodes.unsqueeze(1)).unsqueeze(-1)
        points = ((coords.cpu() - shift.unsqueeze(1)) @ rot) * mask
        return {"moments": moments, "frames": frames, "points": points, "volume": o[:, 0].clone(), "tensor0": t0,
                "first_moments": o[:, 1:4].clone()}

    # -- tanimoto_score over samples x orientations ---------------------------------------------------------------
    def tanimoto(self, ref_pts: torch.Tensor, coords: torch.Tensor, n_nodes: torch.Tensor, frames: torch.Tensor,
                 n: int = GRID_POINTS) -> Dict[str, torch.Tensor]:
        """ref_pts (n_ref,3): reference atoms in their principal frame.  Returns `scores` (B,4) for the orientations of
        orientation_matrices() and `aligned` (B,4,N,3) the candidate coordinates in each of them (device tensors)."""
        e = self.engine
        coords = coords.to(self.device, torch.float32).contiguous()
        B, N = coords.size(0), coords.size(1)
        ref = ref_pts.to(self.device, torch.float32).contiguous()
        if ref.size(0) > MAX_ATOMS:
            raise ValueError("at most 64 reference atoms")
        axes = grid_axes(ref_pts, n).to(self.device).contiguous()
        orient = self._orient.reshape(-1, 9).to(self.device).contiguous()
        nn_dev = torch.as_tensor(n_nodes).to(torch.int32).to(self.device)
        fr = frames.to(self.device, torch.float32).contiguous()
        work = torch.empty(n ** 3 + 1, device=self.device)
        scores = torch.empty(B, orient.size(0), device=self.device)
        aligned = torch.zeros(B, orient.size(0), N, 3, device=self.device)
        e._check(e.lib.mlcg_shape_tanimoto(e.h, _ptr(ref), ref.size(0), _ptr(coords), _ptr(nn_dev), B, N, _ptr(fr),
                                           _ptr(orient), orient.size(0), _ptr(axes), n, self.amplitude, self.atom_radius,
                                           _ptr(work), _ptr(scores), _ptr(aligned), e._stream()), "shape_tanimoto")
        return {"scores": scores, "aligned": aligned}

    # -- the tensor part of evaluate_samples ------------------------------------------------------------------------
    def evaluate(self, reference_coord: torch.Tensor, sample_coords: torch.Tensor, n_nodes: torch.Tensor) -> Dict[str, torch.Tensor]:
        """reference_coord (n_ref,3) heavy-atom coordinates of the reference conformer; sample_coords (B,N,3) / n_nodes
        (B) heavy atoms of the samples.  Returns `shape_tanimoto` (B,) the best of the four orientations (first strict
        maximum, pipeline.py:76-85), `best_orientation` (B,), `aligned_coords` (B,N,3) the sample in that orientation,
        `scores` (B,4), `reference_coords` (n_ref,3) the reference in its principal frame, `moments` (B,3)."""
        ref = reference_coord.to(torch.float32)
        rf = self.principal_frames(ref.unsqueeze(0), torch.tensor([ref.size(0)]))
        ref_pts = rf["points"][0]
        sf = self.principal_frames(sample_coords, n_nodes)
        t = self.tanimoto(ref_pts, sample_coords, n_nodes, sf["frames"])
        scores = t["scores"].cpu()
        best = torch.zeros(scores.size(0), dtype=torch.long)
        best_score = scores[:, 0].clone()
        for k in range(1, scores.size(1)):
            better = scores[:, k] > best_score
            best = torch.where(better, torch.full_like(best, k), best)
            best_score = torch.where(better, scores[:, k], best_score)
        aligned = t["aligned"].cpu()
        idx = best.view(-1, 1, 1, 1).expand(-1, 1, aligned.size(2), 3)
        return {"shape_tanimoto": best_score, "best_orientation": best, "aligned_coords": aligned.gather(1, idx).squeeze(1),
                "scores": scores, "reference_coords": ref_pts, "moments": sf["moments"], "reference_moments": rf["moments"][0]}
