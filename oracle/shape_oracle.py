"""CPU restatement of the reference's Gaussian shape-similarity path (test infrastructure only -- see oracle/README).

Follows /root/reference/src/mlconfgen/cheminformatics/shape_similarity.py (Grant & Pickup Gaussian volumes) and the
tensor part of cheminformatics/pipeline.py:37-86, function by function, in plain torch fp32 on the CPU.  Pinned against
the reference itself by oracle/make_golden.py -> tests/golden/shape.npz (tests/test_oracle_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import itertools
import math
from typing import List, Tuple

import numpy as np
import torch

ATOM_RADIUS = 1.60   # shape_similarity.py:14
AMPLITUDE = 2.70     # shape_similarity.py:15
N_TERMS = 6          # get_shape_quadrupole_for_molecule default (:22)
GRID_N = 40          # tanimoto_score default (:470)


def get_alpha(atom_radius: float = ATOM_RADIUS, amplitude: float = AMPLITUDE) -> float:
    """shape_similarity.py:322-329."""
    lam = 4 * np.pi / 3 / amplitude
    k_a = np.pi / lam ** (2 / 3)
    return k_a / atom_radius ** 2


ALPHA = get_alpha()


def neighbour_matrix(coords: torch.Tensor, threshold: float) -> torch.Tensor:
    """0/1 matrix of pairs with 0 < distance < threshold (get_valid_combinations, :243-257)."""
    d = torch.sqrt(((coords.unsqueeze(1) - coords.unsqueeze(0)) ** 2).sum(2))
    return ((d > 0) & (d < threshold))


def cliques(adj: torch.Tensor, order: int) -> torch.Tensor:
    """All `order`-cliques as increasing index tuples, in the lexicographic order the reference's backtracking emits
    them (find_r_cliques_fast, :267-311)."""
    n = adj.size(0)
    a = adj.numpy()
    found: List[Tuple[int, ...]] = []

    def extend(partial, cands):
        if len(partial) == order:
            found.append(tuple(partial))
            return
        for v in cands:
            extend(partial + [v], [w for w in cands if w > v and a[v, w]])

    extend([], list(range(n)))
    if not found:
        return torch.empty((0, order), dtype=torch.long)
    return torch.tensor(found, dtype=torch.long)


def product_of_gaussians(centers: torch.Tensor, alpha: float, amplitude: float):
    """shape_similarity.py:206-230: centres (C, k, 3) -> new centre (C,3), new alpha (float), new amplitude (C,)."""
    k = centers.size(1)
    new_c = centers.mean(1)
    r2 = (centers ** 2).sum(-1).sum(-1)
    xyz = (centers.sum(1) ** 2).sum(-1) / k
    gamma = r2 - xyz
    return new_c, k * alpha, amplitude ** k * torch.exp(-alpha * gamma)


def _second_moments(points: torch.Tensor, combos, alpha: float, amplitude: float):
    """ii / ij second-moment sums of the inclusion-exclusion series for `points` (:89-125, repeated at :142-181)."""
    vol1 = (np.pi / alpha) ** 1.5
    ii = (amplitude * vol1 * (points ** 2 + 1 / (2 * alpha))).sum(0)
    ij = amplitude * vol1 * torch.stack((points[:, 0] * points[:, 1], points[:, 0] * points[:, 2],
                                         points[:, 1] * points[:, 2]), 0).sum(-1)
    for k, idx in combos.items():
        if idx.numel() == 0:
            continue
        c, a_k, amp = product_of_gaussians(points[idx], alpha, amplitude)
        sign = (-1) ** (k - 1)
        vk = (np.pi / a_k) ** 1.5
        ii = ii + sign * (amp.unsqueeze(-1) * vk * (c ** 2 + 1 / (2 * a_k))).sum(0)
        ij = ij + sign * (amp.unsqueeze(0) * torch.stack((c[:, 0] * c[:, 1], c[:, 0] * c[:, 2], c[:, 1] * c[:, 2]), 0)
                          * vk).sum(-1)
    return ii, ij


def _tensor(ii, ij, volume):
    return torch.tensor([[ii[0].item(), ij[0].item(), ij[1].item()],
                         [ij[0].item(), ii[1].item(), ij[2].item()],
                         [ij[1].item(), ij[2].item(), ii[2].item()]]) / volume


def shape_quadrupole(coords: torch.Tensor, amplitude: float = AMPLITUDE, radius: float = ATOM_RADIUS,
                     n_terms: int = N_TERMS, detail: bool = False):
    """get_shape_quadrupole_for_molecule (:18-203): principal shape-quadrupole moments (descending) and the coordinates
    in that principal frame.  `detail=True` also returns volume, first moments and the pre-rotation tensor."""
    alpha = get_alpha(radius, amplitude)
    thr = 2 * amplitude
    adj = neighbour_matrix(coords, thr)
    combos = {k: cliques(adj, k) for k in range(2, n_terms + 1)}
    vol1 = (np.pi / alpha) ** 1.5
    volume = coords.size(0) * amplitude * vol1
    first = (amplitude * coords * vol1).sum(0)
    for k, idx in combos.items():
        if idx.numel() == 0:
            continue
        c, a_k, amp = product_of_gaussians(coords[idx], alpha, amplitude)
        sign = (-1) ** (k - 1)
        vk = (np.pi / a_k) ** 1.5
        volume = volume + sign * (amp * vk).sum()
        first = first + sign * (amp.unsqueeze(-1) * c * vk).sum(0)
    first = first / volume
    centred = coords - first
    ii0, ij0 = _second_moments(centred, combos, alpha, amplitude)
    t0 = _tensor(ii0, ij0, volume)
    _, vecs = torch.linalg.eigh(t0)
    rotated = centred @ vecs
    ii, ij = _second_moments(rotated, combos, alpha, amplitude)
    t = _tensor(ii, ij, volume)
    moments, order = torch.sort(torch.diag(t), descending=True)
    pts = rotated[:, order]
    if detail:
        return moments, pts, {"volume": float(volume), "first_moments": first, "tensor0": t0, "n_cliques":
                              [int(v.size(0)) for v in combos.values()]}
    return moments, pts


def grid_axes(ref: torch.Tensor, cand: torch.Tensor, n: int = GRID_N, bounds_scale: float = 6, max_sigma: float = ATOM_RADIUS):
    """Grid.__init__ as tanimoto_score calls it (:476-480 + :381-403).  NOTE the reference reduces over dim=1 of the
    (atoms, 3) array, i.e. over x/y/z per ATOM, and then reads entries 0, 1, 2: the x / y / z axis bounds are the min and
    max coordinate component of the first, second and third atom of cat(ref, cand) -- reproduced as is."""
    cat = torch.cat((ref, cand), 0)
    lo = cat.min(dim=1).values - bounds_scale * max_sigma
    hi = cat.max(dim=1).values + bounds_scale * max_sigma
    return [torch.linspace(lo[k], hi[k], n) for k in range(3)]


def density(coords: torch.Tensor, pts: torch.Tensor, alpha: float = ALPHA, amplitude: float = AMPLITUDE) -> torch.Tensor:
    """torch_evaluate_density_on_grid (:406-419)."""
    d2 = torch.cdist(pts, coords) ** 2
    return 1 - torch.prod(1 - amplitude * torch.exp(-d2 * alpha), dim=-1)


def tanimoto(ref: torch.Tensor, cand: torch.Tensor, n: int = GRID_N) -> float:
    """tanimoto_score (:465-492)."""
    xs, ys, zs = grid_axes(ref, cand, n)
    xg, yg, zg = torch.meshgrid(xs, ys, zs, indexing="ij")
    pts = torch.stack([xg.flatten(), yg.flatten(), zg.flatten()], -1)
    f = density(ref, pts)
    g = density(cand, pts)
    fg = (f * g).sum()
    return float(fg / ((f * f).sum() + (g * g).sum() - fg))


def rotation(angles: torch.Tensor) -> torch.Tensor:
    """rotate_coord's matrix product Rx @ Ry @ Rz (:422-436); float32 cos/sin of float32 pi included."""
    c, s = torch.cos(angles), torch.sin(angles)
    rx = torch.tensor([[1, 0, 0], [0, c[0], -s[0]], [0, s[0], c[0]]])
    ry = torch.tensor([[c[1], 0, s[1]], [0, 1, 0], [-s[1], 0, c[1]]])
    rz = torch.tensor([[c[2], -s[2], 0], [s[2], c[2], 0], [0, 0, 1]])
    return rx, ry, rz


def orientations() -> List[torch.Tensor]:
    """Identity + the three pi rotations evaluate_samples tries (pipeline.py:46-51, 76-85)."""
    pi = torch.pi
    return [torch.zeros(3), torch.tensor([pi, 0, 0]), torch.tensor([0, pi, 0]), torch.tensor([0, 0, pi])]


def rotate(coord: torch.Tensor, angles: torch.Tensor) -> torch.Tensor:
    rx, ry, rz = rotation(angles)
    return torch.matmul(torch.matmul(torch.matmul(coord, rx), ry), rz)


def evaluate_shape(ref_coord: torch.Tensor, sample_coord: torch.Tensor):
    """Tensor part of evaluate_samples for one sample (pipeline.py:37-86): both molecules to their principal frames, score
    the sample as is and after each pi rotation, keep the first strict maximum.  Returns (scores[4], best index, best
    coordinates, reference principal coordinates)."""
    ref_c = ref_coord - ref_coord.mean(0)
    _, ref_pts = shape_quadrupole(ref_c)
    s_c = sample_coord - sample_coord.mean(0)
    _, s_pts = shape_quadrupole(s_c)
    scores, best, best_idx, best_coord = [], None, 0, s_pts
    for k, ang in enumerate(orientations()):
        pts = s_pts if k == 0 else rotate(s_pts, ang)
        sc = tanimoto(ref_pts, pts)
        scores.append(sc)
        if best is None or sc > best:
            best, best_idx, best_coord = sc, k, pts
    return scores, best_idx, best_coord, ref_pts
