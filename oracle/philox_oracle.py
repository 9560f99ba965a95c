"""TEST INFRASTRUCTURE ONLY -- host restatement (numpy) of the device noise generator of
`ml_conformer_generator_b200/csrc/mlcg_kernels.cuh` (`philox4x32_10`, `box_muller`, `raw_noise`).

The reference draws its noise with `torch.randn` from the global generator
(reference equivariant_diffusion.py:56-76, 341-363); parity runs inject that tape.  When no tape is injected the CUDA
path draws from Philox4x32-10 (Salmon et al., SC'11) with an explicit key / counter assignment:

    key     = {seed lo, seed hi}
    counter = {3 * draw + k, atom, sample id lo, sample id hi},  k = 0, 1, 2

which gives 12 uint32 per (sample, atom, draw); consecutive pairs go through Box-Muller
(u = (a + 0.5) / 2^32, v = (b + 0.5) / 2^32 -> sqrt(-2 ln u) * (sin 2 pi v, cos 2 pi v)) and the first 11 values are the
raw N(0,1) draws (3 position + 8 feature channels).  This file recomputes them so the GPU tests can check the generator
value by value (known-answer vectors of the Random123 distribution pin the block function itself).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter: np.ndarray, key: np.ndarray) -> np.ndarray:
    """counter (..., 4) uint32, key (..., 2) uint32 -> (..., 4) uint32."""
    c = [counter[..., i].astype(np.uint64) for i in range(4)]
    k0 = key[..., 0].astype(np.uint64)
    k1 = key[..., 1].astype(np.uint64)
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def box_muller(a: np.ndarray, b: np.ndarray):
    u = (a.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float64)
    v = (b.astype(np.float32) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33)).astype(np.float64)
    r = np.sqrt(-2.0 * np.log(u))
    return r * np.sin(2 * np.pi * v), r * np.cos(2 * np.pi * v)


def raw_noise(seed: int, sample_ids: np.ndarray, n_atoms: int, draw: int) -> np.ndarray:
    """(B, n_atoms, 11) float64 raw normals for the given global sample ids and draw index."""
    ids = np.asarray(sample_ids, dtype=np.uint64).reshape(-1)
    B = ids.size
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    out = np.zeros((B, n_atoms, 12))
    atoms = np.arange(n_atoms, dtype=np.uint32)
    for k in range(3):
        ctr = np.zeros((B, n_atoms, 4), dtype=np.uint32)
        ctr[..., 0] = np.uint32((3 * draw + k) & 0xFFFFFFFF)
        ctr[..., 1] = atoms[None, :]
        ctr[..., 2] = (ids & MASK).astype(np.uint32)[:, None]
        ctr[..., 3] = (ids >> np.uint64(32)).astype(np.uint32)[:, None]
        r = philox4x32_10(ctr, np.broadcast_to(key, ctr.shape[:-1] + (2,)))
        a0, a1 = box_muller(r[..., 0], r[..., 1])
        c0, c1 = box_muller(r[..., 2], r[..., 3])
        out[..., 4 * k + 0], out[..., 4 * k + 1], out[..., 4 * k + 2], out[..., 4 * k + 3] = a0, a1, c0, c1
    return out[..., :11]


def combined_noise(seed: int, sample_ids: np.ndarray, n_nodes: np.ndarray, n_max: int, draw: int) -> np.ndarray:
    """Masked, centre-of-gravity-free combined noise (reference equivariant_diffusion.py:341-363) from raw_noise."""
    raw = raw_noise(seed, sample_ids, n_max, draw)
    mask = (np.arange(n_max)[None, :] < np.asarray(n_nodes)[:, None]).astype(np.float64)[..., None]
    raw = raw * mask
    mean = raw[..., :3].sum(axis=1, keepdims=True) / np.asarray(n_nodes, dtype=np.float64)[:, None, None]
    raw[..., :3] = raw[..., :3] - mean * mask
    return raw


# Known-answer vectors of Philox4x32-10 from the Random123 distribution (kat_vectors): counter, key -> output
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
     (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]
