"""CPU: the host-side plan of the fused edge kernel (mlcg_plan_edge_tiles).  No device needed.

Invariants checked for ragged batches: the tiles of a molecule cover its n(n-1) target-major edge rows exactly once and in
order; no tile exceeds 128 rows or 12 targets; a target cut by a tile boundary is marked on both sides, as "carried" iff
both tiles belong to the same CTA and are consecutive there, otherwise with a shared side-buffer id; the owner map equals
the tile ranges the kernel derives from (blockIdx, gridDim, n_tiles)."""
import ctypes as C

import numpy as np
import pytest

from ml_conformer_generator_b200 import _lib

WHOLE, CARRY = -1, -2


def plan(n_nodes, N=39, num_sms=148):
    lib = _lib.load()
    n_nodes = np.ascontiguousarray(n_nodes, dtype=np.int32)
    cap = int((n_nodes.astype(np.int64) * (n_nodes - 1)).sum() // 64 + 2 * len(n_nodes) + 8)
    tiles = np.zeros((cap, 8), np.int32)
    owner = np.zeros(cap, np.int32)
    n_fix = C.c_int32(0)
    nt = lib.mlcg_plan_edge_tiles(n_nodes.ctypes.data, len(n_nodes), N, num_sms, tiles.ctypes.data, owner.ctypes.data, cap,
                                  C.byref(n_fix))
    assert 0 < nt <= cap
    return tiles[:nt], owner[:nt], n_fix.value


def kernel_owner(n_tiles, num_sms, pair=True):
    """The tile ranges of k_tc_edge (pair mode: a contiguous range per CTA pair, first half to CTA 0)."""
    grid = min(n_tiles, num_sms)
    pair = pair and grid >= 2
    own = np.zeros(n_tiles, np.int64)
    if pair:
        grid &= ~1
        npairs = grid // 2
        for pr in range(npairs):
            t0, t1 = pr * n_tiles // npairs, (pr + 1) * n_tiles // npairs
            n_iter = (t1 - t0 + 1) >> 1
            for t in range(t0, t1):
                own[t] = 2 * pr + (1 if t >= t0 + n_iter else 0)
    else:
        for b in range(grid):
            own[b * n_tiles // grid:(b + 1) * n_tiles // grid] = b
    return own


@pytest.mark.parametrize("case", ["all_sizes", "c2_like", "small", "one_sm"])
def test_plan_invariants(case):
    rng = np.random.RandomState(3)
    if case == "all_sizes":
        n_nodes, sms = np.concatenate([np.arange(1, 40), rng.randint(1, 40, 200)]), 148
    elif case == "c2_like":
        n_nodes, sms = np.full(300, 39), 148
    elif case == "small":
        n_nodes, sms = np.array([17, 2, 39]), 148
    else:
        n_nodes, sms = rng.randint(13, 40, 40), 1
    tiles, owner, n_fix = plan(n_nodes, 39, sms)
    assert np.array_equal(owner, kernel_owner(len(tiles), sms))
    t = 0
    fix_ids = set()
    for b, n in enumerate(n_nodes):
        nm1, covered = n - 1, 0
        first = True
        while t < len(tiles) and tiles[t, 0] == b:
            mol, off0, nrows, nn, i0, ng, fa, fb = (int(v) for v in tiles[t])
            assert nn == n and 0 <= nrows <= 128 and 1 <= ng <= 12
            if nm1 > 0:
                assert i0 * nm1 + off0 == covered                      # rows continue where the previous tile stopped
                last_row = covered + nrows - 1
                assert (last_row // nm1) - i0 + 1 == ng if nrows > 0 else True
            covered += nrows
            cut_begin = nm1 > 0 and off0 != 0
            cut_end = nm1 > 0 and covered % nm1 != 0
            assert (fa != WHOLE) == cut_begin and (fb != WHOLE) == cut_end
            assert not first or fa == WHOLE
            if cut_end:
                nxt = tiles[t + 1]
                assert nxt[0] == b and nxt[6] == fb                     # both sides carry the same mark
                if fb == CARRY:
                    assert owner[t] == owner[t + 1]
                else:
                    assert fb >= 0 and owner[t] != owner[t + 1] and fb not in fix_ids
                    fix_ids.add(fb)
            first = False
            t += 1
        assert covered == n * nm1
    assert t == len(tiles) and fix_ids == set(range(n_fix))
    # the side buffer only serves cuts between CTAs: at most one per CTA boundary
    assert n_fix <= max(min(len(tiles), sms) - 1, 0)


def test_plan_occupancy_and_errors():
    tiles, _, _ = plan(np.full(64, 39))
    assert len(tiles) == 64 * 12 and tiles[:, 2].min() >= 123           # 12 tiles of 123/124 rows per 39-atom molecule
    lib = _lib.load()
    bad = np.array([40], np.int32)
    assert lib.mlcg_plan_edge_tiles(bad.ctypes.data, 1, 39, 148, None, None, 0, None) < 0
    assert lib.mlcg_plan_edge_tiles(bad.ctypes.data, 1, 64, 148, None, None, 0, None) < 0   # N > 39 is refused


def test_edge3_block_order():
    """k_tc_edge3: the static order of the 21 (third, K chunk) weight blocks of a tile that producer and MMA
    issuer share.  Both variants visit every block once, keep each third's chunks ascending (the first MMA of a third
    overwrites its accumulator), never multiply chunk k of the second third before chunk k of the first has been generated
    (the first third is the one that waits for the A chunks), and run the last third after the other two (it releases the A
    chunks for the next tile, and its accumulator is the first third's)."""
    lib = _lib.load()
    for equiv in (0, 1):
        out = np.zeros(21, np.int32)
        assert lib.mlcg_edge_block_order(equiv, out.ctypes.data) == 21
        blocks = [(int(v) >> 3, int(v) & 7) for v in out]
        assert sorted(blocks) == [(t, k) for t in range(3) for k in range(7)]
        pos = {b: i for i, b in enumerate(blocks)}
        for t in range(3):
            assert [pos[(t, k)] for k in range(7)] == sorted(pos[(t, k)] for k in range(7))
        for k in range(7):
            assert pos[(0, k)] < pos[(1, k)] < pos[(2, k)]
        assert min(pos[(2, k)] for k in range(7)) > max(pos[(t, k)] for t in (0, 1) for k in range(7))
    seq = np.zeros(21, np.int32)
    lib.mlcg_edge_block_order(0, seq.ctypes.data)
    assert [(int(v) >> 3, int(v) & 7) for v in seq] == [(t, k) for t in range(3) for k in range(7)]
    il = np.zeros(21, np.int32)
    lib.mlcg_edge_block_order(1, il.ctypes.data)
    # the equivariant order starts on the second third after two chunks (its accumulator is released between the second and
    # the third A chunk of the tile) and then alternates
    assert [(int(v) >> 3, int(v) & 7) for v in il[:6]] == [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (1, 2)]
    assert lib.mlcg_edge_block_order(0, None) < 0
