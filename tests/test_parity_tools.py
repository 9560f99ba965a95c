"""CPU: the parity tooling itself -- Philox known-answer vectors of the host restatement, the noise-tape helper, the
bond-agreement metric, and the oracle port against the T = 100 golden runs of the reference (a few of the 101 calls)."""
import numpy as np
import pytest
import torch

from conftest import golden, rel_l2
from ml_conformer_generator_b200.config import CONTEXT_NORMS
from oracle import edm_oracle as O
from oracle import philox_oracle as PH
from tools import parity_check as PC


def test_philox_known_answers():
    for ctr, key, out in PH.KAT:
        r = PH.philox4x32_10(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32))
        assert tuple(int(v) for v in r) == out


def test_philox_draws_do_not_overlap():
    """Every (sample, atom, draw) owns its own counters: no raw value is shared between consecutive draws (the round-1
    generator re-used 9 of 12 uniforms between draw d and d + 1)."""
    a = PH.raw_noise(7, np.arange(50), 39, 3)
    b = PH.raw_noise(7, np.arange(50), 39, 4)
    assert len(np.intersect1d(a.ravel(), b.ravel())) == 0
    x = PH.raw_noise(1, np.arange(3000), 39, 0)
    assert abs(x.mean()) < 5e-3 and abs(x.std() - 1) < 5e-3
    c = PH.combined_noise(1, np.arange(8), np.array([15, 39, 1, 2, 20, 30, 39, 17]), 39, 2)
    assert np.abs(c[..., :3].sum(1)).max() < 1e-9 and np.all(c[0, 15:] == 0)


def test_noise_tape_helper_matches_oracle():
    g = golden("edm_forward_T10")
    B = len(g["n_nodes"])
    ref = O.NoiseTape.draw(int(g["n_pairs"]), B, int(g["n_max"]), int(g["seed"])).stacked()
    assert torch.equal(PC.noise_tape(g), ref)


def test_bond_agreement_counts_only_the_strict_lower_triangle():
    ref = torch.zeros(2, 42, 42, dtype=torch.long)
    ref[0, 3, 1] = 2
    ref[1, 10, 2] = 1
    mine = ref.clone()
    mine[0, 1, 3] = 4       # upper triangle: never consumed (reference utils/mol_utils.py:210-211)
    mine[0, 5, 5] = 1       # diagonal: never consumed
    r = PC.bond_agreement(mine, ref, [17, 15])
    assert r["flips"] == 0 and r["lower_triangle"] == 1.0 and r["lower_triangle_entries"] == 2 * 861
    assert r["real_pair_entries"] == 17 * 16 // 2 + 15 * 14 // 2
    mine[1, 10, 2] = 0      # a real-pair flip
    mine[1, 41, 40] = 1     # a flip between padding slots
    logits = torch.zeros(2, 42, 42, 5)
    logits[1, 10, 2, 1] = 0.25
    r = PC.bond_agreement(mine, ref, [17, 15], logits)
    assert r["flips"] == 2 and r["flips_real_pairs"] == 1
    assert abs(r["real_pairs"] - (1 - 1 / r["real_pair_entries"])) < 1e-6
    assert abs(r["max_ref_margin_at_flips"] - 0.25) < 1e-9


@pytest.mark.parametrize("name", ["edm_forward_T100_n39", "edm_forward_T100_mixed"])
def test_oracle_port_on_T100_golden_calls(state_dicts, name):
    """The oracle port reproduces the reference's recorded denoiser outputs of the T = 100 runs (first, middle, last call)."""
    g = golden(name)
    assert g["traj_z"].shape[0] == 101 and int(g["T"]) == 100
    n_nodes = torch.from_numpy(g["n_nodes"])
    nm, em = O.prepare_masks(n_nodes, int(g["n_max"]))
    ctx = O.batch_context(O.normalise_context(torch.tensor(g["raw_context"], dtype=torch.float32), CONTEXT_NORMS), nm)
    with torch.no_grad():
        for k in (0, 50, 100):
            eps = O.egnn_dynamics(state_dicts[0], torch.from_numpy(g["traj_t"][k]), torch.from_numpy(g["traj_z"][k]), nm, em, ctx)
            assert rel_l2(eps, g["traj_eps"][k]) < 1e-5
    # schedule sanity of the stored times: t = (s+1)/T for s = 99..0, then 0
    t = g["traj_t"][:, 0, 0]
    assert np.allclose(t[:100], (np.arange(99, -1, -1) + 1) / 100.0) and t[100] == 0.0
