"""CPU: the oracle port against the golden vectors produced by the real reference (oracle/make_golden.py)."""
import numpy as np
import torch

from conftest import golden, rel_l2
from ml_conformer_generator_b200.config import CONTEXT_NORMS
from oracle import edm_oracle as O


def _ctx(raw, node_mask):
    return O.batch_context(O.normalise_context(torch.tensor(raw, dtype=torch.float32), CONTEXT_NORMS), node_mask)


def test_schedule_tables_bit_exact():
    g = golden("schedule")
    for T in (10, 100, 1000):
        assert np.array_equal(O.gamma_table(T).numpy(), g["gamma_%d" % T])
    # SURVEY 8a2 known answers for T = 100
    t = O.gamma_table(100)
    assert abs(float(t[0]) + 11.5116) < 1e-3 and abs(float(t[100]) - 11.4741) < 1e-3


def test_masks_bit_exact():
    g = golden("masks")
    nm, em = O.prepare_masks(torch.from_numpy(g["n_nodes"]), int(g["n_max"]))
    assert np.array_equal(nm.numpy(), g["node_mask"])
    assert np.array_equal(em.numpy(), g["edge_mask"])
    assert np.allclose(_ctx(g["raw_context"], nm).numpy(), g["context"], atol=0, rtol=0)


def test_egnn_dynamics(state_dicts):
    sd = state_dicts[0]
    for name in ("egnn_small", "egnn_n39"):
        g = golden(name)
        nm, em = O.prepare_masks(torch.from_numpy(g["n_nodes"]), int(g["n_max"]))
        with torch.no_grad():
            eps = O.egnn_dynamics(sd, torch.from_numpy(g["t"]), torch.from_numpy(g["xh"]), nm, em,
                                  _ctx(g["raw_context"], nm))
        assert rel_l2(eps, g["eps"]) < 1e-5
        assert float(eps[nm.squeeze(-1) == 0].abs().max()) == 0.0  # padded rows exactly zero


def test_edm_forward_golden(state_dicts):
    sd = state_dicts[0]
    g = golden("edm_forward_T6_r1")
    nm, em = O.prepare_masks(torch.from_numpy(g["n_nodes"]), int(g["n_max"]))
    tape = O.NoiseTape.draw(int(g["n_pairs"]), nm.size(0), int(g["n_max"]), int(g["seed"]))
    with torch.no_grad():
        x, h = O.edm_forward(sd, O.gamma_table(int(g["T"])), nm, em, _ctx(g["raw_context"], nm), tape,
                             int(g["resample_steps"]))
    assert rel_l2(x, g["x"]) < 1e-4
    assert np.array_equal(h.numpy(), g["h"])


def test_edm_merge_golden(state_dicts):
    sd = state_dicts[0]
    g = golden("edm_merge_T10_L4")
    nm, em = O.prepare_masks(torch.from_numpy(g["n_nodes"]), int(g["n_max"]))
    tape = O.NoiseTape.draw(int(g["n_pairs"]), nm.size(0), int(g["n_max"]), int(g["seed"]))
    with torch.no_grad():
        x, h = O.edm_merge_fragments(sd, O.gamma_table(int(g["T"])), nm, em, torch.from_numpy(g["fixed_mask"]),
                                     _ctx(g["raw_context"], nm), torch.from_numpy(g["z_known"]), tape,
                                     int(g["diffusion_level"]), int(g["resample_steps"]), int(g["blend_power"]))
    assert rel_l2(x, g["x"]) < 1e-4
    assert np.array_equal(h.numpy(), g["h"])


def test_seer_golden(state_dicts):
    g = golden("seer")
    with torch.no_grad():
        lo = O.seer_forward(state_dicts[1], torch.from_numpy(g["elements"]), torch.from_numpy(g["dist_mat"]),
                            torch.from_numpy(g["adj_mat"]))
    assert rel_l2(lo, g["logits"]) < 1e-5
    assert np.array_equal(O.bond_orders(lo).numpy(), g["bonds"])


def test_step_scalars_match_oracle():
    from ml_conformer_generator_b200.schedule import gamma_table, step_scalars
    gam = gamma_table(100)
    assert torch.equal(gam, O.gamma_table(100))
    for s in (0, 1, 37, 99):
        a, b = step_scalars(gam, s), O.step_coefficients(gam, s)
        assert a["t"] == b["t"] and a["alpha_ts"] == b["alpha_ts"] and a["c_eps"] == b["eps_coef"]
        assert a["c_sigma"] == b["sigma"] and a["alpha_s"] == b["alpha_s"] and a["sigma_s"] == b["sigma_s"]


def test_shape_oracle_against_reference():
    """oracle/shape_oracle.py (Gaussian shape quadrupole frame + grid Tanimoto, reference shape_similarity.py and
    pipeline.py:37-86) against vectors produced by the reference's own functions."""
    from oracle import shape_oracle as S
    g = golden("shape")
    ref = torch.from_numpy(g["ref_xyz"])
    mom, pts = S.shape_quadrupole(ref - ref.mean(0))
    assert np.allclose(mom.numpy(), g["ref_moments"], rtol=1e-5) and np.allclose(pts.numpy(), g["ref_pts"], atol=1e-5)
    for b in range(len(g["n_nodes"])):
        n = int(g["n_nodes"][b])
        scores, best, coord, _ = S.evaluate_shape(ref, torch.from_numpy(g["coords"][b, :n]))
        assert np.allclose(scores, g["scores"][b], atol=2e-6), (b, scores, g["scores"][b])
        assert best == int(g["best_idx"][b])
        assert np.allclose(coord.numpy(), g["best_coord"][b, :n], atol=2e-5)
