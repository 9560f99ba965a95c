"""GPU parity on BASELINE.json's own shapes and step count, in every precision of the CUDA path.

  * teacher-forced per-step eps over ALL 101 denoiser calls of the reference's T = 100 runs (8 x 39 atoms = config 2's
    molecule size; 8 molecules of 15..39 atoms = config 3 / 5's mix), fp32 / tf32 / bf16;
  * free-running samplers with the reference's injected noise, T = 100, against the reference's final x / atom types;
  * free-running tf32 and bf16 against the exact-fp32 CUDA path (pinned to the reference at <= 2e-5 per step by the
    tests above) on > 10 000 atoms, where the CPU reference cannot reach: atom-type agreement;
  * bond orders on the strict lower triangle (what reference utils/mol_utils.py:210-211 consumes), real-atom pairs
    separately, flips reported against the reference's top-2 logit margin.

The numbers printed here (pytest -rP) are collected in profiles/r2_parity.txt.
"""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import edm_oracle as O
from tools import parity_check as PC

pytestmark = pytest.mark.gpu
GOLDENS = ["edm_forward_T100_n39", "edm_forward_T100_mixed"]


# Per-call bound over the whole T = 100 trajectory.  The reference trajectory with random-init weights blows up to
# |x| ~ 1e3 (ordinary molecules: |x| < 10); there the 10-bit-mantissa modes (tf32, fp16) reach 1.0e-3 .. 1.3e-3 -- the
# rounding of the second edge layer's two operands alone gives 1.0e-3 in a float64 emulation.  The
# north-star bar of 1e-3 is asserted on the median and on every call whose input is at ordinary scale (|x| <= 100).
TOL_T100_WORST = {"fp32": 2e-5, "tf32": 1.5e-3, "fp16": 1.5e-3, "bf16": 2e-2}
TOL_T100_ORDINARY = {"fp32": 2e-5, "tf32": 1e-3, "fp16": 1e-3, "bf16": 2e-2}


@pytest.mark.parametrize("mode", ["fp32", "tf32", "fp16", "bf16"])
@pytest.mark.parametrize("name", GOLDENS)
def test_teacher_forced_T100(engines, mode, name):
    g = golden(name)
    errs = PC.teacher_forced(engines(mode), g, chunk=26 if mode == "fp32" else 101)
    assert len(errs) == 101
    worst = int(np.argmax(errs))
    xmax = np.abs(g["traj_z"][:, :, :, :3]).reshape(101, -1).max(1)
    ordinary = [e for e, m in zip(errs, xmax) if m <= 100.0]
    print("teacher-forced T=100 %s %s: worst eps rel-L2 %.3e at call %d (t=%.2f, |x|max %.0f), median %.3e, first %.3e, last %.3e; "
          "%d calls at ordinary scale (|x| <= 100): worst %.3e"
          % (name, mode, errs[worst], worst, float(g["traj_t"][worst].reshape(-1)[0]), xmax[worst], float(np.median(errs)),
             errs[0], errs[-1], len(ordinary), max(ordinary)))
    assert errs[worst] < TOL_T100_WORST[mode]
    assert float(np.median(errs)) < PC.TOL[mode] and max(ordinary) < TOL_T100_ORDINARY[mode]


@pytest.mark.parametrize("mode", ["fp32", "tf32", "fp16", "bf16"])
@pytest.mark.parametrize("name", GOLDENS)
def test_free_running_T100_against_reference(engines, mode, name):
    """Errors compound over 100 steps of a random-weight trajectory (|x| grows to ~1e3), so only the exact-fp32 mode is
    held to a bound on x; the argmax agreement is reported for every mode."""
    g = golden(name)
    xerr, agree, atoms = PC.free_running(engines(mode), g)
    print("free-running T=100 %s %s: final x rel-L2 %.3e, atom-type agreement %.4f on %d atoms" % (name, mode, xerr, agree, atoms))
    if mode == "fp32":
        assert xerr < 2e-2 and agree >= 0.99
    else:
        assert np.isfinite(xerr)


@pytest.mark.parametrize("workload", ["C2", "C3"])
def test_free_running_argmax_agreement_10k_atoms(engines, workload):
    """Atom-type agreement of the tensor-core modes against the exact-fp32 CUDA path on > 10 000 atoms (T = 100, identical
    injected noise): 264 x 39 atoms (config 2's shape) and 384 molecules of 15..39 atoms (a slice of config 3)."""
    rng = np.random.RandomState(3)
    if workload == "C2":
        n_nodes = np.full(264, 39, np.int32)
    else:
        n_nodes = rng.randint(15, 40, 384).astype(np.int32)
    B, N, T = len(n_nodes), 39, 100
    assert int(n_nodes.sum()) > 10000
    tape = O.NoiseTape.draw(1 + T + 1, B, N, 2024).stacked()
    ctx = PC.normed_context([53.6424, 108.3042, 151.4399], B)
    out = {}
    for mode in ("fp32", "tf32", "fp16", "bf16"):
        e = engines(mode)
        e.set_batch(n_nodes, N)
        x, cls = e.sample(ctx, T, "forward", 0, noise_tape=tape)
        out[mode] = (x.cpu(), cls.cpu())
    real = out["fp32"][1] >= 0
    for mode in ("tf32", "fp16", "bf16"):
        agree = float((out[mode][1][real] == out["fp32"][1][real]).float().mean())
        xerr = PC.rel_l2(out[mode][0], out["fp32"][0])
        print("free-running T=100 %s, %s vs exact-fp32 CUDA: atom-type agreement %.5f on %d atoms (%d differ), final x rel-L2 %.3e"
              % (workload, mode, agree, int(real.sum()), int((out[mode][1][real] != out["fp32"][1][real]).sum()), xerr))
        # north-star bar (99.9 %) for the 10-bit-mantissa modes; bf16 (8-bit mantissa) is stated separately: >= 98 %
        assert agree >= (0.98 if mode == "bf16" else 0.999)


def test_bond_orders_strict_lower_triangle(engines):
    """Bond-order agreement where the reference consumes it: on generated samples (384 molecules of 15..39 atoms, T = 20),
    AdjMatSeer of the tensor-core engine (kind::tf32) against the exact-fp32 CUDA AdjMatSeer on identical inputs, and the
    fp32 CUDA AdjMatSeer against the reference's own logits on the golden inputs."""
    g = golden("seer")
    for mode in ("fp32", "tf32"):
        logits, bonds = engines(mode).seer_forward(torch.from_numpy(g["elements"]).int(), torch.from_numpy(g["dist_mat"]),
                                                   torch.from_numpy(g["adj_mat"]))
        r = PC.bond_agreement(bonds, g["bonds"], g["sizes"], g["logits"])
        print("bond orders, golden inputs, %s vs reference:" % mode, r, "logit abs err max %.2e"
              % float((logits.cpu() - torch.from_numpy(g["logits"])).abs().max()))
        if mode == "fp32":
            assert r["flips"] == 0
    rng = np.random.RandomState(4)
    n_nodes = rng.randint(15, 40, 384).astype(np.int32)
    B, N = len(n_nodes), 39
    e = engines("fp16")
    e.set_batch(n_nodes, N)
    x, cls = e.sample(PC.normed_context([53.6424, 108.3042, 151.4399], B), 20, "forward", 0, seed=11)
    el, dist, adj = e.seer_inputs(x, cls)
    lo_tc, b_tc = e.seer_forward(el, dist, adj)
    lo_ref, b_ref = engines("fp32").seer_forward(el, dist, adj)
    r = PC.bond_agreement(b_tc, b_ref, n_nodes, lo_ref)
    err = float((lo_tc - lo_ref).abs().max())
    print("bond orders, 384 generated molecules, tensor-core vs exact-fp32 CUDA:", r, "logit abs err max %.2e" % err)
    # a flip is only an error when the reference margin exceeds the logit error of the tensor-core GEMMs
    assert r["max_ref_margin_at_flips"] <= 4 * err
    assert r["real_pairs"] >= 0.999 and r["lower_triangle"] >= 0.999


def test_inpaint_config4_shape_range(engines):
    """Config 4's shape (fragment inpainting: 21..25 atoms, 8 fixed fragment atoms, resample 1) on 256 molecules.  With
    RANDOM-INIT weights the inpaint trajectory diverges at once -- |z| reaches 5e3 after 2 denoiser calls and 7e5 after 8
    (exact-fp32 CUDA path; at T = 100 it is NaN from call 62 on, as the reference's own fp32 arithmetic would be) -- so
    nothing can be asserted about a T = 100 inpaint run in any precision; parity of the inpaint loop is pinned on the short
    reference goldens (test_free_running_sampler_fp32).  What is asserted here: for T = 4 (9 calls) the modes with fp32's
    exponent range stay finite and agree with the exact path, and the fp16 mode, whose pre-activations leave the fp16 range
    once pairwise distances exceed ~1.6e4, does NOT pass silently: the engine flags the non-finite result."""
    import warnings
    rng = np.random.RandomState(4)
    B, N, T, n_ff = 256, 25, 4, 8
    n_nodes = rng.randint(21, 26, B).astype(np.int32)
    g = torch.Generator().manual_seed(77)
    frag = torch.randn(n_ff, 3, generator=g) * 1.5
    zk = torch.zeros(B, N, 11)
    zk[:, :n_ff, :3] = frag
    for k, c in enumerate([6, 6, 0, 0, 0, 0, 0, 0]):
        zk[:, k, 3 + c] = 1.0
    fm = torch.zeros(B, N)
    fm[:, :n_ff] = 1.0
    tape = O.NoiseTape.draw(1 + T * 3 + 1, B, N, 404).stacked()
    ctx = PC.normed_context([89.8693, 210.7831, 217.7827], B)
    out = {}
    for mode in ("fp32", "tf32", "bf16", "fp16"):
        e = engines(mode)
        e.set_batch(n_nodes, N)
        x, cls = e.sample(ctx, T, "inpaint", 1, z_known=zk, fixed_mask=fm, blend_power=3, noise_tape=tape)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            finite = e.check_finite("inpaint")
        out[mode] = (x.cpu(), cls.cpu(), finite, len(w))
    assert out["fp32"][2] and bool(torch.isfinite(out["fp32"][0]).all())
    print("inpaint T=4 (C4 shape): |x|max in the exact-fp32 path %.3g" % float(out["fp32"][0].abs().max()))
    real = out["fp32"][1] >= 0
    for mode in ("tf32", "bf16"):
        assert out[mode][2] and bool(torch.isfinite(out[mode][0]).all())
        diff = int((out[mode][1][real] != out["fp32"][1][real]).sum())
        print("inpaint T=4 (C4 shape), %s vs exact-fp32 CUDA: atom-type agreement %.5f on %d atoms, final x rel-L2 %.3e"
              % (mode, 1 - diff / int(real.sum()), int(real.sum()), PC.rel_l2(out[mode][0], out["fp32"][0])))
        # at |x| ~ 2e4 the class channels are a numerical accident; only the coordinates are held to a bound
        assert PC.rel_l2(out[mode][0], out["fp32"][0]) < (1e-2 if mode == "tf32" else 1e-1)
    # fp16: either still finite (then it must agree like tf32) or flagged -- never silently wrong
    if out["fp16"][2]:
        assert PC.rel_l2(out["fp16"][0], out["fp32"][0]) < 1e-2
    else:
        assert out["fp16"][3] == 1  # one RuntimeWarning
        print("inpaint T=4 (C4 shape), fp16: trajectory left the fp16 range; flagged by mlcg_nonfinite (RuntimeWarning)")
