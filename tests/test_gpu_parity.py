"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference's golden vectors.

Tolerances (relative L2 of eps against the reference's fp32 torch output):
  fp32 SIMT mode  : 2e-5   (same arithmetic, different summation order)
  tf32 tcgen05    : 1e-3   (north-star bar for the fp32/tf32 mode)
  bf16 tcgen05    : 2e-2   (bf16 operands, fp32 accumulate / residual; stated separately as the north star asks)
"""
import numpy as np
import pytest
import torch

from conftest import golden, rel_l2
from ml_conformer_generator_b200.config import CONTEXT_NORMS
from oracle import edm_oracle as O

pytestmark = pytest.mark.gpu
TOL = {"fp32": 2e-5, "tf32": 1e-3, "fp16": 1e-3, "bf16": 2e-2}


def _norm_ctx(raw, B):
    c = O.normalise_context(torch.tensor(np.asarray(raw), dtype=torch.float32), CONTEXT_NORMS)
    return c.view(1, 3).repeat(B, 1)


@pytest.mark.parametrize("mode", ["tf32", "bf16", "fp16"])
@pytest.mark.parametrize("shape", [(300, 896, 448, 448), (129, 448, 896, 448), (260, 512, 64, 256), (128, 256, 2048, 256)])
def test_tc_gemm(engines, mode, shape):
    M, N, K, bn = shape
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    c = engines(mode).test_gemm(mode, bn, a, w, b).cpu()
    ref = a.double() @ w.double().t() + b.double()
    assert rel_l2(c, ref) < (1e-2 if mode == "bf16" else 2e-3)


@pytest.mark.parametrize("mode", ["fp32", "tf32", "fp16", "bf16"])
@pytest.mark.parametrize("name", ["egnn_small", "egnn_n39"])
def test_egnn_forward_golden(engines, mode, name):
    g = golden(name)
    e = engines(mode)
    B = len(g["n_nodes"])
    e.set_batch(g["n_nodes"], int(g["n_max"]))
    eps = e.egnn_forward(torch.from_numpy(g["t"]).view(-1), torch.from_numpy(g["xh"]), _norm_ctx(g["raw_context"], B)).cpu()
    err = rel_l2(eps, g["eps"])
    print(name, mode, "eps rel-L2", err)
    assert err < TOL[mode]
    nm, _ = O.prepare_masks(torch.from_numpy(g["n_nodes"]), int(g["n_max"]))
    assert float(eps[nm.squeeze(-1) == 0].abs().max()) == 0.0  # padded rows exactly zero (bit-exact mask)
    # x-part is centre-of-gravity free
    assert float((eps[:, :, :3] * nm).sum(1).abs().max()) < 1e-3 * float(eps[:, :, :3].abs().max())


@pytest.mark.parametrize("mode", ["fp32", "tf32", "fp16", "bf16"])
def test_teacher_forced_trajectory(engines, mode):
    """Per-step eps on the reference's own z_t trajectory (11 denoiser calls of a T=10 run)."""
    g = golden("edm_forward_T10")
    e = engines(mode)
    B = len(g["n_nodes"])
    e.set_batch(g["n_nodes"], int(g["n_max"]))
    ctx = _norm_ctx(g["raw_context"], B)
    worst = 0.0
    for k in range(g["traj_z"].shape[0]):
        eps = e.egnn_forward(torch.from_numpy(g["traj_t"][k]).view(-1), torch.from_numpy(g["traj_z"][k]), ctx)
        worst = max(worst, rel_l2(eps, g["traj_eps"][k]))
    print("teacher-forced worst eps rel-L2", mode, worst)
    assert worst < TOL[mode]


_DIST_FP32_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from conftest import golden, rel_l2
from ml_conformer_generator_b200.config import CONTEXT_NORMS
from ml_conformer_generator_b200.engine import Engine
from ml_conformer_generator_b200.weights import random_state_dicts
from oracle import edm_oracle as O
edm_sd, _ = random_state_dicts(0)
e = Engine(torch.device("cuda:0"), "bf16"); e.load_edm_state_dict(edm_sd)
g = golden("edm_forward_T10"); B = len(g["n_nodes"])
e.set_batch(g["n_nodes"], int(g["n_max"]))
c = O.normalise_context(torch.tensor(np.asarray(g["raw_context"]), dtype=torch.float32), CONTEXT_NORMS).view(1, 3).repeat(B, 1)
worst = 0.0
for k in range(g["traj_z"].shape[0]):
    eps = e.egnn_forward(torch.from_numpy(g["traj_t"][k]).view(-1), torch.from_numpy(g["traj_z"][k]), c)
    worst = max(worst, rel_l2(eps, g["traj_eps"][k]))
print("WORST", worst)
"""


def test_teacher_forced_bf16_fp32_distance_terms():
    """MLCG_EDGE_DIST_FP32=1 (read once per process, hence the subprocess): bf16 mode with the first-layer distance terms
    in fp32 is tighter on the random-weight trajectory, whose late steps are dominated by d2 * wc."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, MLCG_EDGE_DIST_FP32="1")
    out = subprocess.run([sys.executable, "-c", _DIST_FP32_SCRIPT, root], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    worst = float([ln for ln in out.stdout.splitlines() if ln.startswith("WORST")][-1].split()[1])
    print("teacher-forced worst eps rel-L2 bf16 (fp32 distance terms)", worst)
    assert worst < 1e-2


_VARIANT_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
from conftest import rel_l2
from ml_conformer_generator_b200.engine import Engine
from ml_conformer_generator_b200.weights import random_state_dicts
from oracle import edm_oracle as O
edm_sd, _ = random_state_dicts(0)
g = torch.Generator().manual_seed(5)
n_nodes = torch.cat([torch.arange(1, 40), torch.tensor([39] * 9)])
B, N = n_nodes.numel(), 39
nm, _ = O.prepare_masks(n_nodes, N)
z = torch.randn(B, N, 11, generator=g) * nm
z[:, :, :3] *= 1.5
ctx = torch.zeros(B, 3)
t = torch.full((B,), 0.3)
out = {}
for m in ("fp32", "tf32", "bf16"):
    e = Engine(torch.device("cuda:0"), m); e.load_edm_state_dict(edm_sd)
    e.set_batch(n_nodes.numpy(), N)
    out[m] = e.egnn_forward(t, z, ctx).cpu()
print("ERR", rel_l2(out["tf32"], out["fp32"]), rel_l2(out["bf16"], out["fp32"]))
"""


@pytest.mark.parametrize("env", [{"MLCG_EDGE_PAIR": "0"}, {"MLCG_EDGE_SPLIT": "0"}, {"MLCG_EDGE_GRID": "37"},
                                 {"MLCG_EDGE_PAIR": "0", "MLCG_EDGE_SPLIT": "0", "MLCG_EDGE_GRID": "5"},
                                 {"MLCG_EDGE_V3": "0"}, {"MLCG_EDGE_V3": "0", "MLCG_EDGE_GRID": "36"},
                                 {"MLCG_EDGE_V3": "1", "MLCG_EDGE_GRID": "6"}, {"MLCG_EDGE_V3": "1", "MLCG_EDGE_SPLIT": "0"}])
def test_kernel_variants_selected_by_environment(env):
    """The documented runtime switches select other variants of the same CUDA path: single-CTA edge kernel,
    whole-target tiles, smaller grids (different CTA tile ranges, hence different carried / side-buffer split targets), the
    single-accumulator pair kernel k_tc_edge (MLCG_EDGE_V3=0) instead of the default k_tc_edge3.
    Each must match the exact fp32 CUDA path on a batch of every size 1..39 (read once per process => subprocess)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT, root], env=dict(os.environ, **env), capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    tf32_err, bf16_err = (float(v) for v in [ln for ln in out.stdout.splitlines() if ln.startswith("ERR")][-1].split()[1:])
    print(env, "tf32 vs fp32", tf32_err, "bf16 vs fp32", bf16_err)
    assert tf32_err < 1e-3 and bf16_err < 2e-2


def _tape(g, B):
    return O.NoiseTape.draw(int(g["n_pairs"]), B, int(g["n_max"]), int(g["seed"])).stacked()


@pytest.mark.parametrize("name,mode", [("edm_forward_T10", "forward"), ("edm_forward_T6_r1", "forward"),
                                       ("edm_inpaint_T6", "inpaint"), ("edm_merge_T10_L4", "merge")])
def test_free_running_sampler_fp32(engines, name, mode):
    """Whole reverse loop with the reference's injected noise, exact-fp32 mode: final x and atom types."""
    g = golden(name)
    e = engines("fp32")
    B = len(g["n_nodes"])
    e.set_batch(g["n_nodes"], int(g["n_max"]))
    kw = {}
    if mode != "forward":
        kw = dict(z_known=torch.from_numpy(g["z_known"]), fixed_mask=torch.from_numpy(g["fixed_mask"]),
                  blend_power=int(g["blend_power"]))
    if mode == "merge":
        kw["diffusion_level"] = int(g["diffusion_level"])
    x, cls, (tz, te) = e.sample(_norm_ctx(g["raw_context"], B), int(g["T"]), mode, int(g["resample_steps"]),
                                noise_tape=_tape(g, B), trace=True, **kw)
    assert tz.shape[0] == g["traj_z"].shape[0]  # same number of denoiser calls as the reference
    err_first = rel_l2(tz[0], g["traj_z"][0])
    assert err_first < 1e-6  # initial latent from the injected noise
    err = rel_l2(x.cpu(), g["x"])
    print(name, "free-running x rel-L2", err)
    assert err < 5e-3
    ref_cls = torch.from_numpy(g["h"]).argmax(-1)
    real = torch.from_numpy(g["h"]).sum(-1) > 0
    assert bool((cls.cpu().long()[real] == ref_cls[real]).all())
    assert bool((cls.cpu()[~real] == -1).all())


def test_step_kernels_against_oracle(engines):
    """noise_init / step / reinject / decode with injected noise, against the oracle's arithmetic."""
    import ctypes as C
    from ml_conformer_generator_b200 import _lib
    from ml_conformer_generator_b200.schedule import decode_scalars, gamma_table, step_scalars
    e = engines("fp32")
    n_nodes = torch.tensor([15, 39, 23, 33, 1, 2])
    N, B = 39, 6
    e.set_batch(n_nodes.numpy(), N)
    nm, em = O.prepare_masks(n_nodes, N)
    tape = O.NoiseTape.draw(4, B, N, 5)
    raw = tape.stacked().cuda()
    gam = gamma_table(100)
    lib, st = e.lib, e._stream()

    def noise(k):
        return _lib.Noise(raw[k].data_ptr(), 0, k, 0)
    z = torch.empty(B, N, 11, device="cuda")
    assert lib.mlcg_noise_init(e.h, z.data_ptr(), C.byref(noise(0)), st) == 0
    z_ref = O.combined_noise(tape, nm)
    assert rel_l2(z.cpu(), z_ref) < 1e-6 and float(z.cpu()[nm.squeeze(-1) == 0].abs().max()) == 0
    eps = (torch.randn(B, N, 11, generator=torch.Generator().manual_seed(3)) * nm)
    sc = step_scalars(gam, 41)
    s = _lib.StepScalars(sc["t"], sc["alpha_ts"], sc["c_eps"], sc["c_sigma"], sc["alpha_s"], sc["sigma_s"], 0.3)
    epsd = eps.cuda()
    assert lib.mlcg_step(e.h, z.data_ptr(), epsd.data_ptr(), C.byref(s), C.byref(noise(1)), st) == 0
    mu = z_ref / sc["alpha_ts"] - sc["c_eps"] * eps
    zs = mu + sc["c_sigma"] * O.combined_noise(tape, nm)
    zs = torch.cat([O.remove_mean_with_mask(zs[:, :, :3], nm), zs[:, :, 3:]], 2)
    assert rel_l2(z.cpu(), zs) < 1e-6
    # reinject
    fm = torch.zeros(B, N, 1)
    fm[:, :1] = 1.0
    fm[1:4, :8] = 1.0
    zk = torch.randn(B, N, 11, generator=torch.Generator().manual_seed(4)) * nm
    zkd, fmd = zk.cuda(), fm.view(B, N).contiguous().cuda()
    assert lib.mlcg_reinject(e.h, z.data_ptr(), zkd.data_ptr(), fmd.data_ptr(), C.byref(s), C.byref(noise(2)), st) == 0
    zkn = sc["alpha_s"] * zk + sc["sigma_s"] * O.combined_noise(tape, nm)
    zkn = O.align_fragment_com(zkn, zs, fm)
    zr = 0.3 * zkn * fm + (1 - 0.3) * zs * fm + zs * (1 - fm)
    assert rel_l2(z.cpu(), zr) < 1e-6
    # decode
    d = decode_scalars(gam)
    x = torch.empty(B, N, 3, device="cuda")
    cls = torch.empty(B, N, dtype=torch.int32, device="cuda")
    assert lib.mlcg_decode(e.h, z.data_ptr(), epsd.data_ptr(), d["sigma_0"], d["alpha_0"], d["sigma_x"],
                           C.byref(noise(3)), x.data_ptr(), cls.data_ptr(), st) == 0
    mu_x = 1.0 / d["alpha_0"] * (zr - d["sigma_0"] * eps)
    xr = (mu_x + d["sigma_x"] * O.combined_noise(tape, nm))[:, :, :3]
    assert rel_l2(x.cpu(), xr) < 1e-6
    ref_cls = torch.argmax(zr[:, :, 3:10] * 9 * nm, dim=2)
    real = nm.squeeze(-1) > 0
    assert bool((cls.cpu().long()[real] == ref_cls[real]).all()) and bool((cls.cpu()[~real] == -1).all())


@pytest.mark.parametrize("mode", ["fp32", "tf32"])
def test_seer_golden(engines, mode):
    g = golden("seer")
    e = engines(mode)
    logits, bonds = e.seer_forward(torch.from_numpy(g["elements"]).int(), torch.from_numpy(g["dist_mat"]),
                                   torch.from_numpy(g["adj_mat"]))
    err = rel_l2(logits.cpu(), g["logits"])
    print("seer logits rel-L2", mode, err)
    assert err < (1e-4 if mode == "fp32" else 2e-3)
    agree = float((bonds.cpu().long() == torch.from_numpy(g["bonds"])).float().mean())
    print("bond-order agreement", mode, agree)
    assert agree >= 0.999


def test_seer_inputs_against_oracle(engines):
    e = engines("fp32")
    n_nodes = torch.tensor([17, 39, 15, 30])
    e.set_batch(n_nodes.numpy(), 39)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(4, 39, 3, generator=g) * 2.0
    cls = torch.randint(0, 7, (4, 39), generator=g)
    el, dist, adj = e.seer_inputs(x, cls.int())
    el_r, dist_r, adj_r = O.seer_inputs_from_samples(x, cls, n_nodes)
    assert torch.equal(el.cpu().long(), el_r)
    assert float((dist.cpu() - dist_r).abs().max()) < 1e-5
    # connectivity may differ only for pairs sitting exactly on the threshold
    assert float((adj.cpu() != adj_r).float().mean()) < 1e-3


@pytest.mark.parametrize("mode", ["tf32", "fp16", "bf16"])
def test_modes_agree_at_full_width(engines, mode):
    """N = 39 molecules (13 edge tiles each) and ragged sizes: tensor-core modes against the exact-fp32 CUDA mode."""
    g = torch.Generator().manual_seed(21)
    n_nodes = torch.tensor([39] * 6 + [15, 16, 22, 27, 31, 38, 2, 1, 5, 9])
    B, N = len(n_nodes), 39
    nm, _ = O.prepare_masks(n_nodes, N)
    z = torch.randn(B, N, 11, generator=g) * nm
    z[:, :, :3] *= 1.5
    ctx = _norm_ctx([53.6424, 108.3042, 151.4399], B)
    t = torch.full((B,), 0.5)
    ref = None
    out = {}
    for m in ("fp32", mode):
        e = engines(m)
        e.set_batch(n_nodes.numpy(), N)
        out[m] = e.egnn_forward(t, z, ctx).cpu()
    err = rel_l2(out[mode], out["fp32"])
    print("full-width", mode, "vs fp32 CUDA:", err)
    assert err < TOL[mode]


def test_generate_host_and_shard_invariance(engines):
    """End-to-end host-buffer API; sharding a batch (sample_offset) reproduces the unsharded result bit-for-bit."""
    e = engines("bf16")
    rng = np.random.RandomState(0)
    n_nodes = rng.randint(15, 40, size=12).astype(np.int32)
    ctx = np.tile(np.asarray(_norm_ctx([53.6424, 108.3042, 151.4399], 1)), (12, 1))
    x, cls, bonds = [t.clone() for t in e.generate_host(n_nodes, 39, ctx, T=4, seed=7)]
    assert torch.isfinite(x).all()
    for b in range(12):
        assert bool((cls[b, : n_nodes[b]] >= 0).all()) and bool((cls[b, n_nodes[b]:] == -1).all())
        assert float(x[b, n_nodes[b]:].abs().max()) == 0.0 if n_nodes[b] < 39 else True
    assert int(bonds.max()) <= 4 and int(bonds.min()) >= 0
    assert bool((torch.triu(bonds.long()) == 0).all())
    xa, ca, ba = [t.clone() for t in e.generate_host(n_nodes[:5], 39, ctx[:5], T=4, seed=7, sample_offset=0)]
    xb, cb, bb = [t.clone() for t in e.generate_host(n_nodes[5:], 39, ctx[5:], T=4, seed=7, sample_offset=5)]
    assert torch.equal(torch.cat([xa, xb]), x) and torch.equal(torch.cat([ca, cb]), cls)
    assert torch.equal(torch.cat([ba, bb]), bonds)


def test_argument_errors(engines):
    e = engines("bf16")
    with pytest.raises(ValueError):
        e.set_batch([40], 40)  # N > 39
    with pytest.raises(ValueError):
        e.set_batch([0, 5], 10)  # empty molecule
    e.set_batch([15, 16], 16)
    with pytest.raises(ValueError):
        e.egnn_forward(torch.zeros(3), torch.zeros(3, 16, 11), torch.zeros(3, 3))


@pytest.mark.parametrize("mode", ["tf32", "fp16", "bf16"])
def test_every_molecule_size_against_fp32_cuda(engines, mode):
    """All sizes 1..39 in one batch (twice, shuffled): exercises whole-target tiles (n < 13), split-target tiles with and
    without a target cut at the tile boundary, single-tile molecules and the n = 1 / n = 2 corner cases, against the exact
    fp32 CUDA path (itself pinned to the reference at 5e-7)."""
    g = torch.Generator().manual_seed(77)
    n_nodes = torch.cat([torch.arange(1, 40), torch.arange(1, 40)])[torch.randperm(78, generator=g)]
    B, N = n_nodes.numel(), 39
    nm, _ = O.prepare_masks(n_nodes, N)
    z = torch.randn(B, N, 11, generator=g) * nm
    z[:, :, :3] *= 1.5
    ctx = _norm_ctx([53.6424, 108.3042, 151.4399], B)
    t = torch.full((B,), 0.3)
    out = {}
    for m in ("fp32", mode):
        e = engines(m)
        e.set_batch(n_nodes.numpy(), N)
        out[m] = e.egnn_forward(t, z, ctx).cpu()
    worst = 0.0
    for b in range(B):
        n = int(n_nodes[b])
        ref = out["fp32"][b, :n]
        if float(ref.abs().max()) == 0.0:   # n = 1: the velocity of a single atom is exactly zero after COM removal
            assert float(out[mode][b, :n, :3].abs().max()) == 0.0
            continue
        worst = max(worst, rel_l2(out[mode][b, :n], ref))
        assert float(out[mode][b, n:].abs().max()) == 0.0 if n < N else True
    print("all sizes 1..39,", mode, "worst per-molecule rel-L2 vs fp32 CUDA:", worst)
    assert worst < (3e-2 if mode == "bf16" else 2e-3)
