"""GPU tests of the device noise generator (the default path whenever no noise tape is injected) and of the sharding
contract built on it: value-by-value against the host restatement (oracle/philox_oracle.py), moments, independence of
consecutive draws, centre of gravity of the initial latent, and sharded == unsharded for arbitrary id subsets."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from ml_conformer_generator_b200 import _lib
from oracle import philox_oracle as PH

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_noise(e, n_nodes, N, seed, draw, ids=None, offset=0):
    e.set_batch(n_nodes, N)
    B = len(n_nodes)
    z = torch.empty(B, N, 11, device="cuda")
    ids_t = None if ids is None else torch.as_tensor(np.asarray(ids, dtype=np.int64)).cuda()
    n = _lib.Noise(None, seed, draw, offset, None if ids_t is None else ids_t.data_ptr())
    assert e.lib.mlcg_noise_init(e.h, z.data_ptr(), C.byref(n), e._stream()) == 0
    torch.cuda.synchronize()
    return z.cpu()


def test_device_noise_matches_host_restatement(engines):
    e = engines("fp32")
    rng = np.random.RandomState(0)
    n_nodes = rng.randint(1, 40, 64).astype(np.int32)
    ids = rng.permutation(10 ** 6)[:64] + (1 << 33)  # exercises the high id word
    for seed, draw in ((0, 0), (12345678901234567, 7), (2 ** 63 + 5, 101)):
        z = _device_noise(e, n_nodes, 39, seed, draw, ids=ids)
        ref = PH.combined_noise(seed, ids, n_nodes, 39, draw)
        assert float(np.abs(z.numpy() - ref).max()) < 2e-5
    # contiguous shard addressed by sample_offset == the same ids given explicitly
    za = _device_noise(e, n_nodes, 39, 99, 3, offset=1000)
    zb = _device_noise(e, n_nodes, 39, 99, 3, ids=np.arange(1000, 1064))
    assert torch.equal(za, zb)


def test_device_noise_moments_and_independence(engines):
    e = engines("fp32")
    B, N = 4096, 39
    n_nodes = np.full(B, N, np.int32)
    d = [_device_noise(e, n_nodes, N, 2024, k) for k in range(4)]
    h = torch.stack([z[:, :, 3:] for z in d])            # feature channels: raw N(0,1), (4, B, N, 8)
    n = h[0].numel()
    assert abs(float(h.mean())) < 4 / np.sqrt(4 * n) and abs(float(h.var()) - 1.0) < 4 * np.sqrt(2 / (4 * n))
    assert abs(float((h ** 4).mean()) - 3.0) < 0.02       # kurtosis of a normal
    # centre of gravity of the position part is removed per molecule; its variance is (1 - 1/N)
    x = d[0][:, :, :3]
    assert float(x.sum(1).abs().max()) < 1e-4
    assert abs(float(x.var()) - (1 - 1 / N)) < 0.01
    # consecutive draws are independent: no linear, radius / angle or squared correlation between any channel pair
    flat = [z.reshape(-1, 11).double() for z in d]
    lim = 5 / np.sqrt(flat[0].shape[0])
    for a in range(3):
        u, v = flat[a], flat[a + 1]
        for f in (lambda t: t, lambda t: t * t, ):
            fu, fv = f(u) - f(u).mean(0), f(v) - f(v).mean(0)
            corr = (fu.t() @ fv) / (fu.norm(dim=0).view(-1, 1) * fv.norm(dim=0).view(1, -1))
            assert float(corr.abs().max()) < lim, (a, float(corr.abs().max()), lim)
        # the advisor's signature of shared uniforms: Box-Muller radius of one draw vs angle of the next
        rad = (v[:, 3] ** 2 + v[:, 4] ** 2)
        ang = torch.atan2(u[:, 5], u[:, 6])
        c = torch.corrcoef(torch.stack([rad, ang]))[0, 1]
        assert abs(float(c)) < lim
    # no value is shared between draws / atoms (a shifted stream would repeat values)
    vals = torch.cat([z[:64, :, 3:].reshape(-1) for z in d]).numpy()
    assert len(np.unique(vals)) > 0.999 * len(vals)


def test_sharding_by_arbitrary_id_subsets_is_bit_exact(engines):
    e = engines("bf16")
    rng = np.random.RandomState(1)
    n_nodes = rng.randint(15, 40, size=14).astype(np.int32)
    from tools.parity_check import normed_context
    ctx = np.tile(normed_context([53.6424, 108.3042, 151.4399], 1).numpy(), (14, 1))
    full = [t.clone() for t in e.generate_host(n_nodes, 39, ctx, T=3, seed=5)]
    perm = rng.permutation(14)
    parts = [np.sort(perm[:5]), np.sort(perm[5:6]), np.sort(perm[6:])]
    outs = [torch.empty_like(t) for t in full]
    for ids in parts:
        res = e.generate_host(n_nodes[ids], 39, ctx[ids], T=3, seed=5, sample_ids=ids, device_out=True)
        for o, r in zip(outs, res):
            o[torch.as_tensor(ids)] = r.cpu()
    for a, b in zip(full, outs):
        assert torch.equal(a, b)


_TWO_RANK = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from ml_conformer_generator_b200.engine import Engine
from ml_conformer_generator_b200.weights import random_state_dicts
from ml_conformer_generator_b200 import parallel as P
from tools.parity_check import normed_context
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
sd, ssd = random_state_dicts(0)
e = Engine(dev, "bf16"); e.load_edm_state_dict(sd); e.load_seer_state_dict(ssd)
n_nodes = np.random.RandomState(2).randint(15, 40, 41).astype(np.int32)
ctx = np.tile(normed_context([53.6424, 108.3042, 151.4399], 1).numpy(), (41, 1))
stats = {}
x, cls, bonds = P.generate_sharded_engine(e, n_nodes, 39, ctx, T=5, seed=9, stats=stats)
assert stats["calls"] == 1
if rank == 0:
    xs, cs, bs = e.generate_host(n_nodes, 39, ctx, T=5, seed=9)
    ok = torch.equal(x.cpu(), xs) and torch.equal(cls.cpu(), cs) and torch.equal(bonds.cpu(), bs)
    print("SHARDED_EQUALS_SINGLE", ok)
dist.barrier()
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_sharded_generation_equals_single_gpu(tmp_path):
    """The multi-GPU product path (parallel.generate_sharded_engine -> Engine.generate_host per rank, one packed NCCL
    all-gather) on 2 GPUs reproduces the single-GPU result bit for bit."""
    script = tmp_path / "two_rank.py"
    script.write_text(_TWO_RANK)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
                          "127.0.0.1", "--master-port", "29533", str(script), ROOT], capture_output=True, text=True,
                         timeout=900)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "SHARDED_EQUALS_SINGLE True" in out.stdout
