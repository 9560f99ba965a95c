#!/usr/bin/env python
"""Headline benchmark: generated molecules / second through the hot path (T=100 reverse steps = 101 EGNN forwards,
then GCN-input build + AdjMatSeer + bond argmax), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|tf32] [--workload C2|C3|C1|C4|C5]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)
  python bench.py --impl reference ...        (reference CPU path = oracle port, timed on the host cores)

A "step" is one pass of the hot path over one batch of synthetic molecules (random-init weights of the named
architecture; the HuggingFace checkpoints are not available offline).  `value` is timed with inputs resident in HBM;
`e2e` goes through the host-buffer API (pinned host -> device -> pinned host inside the timed region)."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ONNX_CONTEXT = [53.6424, 108.3042, 151.4399]  # raw context of configs C2-C5
CEYYAG_CONTEXT = [50.5897, 105.3132, 133.5223]
T_STEPS = 100
MAC_PER_EDGE = 420 * 420 + 420 + 2 * 420  # second edge layer + gate/coord head + rank-2 distance terms
MAC_PER_NODE = 19_061_280                 # per node per forward


def workload(name, seed_shift=0, world=1, rank=0):
    rng = np.random.RandomState(1234 + seed_shift)
    if name == "C5":
        # 65 536 samples of 15-39 atoms in total, sharded contiguously over the ranks (strong scaling); every
        # rank works through its share in sub-batches of 8192.  Sizes are drawn for the GLOBAL sample ids, so the job is the
        # same whatever the number of ranks.
        total, chunk = 65536, 8192
        sizes = np.random.RandomState(5).randint(15, 40, total).astype(np.int32)
        per = total // world
        mine = sizes[rank * per:(rank + 1) * per]
        chunks = [mine[i:i + chunk] for i in range(0, per, chunk)]
        return dict(B=len(chunks[0]), N=39, n_nodes=chunks[0], chunks=chunks, total=total, per_rank=per, ctx=ONNX_CONTEXT,
                    desc="C5: 65536 samples of 15-39 atoms in total over %d GPU(s) (strong scaling, sub-batches of %d), "
                         "T=100 + AdjMatSeer GCN" % (world, len(chunks[0])))
    if name == "C2":
        return dict(B=1024, N=39, n_nodes=np.full(1024, 39, np.int32), ctx=ONNX_CONTEXT,
                    desc="C2: B=1024 samples x 39 atoms, T=100 (101 EGNN forwards) + AdjMatSeer GCN")
    if name == "C3":
        return dict(B=8192, N=39, n_nodes=rng.randint(15, 40, 8192).astype(np.int32), ctx=ONNX_CONTEXT,
                    desc="C3: B=8192 samples, 15-39 atoms, T=100 + AdjMatSeer GCN")
    if name == "C4":
        # simple inpainting around the 8 heavy atoms of frag_yibfeu (Cl, Cl, C x 6), resample_steps = 1
        # => 201 EGNN forwards per sample.  Fragment coordinates are synthetic (seeded), one-hot raw 0/1.
        B, N = 4096, 25
        frag = rng.randn(8, 3).astype(np.float32) * 1.5
        zk = np.zeros((B, N, 11), np.float32)
        zk[:, :8, :3] = frag
        for k, c in enumerate([6, 6, 0, 0, 0, 0, 0, 0]):
            zk[:, k, 3 + c] = 1.0
        fm = np.zeros((B, N), np.float32)
        fm[:, :8] = 1.0
        return dict(B=B, N=N, n_nodes=rng.randint(21, 26, B).astype(np.int32), ctx=[89.8693, 210.7831, 217.7827],
                    mode="inpaint", resample=1, z_known=zk, fixed_mask=fm, n_forwards=201,
                    desc="C4: inpaint, B=4096 samples, 21-25 atoms, 8 fixed fragment atoms, T=100, resample 1 (201 EGNN "
                         "forwards) + AdjMatSeer GCN")
    if name == "C1":
        return dict(B=20, N=19, n_nodes=rng.randint(15, 20, 20).astype(np.int32), ctx=CEYYAG_CONTEXT,
                    desc="C1: B=20 samples, 15-19 atoms (ceyyag), T=100 + AdjMatSeer GCN")
    raise SystemExit("unknown workload " + name)


def alg_flops_forward(n_nodes):
    n = n_nodes.astype(np.float64)
    return float((2 * 27 * MAC_PER_EDGE * n * (n - 1) + 2 * MAC_PER_NODE * n).sum())


def normed_ctx(raw, B):
    from ml_conformer_generator_b200.config import CONTEXT_NORMS
    c = (np.asarray(raw, np.float32) - np.asarray(CONTEXT_NORMS["mean"], np.float32)) / np.asarray(
        CONTEXT_NORMS["mad"], np.float32)
    return np.tile(c.reshape(1, 3), (B, 1)).astype(np.float32)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(burst=d["bf16_tflops"], sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured")
    return dict(burst=1590.0, sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        clk, mx, reasons = [], None, set()
        for r in rows:
            try:
                c, m = float(r[1]), float(r[2])
            except (ValueError, IndexError):
                continue
            mx = m
            clk.append(c)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(name)
        load = [c for c in clk if mx and c > 0.5 * mx] or clk
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": mx, "samples": len(clk),
                "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------------------
# reference CPU arm (oracle port on the host cores)
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(wl, n_mols, n_steps, threads):
    """Times the reference formulation (oracle port) on host cores: n_steps calls of sample_p_zs_given_zt on the first
    n_mols molecules of the workload + one AdjMatSeer pass; extrapolates to 101 forwards per molecule."""
    from ml_conformer_generator_b200.config import CONTEXT_NORMS
    from ml_conformer_generator_b200.weights import random_state_dicts
    from oracle import edm_oracle as O
    torch.set_num_threads(threads)
    sd, ssd = random_state_dicts(0)
    n_nodes = torch.from_numpy(wl["n_nodes"][:n_mols].astype(np.int64))
    N = wl["N"]
    nm, em = O.prepare_masks(n_nodes, N)
    ctx = O.batch_context(O.normalise_context(torch.tensor(wl["ctx"]), CONTEXT_NORMS), nm)
    gamma = O.gamma_table(T_STEPS)
    tape = O.NoiseTape.draw(n_steps + 2, n_mols, N, 99)
    per_step = []
    with torch.no_grad():
        z = O.combined_noise(tape, nm)
        for k in range(n_steps):
            s = T_STEPS - 1 - k
            s_arr, t_arr = O._times(s, T_STEPS, n_mols)
            t0 = time.perf_counter()
            z = O.sample_p_zs_given_zt(sd, gamma, s_arr, t_arr, z, nm, em, ctx, tape)
            per_step.append(time.perf_counter() - t0)
        x = z[:, :, :3]
        cls = torch.argmax(z[:, :, 3:10], dim=2)
        el, dist, adj = O.seer_inputs_from_samples(x, cls, n_nodes)
        t0 = time.perf_counter()
        O.bond_orders(O.seer_forward(ssd, el, dist, adj))
        t_seer = time.perf_counter() - t0
    step = min(per_step)
    total = step * wl.get("n_forwards", T_STEPS + 1) + t_seer  # the denoiser call dominates; re-injection steps are negligible
    return n_mols / total, dict(step_s=step, seer_s=t_seer, n_mols=n_mols, n_steps=n_steps)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = workload(args.workload)
    threads = os.cpu_count() or 1
    n_mols = 8 if wl["N"] > 30 else 16
    vals = []
    for _ in range(args.warmup):
        cpu_reference_rate(wl, n_mols, 1, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, info = cpu_reference_rate(wl, n_mols, 1, threads)
        vals.append(v)
    ms = (time.perf_counter() - t0) / max(args.steps, 1) * 1e3
    value = statistics.median(vals)
    sample = ("oracle port of the reference CPU path (reference formulation, torch fp32): 1 reverse step + AdjMatSeer on "
              "the first %d molecules of the workload per bench step, extrapolated to 101 EGNN forwards / molecule"
              % n_mols)
    print(json.dumps({
        "impl": "reference", "metric": "generated mols/sec (100 EGNN steps + GCN, <=39 atoms)", "value": value,
        "unit": "mols/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "precision": "fp32 (torch CPU)"},
        "cpu_baseline": {"value": value, "unit": "mols/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "mols/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    from ml_conformer_generator_b200.engine import Engine
    from ml_conformer_generator_b200.weights import random_state_dicts

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the CUDA extension has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = workload(args.workload, seed_shift=rank, world=world, rank=rank)
    B, N = wl["B"], wl["N"]
    chunks = wl.get("chunks")  # strong-scaling workload: this rank's share, processed in sub-batches
    eng = Engine(dev, args.precision)
    sd, ssd = random_state_dicts(0)
    eng.load_edm_state_dict(sd)
    eng.load_seer_state_dict(ssd)
    del sd, ssd
    ctx_np = normed_ctx(wl["ctx"], B)
    ctx_dev = torch.from_numpy(ctx_np).to(dev)
    eng.set_batch(wl["n_nodes"], N)
    gather = None
    if world > 1:
        gather = [torch.empty(world * B, N, 3, device=dev), torch.empty(world * B, N, dtype=torch.int32, device=dev),
                  torch.empty(world * B, 42, 42, dtype=torch.int8, device=dev)]

    mode = wl.get("mode", "forward")
    n_forwards = wl.get("n_forwards", T_STEPS + 1)
    zk_dev = torch.from_numpy(wl["z_known"]).to(dev) if mode != "forward" else None
    fm_dev = torch.from_numpy(wl["fixed_mask"]).to(dev) if mode != "forward" else None

    def device_batch(seed, offset):
        x, cls = eng.sample(ctx_dev, T_STEPS, mode, wl.get("resample", 0), z_known=zk_dev, fixed_mask=fm_dev, seed=seed,
                            sample_offset=offset)
        el, dmat, adj = eng.seer_inputs(x, cls)
        _, bonds = eng.seer_forward(el, dmat, adj, want_logits=False)
        if world > 1:  # the single collective of the path: final gather of coordinates / types / bonds
            dist.all_gather_into_tensor(gather[0], x)
            dist.all_gather_into_tensor(gather[1], cls)
            dist.all_gather_into_tensor(gather[2], bonds)
        return x, cls, bonds

    def device_step(seed):
        if chunks is None:
            return device_batch(seed, rank * B)
        for ci, nn in enumerate(chunks):   # the batch plan (edge-tile table) is rebuilt per sub-batch: part of the job
            eng.set_batch(nn, N)
            out = device_batch(seed, rank * wl["per_rank"] + ci * B)
        return out

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(args.warmup):
        device_step(100 + i)
    sync()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = eng.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        device_step(200 + i)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1) / args.steps
    launches = eng.kernel_launches() - l0
    clocks = sampler.stop() if sampler else None

    # end-to-end through the host-buffer API (pinned host in, pinned host out, copies inside the timed region); the
    # host-buffer entry point covers plain generation, so the fragment workload reports the device-timed value only
    e2e_ms = float("nan")
    if mode == "forward":
        parts = chunks if chunks is not None else [wl["n_nodes"]]
        base = rank * (wl["per_rank"] if chunks is not None else B)
        out = eng.generate_host(parts[0], N, ctx_np, T_STEPS, 0, seed=300, sample_offset=base)
        sync()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 2)) if chunks is None else 1
        for i in range(e2e_steps):
            for ci, nn in enumerate(parts):
                eng.generate_host(nn, N, ctx_np, T_STEPS, 0, seed=400 + i, sample_offset=base + ci * B, out=out)
        sync()
        e2e_ms = (time.perf_counter() - t0) / e2e_steps * 1e3

    # dominant kernel: fused edge kernel (GCL variant), timed live with CUDA events on the launching stream
    edge_ms = eng.time_edge_kernel(layer=0, iters=20)
    n_edges = eng.num_edges()

    t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(t[0]), float(t[1])

    if rank == 0:
        pk = peaks()
        edge_flops = 2.0 * MAC_PER_EDGE * n_edges
        achieved = edge_flops / (edge_ms * 1e-3) / 1e12
        mine = np.concatenate(chunks) if chunks is not None else wl["n_nodes"]
        step_flops = alg_flops_forward(mine) * n_forwards + 1.871e9 * len(mine)
        n_job = wl["total"] if chunks is not None else world * B  # molecules per step over all ranks
        prof = os.path.join(ROOT, "profiles", "edge_kernel_traffic.json")
        traffic = json.load(open(prof)).get(args.precision) if os.path.exists(prof) else None
        res = {
            "metric": "generated mols/sec (100 EGNN steps + GCN, <=39 atoms)",
            "value": n_job / (ms * 1e-3), "unit": "mols/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong" if chunks is not None else "weak",
            "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
            "config": {"workload": wl["desc"], "per_gpu_batch": B, "max_atoms": N, "diffusion_steps": T_STEPS,
                       "weights": "random-init (seed 0) of the reference architecture",
                       "noise": "device Philox keyed by global sample id",
                       "l2": "per-step working set (PQ projections %.0f MB + operands) exceeds the 126 MB L2; no flush"
                             % (eng.n_nodes.sum().item() * 896 * 4 / 1e6),
                       "parallelism": "dp%d, no collective inside the loop, one NCCL all-gather of results" % world},
            "e2e": ({"value": n_job / (e2e_ms * 1e-3), "unit": "mols/s",
                     "h2d_bytes_per_step": int(len(mine) * 4 + len(mine) * 3 * 4),
                     "d2h_bytes_per_step": int(len(mine) * (N * 3 * 4 + N * 4 + 42 * 42))} if e2e_ms == e2e_ms else None),
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "k_tc_edge (GCL sub-layer, %s)" % args.precision,
                         "achieved": achieved, "peak": pk["burst"], "unit": "TFLOP/s", "frac": achieved / pk["burst"],
                         "traffic": traffic, "peak_source": pk["source"] + " bf16 dense burst",
                         "launch_ms": edge_ms, "alg_flops_per_launch": edge_flops,
                         "step_achieved": step_flops / (ms * 1e-3) / 1e12,
                         "step_frac_of_sustained": step_flops / (ms * 1e-3) / 1e12 / pk["sustained"]},
        }
        if not args.no_cpu_baseline and world == 1:
            n_mols = 8 if N > 30 else 16
            v, info = cpu_reference_rate(wl, n_mols, 2, os.cpu_count() or 1)
            res["cpu_baseline"] = {
                "value": v, "unit": "mols/s", "cores": os.cpu_count() or 1, "kind": "port",
                "sample": "oracle port (reference formulation, torch fp32): 2 reverse steps (best taken) + AdjMatSeer on "
                          "%d molecules of this workload, %.2f s / step, extrapolated to 101 forwards" % (n_mols, info["step_s"])}
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
