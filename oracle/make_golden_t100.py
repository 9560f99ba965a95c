"""TEST INFRASTRUCTURE ONLY.  Full-length (T = 100) golden runs of the REAL reference (imported from /root/reference with
rdkit stubbed, oracle/reference_loader.py) at BASELINE.json's molecule sizes:

    python -m oracle.make_golden_t100

  edm_forward_T100_n39    8 molecules x 39 atoms (config 2's shape), T = 100, plain forward, noise seed 12
  edm_forward_T100_mixed  8 molecules of 15..39 atoms (config 3 / 5's shape), T = 100, noise seed 13

Both store the reference's complete trajectory: the 101 inputs z_t and outputs eps of `EGNNDynamics.forward`
(reference egnn.py:472-513, recorded with a forward hook), the final x / one-hot h of
`EquivariantDiffusion.forward` (reference equivariant_diffusion.py:365-421), and the seed of the global CPU generator
that reproduces the noise tape (`NoiseTape.draw(1 + T + 1, B, N, seed)`, checked here against the oracle port).
Weights: `ml_conformer_generator_b200.weights.random_state_dicts(0)`, loaded with strict=True.
About 5 minutes of CPU time on 8 cores.
"""
import time

import torch

from oracle import edm_oracle as O
from oracle.make_golden import ONNX_CONTEXT, Recorder, build_reference, inputs, save


@torch.no_grad()
def main():
    T = 100
    for tag, sizes, seed in (("edm_forward_T100_n39", [39] * 8, 12),
                             ("edm_forward_T100_mixed", [15, 22, 27, 31, 36, 39, 18, 25], 13)):
        edm, _ = build_reference(T)
        rec = Recorder(edm.dynamics)
        n_max = 39
        n_nodes, nm, em, ctx = inputs(sizes, n_max, ONNX_CONTEXT)
        torch.manual_seed(seed)
        t0 = time.time()
        x, h = edm(nm, em, ctx, 0)
        dt = time.time() - t0
        n_pairs = 1 + T + 1
        tape = O.NoiseTape.draw(n_pairs, len(sizes), n_max, seed)
        xo, ho = O.edm_forward(edm.state_dict(), edm.gamma.gamma, nm, em, ctx, tape, 0)
        print(tag, "reference run %.1f s; port vs reference: x" % dt, (x - xo).abs().max().item(), "h equal",
              torch.equal(h, ho), "|x|max", x.abs().max().item(), flush=True)
        p = rec.pack()
        save(tag, n_nodes=n_nodes, n_max=n_max, T=T, resample_steps=0, seed=seed, n_pairs=n_pairs,
             raw_context=ONNX_CONTEXT, x=x, h=h, traj_z=p["traj_z"], traj_t=p["traj_t"], traj_eps=p["traj_eps"])


if __name__ == "__main__":
    main()
