"""Writes a synthetic 17-heavy-atom V2000 file (a jittered chain; not a real molecule) for trying the example offline."""
import sys
import torch
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from ml_conformer_generator_b200.mol_utils import samples_to_sdf_blocks
g = torch.Generator().manual_seed(0)
pts = [torch.zeros(3)]
while len(pts) < 17:
    step = torch.randn(3, generator=g)
    cand = pts[-1] + 1.5 * step / step.norm()
    if min(float((cand - p).norm()) for p in pts) > 1.3:
        pts.append(cand)
x = torch.stack(pts).unsqueeze(0)
cls = torch.zeros(1, 17, dtype=torch.int32)
bonds = torch.zeros(1, 42, 42, dtype=torch.int32)
for i in range(1, 17):
    bonds[0, i, i - 1] = 1
open(sys.argv[1], "w").write(samples_to_sdf_blocks(x, cls, bonds, torch.tensor([17]), names=["probe"])[0].replace("$$$$\n", ""))
