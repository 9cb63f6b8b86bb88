"""Launches dpfhe_ct_mul_plain_inner at the config-4 shape on a small batch (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
NB, NG, L, N = 32, 24, 4, 8192
c = dp.Context(13, L)
steps = torch.empty((NB, B, 2, L, N), dtype=torch.int64, device="cuda")
pts = torch.empty((NG, NB, L, N), dtype=torch.int64, device="cuda")
out = torch.empty((NG, B, 2, L, N), dtype=torch.int64, device="cuda")
c.fill_uniform(1, steps, NB * B * 2); c.fill_uniform(2, pts, NG * NB)
for _ in range(3):
    c.ct_mul_plain_inner(steps, pts, out, NB, NG, B)
torch.cuda.synchronize()
print("done")
