"""Launches the hoisted-rotation kernels a few times on a small batch (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp
log_n, L, B, R = 13, 4, 444, 3
c = dp.Context(log_n, L)
N = 1 << log_n
ct = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(1, ct, 2 * B)
keys = torch.empty((R, L, 2, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(2, keys, R * 2 * L)
out = torch.empty((R, B, 2, L, N), dtype=torch.int64, device="cuda")
for _ in range(2):
    c.rotate_hoisted(ct, [c.galois_elt(k + 1) for k in range(R)], [keys[r] for r in range(R)], out, B)
torch.cuda.synchronize()
print("done")
