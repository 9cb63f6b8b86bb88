"""Launches the hybrid key-switching kernel and the two modulus-switching kernels a few times on a small batch (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
Lq = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B = int(sys.argv[3]) if len(sys.argv) > 3 else 440
L = Lq + 1
c = dp.Context(log_n, L)
cq = dp.Context(log_n, Lq, c.moduli[:Lq])
N = 1 << log_n
a = torch.empty((B, 2, Lq, N), dtype=torch.int64, device="cuda")
b = torch.empty_like(a); out = torch.empty_like(a)
hk = torch.empty((Lq, 2, L, N), dtype=torch.int64, device="cuda")
cq.fill_uniform(1, a, 2 * B); cq.fill_uniform(2, b, 2 * B); c.fill_uniform(3, hk, 2 * Lq)
x = torch.empty((2 * B, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(5, x, 2 * B)
y = torch.empty((2 * B, Lq, N), dtype=torch.int64, device="cuda")
for _ in range(3):
    c.ct_mul_relin_hybrid(a, b, hk, out, B, 65537)
    c.mod_switch_down(x, y, 2 * B, 65537)
torch.cuda.synchronize()
print("done")
