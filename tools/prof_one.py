"""Launches each hot kernel a few times on a small batch (for ncu captures)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
B = int(sys.argv[3]) if len(sys.argv) > 3 else 296
c = dp.Context(log_n, L)
N = 1 << log_n
a = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
b = torch.empty_like(a); out = torch.empty_like(a)
evk = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
c.fill_uniform(1, a, 2 * B); c.fill_uniform(2, b, 2 * B); c.fill_uniform(3, evk, 2 * L)
for _ in range(3):
    c.ntt_fwd(a, 2 * B)
    c.ntt_inv(a, 2 * B)
    c.ct_mul_relin(a, b, evk, out, B)
torch.cuda.synchronize()
print("done")
