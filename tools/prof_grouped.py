"""Launches the grouped special-prime kernels a few times on a small batch (for ncu): ks_grouped_kernel, ks_hoistg_kernel,
rot_apply_grouped_kernel, md_tau_kernel, md_limb_kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp
log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
Lq = int(sys.argv[2]) if len(sys.argv) > 2 else 4
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
B = int(sys.argv[4]) if len(sys.argv) > 4 else 444
L = Lq + K
c = dp.Context(log_n, L)
cq = dp.Context(log_n, Lq, c.moduli[:Lq])
N = 1 << log_n
dn = c.grouped_digits(K)
a = torch.empty((B, 2, Lq, N), dtype=torch.int64, device="cuda")
b = torch.empty_like(a); out = torch.empty_like(a)
key = torch.empty((dn, 2, L, N), dtype=torch.int64, device="cuda")
cq.fill_uniform(1, a, 2 * B); cq.fill_uniform(2, b, 2 * B); c.fill_uniform(3, key, 2 * dn)
rot = torch.empty((1, B, 2, Lq, N), dtype=torch.int64, device="cuda")
for _ in range(2):
    c.ct_mul_relin_grouped(K, a, b, key, out, B, 65537)
    c.rotate_hoisted_grouped(K, a, [c.galois_elt(1)], [key], rot, B, 65537)
torch.cuda.synchronize()
print("done")
