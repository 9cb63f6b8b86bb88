"""small ct_mul_relin + NTT run for compute-sanitizer racecheck/memcheck"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp
for log_n, L, B in ((13, 4, 3), (12, 2, 2), (14, 2, 1)):
    c = dp.Context(log_n, L)
    N = 1 << log_n
    a = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda"); b = torch.empty_like(a); out = torch.empty_like(a)
    evk = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(1, a, 2 * B); c.fill_uniform(2, b, 2 * B); c.fill_uniform(3, evk, 2 * L)
    c.ntt_fwd(a, 2 * B); c.ntt_inv(a, 2 * B)
    if True:
        c.ct_mul_relin(a, b, evk, out, B)
        c.rotate(a, 5, evk, out, B)
    if L >= 2:   # hybrid variants (last limb = special prime) and modulus switching
        Lq = L - 1
        ha = torch.empty((B, 2, Lq, N), dtype=torch.int64, device="cuda"); hb = torch.empty_like(ha); ho = torch.empty_like(ha)
        cq = dp.Context(log_n, Lq, c.moduli[:Lq])
        cq.fill_uniform(4, ha, 2 * B); cq.fill_uniform(5, hb, 2 * B)
        cq.close()
        hk = torch.empty((Lq, 2, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(6, hk, 2 * Lq)
        c.ct_mul_relin_hybrid(ha, hb, hk, ho, B, 65537)
        c.rotate_hybrid(ha, 5, hk, ho, B, 65537)
        c.mod_switch_down(a, ho, 2 * B, 65537)
    # hoisted rotations (zero digits included: the filtered fallback kernel runs) and the fused plaintext inner products
    a[B - 1, 1] = 0
    keys = torch.empty((2, L, 2, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(7, keys, 4 * L)
    hout = torch.empty((2, B, 2, L, N), dtype=torch.int64, device="cuda")
    c.rotate_hoisted(a, [c.galois_elt(1), c.galois_elt(-2)], [keys[0], keys[1]], hout, B)
    pts = torch.empty((3, 2, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(8, pts, 6)
    iout = torch.empty((3, B, 2, L, N), dtype=torch.int64, device="cuda")
    c.ct_mul_plain_inner(hout, pts, iout, 2, 3, B)
    torch.cuda.synchronize()
    c.close()
# round 2: the generic arithmetic variant on the same basis, single-buffered digit slots, logical multi-device shards with the
# peer-store gather (outputs written once into another context's buffer), the linear-layer object
import numpy as np
for env in ({"DPFHE_FORCE_GENERIC": "1"}, {"DPFHE_KS_SINGLE": "1"}):
    os.environ.update(env)
    c = dp.Context(13, 4)
    N, L, B = 8192, 4, 3
    a = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda"); b = torch.empty_like(a); out = torch.empty_like(a)
    evk = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(1, a, 2 * B); c.fill_uniform(2, b, 2 * B); c.fill_uniform(3, evk, 2 * L)
    c.ntt_fwd(a, 2 * B); c.ntt_inv(a, 2 * B)
    c.ct_mul_relin(a, b, evk, out, B); c.ct_mul_relin(a, b, evk, out, B); c.rotate(a, 5, evk, out, B)
    torch.cuda.synchronize()
    c.close()
    for k in env:
        del os.environ[k]
# grouped hybrid key switching: digits of two limbs, two special primes (ks_grouped_kernel), ragged last digit as well
for L, K in ((6, 2), (5, 2)):
    c = dp.Context(12, L)
    N, B, Lq = 4096, 3, L - K
    dn = c.grouped_digits(K)
    cq = dp.Context(12, Lq, c.moduli[:Lq])
    ga = torch.empty((B, 2, Lq, N), dtype=torch.int64, device="cuda"); gb = torch.empty_like(ga); go = torch.empty_like(ga)
    cq.fill_uniform(11, ga, 2 * B); cq.fill_uniform(12, gb, 2 * B)
    cq.close()
    gk = torch.empty((dn, 2, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(13, gk, 2 * dn)
    c.ct_mul_relin_grouped(K, ga, gb, gk, go, B, 65537); c.ct_mul_relin_grouped(K, ga, gb, gk, go, B, 65537)
    c.rotate_grouped(K, ga, 5, gk, go, B, 0)
    ho = torch.empty((2, B, 2, Lq, N), dtype=torch.int64, device="cuda")
    c.rotate_hoisted_grouped(K, ga, [c.galois_elt(1), c.galois_elt(-2)], [gk, gk], ho, B, 65537)
    md = torch.empty((2 * B, Lq, N), dtype=torch.int64, device="cuda"); full = torch.empty((2 * B, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(14, full, 2 * B); c.mod_down_special(K, full, md, 2 * B, 65537)
    torch.cuda.synchronize()
    c.close()
m = dp.MultiContext(12, 2, devices=[0, 0])
B, L, N = 5, 2, 4096
ha = np.zeros((B, 2, L, N), dtype=np.uint64); hk = np.zeros((L, 2, L, N), dtype=np.uint64); ho = np.zeros_like(ha)
m.ct_mul_relin_host(ha, ha, hk, ho)
sh = [m.shard(B, r) for r in range(2)]
da = [torch.zeros((cnt, 2, L, N), dtype=torch.int64, device="cuda") for _, cnt in sh]
dk = [torch.zeros((L, 2, L, N), dtype=torch.int64, device="cuda") for _ in sh]
root = torch.zeros((B, 2, L, N), dtype=torch.int64, device="cuda")
torch.cuda.synchronize()
m.ct_mul_relin_gather(da, da, dk, root, 1, B)
lay = dp.LinearLayer(m.contexts[0], np.zeros((4, L, N), dtype=np.uint64), 2, np.zeros((1, L, 2, L, N), dtype=np.uint64), hk)
lay.apply_host(ha, ho)
lay.close()
m.close()
print("ok")
