"""ct x ct with special primes on one GPU: one special prime (ks_hybrid_kernel) against grouped digits (ks_grouped_kernel) at the
same ciphertext modulus (Lq = 4 limbs, N = 8192), CUDA-event timed, inputs resident in HBM.  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp

LOG_N, LQ = 13, 4
BATCH = int(os.environ.get("BATCH", "4096"))
STEPS = int(os.environ.get("STEPS", "5"))
N = 1 << LOG_N
res = {}
for K in (1, 2, 4):
    L = LQ + K
    c = dp.Context(LOG_N, L)
    dn = c.grouped_digits(K)
    cq = dp.Context(LOG_N, LQ, c.moduli[:LQ])
    a = torch.empty((BATCH, 2, LQ, N), dtype=torch.int64, device="cuda"); b = torch.empty_like(a); out = torch.empty_like(a)
    cq.fill_uniform(1, a, 2 * BATCH); cq.fill_uniform(2, b, 2 * BATCH)
    cq.close()
    key = torch.empty((dn, 2, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(3, key, 2 * dn)
    def run():
        if K == 1 and not os.environ.get("GROUPED_K1"):
            c.ct_mul_relin_hybrid(a, b, key, out, BATCH, 65537)
        else:
            c.ct_mul_relin_grouped(K, a, b, key, out, BATCH, 65537)
    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    res["K=%d" % K] = {"context_limbs": L, "digits": dn, "key_MiB": dn * 2 * L * N * 8 / 2**20, "ms_per_step": ms, "ct_mult_per_s": BATCH / ms * 1e3,
                       "transforms_per_ct": LQ + dn * L - LQ + 2 * K + 2 * LQ}
    c.close()
print(json.dumps({"workload": "ct x ct + relinearise with special primes, N=8192, 4 ciphertext limbs, batch=%d, t=65537" % BATCH, "results": res}))
