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
# rotation sweep (26 rotations of one batch, as BASELINE.json config 3 does with per-limb digits): independent rotations against hoisting
sweep = {}
RB = int(os.environ.get("ROT_BATCH", "512"))
for K in (1, 2):
    L = LQ + K
    c = dp.Context(LOG_N, L)
    dn = c.grouped_digits(K)
    cq = dp.Context(LOG_N, LQ, c.moduli[:LQ])
    a = torch.empty((RB, 2, LQ, N), dtype=torch.int64, device="cuda")
    cq.fill_uniform(1, a, 2 * RB)
    cq.close()
    steps = [s for s in range(1, 14)] + [-s for s in range(1, 14)]
    galois = [c.galois_elt(s) for s in steps]
    keys = []
    for r in range(len(steps)):
        k = torch.empty((dn, 2, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(100 + r, k, 2 * dn); keys.append(k)
    out = torch.empty((len(steps), RB, 2, LQ, N), dtype=torch.int64, device="cuda")
    def indep():
        for r, g in enumerate(galois):
            c.rotate_grouped(K, a, g, keys[r], out[r], RB, 65537)
    def hoisted():
        c.rotate_hoisted_grouped(K, a, galois, keys, out, RB, 65537)
    row = {}
    for name, fn in (("independent", indep), ("hoisted", hoisted)):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        row[name] = {"ms_per_sweep": ms, "rotations_per_s": len(steps) * RB / ms * 1e3}
    sweep["K=%d" % K] = row
    c.close()
res["rotation_sweep_26x%d" % RB] = sweep
# BASELINE.json config 4's layer (768 x 768, 32 baby x 24 giant steps) with special-prime keys: hoisted baby steps, fused inner products on
# the ciphertext moduli, 23 giant-step rotations (deeppowers_b200.linear_bsgs_grouped), next to the per-limb-digit composition
LB = int(os.environ.get("LAYER_BATCH", "256"))
layer = {}
for K in (0, 1, 2):
    L = LQ + K
    c = dp.Context(LOG_N, L)
    cq = dp.Context(LOG_N, LQ, c.moduli[:LQ])
    x = torch.empty((LB, 2, LQ, N), dtype=torch.int64, device="cuda"); cq.fill_uniform(1, x, 2 * LB)
    diags = torch.empty((768, LQ, N), dtype=torch.int64, device="cuda"); cq.fill_uniform(2, diags, 768)
    dn = c.grouped_digits(K) if K else LQ
    keys = []
    for r in range(32):
        k = torch.empty((dn, 2, L, N), dtype=torch.int64, device="cuda"); c.fill_uniform(200 + r, k, 2 * dn); keys.append(k)
    out = torch.empty_like(x)
    scratch = torch.empty((32 + 24 + 1, LB, 2, LQ, N), dtype=torch.int64, device="cuda")
    if K:
        fn = lambda: dp.linear_bsgs_grouped(c, cq, K, x, diags, keys[:31], keys[31], 32, out, LB, 65537, scratch)
    else:
        fn = lambda: c.linear_bsgs(x, diags, keys[:31], keys[31], 32, out, LB, scratch)
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    layer["per-limb digits" if K == 0 else "K=%d" % K] = {"ms_per_batch": ms, "prompts_per_s": LB / ms * 1e3}
    del scratch, keys, diags
    c.close(); cq.close()
res["config4_layer_batch%d" % LB] = layer
print(json.dumps({"workload": "ct x ct + relinearise with special primes, N=8192, 4 ciphertext limbs, batch=%d, t=65537" % BATCH, "results": res}))
