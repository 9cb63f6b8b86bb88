"""Developer timing (not the bench contract): special-prime hybrid key switching and modulus switching, kernel-only.

usage: bench_hybrid.py [log_n] [L_ct] [batch]   -- the context gets L_ct + 1 limbs (last = special prime)
"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp
from quickbench import time_op


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    Lq = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1184
    L = Lq + 1
    N = 1 << log_n
    c = dp.Context(log_n, L)          # hybrid: L_ct limbs + special
    cq = dp.Context(log_n, Lq, c.moduli[:Lq])   # per-limb-digit baseline at the same ciphertext size
    a = torch.empty((B, 2, Lq, N), dtype=torch.int64, device="cuda")
    b = torch.empty_like(a)
    out = torch.empty_like(a)
    cq.fill_uniform(1, a, 2 * B)
    cq.fill_uniform(2, b, 2 * B)
    hk = torch.empty((Lq, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(3, hk, 2 * Lq)
    bk = torch.empty((Lq, 2, Lq, N), dtype=torch.int64, device="cuda")
    cq.fill_uniform(4, bk, 2 * Lq)
    res = {"log_n": log_n, "L_ct": Lq, "batch": B}
    ct_bytes = 2 * Lq * N * 8
    for name, fn, io in (
        ("ct_mul_relin_hybrid", lambda: c.ct_mul_relin_hybrid(a, b, hk, out, B, 65537), 3),
        ("ct_mul_relin_bv", lambda: cq.ct_mul_relin(a, b, bk, out, B), 3),
        ("rotate_hybrid", lambda: c.rotate_hybrid(a, 5, hk, out, B, 65537), 2),
        ("rotate_bv", lambda: cq.rotate(a, 5, bk, out, B), 2),
    ):
        ms = time_op(fn)
        res[name] = {"ms": round(ms, 4), "per_s": round(B / ms * 1e3, 1), "GBps": round(B * io * ct_bytes / ms / 1e6, 1)}
    # modulus switching of the L-limb context: [2B][L][N] -> [2B][L-1][N]
    x = torch.empty((2 * B, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(5, x, 2 * B)
    y = torch.empty((2 * B, Lq, N), dtype=torch.int64, device="cuda")
    ms = time_op(lambda: c.mod_switch_down(x, y, 2 * B, 65537))
    res["mod_switch_down"] = {"ms": round(ms, 4), "ct_per_s": round(B / ms * 1e3, 1), "GBps": round(2 * B * (L + Lq) * N * 8 / ms / 1e6, 1)}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
