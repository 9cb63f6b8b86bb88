"""Developer timing harness (not the bench contract): kernel-only CUDA-event timings."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp


def time_op(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 1184
    c = dp.Context(log_n, L)
    N = 1 << log_n
    a = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    b = torch.empty_like(a)
    out = torch.empty_like(a)
    evk = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(1, a, 2 * B)
    c.fill_uniform(2, b, 2 * B)
    c.fill_uniform(3, evk, 2 * L)
    res = {"log_n": log_n, "L": L, "batch": B}
    ms = time_op(lambda: c.ntt_fwd(a, 2 * B))
    n_ntt = 2 * B * L
    res["ntt_fwd"] = {"ms": ms, "ntt_per_s": n_ntt / ms * 1e3, "GBps": n_ntt * N * 16 / ms / 1e6}
    ms = time_op(lambda: c.ntt_inv(a, 2 * B))
    res["ntt_inv"] = {"ms": ms, "ntt_per_s": n_ntt / ms * 1e3, "GBps": n_ntt * N * 16 / ms / 1e6}
    ms = time_op(lambda: c.ct_mul_plain(a, evk, out, B))
    res["ct_mul_plain"] = {"ms": ms, "per_s": B / ms * 1e3, "GBps": B * 4 * L * N * 8 / ms / 1e6}
    if True:
        ms = time_op(lambda: c.ct_mul_relin(a, b, evk, out, B))
        res["ct_mul_relin"] = {"ms": ms, "per_s": B / ms * 1e3, "GBps": B * 6 * L * N * 8 / ms / 1e6}
        ms = time_op(lambda: c.rotate(a, 5, evk, out, B))
        res["rotate"] = {"ms": ms, "per_s": B / ms * 1e3, "GBps": B * 4 * L * N * 8 / ms / 1e6}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
