"""Developer timing: n rotations of the same batch, independent (dpfhe_rotate) vs hoisted (dpfhe_rotate_hoisted)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp
from quickbench import time_op


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 13
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    R = int(sys.argv[4]) if len(sys.argv) > 4 else 31
    c = dp.Context(log_n, L)
    N = 1 << log_n
    ct = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(1, ct, 2 * B)
    keys = torch.empty((R, L, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(2, keys, R * 2 * L)
    out = torch.empty((R, B, 2, L, N), dtype=torch.int64, device="cuda")
    galois = [c.galois_elt(k + 1) for k in range(R)]
    klist = [keys[r] for r in range(R)]
    print("setup done", flush=True)
    ms_h = time_op(lambda: c.rotate_hoisted(ct, galois, klist, out, B), iters=3, warm=1)
    print("hoisted", ms_h, flush=True)
    ref = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")

    def indep():
        for r in range(R):
            c.rotate(ct, galois[r], klist[r], ref, B)

    ms_i = time_op(indep, iters=2, warm=1)
    same = bool(torch.equal(out[R - 1], ref))
    print(json.dumps({"log_n": log_n, "L": L, "batch": B, "rotations": R, "hoisted_ms": ms_h, "independent_ms": ms_i,
                      "hoisted_rot_per_s": R * B / ms_h * 1e3, "independent_rot_per_s": R * B / ms_i * 1e3, "identical": same}))


if __name__ == "__main__":
    main()
