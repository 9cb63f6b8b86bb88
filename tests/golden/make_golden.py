"""Regenerates tests/golden/kat.json from the CPU oracle.

The reference tree holds no golden vectors for this path, so these are
self-generated pins: they freeze the spec as implemented by the oracle at the
commit that introduced them, so any later change of convention (prime choice, root choice,
bit-reversed ordering, digit decomposition) is caught on CPU.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import Oracle  # noqa: E402


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def main():
    out = {"params": {}, "cases": []}
    for log_n, L in [(12, 1), (12, 3), (13, 4), (13, 5), (13, 6), (14, 8)]:
        o = Oracle(log_n, L)
        out["params"]["%d,%d" % (log_n, L)] = {"moduli": [str(q) for q in o.moduli], "psi": [str(p) for p in o.psi]}
    # config 1 of BASELINE.json: single forward NTT, N=4096, one 60-bit modulus
    o = Oracle(12, 1)
    x = o.fill_uniform(0xD3390001, 1)
    y = o.ntt_fwd(x)
    out["cases"].append({"name": "cfg1_ntt_fwd_n4096_l1", "log_n": 12, "L": 1, "seed": 0xD3390001, "n_polys": 1,
                         "in_sha256": h(x), "out_sha256": h(y), "in_head": [str(v) for v in x.reshape(-1)[:8]],
                         "out_head": [str(v) for v in y.reshape(-1)[:8]]})
    for log_n, L, seed in [(13, 4, 0xD3390002), (14, 8, 0xD3390003)]:
        o = Oracle(log_n, L)
        x = o.fill_uniform(seed, 2)
        y = o.ntt_fwd(x)
        out["cases"].append({"name": "ntt_fwd_n%d_l%d" % (1 << log_n, L), "log_n": log_n, "L": L, "seed": seed, "n_polys": 2,
                             "in_sha256": h(x), "out_sha256": h(y), "in_head": [str(v) for v in x.reshape(-1)[:8]],
                             "out_head": [str(v) for v in y.reshape(-1)[:8]]})
    # ct x ct + relin and rotate at config-2 shape, keys from the seeded test scheme
    o = Oracle(13, 4)
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    a = o.fill_uniform(0xD3390002, 4).reshape(2, 2, 4, o.N)
    b = o.fill_uniform(0xD3390002, 4, first_poly=4).reshape(2, 2, 4, o.N)
    r = o.ct_mul_relin(a, b, evk)
    g = o.galois_elt(1)
    gk = o.keygen_galois(3, 65537, s, g)
    rot = o.rotate(a, g, gk)
    out["cases"].append({"name": "cfg2_ct_mul_relin_n8192_l4", "log_n": 13, "L": 4, "seed": 0xD3390002,
                         "secret_sha256": h(s), "evk_sha256": h(evk), "out_sha256": h(r),
                         "out_head": [str(v) for v in r.reshape(-1)[:8]]})
    out["cases"].append({"name": "rotate1_n8192_l4", "log_n": 13, "L": 4, "seed": 0xD3390002, "galois": int(g),
                         "gk_sha256": h(gk), "out_sha256": h(rot), "out_head": [str(v) for v in rot.reshape(-1)[:8]]})
    # modulus switching (BGV, t = 65537) of the same four polynomials, and the special-prime hybrid variants:
    # context (13, 5) = the four ciphertext moduli above + the next prime as the special one
    ms = o.mod_switch_down(a.reshape(4, 4, o.N), 65537)
    out["cases"].append({"name": "mod_switch_down_n8192_l4", "log_n": 13, "L": 4, "seed": 0xD3390002, "t": 65537,
                         "out_sha256": h(ms), "out_head": [str(v) for v in ms.reshape(-1)[:8]]})
    o5 = Oracle(13, 5)
    assert o5.moduli[:4] == o.moduli
    s5 = o5.keygen_secret(1)
    hk = o5.keygen_relin_hybrid(2, 65537, s5)
    hr = o5.ct_mul_relin_hybrid(a, b, hk, 65537)
    hgk = o5.keygen_galois_hybrid(3, 65537, s5, g)
    hrot = o5.rotate_hybrid(a, g, hgk, 65537)
    out["cases"].append({"name": "ct_mul_relin_hybrid_n8192_l4p1", "log_n": 13, "L": 5, "seed": 0xD3390002, "t": 65537,
                         "evk_sha256": h(hk), "out_sha256": h(hr), "out_head": [str(v) for v in hr.reshape(-1)[:8]]})
    out["cases"].append({"name": "rotate1_hybrid_n8192_l4p1", "log_n": 13, "L": 5, "seed": 0xD3390002, "t": 65537, "galois": int(g),
                         "gk_sha256": h(hgk), "out_sha256": h(hrot), "out_head": [str(v) for v in hrot.reshape(-1)[:8]]})
    # grouped hybrid key switching: context (13, 6) = the same four ciphertext moduli + two special primes, digits of
    # two limbs; the plain and the hoisted rotation (2.11b) are different bits by definition, both pinned
    o6 = Oracle(13, 6)
    assert o6.moduli[:4] == o.moduli
    s6 = o6.keygen_secret(1)
    gek = o6.keygen_relin_grouped(2, 2, 65537, s6)
    ger = o6.ct_mul_relin_grouped(2, a, b, gek, 65537)
    ggk = o6.keygen_galois_grouped(2, 3, 65537, s6, g)
    grot = o6.rotate_grouped(2, a, g, ggk, 65537)
    ghoist = o6.rotate_hoisted_grouped(2, a, [g], ggk[None], 65537)
    out["cases"].append({"name": "ct_mul_relin_grouped_n8192_l4p2", "log_n": 13, "L": 6, "K": 2, "seed": 0xD3390002, "t": 65537,
                         "evk_sha256": h(gek), "out_sha256": h(ger), "out_head": [str(v) for v in ger.reshape(-1)[:8]]})
    out["cases"].append({"name": "rotate1_grouped_n8192_l4p2", "log_n": 13, "L": 6, "K": 2, "seed": 0xD3390002, "t": 65537, "galois": int(g),
                         "gk_sha256": h(ggk), "out_sha256": h(grot), "hoisted_sha256": h(ghoist),
                         "out_head": [str(v) for v in grot.reshape(-1)[:8]]})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
