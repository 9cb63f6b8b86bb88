"""Regenerates tests/golden/configs.json: one SHA-256 per BASELINE.json configuration at its FULL batch size.

SURVEY.md §8c item 6 asked for a hash per BASELINE-config batch so that a GPU run compares one digest per batch; the
reference holds no vectors for this path, so these are self-generated from the CPU oracle (parity stays "unpinned", see
oracle/dpfhe_oracle.c).  Inputs are synthetic uniform residues from the counter-based generator, which the
GPU reproduces bit for bit (dpfhe_fill_uniform) — the tests hash the regenerated inputs too — so no oracle arithmetic
runs on the GPU box.  About ten minutes on eight cores and 20 GiB of RAM:
    python tests/golden/make_config_golden.py
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import Oracle  # noqa: E402

SEED2, SEED3, SEED4 = 0xD3390002, 0xD3390003, 0xD3390004


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


def cfg3_galois(o):
    """the 26 rotation indices +-1, +-2, ..., +-2^12 of the config-3 sweep"""
    return [o.galois_elt(sgn * (1 << j)) for j in range(13) for sgn in (1, -1)]


def main(only=None):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs.json")
    try:
        out = json.load(open(path))
    except Exception:
        out = {}
    if only in (None, "2"):
        t0 = time.time()
        L, B = 4, 4096
        o = Oracle(13, L)
        a = o.fill_uniform(SEED2, 2 * B).reshape(B, 2, L, o.N)
        b = o.fill_uniform(SEED2 + 1, 2 * B).reshape(B, 2, L, o.N)
        evk = o.fill_uniform(SEED2 + 2, 2 * L).reshape(L, 2, L, o.N)
        r = o.ct_mul_relin(a, b, evk)
        out["config2"] = {"op": "ct_mul_relin", "log_n": 13, "L": L, "batch": B, "seeds": {"a": SEED2, "b": SEED2 + 1, "evk": SEED2 + 2},
                          "in_sha256": {"a": h(a), "b": h(b), "evk": h(evk)}, "out_sha256": h(r)}
        print("config 2 done in %.0f s" % (time.time() - t0), flush=True)
        del a, b, r
    if only in (None, "3"):
        t0 = time.time()
        L, B = 8, 1024
        o = Oracle(14, L)
        ct = o.fill_uniform(SEED3, 2 * B).reshape(B, 2, L, o.N)
        gs = cfg3_galois(o)
        digests = []
        for r, g in enumerate(gs):
            gk = o.fill_uniform(SEED3 + 100 + r, 2 * L).reshape(L, 2, L, o.N)
            digests.append(h(o.rotate(ct, g, gk)))
            print("  config 3 rotation %d/%d (%.0f s)" % (r + 1, len(gs), time.time() - t0), flush=True)
        out["config3"] = {"op": "rotate sweep (hoisted)", "log_n": 14, "L": L, "batch": B, "galois": [int(g) for g in gs],
                          "seeds": {"ct": SEED3, "gk_r": "SEED3 + 100 + r"}, "in_sha256": {"ct": h(ct)},
                          "out_sha256_per_rotation": digests, "out_sha256": hashlib.sha256("".join(digests).encode()).hexdigest()}
        print("config 3 done in %.0f s" % (time.time() - t0), flush=True)
        del ct
    if only in (None, "4"):
        t0 = time.time()
        L, B, n, baby = 4, 512, 768, 32
        giant = n // baby
        o = Oracle(13, L)
        x = o.fill_uniform(SEED4, 2 * B).reshape(B, 2, L, o.N)
        diags = o.fill_uniform(SEED4 + 1, n).reshape(n, L, o.N)
        steps = np.empty((baby, B, 2, L, o.N), dtype=np.uint64)
        steps[0] = x
        for b in range(1, baby):
            gk = o.fill_uniform(SEED4 + 10 + b, 2 * L).reshape(L, 2, L, o.N)
            steps[b] = o.rotate(x, o.galois_elt(b), gk)
        print("  config 4 baby steps done (%.0f s)" % (time.time() - t0), flush=True)
        inner = o.ct_mul_plain_inner(steps, diags.reshape(giant, baby, L, o.N))
        print("  config 4 inner products done (%.0f s)" % (time.time() - t0), flush=True)
        gkg = o.fill_uniform(SEED4 + 99, 2 * L).reshape(L, 2, L, o.N)
        gb = o.galois_elt(baby)
        acc = inner[giant - 1]
        for g in range(giant - 2, -1, -1):
            acc = o.poly_add(o.rotate(acc, gb, gkg).reshape(2 * B, L, o.N), inner[g].reshape(2 * B, L, o.N)).reshape(B, 2, L, o.N)
        out["config4"] = {"op": "768x768 linear layer, baby-step/giant-step diagonals (32 x 24)", "log_n": 13, "L": L, "batch": B, "diagonals": n,
                          "baby": baby, "seeds": {"x": SEED4, "diags": SEED4 + 1, "gk_baby_b": "SEED4 + 10 + b", "gk_giant": SEED4 + 99},
                          "in_sha256": {"x": h(x), "diags": h(diags)}, "out_sha256": h(acc)}
        print("config 4 done in %.0f s" % (time.time() - t0), flush=True)
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
