"""GPU results against the COMMITTED golden fixtures (tests/golden/kat.json) — no oracle arithmetic in the loop
(the oracle only regenerates the seeded keys, whose hashes are themselves pinned by the fixture) — and the
shard-equality property of SURVEY.md §8e on one GPU with logical shards."""
import hashlib
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def sha(t):
    return hashlib.sha256(t.cpu().numpy().view(np.uint64).tobytes()).hexdigest()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(HERE, "golden", "kat.json")) as f:
        return json.load(f)


def test_golden_transforms(kat):
    import deeppowers_b200 as dp
    for case in kat["cases"]:
        if "ntt_fwd" not in case["name"]:
            continue
        c = dp.Context(case["log_n"], case["L"])
        assert [str(q) for q in c.moduli] == kat["params"]["%d,%d" % (case["log_n"], case["L"])]["moduli"]
        assert [str(p) for p in c.psi] == kat["params"]["%d,%d" % (case["log_n"], case["L"])]["psi"]
        d = torch.empty((case["n_polys"], case["L"], 1 << case["log_n"]), dtype=torch.int64, device="cuda")
        c.fill_uniform(case["seed"], d, case["n_polys"])
        assert sha(d) == case["in_sha256"]                       # the GPU generator reproduces the fixture's input
        c.ntt_fwd(d, case["n_polys"])
        assert sha(d) == case["out_sha256"]
        assert [str(v) for v in d.cpu().numpy().view(np.uint64).reshape(-1)[:8]] == case["out_head"]
        c.ntt_inv(d, case["n_polys"])
        assert sha(d) == case["in_sha256"]
        c.close()


def test_golden_ct_mul_relin_and_rotate(kat, oracle_mod):
    import deeppowers_b200 as dp
    mul = next(c for c in kat["cases"] if c["name"] == "cfg2_ct_mul_relin_n8192_l4")
    rot = next(c for c in kat["cases"] if c["name"] == "rotate1_n8192_l4")
    o = oracle_mod.Oracle(13, 4)
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    gk = o.keygen_galois(3, 65537, s, rot["galois"])
    assert hashlib.sha256(evk.tobytes()).hexdigest() == mul["evk_sha256"]
    assert hashlib.sha256(gk.tobytes()).hexdigest() == rot["gk_sha256"]
    c = dp.Context(13, 4)
    a = torch.empty((2, 2, 4, 8192), dtype=torch.int64, device="cuda")
    b = torch.empty_like(a)
    c.fill_uniform(mul["seed"], a, 4, first_poly=0)
    c.fill_uniform(mul["seed"], b, 4, first_poly=4)
    out = torch.zeros_like(a)
    c.ct_mul_relin(a, b, dev(evk), out, 2)
    assert sha(out) == mul["out_sha256"]
    c.rotate(a, rot["galois"], dev(gk), out, 2)
    assert sha(out) == rot["out_sha256"]
    c.close()


def test_golden_mod_switch_and_hybrid(kat, oracle_mod):
    """the level-dropping and special-prime paths against the committed hashes (keys regenerated, their hashes pinned)"""
    import deeppowers_b200 as dp
    ms = next(c for c in kat["cases"] if c["name"] == "mod_switch_down_n8192_l4")
    mul = next(c for c in kat["cases"] if c["name"] == "ct_mul_relin_hybrid_n8192_l4p1")
    rot = next(c for c in kat["cases"] if c["name"] == "rotate1_hybrid_n8192_l4p1")
    c4, c5 = dp.Context(13, 4), dp.Context(13, 5)
    assert [str(q) for q in c5.moduli] == kat["params"]["13,5"]["moduli"] and c5.moduli[:4] == c4.moduli
    a = torch.empty((2, 2, 4, 8192), dtype=torch.int64, device="cuda")
    b = torch.empty_like(a)
    c4.fill_uniform(ms["seed"], a, 4, first_poly=0)
    c4.fill_uniform(ms["seed"], b, 4, first_poly=4)
    low = torch.zeros((4, 3, 8192), dtype=torch.int64, device="cuda")
    c4.mod_switch_down(a, low, 4, ms["t"])
    assert sha(low) == ms["out_sha256"]
    o5 = oracle_mod.Oracle(13, 5)
    s5 = o5.keygen_secret(1)
    evk = o5.keygen_relin_hybrid(2, mul["t"], s5)
    gk = o5.keygen_galois_hybrid(3, rot["t"], s5, rot["galois"])
    assert hashlib.sha256(evk.tobytes()).hexdigest() == mul["evk_sha256"]
    assert hashlib.sha256(gk.tobytes()).hexdigest() == rot["gk_sha256"]
    out = torch.zeros_like(a)
    c5.ct_mul_relin_hybrid(a, b, dev(evk), out, 2, mul["t"])
    assert sha(out) == mul["out_sha256"]
    c5.rotate_hybrid(a, rot["galois"], dev(gk), out, 2, rot["t"])
    assert sha(out) == rot["out_sha256"]
    c4.close()
    c5.close()


def test_shard_equality_on_one_gpu():
    """G-way sharded evaluation == 1-way evaluation, byte for byte (logical shards on one device)"""
    import deeppowers_b200 as dp
    from deeppowers_b200.sharding import shard_range
    c = dp.Context(13, 4)
    B = 37
    a = torch.empty((B, 2, 4, 8192), dtype=torch.int64, device="cuda")
    b = torch.empty_like(a)
    evk = torch.empty((4, 2, 4, 8192), dtype=torch.int64, device="cuda")
    c.fill_uniform(1, a, 2 * B)
    c.fill_uniform(2, b, 2 * B)
    c.fill_uniform(3, evk, 8)
    whole = torch.zeros_like(a)
    c.ct_mul_relin(a, b, evk, whole, B)
    for world in (2, 4, 8):
        parts = torch.zeros_like(a)
        for r in range(world):
            lo, hi = shard_range(B, r, world)
            if hi > lo:
                c.ct_mul_relin(a[lo:hi], b[lo:hi], evk, parts[lo:hi], hi - lo)
        assert torch.equal(parts, whole)
    c.close()
