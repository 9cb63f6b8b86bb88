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


def test_golden_grouped(kat, oracle_mod):
    """digits of two limbs and two special primes against the committed hashes: ct x ct, rotate, and the hoisted rotate"""
    import deeppowers_b200 as dp
    mul = next(c for c in kat["cases"] if c["name"] == "ct_mul_relin_grouped_n8192_l4p2")
    rot = next(c for c in kat["cases"] if c["name"] == "rotate1_grouped_n8192_l4p2")
    c4, c6 = dp.Context(13, 4), dp.Context(13, 6)
    assert [str(q) for q in c6.moduli] == kat["params"]["13,6"]["moduli"] and c6.moduli[:4] == c4.moduli
    a = torch.empty((2, 2, 4, 8192), dtype=torch.int64, device="cuda")
    b = torch.empty_like(a)
    c4.fill_uniform(mul["seed"], a, 4, first_poly=0)
    c4.fill_uniform(mul["seed"], b, 4, first_poly=4)
    o6 = oracle_mod.Oracle(13, 6)
    s6 = o6.keygen_secret(1)
    evk = o6.keygen_relin_grouped(2, 2, mul["t"], s6)
    gk = o6.keygen_galois_grouped(2, 3, rot["t"], s6, rot["galois"])
    assert hashlib.sha256(evk.tobytes()).hexdigest() == mul["evk_sha256"]
    assert hashlib.sha256(gk.tobytes()).hexdigest() == rot["gk_sha256"]
    out = torch.zeros_like(a)
    c6.ct_mul_relin_grouped(2, a, b, dev(evk), out, 2, mul["t"])
    assert sha(out) == mul["out_sha256"]
    c6.rotate_grouped(2, a, rot["galois"], dev(gk), out, 2, rot["t"])
    assert sha(out) == rot["out_sha256"]
    c6.rotate_hoisted_grouped(2, a, [rot["galois"]], [dev(gk)], out, 2, rot["t"])
    assert sha(out) == rot["hoisted_sha256"]
    c4.close()
    c6.close()


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


# ---- one digest per BASELINE.json configuration at its full batch size (tests/golden/configs.json) ----
@pytest.fixture(scope="module")
def cfg_golden():
    path = os.path.join(HERE, "golden", "configs.json")
    if not os.path.exists(path):
        pytest.skip("tests/golden/configs.json not generated")
    with open(path) as f:
        return json.load(f)


def _sha_parallel(tensors):
    """SHA-256 of several device tensors, downloads and hashes overlapped on a few host threads"""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(8) as ex:
        return list(ex.map(sha, tensors))


def test_config2_full_batch_digest(cfg_golden):
    """ct x ct + relinearise on all 4096 ciphertexts of config 2: one digest, no oracle in the loop"""
    import deeppowers_b200 as dp
    g = cfg_golden["config2"]
    L, B, N = g["L"], g["batch"], 1 << g["log_n"]
    c = dp.Context(g["log_n"], L)
    a = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    b, out = torch.empty_like(a), torch.empty_like(a)
    evk = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(g["seeds"]["a"], a, 2 * B)
    c.fill_uniform(g["seeds"]["b"], b, 2 * B)
    c.fill_uniform(g["seeds"]["evk"], evk, 2 * L)
    assert sha(evk) == g["in_sha256"]["evk"] and sha(a) == g["in_sha256"]["a"]
    c.ct_mul_relin(a, b, evk, out, B)
    torch.cuda.synchronize()
    assert sha(out) == g["out_sha256"]
    # and through the host-buffer pipeline (chunked): the same digest
    ha, hb, hk, ho = (t.cpu().numpy().view(np.uint64) for t in (a, b, evk, torch.zeros_like(out)))
    c.ct_mul_relin_host(ha, hb, hk, ho)
    assert hashlib.sha256(ho.tobytes()).hexdigest() == g["out_sha256"]
    c.close()


def test_config3_hoisted_sweep_digest(cfg_golden):
    """the 26-index rotation sweep of config 3 (N=16384, L=8, 1024 ciphertexts) through dpfhe_rotate_hoisted at full size"""
    import deeppowers_b200 as dp
    g = cfg_golden["config3"]
    L, B, N = g["L"], g["batch"], 1 << g["log_n"]
    c = dp.Context(g["log_n"], L)
    ct = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(0xD3390003, ct, 2 * B)
    assert sha(ct) == g["in_sha256"]["ct"]
    gs = g["galois"]
    keys = []
    for r in range(len(gs)):
        k = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
        c.fill_uniform(0xD3390003 + 100 + r, k, 2 * L)
        keys.append(k)
    out = torch.empty((len(gs), B, 2, L, N), dtype=torch.int64, device="cuda")
    c.rotate_hoisted(ct, gs, keys, out, B)
    torch.cuda.synchronize()
    digests = _sha_parallel([out[r] for r in range(len(gs))])
    assert digests == g["out_sha256_per_rotation"]
    assert hashlib.sha256("".join(digests).encode()).hexdigest() == g["out_sha256"]
    # one index of the sweep through the ordinary rotate kernel as well
    one = torch.empty_like(ct)
    c.rotate(ct, gs[5], keys[5], one, B)
    torch.cuda.synchronize()
    assert sha(one) == g["out_sha256_per_rotation"][5]
    c.close()


def test_config4_linear_layer_digest(cfg_golden):
    """the 768 x 768 layer of config 4 on all 512 prompts: baby steps hoisted, inner products fused, Horner over the giant steps"""
    import deeppowers_b200 as dp
    g = cfg_golden["config4"]
    L, B, N, n, baby = g["L"], g["batch"], 1 << g["log_n"], g["diagonals"], g["baby"]
    c = dp.Context(g["log_n"], L)
    x = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    diags = torch.empty((n, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(0xD3390004, x, 2 * B)
    c.fill_uniform(0xD3390004 + 1, diags, n)
    assert sha(x) == g["in_sha256"]["x"] and sha(diags) == g["in_sha256"]["diags"]
    key = lambda seed: (lambda k: (c.fill_uniform(seed, k, 2 * L), k)[1])(torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda"))
    gk_baby = [key(0xD3390004 + 10 + b) for b in range(1, baby)]
    gk_giant = key(0xD3390004 + 99)
    out = torch.empty_like(x)
    c.linear_bsgs(x, diags, gk_baby, gk_giant, baby, out, B)
    torch.cuda.synchronize()
    assert sha(out) == g["out_sha256"]
    # the same layer as a library object (dpfhe_linear_*): weights and keys uploaded once, device and host-buffer forms
    h = lambda t: t.cpu().numpy().view(np.uint64)
    layer = dp.LinearLayer(c, h(diags), baby, h(torch.stack(gk_baby)), h(gk_giant))
    out2 = torch.zeros_like(x)
    layer.apply(x, out2, B)
    torch.cuda.synchronize()
    assert sha(out2) == g["out_sha256"]
    hx, ho = h(x), np.zeros((B, 2, L, N), dtype=np.uint64)
    layer.apply_host(hx, ho)
    assert hashlib.sha256(ho.tobytes()).hexdigest() == g["out_sha256"]
    layer.apply_host(hx[:37], ho[:37])                      # ragged batch, one chunk
    assert np.array_equal(ho[:37], h(out2)[:37])
    layer.close()
    c.close()
