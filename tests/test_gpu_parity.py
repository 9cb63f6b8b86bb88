"""GPU parity tests: every entry point of the C ABI (include/dpfhe.h) against the CPU oracle, bit-exact.

The reference has no tests for this path, so the cases follow the oracle's own
pins: seeded uniform residues, edge values (0, q-1), ragged batches, every supported N, custom moduli,
and size-independent properties at the BASELINE.json config-2 shape.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.fixture(scope="module")
def dp():
    import deeppowers_b200
    return deeppowers_b200


@pytest.fixture(scope="module")
def ctxs(dp, oracle_mod):
    cache = {}

    def get(log_n, L, moduli=None):
        key = (log_n, L, tuple(moduli) if moduli else None)
        if key not in cache:
            cache[key] = (dp.Context(log_n, L, moduli), oracle_mod.Oracle(log_n, L, moduli))
        return cache[key]

    yield get
    for c, _ in cache.values():
        c.close()


def edge_polys(o, n_polys, seed):
    """uniform residues with a few adversarial rows: all zero, all q-1, alternating 0 / q-1"""
    x = o.fill_uniform(seed, n_polys)
    q = np.array(o.moduli, dtype=np.uint64)
    if n_polys >= 3:
        x[0] = 0
        x[1] = (q - 1)[:, None]
        x[2, :, ::2] = 0
        x[2, :, 1::2] = (q - 1)[:, None]
    return x


@pytest.mark.parametrize("log_n,L", [(12, 1), (12, 3), (13, 4), (14, 2), (14, 8)])
def test_tables_match_oracle(ctxs, log_n, L):
    c, o = ctxs(log_n, L)
    assert c.moduli == o.moduli
    assert c.psi == o.psi
    for l in (0, L - 1):
        assert np.array_equal(c.root_powers(l), o.root_powers(l))
        assert np.array_equal(c.root_powers(l, inverse=True), o.inv_root_powers(l))


@pytest.mark.parametrize("log_n,L,n_polys", [(12, 1, 1), (12, 3, 5), (13, 4, 7), (14, 2, 3), (14, 8, 2)])
def test_ntt_fwd_inv(ctxs, log_n, L, n_polys):
    c, o = ctxs(log_n, L)
    x = edge_polys(o, n_polys, 0xD3390001)
    d = dev(x)
    c.ntt_fwd(d, n_polys)
    y = host(d).reshape(x.shape)
    assert np.array_equal(y, o.ntt_fwd(x))
    c.ntt_inv(d, n_polys)
    assert np.array_equal(host(d).reshape(x.shape), x)


@pytest.mark.parametrize("log_n,L,n_polys", [(12, 2, 5), (13, 4, 9)])
def test_ntt_inv_thread_copy_and_tma_paths(dp, oracle_mod, monkeypatch, log_n, L, n_polys):
    """the inverse transform fetches its limb through a TMA tensor map by default and by thread copies with DPFHE_NTT_TMA=0:
    both must give the oracle's bits (and an offset, unaligned-to-the-allocation view exercises the tensor map's base address)"""
    o = oracle_mod.Oracle(log_n, L)
    x = edge_polys(o, n_polys, 0xD3390011)
    exp = o.ntt_inv(x)
    for env in ("1", "0"):
        monkeypatch.setenv("DPFHE_NTT_TMA", env)
        c = dp.Context(log_n, L)
        big = torch.zeros((n_polys + 1, L, o.N), dtype=torch.int64, device="cuda")
        d = big[1:]                                   # starts one polynomial into the allocation
        d.copy_(dev(x))
        c.ntt_inv(d, n_polys)
        assert np.array_equal(host(d).reshape(x.shape), exp), env
        assert int(big[0].abs().sum()) == 0           # nothing written before the view
        c.close()
    monkeypatch.delenv("DPFHE_NTT_TMA")


def test_ntt_custom_moduli(ctxs, oracle_mod):
    # 50-bit and 36-bit NTT-friendly primes exercise the generic Barrett / word-reduce constants
    mods = []
    two_n = 2 << 12
    for start in ((1 << 50), (1 << 36), (1 << 59)):
        cand = (start // two_n) * two_n + 1
        while True:
            cand -= two_n
            if oracle_mod.lib().dpo_is_prime(cand):
                mods.append(cand)
                break
    c, o = ctxs(12, 3, mods)
    x = edge_polys(o, 4, 77)
    d = dev(x)
    c.ntt_fwd(d, 4)
    assert np.array_equal(host(d).reshape(x.shape), o.ntt_fwd(x))
    c.ntt_inv(d, 4)
    assert np.array_equal(host(d).reshape(x.shape), x)
    # ct x ct with mixed-size moduli: the digit lift t mod q_i really reduces here
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    a = o.fill_uniform(3, 2 * 3).reshape(3, 2, 3, o.N)
    b = o.fill_uniform(4, 2 * 3).reshape(3, 2, 3, o.N)
    out = torch.zeros(a.shape, dtype=torch.int64, device="cuda")
    c.ct_mul_relin(dev(a), dev(b), dev(evk), out, 3)
    assert np.array_equal(host(out).reshape(a.shape), o.ct_mul_relin(a, b, evk))


@pytest.mark.parametrize("log_n,L", [(12, 2), (13, 4), (14, 2)])
def test_pointwise_and_tensor_and_plain(ctxs, log_n, L):
    c, o = ctxs(log_n, L)
    B = 3
    a = edge_polys(o, 2 * B, 11).reshape(B, 2, L, o.N)
    b = edge_polys(o, 2 * B, 12)[::-1].copy().reshape(B, 2, L, o.N)
    da, db = dev(a), dev(b)
    out = torch.empty_like(da)
    c.poly_mul_pointwise(da, db, out, 2 * B)
    assert np.array_equal(host(out).reshape(a.shape), o.poly_mul_pointwise(a, b))
    d = torch.empty((B, 3, L, o.N), dtype=torch.int64, device="cuda")
    c.ct_tensor(da, db, d, B)
    assert np.array_equal(host(d).reshape(B, 3, L, o.N), o.ct_tensor(a, b))
    pt = o.fill_uniform(13, 1)[0]
    c.ct_mul_plain(da, dev(pt), out, B)
    assert np.array_equal(host(out).reshape(a.shape), o.ct_mul_plain(a, pt))


@pytest.mark.parametrize("log_n,L,batch", [(12, 2, 1), (12, 3, 5), (13, 4, 3), (13, 1, 2), (13, 4, 41), (14, 2, 3), (14, 8, 2), (12, 9, 3), (12, 16, 2)])
def test_ct_mul_relin(ctxs, log_n, L, batch):
    c, o = ctxs(log_n, L)
    s = o.keygen_secret(21)
    evk = o.keygen_relin(22, 65537, s)
    a = edge_polys(o, 2 * batch, 23).reshape(batch, 2, L, o.N)
    b = o.fill_uniform(24, 2 * batch).reshape(batch, 2, L, o.N)
    out = torch.zeros(a.shape, dtype=torch.int64, device="cuda")
    c.ct_mul_relin(dev(a), dev(b), dev(evk), out, batch)
    got = host(out).reshape(a.shape)
    ref = o.ct_mul_relin(a, b, evk)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("log_n,L,batch", [(12, 3, 4), (13, 4, 5), (14, 8, 2)])
def test_keyswitch_and_rotate(ctxs, log_n, L, batch):
    c, o = ctxs(log_n, L)
    s = o.keygen_secret(31)
    evk = o.keygen_relin(32, 65537, s)
    d = edge_polys(o, batch, 33)
    out = torch.zeros((batch, 2, L, o.N), dtype=torch.int64, device="cuda")
    c.keyswitch(dev(d), dev(evk), out, batch)
    ref = np.stack([np.stack(o.keyswitch(d[k], evk)) for k in range(batch)])
    assert np.array_equal(host(out).reshape(ref.shape), ref)
    ct = o.fill_uniform(34, 2 * batch).reshape(batch, 2, L, o.N)
    for k in (1, -1, 7):
        g = o.galois_elt(k)
        assert g == c.galois_elt(k)
        gk = o.keygen_galois(35 + k, 65537, s, g)
        c.rotate(dev(ct), g, dev(gk), out, batch)
        assert np.array_equal(host(out).reshape(ct.shape), o.rotate(ct, g, gk))


@pytest.mark.parametrize("log_n,L,batch", [(12, 1, 2), (12, 3, 5), (13, 4, 37), (14, 8, 2), (12, 16, 2)])
def test_rotate_hoisted(ctxs, log_n, L, batch):
    """n rotations of the same batch sharing the digit transforms == n independent rotations, bit for bit"""
    c, o = ctxs(log_n, L)
    ct = edge_polys(o, 2 * batch, 91).reshape(batch, 2, L, o.N)      # includes an all-zero c0 and extreme rows
    ct[batch - 1, 1] = 0                                              # c1 = 0: zero digits, the fallback path
    if batch > 2:
        ct[2, 1, L - 1] = 0
    galois = [o.galois_elt(k) for k in (1, -1, 7)] + [2 * o.N - 1]
    keys = [o.fill_uniform(100 + r, 2 * L).reshape(L, 2, L, o.N) for r in range(len(galois))]
    d_keys = [dev(k) for k in keys]
    out = torch.full((len(galois), batch, 2, L, o.N), -1, dtype=torch.int64, device="cuda")
    d_ct = dev(ct)
    c.rotate_hoisted(d_ct, galois, d_keys, out, batch)
    single = torch.empty((batch, 2, L, o.N), dtype=torch.int64, device="cuda")
    for r, g in enumerate(galois):
        assert np.array_equal(host(out[r]).reshape(ct.shape), o.rotate(ct, g, keys[r])), "rotation %d vs oracle" % r
        c.rotate(d_ct, g, d_keys[r], single, batch)
        assert torch.equal(out[r], single), "rotation %d vs dpfhe_rotate" % r
    with pytest.raises(RuntimeError, match="galois"):
        c.rotate_hoisted(d_ct, [2], d_keys[:1], out, batch)


def test_rotate_hoisted_in_chunks(dp, oracle_mod, monkeypatch):
    """a scratch cap smaller than the batch: the shared transforms are produced and consumed chunk by chunk"""
    monkeypatch.setenv("DPFHE_HOIST_CAP_MB", "1")          # 288 KiB of shared transforms per ciphertext at (12, 3): 3 per chunk
    o = oracle_mod.Oracle(12, 3)
    c = dp.Context(12, 3)
    batch = 8
    ct = o.fill_uniform(93, 2 * batch).reshape(batch, 2, 3, o.N)
    ct[4, 1] = 0
    galois = [o.galois_elt(k) for k in (2, -3)]
    keys = [o.fill_uniform(110 + r, 6).reshape(3, 2, 3, o.N) for r in range(2)]
    out = torch.zeros((2, batch, 2, 3, o.N), dtype=torch.int64, device="cuda")
    c.rotate_hoisted(dev(ct), galois, [dev(k) for k in keys], out, batch)
    for r in range(2):
        assert np.array_equal(host(out[r]).reshape(ct.shape), o.rotate(ct, galois[r], keys[r]))
    c.close()


def test_semantics_decrypt_of_product(ctxs):
    """Dec(GPU ct x ct) == a*b mod (X^N+1, t): checks the scheme meaning, not just oracle agreement."""
    c, o = ctxs(12, 3)
    t = 65537
    rng = np.random.default_rng(5)
    s = o.keygen_secret(41)
    evk = o.keygen_relin(42, t, s)
    m1 = rng.integers(0, t, o.N).astype(np.uint64)
    m2 = np.zeros(o.N, dtype=np.uint64)
    m2[1] = 3   # multiply by 3X: a negacyclic shift scaled by 3
    c1, c2 = o.encrypt(43, t, s, m1), o.encrypt(44, t, s, m2)
    out = torch.zeros((1, 2, 3, o.N), dtype=torch.int64, device="cuda")
    c.ct_mul_relin(dev(c1[None]), dev(c2[None]), dev(evk), out, 1)
    dec = o.decrypt(s, host(out).reshape(2, 3, o.N), t)
    exp = np.empty_like(m1)
    exp[1:] = (3 * m1[:-1]) % t
    exp[0] = (t - (3 * m1[-1]) % t) % t
    assert np.array_equal(dec, exp)


@pytest.mark.parametrize("log_n,L,n_polys,t", [(12, 2, 3, 0), (12, 3, 6, 65537), (13, 4, 10, 65537), (13, 4, 4, 167772161), (14, 8, 4, 65537), (12, 16, 2, 0)])
def test_mod_switch_down(ctxs, log_n, L, n_polys, t):
    c, o = ctxs(log_n, L)
    x = edge_polys(o, n_polys, 71)
    out = torch.full((n_polys, L - 1, o.N), -1, dtype=torch.int64, device="cuda")
    c.mod_switch_down(dev(x), out, n_polys, t)
    assert np.array_equal(host(out).reshape(n_polys, L - 1, o.N), o.mod_switch_down(x, t))
    # a second call with more polynomials regrows the scratch
    x2 = o.fill_uniform(72, 3 * n_polys)
    out2 = torch.empty((3 * n_polys, L - 1, o.N), dtype=torch.int64, device="cuda")
    c.mod_switch_down(dev(x2), out2, 3 * n_polys, t)
    assert np.array_equal(host(out2).reshape(3 * n_polys, L - 1, o.N), o.mod_switch_down(x2, t))


def test_semantics_multiply_then_mod_switch(ctxs, oracle_mod):
    """Dec_{L-1}(modswitch(GPU ct x ct)) == m1*m2 * q_last^-1 mod t: the level-dropping step keeps the BGV meaning."""
    c, o = ctxs(12, 3)
    t = 65537
    rng = np.random.default_rng(6)
    s = o.keygen_secret(51)
    evk = o.keygen_relin(52, t, s)
    m1 = rng.integers(0, t, o.N).astype(np.uint64)
    m2 = np.zeros(o.N, dtype=np.uint64)
    m2[0] = 7
    c1, c2 = o.encrypt(53, t, s, m1), o.encrypt(54, t, s, m2)
    prod = torch.zeros((1, 2, 3, o.N), dtype=torch.int64, device="cuda")
    c.ct_mul_relin(dev(c1[None]), dev(c2[None]), dev(evk), prod, 1)
    low = torch.zeros((2, 2, o.N), dtype=torch.int64, device="cuda")
    c.mod_switch_down(prod, low, 2, t)
    o2 = oracle_mod.Oracle(12, 2, o.moduli[:2])
    dec = o2.decrypt(np.ascontiguousarray(s[:2]), host(low).reshape(2, 2, o.N), t)
    scale = pow(o.moduli[2], -1, t)
    assert np.array_equal(dec, (7 * m1.astype(object) * scale % t).astype(np.uint64))


def hybrid_inputs(o, batch, seed):
    """[batch][2][L-1][N] residues (with edge rows) under the first L-1 moduli, and a uniform hybrid key"""
    Lq = o.L - 1
    x = o.fill_uniform(seed, 2 * batch)[:, :Lq].reshape(batch, 2, Lq, o.N).copy()
    q = np.array(o.moduli[:Lq], dtype=np.uint64)
    x[0, 0] = (q - 1)[:, None]
    x[0, 1, :, ::2] = 0
    key = o.fill_uniform(seed + 1, 2 * Lq).reshape(Lq, 2, o.L, o.N)
    return x, key


@pytest.mark.parametrize("log_n,L,batch,t", [(12, 2, 3, 65537), (12, 4, 7, 65537), (13, 5, 4, 65537), (13, 5, 131, 0), (14, 3, 3, 65537),
                                             (14, 9, 2, 65537), (12, 16, 2, 0)])
def test_hybrid_keyswitch_family(ctxs, log_n, L, batch, t):
    """special-prime key switching (context limb L-1 = special prime): ct x ct, rotate and bare key switch, bit-exact"""
    c, o = ctxs(log_n, L)
    Lq = L - 1
    a, key = hybrid_inputs(o, batch, 81)
    b, _ = hybrid_inputs(o, batch, 83)
    out = torch.full((batch, 2, Lq, o.N), -1, dtype=torch.int64, device="cuda")
    c.ct_mul_relin_hybrid(dev(a), dev(b), dev(key), out, batch, t)
    assert np.array_equal(host(out).reshape(a.shape), o.ct_mul_relin_hybrid(a, b, key, t))
    if batch > 16:
        return   # the large ragged batch exercises scheduling; one mode is enough
    g = o.galois_elt(-2)
    c.rotate_hybrid(dev(a), g, dev(key), out, batch, t)
    assert np.array_equal(host(out).reshape(a.shape), o.rotate_hybrid(a, g, key, t))
    d = np.ascontiguousarray(a[:, 1])
    c.keyswitch_hybrid(dev(d), dev(key), out, batch, t)
    got = host(out).reshape(a.shape)
    for k in range(batch):
        c0, c1 = o.keyswitch_hybrid(d[k], key, t)
        assert np.array_equal(got[k, 0], c0) and np.array_equal(got[k, 1], c1)


def test_semantics_hybrid_product_and_noise(ctxs, oracle_mod):
    """Dec(GPU hybrid ct x ct) == m1*m2, and its noise is far below the per-limb-digit variant's (the point of the special prime)"""
    c, o = ctxs(12, 4)
    o3 = oracle_mod.Oracle(12, 3, o.moduli[:3])
    c3, _ = ctxs(12, 3, o.moduli[:3])
    t = 65537
    rng = np.random.default_rng(8)
    s = o.keygen_secret(61)
    s3 = np.ascontiguousarray(s[:3])
    m1 = rng.integers(0, t, o.N).astype(np.uint64)
    m2 = np.zeros(o.N, dtype=np.uint64)
    m2[2] = 5
    c1, c2 = o3.encrypt(63, t, s3, m1), o3.encrypt(64, t, s3, m2)
    exp = np.empty_like(m1)
    exp[2:] = (5 * m1[:-2]) % t
    exp[:2] = (t - (5 * m1[-2:]) % t) % t

    def noise_bits(ct):
        ph = o3.phase(s3, ct)
        Q = 1
        for q in o3.moduli:
            Q *= q
        coef = [(Q // q) * pow(Q // q, -1, q) for q in o3.moduli]
        worst = 0
        for n in range(0, o.N, 61):
            v = sum(int(ph[l][n]) * coef[l] for l in range(3)) % Q
            worst = max(worst, min(v, Q - v))
        return worst.bit_length()

    out = torch.zeros((1, 2, 3, o.N), dtype=torch.int64, device="cuda")
    c.ct_mul_relin_hybrid(dev(c1[None]), dev(c2[None]), dev(o.keygen_relin_hybrid(62, t, s)), out, 1, t)
    hyb = host(out).reshape(2, 3, o.N).copy()
    assert np.array_equal(o3.decrypt(s3, hyb, t), exp)
    c3.ct_mul_relin(dev(c1[None]), dev(c2[None]), dev(o3.keygen_relin(62, t, s3)), out, 1)
    bv = host(out).reshape(2, 3, o.N).copy()
    assert np.array_equal(o3.decrypt(s3, bv, t), exp)
    assert noise_bits(hyb) + 30 < noise_bits(bv)


def test_two_level_pipeline_on_gpu(ctxs, oracle_mod):
    """multiply (hybrid) -> mod switch -> multiply at the lower level with that level's context (custom moduli q0 q1 | p) ->
    decrypt: the modulus chain of SURVEY.md section 8 row f-2 through the C ABI (tests/test_levels_cpu.py is the oracle-only twin)"""
    from test_levels_cpu import build_chain
    from test_oracle_kat import negacyclic_mod_t
    ch = build_chain(oracle_mod, 12)
    top, l3, low, l2 = ch["top"], ch["l3"], ch["low"], ch["l2"]
    g_top, _ = ctxs(12, 4)
    g_l3, _ = ctxs(12, 3, l3.moduli)
    g_low, _ = ctxs(12, 3, low.moduli)
    t = 65537
    rng = np.random.default_rng(22)
    s_top, s3, s2, s_low = top.keygen_secret(31), l3.keygen_secret(31), l2.keygen_secret(31), low.keygen_secret(31)
    m1 = rng.integers(0, t, top.N).astype(np.uint64)
    m2 = np.zeros(top.N, dtype=np.uint64); m2[1] = 2          # sparse factors keep the host-side expectation cheap
    m3 = np.zeros(top.N, dtype=np.uint64); m3[3] = 5
    c1, c2, c3 = l3.encrypt(41, t, s3, m1), l3.encrypt(42, t, s3, m2), l2.encrypt(44, t, s2, m3)
    prod = torch.zeros((1, 2, 3, top.N), dtype=torch.int64, device="cuda")
    g_top.ct_mul_relin_hybrid(dev(c1[None]), dev(c2[None]), dev(top.keygen_relin_hybrid(43, t, s_top)), prod, 1, t)
    down = torch.zeros((1, 2, 2, top.N), dtype=torch.int64, device="cuda")
    g_l3.mod_switch_down(prod, down, 2, t)
    prod2 = torch.zeros_like(down)
    g_low.ct_mul_relin_hybrid(down, dev(c3[None]), dev(low.keygen_relin_hybrid(45, t, s_low)), prod2, 1, t)
    scale = pow(top.moduli[2], -1, t)
    exp = np.zeros(top.N, dtype=object)
    for k in range(top.N):                                     # m1 * 2X * 5X^3 = 10 X^4 m1, negacyclic
        v = 10 * int(m1[k]) * scale
        exp[(k + 4) % top.N] = (v if k + 4 < top.N else -v) % t
    assert np.array_equal(l2.decrypt(s2, host(prod2).reshape(2, 2, top.N), t), exp.astype(np.uint64))


def test_hybrid_errors(ctxs):
    c, o = ctxs(12, 1)
    x = torch.zeros((1, 2, 1, o.N), dtype=torch.int64, device="cuda")
    with pytest.raises(RuntimeError, match="special prime"):
        c.ct_mul_relin_hybrid(x, x.clone(), x.clone(), x.clone(), 1)


def test_fill_uniform_matches_oracle(ctxs):
    c, o = ctxs(13, 4)
    d = torch.empty((3, 4, o.N), dtype=torch.int64, device="cuda")
    c.fill_uniform(0xD3390002, d, 3, first_poly=5)
    assert np.array_equal(host(d).reshape(3, 4, o.N), o.fill_uniform(0xD3390002, 3, first_poly=5))


def test_host_entry_points(ctxs):
    c, o = ctxs(13, 4)
    s = o.keygen_secret(51)
    evk = o.keygen_relin(52, 65537, s)
    B = 9
    a = o.fill_uniform(53, 2 * B).reshape(B, 2, 4, o.N)
    b = o.fill_uniform(54, 2 * B).reshape(B, 2, 4, o.N)
    out = np.zeros_like(a)
    c.ct_mul_relin_host(a, b, evk, out)
    assert np.array_equal(out, o.ct_mul_relin(a, b, evk))
    x = o.fill_uniform(55, 6)
    y = x.copy()
    c.ntt_fwd_host(y)
    assert np.array_equal(y, o.ntt_fwd(x))
    c.ntt_inv_host(y)
    assert np.array_equal(y, x)
    pt = o.fill_uniform(56, 1)[0]
    c.ct_mul_plain_host(a, pt, out)
    assert np.array_equal(out, o.ct_mul_plain(a, pt))
    g = o.galois_elt(2)
    gk = o.keygen_galois(57, 65537, s, g)
    c.rotate_host(a, g, gk, out)
    assert np.array_equal(out, o.rotate(a, g, gk))


def test_mod_switch_errors(dp, ctxs):
    c, o = ctxs(12, 1)
    x = torch.zeros((1, 1, o.N), dtype=torch.int64, device="cuda")
    with pytest.raises(RuntimeError, match="two limbs"):
        c.mod_switch_down(x, x.clone(), 1)
    c3, _ = ctxs(12, 3)
    y = torch.zeros((1, 3, o.N), dtype=torch.int64, device="cuda")
    with pytest.raises(RuntimeError, match="overlap"):
        c3.mod_switch_down(y, y, 1)


def test_hybrid_and_mod_switch_host_entry_points(ctxs):
    """the host-buffer forms run the same kernels behind the staging pipeline (ragged batch, several chunks)"""
    c, o = ctxs(12, 4)
    batch = 700   # three staging chunks at N=4096 (64 MiB / 192 KiB per ciphertext = 341 per chunk), the last one ragged
    a, key = hybrid_inputs(o, batch, 91)
    b, _ = hybrid_inputs(o, batch, 93)
    out = np.zeros_like(a)
    c.ct_mul_relin_hybrid_host(a, b, key, out, 65537)
    assert np.array_equal(out, o.ct_mul_relin_hybrid(a, b, key, 65537))
    g = o.galois_elt(7)
    c.rotate_hybrid_host(a[:5], g, key, out[:5], 65537)
    assert np.array_equal(out[:5], o.rotate_hybrid(a[:5], g, key, 65537))
    x = o.fill_uniform(95, 9)
    low = np.zeros((9, 3, o.N), dtype=np.uint64)
    c.mod_switch_down_host(x, low, 65537)
    assert np.array_equal(low, o.mod_switch_down(x, 65537))


def test_errors_are_reported(dp, ctxs):
    c, o = ctxs(12, 2)
    with pytest.raises(dp.DpfheError):
        dp.Context(9, 2)                       # unsupported N
    with pytest.raises(dp.DpfheError):
        dp.Context(12, 2, [97, 193])           # moduli out of range
    d = torch.zeros((1, 2, 2, o.N), dtype=torch.int64, device="cuda")
    with pytest.raises(dp.DpfheError):
        c.rotate(d, 4, d, torch.zeros_like(d), 1)   # even Galois element
    with pytest.raises(dp.DpfheError):
        c.ct_mul_relin(d, d, d, d, 1)          # aliasing output
    # partial overlap is refused too: the output of ciphertexts 1..2 lands on input ciphertexts 2..3
    big = torch.zeros((4, 2, 2, o.N), dtype=torch.int64, device="cuda")
    with pytest.raises(dp.DpfheError, match="overlap"):
        c.ct_mul_relin(big[1:3], torch.zeros_like(big[1:3]), d, big[2:4], 2)
    with pytest.raises(dp.DpfheError, match="overlap"):
        c.rotate(big[0:2], 5, d, big[1:3], 2)
    c.ntt_fwd(d, 0)                            # empty batch is a no-op


def test_config2_shape_properties(ctxs):
    """BASELINE.json config 2 shape (N=8192, L=4) on a larger batch than the oracle can chew through:
    size-independent properties + spot checks of individual ciphertexts against the oracle."""
    c, o = ctxs(13, 4)
    B = 1184   # 8 full waves of 148 CTAs x 4 limbs
    P2 = 2 * o.P
    a = torch.empty((B, 2, 4, o.N), dtype=torch.int64, device="cuda")
    b = torch.empty_like(a)
    c.fill_uniform(0xD3390002, a, 2 * B, first_poly=0)
    c.fill_uniform(0xD3390002, b, 2 * B, first_poly=2 * B)
    s = o.keygen_secret(61)
    evk = o.keygen_relin(62, 65537, s)
    devk = dev(evk)
    out = torch.zeros_like(a)
    c.ct_mul_relin(a, b, devk, out, B)
    # (1) commutativity: a (x) b == b (x) a, bit for bit
    out2 = torch.zeros_like(a)
    c.ct_mul_relin(b, a, devk, out2, B)
    assert torch.equal(out, out2)
    # (2) all residues canonical
    q = torch.tensor(np.array(o.moduli, dtype=np.uint64).view(np.int64), device="cuda").view(1, 1, 4, 1)
    assert bool(((out >= 0) & (out < q)).all())
    # (3) spot-check ciphertexts across waves against the oracle
    for k in (0, 1, 147, 148, 591, 592, B - 1):
        ak = o.fill_uniform(0xD3390002, 2, first_poly=2 * k).reshape(1, 2, 4, o.N)
        bk = o.fill_uniform(0xD3390002, 2, first_poly=2 * B + 2 * k).reshape(1, 2, 4, o.N)
        assert np.array_equal(host(a[k]).reshape(1, 2, 4, o.N), ak)
        assert np.array_equal(host(out[k]).reshape(1, 2, 4, o.N), o.ct_mul_relin(ak, bk, evk)), k
    # (4) NTT round trip and linearity at this size
    x = a.clone()
    c.ntt_inv(x, 2 * B)
    c.ntt_fwd(x, 2 * B)
    assert torch.equal(x, a)


def test_host_pipeline_wraps_staging_slots(ctxs):
    """host-buffer ct x ct on a batch spanning more chunks than staging slots (3): result must equal the device path"""
    c, o = ctxs(13, 4)
    B = 700
    a = torch.empty((B, 2, 4, o.N), dtype=torch.int64, device="cuda")
    b = torch.empty_like(a)
    c.fill_uniform(71, a, 2 * B)
    c.fill_uniform(72, b, 2 * B)
    evk = o.fill_uniform(73, 8).reshape(4, 2, 4, o.N)
    out_dev = torch.zeros_like(a)
    c.ct_mul_relin(a, b, dev(evk), out_dev, B)
    ha, hb = host(a).reshape(B, 2, 4, o.N).copy(), host(b).reshape(B, 2, 4, o.N).copy()
    out_host = np.zeros_like(ha)
    c.ct_mul_relin_host(ha, hb, evk, out_host)
    assert np.array_equal(out_host, host(out_dev).reshape(out_host.shape))
    for k in (0, 299, 699):
        assert np.array_equal(out_host[k:k + 1], o.ct_mul_relin(ha[k:k + 1], hb[k:k + 1], evk))


def test_explicit_stream_and_two_contexts(dp, ctxs):
    """ops run on a caller-provided non-default stream; two contexts on one device do not share scratch"""
    c, o = ctxs(12, 2)
    c2 = dp.Context(12, 2)
    s = o.keygen_secret(81)
    evk = o.keygen_relin(82, 65537, s)
    a = o.fill_uniform(83, 6).reshape(3, 2, 2, o.N)
    b = o.fill_uniform(84, 6).reshape(3, 2, 2, o.N)
    ref = o.ct_mul_relin(a, b, evk)
    st = torch.cuda.Stream()
    da, db, dk = dev(a), dev(b), dev(evk)
    out1, out2 = torch.zeros_like(da), torch.zeros_like(da)
    torch.cuda.synchronize()
    with torch.cuda.stream(st):
        c.ct_mul_relin(da, db, dk, out1, 3, stream=st)
        c2.ct_mul_relin(da, db, dk, out2, 3, stream=st)
    st.synchronize()
    assert np.array_equal(host(out1).reshape(ref.shape), ref)
    assert np.array_equal(host(out2).reshape(ref.shape), ref)
    c2.close()


def test_generic_kernels_on_the_default_basis(dp, oracle_mod, monkeypatch):
    """The default basis (k * 2^32 + 1 moduli) selects the dpfhe::fast kernels; DPFHE_FORCE_GENERIC runs the same basis through
    dpfhe::gen, which must give the same bits (both variants are products, not one a fallback of the other), and
    DPFHE_LIFT_REDUCE keeps the word reduction of the digit lift that the default basis may skip."""
    log_n, L, B = 13, 4, 40
    o = oracle_mod.Oracle(log_n, L)
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, 65537, s)
    g = o.galois_elt(-3)
    gk = o.keygen_galois(4, 65537, s, g)
    a = edge_polys(o, 2 * B, 51).reshape(B, 2, L, o.N)
    b = edge_polys(o, 2 * B, 52).reshape(B, 2, L, o.N)
    want_mul, want_rot, want_ntt = o.ct_mul_relin(a, b, evk), o.rotate(a, g, gk), o.ntt_fwd(a.reshape(2 * B, L, o.N))
    for env in ({"DPFHE_FORCE_GENERIC": "1"}, {"DPFHE_LIFT_REDUCE": "1"}, {"DPFHE_KS_SINGLE": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        c = dp.Context(log_n, L)
        out = torch.zeros((B, 2, L, o.N), dtype=torch.int64, device="cuda")
        c.ct_mul_relin(dev(a), dev(b), dev(evk), out, B)
        assert np.array_equal(host(out), want_mul), env
        c.rotate(dev(a), g, dev(gk), out, B)
        assert np.array_equal(host(out), want_rot), env
        d = dev(a)
        c.ntt_fwd(d, 2 * B)
        assert np.array_equal(host(d).reshape(want_ntt.shape), want_ntt), env
        c.ntt_inv(d, 2 * B)
        assert np.array_equal(host(d), a), env
        c.close()
        for k in env:
            monkeypatch.delenv(k)
