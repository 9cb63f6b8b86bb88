"""CPU tests pinning the oracle (PARITY UNPINNED against the reference: SURVEY.md §8c — the reference
has no implementation, tests or vectors for this path, so the oracle is pinned by definitions and by
the committed self-generated fixtures in tests/golden/kat.json)."""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.uint64).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def kat():
    with open(os.path.join(HERE, "golden", "kat.json")) as f:
        return json.load(f)


def test_mulmod_variants_agree(oracle_mod):
    lib = oracle_mod.lib()
    o = oracle_mod.Oracle(12, 3)
    rng = np.random.default_rng(1)
    for q in o.moduli + [(1 << 34) + 0x6001, 998244353 * 2**3 * 1 + 0]:
        if q < (1 << 33):
            continue
        edge = [0, 1, 2, q - 1, q - 2, q // 2, q // 2 + 1]
        vals = edge + [int(v) for v in rng.integers(0, q, 3000, dtype=np.uint64)]
        for a, b in zip(vals, reversed(vals)):
            ref = lib.dpo_mulmod_ref(a, b, q)
            assert ref == (a * b) % q
            assert lib.dpo_mulmod_barrett(a, b, q) == ref
            assert lib.dpo_mulmod_shoup(a, b, lib.dpo_shoup_precompute(b, q), q) == ref
    # Shoup accepts any 64-bit left operand
    q = o.moduli[0]
    w = q - 5
    ws = lib.dpo_shoup_precompute(w, q)
    for x in (2**64 - 1, 2**63, 15 * q + 3):
        assert lib.dpo_mulmod_shoup(x, w, ws, q) == (x * w) % q


def test_parameter_derivation(oracle_mod, kat):
    for key, ref in kat["params"].items():
        log_n, L = map(int, key.split(","))
        o = oracle_mod.Oracle(log_n, L)
        assert [str(q) for q in o.moduli] == ref["moduli"]
        assert [str(p) for p in o.psi] == ref["psi"]
        two_n = 2 << log_n
        for q, psi in zip(o.moduli, o.psi):
            assert q < (1 << 60) and q % two_n == 1 and oracle_mod.lib().dpo_is_prime(q)
            assert pow(psi, two_n // 2, q) == q - 1          # primitive 2N-th root
        assert o.moduli == sorted(o.moduli, reverse=True)
    # the moduli are the LARGEST primes k * 2^32 + 1 below 2^60: nothing prime of that form in between
    o = oracle_mod.Oracle(13, 4)
    cand, found = (1 << 60) + 1, []
    while len(found) < 4:
        cand -= 1 << 32
        if oracle_mod.lib().dpo_is_prime(cand):
            found.append(cand)
    assert found == o.moduli


def test_root_power_tables(oracle_mod):
    o = oracle_mod.Oracle(12, 2)
    for l in range(2):
        q, psi = o.moduli[l], o.psi[l]
        rp, irp = o.root_powers(l), o.inv_root_powers(l)
        for i in (0, 1, 2, 3, 5, 1000, 4095):
            br = int(format(i, "012b")[::-1], 2)
            assert int(rp[i]) == pow(psi, br, q)
            assert (int(rp[i]) * int(irp[i])) % q == 1
        assert (o.inv_n(l) * o.N) % q == 1


@pytest.mark.parametrize("log_n", [3, 6, 10])
def test_ntt_against_definitions(oracle_mod, log_n):
    o = oracle_mod.Oracle(log_n, 2)
    a = o.fill_uniform(5, 1)[0]
    b = o.fill_uniform(6, 1)[0]
    fa, fb = o.ntt_fwd(a), o.ntt_fwd(b)
    for l in range(2):
        # (2) table-free O(N^2) evaluation at psi^(2*bitrev(i)+1)
        assert np.array_equal(fa[l], o.ntt_fwd_limb_slow(l, a[l]))
    # (1) inverse o forward = identity
    assert np.array_equal(o.ntt_inv(fa), a)
    # (3) convolution theorem against the O(N^2) negacyclic schoolbook product
    prod = o.ntt_inv(o.poly_mul_pointwise(fa, fb))
    for l in range(2):
        assert np.array_equal(prod[l], o.schoolbook(l, a[l], b[l]))


def test_ntt_of_monomial(oracle_mod):
    """NTT(X^k)[i] = psi^(k * (2*bitrev(i) + 1))"""
    o = oracle_mod.Oracle(12, 1)
    q, psi, N = o.moduli[0], o.psi[0], o.N
    for k in (0, 1, 2, 77, N - 1):
        x = np.zeros((1, 1, N), dtype=np.uint64)
        x[0, 0, k] = 1
        y = o.ntt_fwd(x)[0, 0]
        for i in (0, 1, 2, 3, 1234, N - 1):
            br = int(format(i, "012b")[::-1], 2)
            assert int(y[i]) == pow(psi, k * (2 * br + 1), q)


def test_ntt_4096_schoolbook(oracle_mod):
    o = oracle_mod.Oracle(12, 1)
    a = o.fill_uniform(0xD3390001, 1)[0]
    b = np.zeros_like(a)
    b[0, [0, 1, 17, 4095]] = [3, q_minus := o.moduli[0] - 1, 5, 7]     # sparse, keeps O(N^2) cheap
    prod = o.ntt_inv(o.poly_mul_pointwise(o.ntt_fwd(a), o.ntt_fwd(b)))
    assert np.array_equal(prod[0], o.schoolbook(0, b[0], a[0]))


def test_golden_vectors(oracle_mod, kat):
    for case in kat["cases"]:
        o = oracle_mod.Oracle(case["log_n"], case["L"])
        if "ntt_fwd" in case["name"]:
            x = o.fill_uniform(case["seed"], case["n_polys"])
            assert sha(x) == case["in_sha256"]
            assert [str(v) for v in x.reshape(-1)[:8]] == case["in_head"]
            y = o.ntt_fwd(x)
            assert sha(y) == case["out_sha256"]
            assert [str(v) for v in y.reshape(-1)[:8]] == case["out_head"]
        elif "grouped" in case["name"]:
            o4 = oracle_mod.Oracle(13, 4)
            K = case["K"]
            s = o.keygen_secret(1)
            a = o4.fill_uniform(case["seed"], 4).reshape(2, 2, 4, o.N)
            b = o4.fill_uniform(case["seed"], 4, first_poly=4).reshape(2, 2, 4, o.N)
            if "rotate" in case["name"]:
                gk = o.keygen_galois_grouped(K, 3, case["t"], s, case["galois"])
                assert sha(gk) == case["gk_sha256"]
                assert sha(o.rotate_grouped(K, a, case["galois"], gk, case["t"])) == case["out_sha256"]
                assert sha(o.rotate_hoisted_grouped(K, a, [case["galois"]], gk[None], case["t"])) == case["hoisted_sha256"]
            else:
                evk = o.keygen_relin_grouped(K, 2, case["t"], s)
                assert sha(evk) == case["evk_sha256"]
                assert sha(o.ct_mul_relin_grouped(K, a, b, evk, case["t"])) == case["out_sha256"]
        elif "hybrid" in case["name"]:
            o4 = oracle_mod.Oracle(13, 4)
            s = o.keygen_secret(1)
            a = o4.fill_uniform(case["seed"], 4).reshape(2, 2, 4, o.N)
            b = o4.fill_uniform(case["seed"], 4, first_poly=4).reshape(2, 2, 4, o.N)
            if "rotate" in case["name"]:
                gk = o.keygen_galois_hybrid(3, case["t"], s, case["galois"])
                assert sha(gk) == case["gk_sha256"]
                assert sha(o.rotate_hybrid(a, case["galois"], gk, case["t"])) == case["out_sha256"]
            else:
                evk = o.keygen_relin_hybrid(2, case["t"], s)
                assert sha(evk) == case["evk_sha256"]
                assert sha(o.ct_mul_relin_hybrid(a, b, evk, case["t"])) == case["out_sha256"]
        elif "mod_switch" in case["name"]:
            x = o.fill_uniform(case["seed"], 4)
            assert sha(o.mod_switch_down(x, case["t"])) == case["out_sha256"]
        elif "ct_mul_relin" in case["name"]:
            s = o.keygen_secret(1)
            evk = o.keygen_relin(2, 65537, s)
            assert sha(s) == case["secret_sha256"] and sha(evk) == case["evk_sha256"]
            a = o.fill_uniform(case["seed"], 4).reshape(2, 2, 4, o.N)
            b = o.fill_uniform(case["seed"], 4, first_poly=4).reshape(2, 2, 4, o.N)
            assert sha(o.ct_mul_relin(a, b, evk)) == case["out_sha256"]
        elif "rotate" in case["name"]:
            s = o.keygen_secret(1)
            gk = o.keygen_galois(3, 65537, s, case["galois"])
            assert sha(gk) == case["gk_sha256"]
            a = o.fill_uniform(case["seed"], 4).reshape(2, 2, 4, o.N)
            assert sha(o.rotate(a, case["galois"], gk)) == case["out_sha256"]


def negacyclic_mod_t(a, b, t):
    n = len(a)
    r = np.convolve(np.array([int(v) for v in a], dtype=object), np.array([int(v) for v in b], dtype=object))
    out = [0] * n
    for i, v in enumerate(r):
        if i < n:
            out[i] += v
        else:
            out[i - n] -= v
    return np.array([v % t for v in out], dtype=np.uint64)


def test_scheme_semantics(oracle_mod):
    """Dec(Enc m) = m;  Dec(ct x ct) = m1*m2;  Dec(rotate) = sigma_g(m)  — the meaning of the hot path."""
    o = oracle_mod.Oracle(10, 3)
    t = 65537
    rng = np.random.default_rng(3)
    s = o.keygen_secret(7)
    evk = o.keygen_relin(8, t, s)
    m1 = rng.integers(0, t, o.N).astype(np.uint64)
    m2 = rng.integers(0, t, o.N).astype(np.uint64)
    c1, c2 = o.encrypt(9, t, s, m1), o.encrypt(10, t, s, m2)
    assert np.array_equal(o.decrypt(s, c1, t), m1)
    d = o.ct_tensor(c1[None], c2[None])[0]
    assert np.array_equal(o.decrypt(s, d, t), negacyclic_mod_t(m1, m2, t))          # 3-component decrypt
    c3 = o.ct_mul_relin(c1[None], c2[None], evk)[0]
    assert np.array_equal(o.decrypt(s, c3, t), negacyclic_mod_t(m1, m2, t))          # after relinearisation
    pt = np.zeros((3, o.N), dtype=np.uint64)
    pt[:, 0] = 5
    pt_eval = o.ntt_fwd(pt[None])[0]
    assert np.array_equal(o.decrypt(s, o.ct_mul_plain(c1[None], pt_eval)[0], t), (5 * m1) % t)
    for k in (1, -1, 3):
        g = o.galois_elt(k)
        gk = o.keygen_galois(11, t, s, g)
        r = o.rotate(c1[None], g, gk)[0]
        exp = np.zeros(o.N, dtype=np.uint64)
        for i in range(o.N):
            e = (i * g) % (2 * o.N)
            if e < o.N:
                exp[e] = m1[i]
            else:
                exp[e - o.N] = (t - m1[i]) % t
        assert np.array_equal(o.decrypt(s, r, t), exp)


@pytest.mark.parametrize("t", [0, 65537])
def test_mod_switch_down_is_exact_rounded_division(oracle_mod, t):
    """out == (X - w) / q_last over the integers, w = the centred representative of X (t = 0) or t * centred(X / t) (BGV)
    modulo q_last: checked coefficient by coefficient with Python integers through the CRT."""
    o = oracle_mod.Oracle(8, 3)
    lo = oracle_mod.Oracle(8, 2, o.moduli[:2])
    x = o.fill_uniform(31, 2)
    x[1, 2] = 0   # divisible by q_last already: nothing to round
    y = lo.ntt_inv(o.mod_switch_down(x, t))
    xc = o.ntt_inv(x)
    ql = o.moduli[2]
    Q = o.moduli[0] * o.moduli[1] * ql
    coef = [(Q // q) * pow(Q // q, -1, q) for q in o.moduli]
    for pidx in range(2):
        for n in range(o.N):
            X = sum(int(xc[pidx, l, n]) * coef[l] for l in range(3)) % Q
            r = (X * pow(t, -1, ql)) % ql if t else X % ql
            if r > ql // 2:
                r -= ql
            w = r * (t if t else 1)
            assert (X - w) % ql == 0
            Y = (X - w) // ql
            assert [int(y[pidx, l, n]) for l in range(2)] == [Y % q for q in o.moduli[:2]]
            if t:
                assert (Y * ql - X) % t == 0   # the plaintext moves by exactly q_last^-1 mod t


def test_hybrid_key_switching_semantics(oracle_mod):
    """special-prime variant: Dec(ct x ct) = m1*m2 and Dec(rotate) = sigma_g(m) under the L-1 ciphertext moduli, with
    far less noise than the per-limb-digit variant; keyswitch_hybrid == mod_switch_down of the (L)-limb inner product."""
    o = oracle_mod.Oracle(10, 4)
    o3 = oracle_mod.Oracle(10, 3, o.moduli[:3])
    t = 65537
    rng = np.random.default_rng(4)
    s = o.keygen_secret(7)
    s3 = np.ascontiguousarray(s[:3])
    assert np.array_equal(o3.keygen_secret(7), s3)
    m1 = rng.integers(0, t, o.N).astype(np.uint64)
    m2 = rng.integers(0, t, o.N).astype(np.uint64)
    c1, c2 = o3.encrypt(9, t, s3, m1), o3.encrypt(10, t, s3, m2)
    hyb = o.ct_mul_relin_hybrid(c1[None], c2[None], o.keygen_relin_hybrid(8, t, s), t)[0]
    bv = o3.ct_mul_relin(c1[None], c2[None], o3.keygen_relin(8, t, s3))[0]
    exp = negacyclic_mod_t(m1, m2, t)
    assert np.array_equal(o3.decrypt(s3, hyb, t), exp) and np.array_equal(o3.decrypt(s3, bv, t), exp)

    def noise_bits(ct):
        ph = o3.phase(s3, ct)
        Q = o3.moduli[0] * o3.moduli[1] * o3.moduli[2]
        coef = [(Q // q) * pow(Q // q, -1, q) for q in o3.moduli]
        worst = 0
        for n in range(0, o.N, 7):
            v = sum(int(ph[l][n]) * coef[l] for l in range(3)) % Q
            worst = max(worst, min(v, Q - v))
        return worst.bit_length()

    assert noise_bits(hyb) + 30 < noise_bits(bv)
    g = o.galois_elt(1)
    r = o.rotate_hybrid(c1[None], g, o.keygen_galois_hybrid(11, t, s, g), t)[0]
    r_bv = o3.rotate(c1[None], g, o3.keygen_galois(11, t, s3, g))[0]
    assert np.array_equal(o3.decrypt(s3, r, t), o3.decrypt(s3, r_bv, t))
    # definition check of the bare key switch against its parts
    d = o3.fill_uniform(12, 1)[0]
    key = o.fill_uniform(13, 6).reshape(3, 2, 4, o.N)
    acc = np.zeros((2, 4, o.N), dtype=np.uint64)
    for j in range(3):
        tj = o3.ntt_inv(d[None])[0][j]
        lifted = np.stack([tj % np.uint64(q) for q in o.moduli])
        u = o.ntt_fwd(lifted[None])[0]
        u[j] = d[j]
        for c in range(2):
            acc[c] = o.poly_add(acc[c][None], o.poly_mul_pointwise(u[None], key[j, c][None]))[0]
    c0, c1_ = o.keyswitch_hybrid(d, key, t)
    low = o.mod_switch_down(acc, t)
    assert np.array_equal(c0, low[0]) and np.array_equal(c1_, low[1])


def test_galois_perm_matches_coefficient_automorphism(oracle_mod):
    o = oracle_mod.Oracle(8, 2)
    a = o.fill_uniform(4, 1)[0]
    fa = o.ntt_fwd(a)
    for g in (3, 5, 25, 2 * o.N - 1):
        perm = o.galois_perm(g)
        sig = np.stack([o.galois_coeff(l, g, a[l]) for l in range(2)])
        assert np.array_equal(o.ntt_fwd(sig), fa[:, perm])


def test_fill_uniform_is_counter_based(oracle_mod):
    o = oracle_mod.Oracle(10, 2)
    full = o.fill_uniform(99, 6)
    assert np.array_equal(full[2:5], o.fill_uniform(99, 3, first_poly=2))
    q = np.array(o.moduli, dtype=np.uint64)[None, :, None]
    assert (full < q).all()
    lib = oracle_mod.lib()
    k = 3 * o.P + 1 * o.N + 7
    assert int(full[3, 1, 7]) == (lib.dpo_splitmix64(99 + k) * o.moduli[1]) >> 64
