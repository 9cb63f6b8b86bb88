"""Grouped hybrid key switching (dnum < L) on the CPU oracle: with one special prime it is the hybrid
variant bit for bit; with two it still decrypts to the product / the rotated message, with far less noise than per-limb digits;
and the division by P is checked against exact big-integer arithmetic."""
import numpy as np
import pytest

from test_oracle_kat import negacyclic_mod_t


def crt(o, rows, n, limbs):
    Q = 1
    for l in limbs:
        Q *= o.moduli[l]
    v = sum(int(rows[l][n]) * (Q // o.moduli[l]) * pow(Q // o.moduli[l], -1, o.moduli[l]) for l in limbs) % Q
    return v, Q


def test_one_special_prime_is_the_hybrid_variant(oracle_mod):
    o = oracle_mod.Oracle(10, 4)
    t = 65537
    s = o.keygen_secret(5)
    assert o.grouped_digits(1) == 3
    assert np.array_equal(o.keygen_relin_grouped(1, 6, t, s), o.keygen_relin_hybrid(6, t, s))
    g = 5
    assert np.array_equal(o.keygen_galois_grouped(1, 7, t, s, g), o.keygen_galois_hybrid(7, t, s, g))
    o3 = oracle_mod.Oracle(10, 3, o.moduli[:3])
    a, b = o3.fill_uniform(1, 4).reshape(2, 2, 3, o.N), o3.fill_uniform(2, 4).reshape(2, 2, 3, o.N)
    key = o.fill_uniform(3, 6).reshape(3, 2, 4, o.N)
    for tp in (0, t):
        assert np.array_equal(o.ct_mul_relin_grouped(1, a, b, key, tp), o.ct_mul_relin_hybrid(a, b, key, tp))
        assert np.array_equal(o.rotate_grouped(1, a, g, key, tp), o.rotate_hybrid(a, g, key, tp))
        x = o.fill_uniform(4, 3).reshape(3, 4, o.N)
        assert np.array_equal(o.mod_down_special(1, x, tp), o.mod_switch_down(x, tp))


@pytest.mark.parametrize("L,K", [(6, 2), (5, 2), (7, 3)])
def test_mod_down_is_an_exact_division(oracle_mod, L, K):
    """out = (x - t*delta) / P over the integers, delta = t^-1 x mod P lifted into (-K P/2, K P/2): checked coefficient-wise by CRT"""
    o = oracle_mod.Oracle(8, L)
    Lq = L - K
    t = 65537
    x = o.fill_uniform(11, 1).reshape(1, L, o.N)
    for tp in (0, t):
        out = o.mod_down_special(K, x, tp)[0]
        xc = o.ntt_inv(x)[0]                                # coefficient form, every limb
        oq = oracle_mod.Oracle(8, Lq, o.moduli[:Lq])
        oc = oq.ntt_inv(out.reshape(1, Lq, o.N))[0]
        P = 1
        for p in o.moduli[Lq:]:
            P *= p
        s = tp if tp else 1
        for n in range(0, o.N, 17):
            X, Qall = crt(o, xc, n, range(L))
            got, Q = crt(o, oc, n, range(Lq))
            # some representative of x (mod Q P) satisfies X' - s*delta = P*got' with |delta| <= K P / 2
            ok = False
            for lift in (X, X - Qall):
                for gl in (got, got - Q):
                    d = lift - P * gl
                    if d % s == 0 and abs(d // s) <= K * P // 2 + 1 and (d // s - (pow(s, -1, P) * lift)) % P == 0:
                        ok = True
            assert ok, n


@pytest.mark.parametrize("L,K", [(6, 2), (5, 2)])
def test_semantics_and_noise(oracle_mod, L, K):
    o = oracle_mod.Oracle(11, L)
    Lq = L - K
    oq = oracle_mod.Oracle(11, Lq, o.moduli[:Lq])
    t = 65537
    rng = np.random.default_rng(3)
    s = o.keygen_secret(61)
    sq = np.ascontiguousarray(s[:Lq])
    m1, m2 = (rng.integers(0, t, o.N).astype(np.uint64) for _ in range(2))
    c1, c2 = oq.encrypt(63, t, sq, m1), oq.encrypt(64, t, sq, m2)
    evk = o.keygen_relin_grouped(K, 62, t, s)
    assert evk.shape == (-(-Lq // K), 2, L, o.N)
    prod = o.ct_mul_relin_grouped(K, c1[None], c2[None], evk, t)[0]
    assert np.array_equal(oq.decrypt(sq, prod, t), negacyclic_mod_t(m1, m2, t))

    def noise_bits(ct):
        ph = oq.phase(sq, ct)
        worst = 0
        for n in range(0, o.N, 61):
            v, Q = crt(oq, ph, n, range(Lq))
            worst = max(worst, min(v, Q - v))
        return worst.bit_length()

    bv = oq.ct_mul_relin(c1[None], c2[None], oq.keygen_relin(62, t, sq))[0]
    assert np.array_equal(oq.decrypt(sq, bv, t), negacyclic_mod_t(m1, m2, t))
    assert noise_bits(prod) + 30 < noise_bits(bv)
    # rotation by one slot: the plaintext polynomial m(X) -> m(X^g)
    g = 5
    rot = o.rotate_grouped(K, c1[None], g, o.keygen_galois_grouped(K, 65, t, s, g), t)[0]
    exp = np.zeros(o.N, dtype=np.uint64)
    for k in range(o.N):
        e = (k * g) % (2 * o.N)
        exp[e % o.N] = m1[k] if e < o.N else (t - m1[k]) % t
    assert np.array_equal(oq.decrypt(sq, rot, t), exp)


@pytest.mark.parametrize("L,K", [(6, 2), (4, 1)])
def test_hoisted_rotations_decrypt_to_the_rotated_message(oracle_mod, L, K):
    """hoisting shares the mod-up of the unpermuted c1: not the bits of rotate_grouped, but the same plaintext and noise level"""
    o = oracle_mod.Oracle(11, L)
    Lq = L - K
    oq = oracle_mod.Oracle(11, Lq, o.moduli[:Lq])
    t = 65537
    rng = np.random.default_rng(5)
    s = o.keygen_secret(71)
    sq = np.ascontiguousarray(s[:Lq])
    msgs = [rng.integers(0, t, o.N).astype(np.uint64) for _ in range(2)]
    cts = np.stack([oq.encrypt(72 + k, t, sq, m) for k, m in enumerate(msgs)])
    galois = [o.galois_elt(1), o.galois_elt(-3), 2 * o.N - 1]
    gks = np.stack([o.keygen_galois_grouped(K, 80 + r, t, s, g) for r, g in enumerate(galois)])
    out = o.rotate_hoisted_grouped(K, cts, galois, gks, t)
    assert out.shape == (3, 2, 2, Lq, o.N)

    def noise_bits(ct):
        ph = oq.phase(sq, ct)
        worst = 0
        for n in range(0, o.N, 61):
            v, Q = crt(oq, ph, n, range(Lq))
            worst = max(worst, min(v, Q - v))
        return worst.bit_length()

    for r, g in enumerate(galois):
        plain = o.rotate_grouped(K, cts, g, gks[r], t)
        for b, m in enumerate(msgs):
            exp = np.zeros(o.N, dtype=np.uint64)
            for k in range(o.N):
                e = (k * g) % (2 * o.N)
                exp[e % o.N] = m[k] if e < o.N else (t - m[k]) % t
            assert np.array_equal(oq.decrypt(sq, out[r, b], t), exp)
            assert abs(noise_bits(out[r, b]) - noise_bits(plain[b])) <= 2
        assert not np.array_equal(out[r], plain)      # a different lift of the rotated digits: different bits, same plaintext
