"""A second, independent derivation of the key-switch family, straight from the definitions of DESIGN.md §2.3-2.8 in
pure Python integers ("closable parity gaps" item c).  Nothing here shares code or method with the oracle:
no butterflies, no Shoup or Barrett reduction, no bit tricks — polynomials are evaluated at the points
psi^(2*bitrev(i)+1) by Horner's rule, interpolated back by the inverse sum, the automorphism is applied to coefficient
vectors, and every product is a Python big-int `%`.  The oracle (and through it the GPU) must agree bit for bit at the
small ring degrees where this O(N^2) arithmetic is affordable.  The reference holds no vectors for this path
(SURVEY.md §8c): parity stays "unpinned", this narrows what can be wrong on our side."""
import numpy as np
import pytest


def bitrev(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


class Ring:
    """Z_q[X]/(X^N+1) in evaluation form at the points psi^(2*br(i)+1), i < N, for every limb"""

    def __init__(self, log_n, moduli, psi):
        self.log_n, self.N, self.q, self.psi = log_n, 1 << log_n, [int(q) for q in moduli], [int(p) for p in psi]
        self.points = [[pow(p, 2 * bitrev(i, log_n) + 1, q) for i in range(self.N)] for p, q in zip(self.psi, self.q)]

    def evaluate(self, coeffs, l):
        q = self.q[l]
        out = []
        for x in self.points[l]:
            acc = 0
            for c in reversed(coeffs):
                acc = (acc * x + c) % q
            out.append(acc)
        return out

    def interpolate(self, values, l):
        """coefficients in [0, q_l) of the polynomial of degree < N taking `values` at the points of limb l:
        a_k = N^-1 * sum_i y_i * x_i^-k, because the x_i are the N roots of X^N + 1"""
        q, n_inv = self.q[l], pow(self.N, -1, self.q[l])
        inv_pts = [pow(x, -1, q) for x in self.points[l]]
        return [n_inv * sum(y * pow(ix, k, q) for y, ix in zip(values, inv_pts)) % q for k in range(self.N)]

    def automorphism(self, coeffs, g, q):
        """X -> X^g on a coefficient vector mod q: X^k -> (+-) X^(k*g mod N), sign from X^N = -1"""
        out = [0] * self.N
        for k, c in enumerate(coeffs):
            e = k * g % (2 * self.N)
            out[e % self.N] = (q - c) % q if e >= self.N else c
        return out

    # ---- the operations, limb by limb, exactly as DESIGN.md states them
    def keyswitch(self, d, key):
        """d: [L][N] evaluation form; key [L digits][2][L][N] -> (c0, c1), each [L][N]"""
        L = len(self.q)
        c = [[[0] * self.N for _ in range(L)] for _ in range(2)]
        for j in range(L):
            t = self.interpolate(d[j], j)                    # integers in [0, q_j)
            for i in range(L):
                u = d[j] if i == j else self.evaluate([x % self.q[i] for x in t], i)
                for comp in range(2):
                    row = c[comp][i]
                    for n in range(self.N):
                        row[n] = (row[n] + u[n] * key[j][comp][i][n]) % self.q[i]
        return c

    def ct_mul_relin(self, a, b, evk):
        L = len(self.q)
        d0 = [[a[0][l][n] * b[0][l][n] % self.q[l] for n in range(self.N)] for l in range(L)]
        d1 = [[(a[0][l][n] * b[1][l][n] + a[1][l][n] * b[0][l][n]) % self.q[l] for n in range(self.N)] for l in range(L)]
        d2 = [[a[1][l][n] * b[1][l][n] % self.q[l] for n in range(self.N)] for l in range(L)]
        ks = self.keyswitch(d2, evk)
        return [[[(x + y) % self.q[l] for x, y in zip(d[l], ks[comp][l])] for l in range(L)] for comp, d in enumerate((d0, d1))]

    def rotate(self, ct, g, gk):
        """(sigma_g(c0) + ks0, ks1), ks = keyswitch(sigma_g(c1), gk); sigma_g through the coefficient form"""
        L = len(self.q)
        sig = [[self.evaluate(self.automorphism(self.interpolate(ct[comp][l], l), g, self.q[l]), l) for l in range(L)] for comp in range(2)]
        ks = self.keyswitch(sig[1], gk)
        return [[[(x + y) % self.q[l] for x, y in zip(sig[0][l], ks[0][l])] for l in range(L)], ks[1]]


def as_lists(a):
    return np.asarray(a, dtype=np.uint64).astype(object).tolist()


@pytest.mark.parametrize("log_n,L", [(3, 1), (4, 2), (5, 3), (6, 2)])
def test_key_switch_family_from_the_definitions(oracle_mod, log_n, L):
    o = oracle_mod.Oracle(log_n, L)
    ring = Ring(log_n, o.moduli, o.psi)
    rng = np.random.default_rng(100 + log_n)
    # the transform itself: the oracle's NTT is evaluation at psi^(2*br(i)+1), its inverse the interpolation
    x = o.fill_uniform(7, 1)
    y = o.ntt_fwd(x)
    for l in range(L):
        assert [int(v) for v in y[0, l]] == ring.evaluate([int(v) for v in x[0, l]], l)
        assert ring.interpolate([int(v) for v in y[0, l]], l) == [int(v) for v in x[0, l]]
    # uniform residues as ciphertexts and keys (the operations are defined for any inputs), plus edge rows
    a = o.fill_uniform(11, 2).reshape(2, L, o.N)
    b = o.fill_uniform(12, 2).reshape(2, L, o.N)
    q = np.array(o.moduli, dtype=np.uint64)
    a[1, :, 0] = q - 1
    b[1, :, 0] = q - 1
    b[0, :, 1] = 0
    key = o.fill_uniform(13, 2 * L).reshape(L, 2, L, o.N)
    got = o.ct_mul_relin(a[None], b[None], key)[0]
    want = ring.ct_mul_relin(as_lists(a), as_lists(b), as_lists(key))
    assert as_lists(got) == want
    d = o.fill_uniform(14, 1)[0]
    c0, c1 = o.keyswitch(d, key)
    want = ring.keyswitch(as_lists(d), as_lists(key))
    assert as_lists(c0) == want[0] and as_lists(c1) == want[1]
    for k in (1, -1, 3):
        g = o.galois_elt(k)
        assert as_lists(o.rotate(a[None], g, key)[0]) == ring.rotate(as_lists(a), g, as_lists(key))
    g = 2 * o.N - 1                                           # the conjugation element
    assert as_lists(o.rotate(a[None], g, key)[0]) == ring.rotate(as_lists(a), g, as_lists(key))
    del rng


def test_real_keys_decrypt_through_the_bigint_path(oracle_mod):
    """with a genuine relinearisation key the big-int ct x ct + relin result decrypts (oracle's decrypt) to the product of the
    plaintexts: the independent derivation is not just self-consistent, it is the scheme's operation"""
    log_n, L, t = 5, 2, 257
    o = oracle_mod.Oracle(log_n, L)
    ring = Ring(log_n, o.moduli, o.psi)
    s = o.keygen_secret(1)
    evk = o.keygen_relin(2, t, s)
    rng = np.random.default_rng(5)
    m1 = rng.integers(0, t, o.N).astype(np.uint64)
    m2 = rng.integers(0, t, o.N).astype(np.uint64)
    c1, c2 = o.encrypt(3, t, s, m1), o.encrypt(4, t, s, m2)
    prod = np.array(ring.ct_mul_relin(as_lists(c1), as_lists(c2), as_lists(evk)), dtype=np.uint64)
    assert np.array_equal(prod, o.ct_mul_relin(c1[None], c2[None], evk)[0])
    neg = [0] * o.N                                           # negacyclic product of the plaintexts mod t
    for i in range(o.N):
        for j in range(o.N):
            k = i + j
            v = int(m1[i]) * int(m2[j])
            neg[k % o.N] = (neg[k % o.N] + (v if k < o.N else -v)) % t
    assert [int(v) for v in o.decrypt(s, prod, t)] == neg


class SpecialPrimeRing(Ring):
    """The special-prime family over the same evaluation points: the last K moduli are special primes, the
    first Lq = L - K carry the ciphertexts.  Digits, lifts and the division are written with Python integers in COEFFICIENT form
    (the oracle and the kernels work limb-wise in evaluation form): a digit is the integer vector v = sum_j y_j * Qhat_j, a lift is
    v mod q_i, the division is ((acc - s * delta) * P^-1) mod q_i with delta the centred recombination of the special residues."""

    def __init__(self, log_n, moduli, psi, K):
        super().__init__(log_n, moduli, psi)
        self.K, self.L, self.Lq = K, len(self.q), len(self.q) - K
        self.dnum = -(-self.Lq // K)
        self.P = 1
        for p in self.q[self.Lq:]:
            self.P *= p

    def group(self, g):
        return range(g * self.K, min((g + 1) * self.K, self.Lq))

    def digit_vectors(self, d):
        """d [Lq][N] evaluation form -> per digit g the integer coefficient vector sum_j y_j * Qhat_j (a lift of the digit: congruent
        to its value modulo Q_g, below |g| * Q_g)"""
        out = []
        for g in range(self.dnum):
            Qg = 1
            for j in self.group(g):
                Qg *= self.q[j]
            v = [0] * self.N
            for j in self.group(g):
                Qhat = Qg // self.q[j]
                y = [c * pow(Qhat, -1, self.q[j]) % self.q[j] for c in self.interpolate(d[j], j)]
                v = [a + b * Qhat for a, b in zip(v, y)]
            out.append(v)
        return out

    def lifted(self, d, vectors):
        """U[g][i]: evaluation form in limb i of digit g's lift; the member limbs keep the ciphertext's own residues"""
        return [[d[i] if i in self.group(g) else self.evaluate([x % self.q[i] for x in vectors[g]], i) for i in range(self.L)]
                for g in range(self.dnum)]

    def divide_by_P(self, acc, t_plain):
        """acc [L][N] evaluation form -> [Lq][N]: (acc - s * delta) / P, delta = t^-1 acc mod P recombined from centred residues"""
        s = t_plain if t_plain else 1
        delta = [0] * self.N
        for k in range(self.K):
            l = self.Lq + k
            p = self.q[l]
            Phat = self.P // p
            f = pow(s * Phat, -1, p)
            for n, c in enumerate(self.interpolate(acc[l], l)):
                y = c * f % p
                delta[n] += (y - p if y > p // 2 else y) * Phat
        out = []
        for i in range(self.Lq):
            q = self.q[i]
            Pinv = pow(self.P, -1, q)
            coeffs = [(c - s * dl) * Pinv % q for c, dl in zip(self.interpolate(acc[i], i), delta)]
            out.append(self.evaluate(coeffs, i))
        return out

    def mac(self, U, key, perm=None):
        acc = [[[0] * self.N for _ in range(self.L)] for _ in range(2)]
        for g in range(self.dnum):
            for i in range(self.L):
                u = U[g][i] if perm is None else [U[g][i][perm[n]] for n in range(self.N)]
                for comp in range(2):
                    row = acc[comp][i]
                    for n in range(self.N):
                        row[n] = (row[n] + u[n] * key[g][comp][i][n]) % self.q[i]
        return acc

    def keyswitch(self, d, key, t_plain, perm=None, vectors=None):
        U = self.lifted(d, vectors if vectors is not None else self.digit_vectors(d))
        acc = self.mac(U, key, perm)
        return [self.divide_by_P(acc[0], t_plain), self.divide_by_P(acc[1], t_plain)]

    def ct_mul_relin(self, a, b, evk, t_plain):
        Lq = self.Lq
        d0 = [[a[0][l][n] * b[0][l][n] % self.q[l] for n in range(self.N)] for l in range(Lq)]
        d1 = [[(a[0][l][n] * b[1][l][n] + a[1][l][n] * b[0][l][n]) % self.q[l] for n in range(self.N)] for l in range(Lq)]
        d2 = [[a[1][l][n] * b[1][l][n] % self.q[l] for n in range(self.N)] for l in range(Lq)]
        ks = self.keyswitch(d2, evk, t_plain)
        return [[[(x + y) % self.q[l] for x, y in zip(d[l], ks[comp][l])] for l in range(Lq)] for comp, d in enumerate((d0, d1))]

    def rotate(self, ct, g, gk, t_plain):
        Lq = self.Lq
        sig = [[self.evaluate(self.automorphism(self.interpolate(ct[comp][l], l), g, self.q[l]), l) for l in range(Lq)] for comp in range(2)]
        ks = self.keyswitch(sig[1], gk, t_plain)
        return [[[(x + y) % self.q[l] for x, y in zip(sig[0][l], ks[0][l])] for l in range(Lq)], ks[1]]

    def rotate_hoisted(self, ct, g, gk, t_plain):
        """the automorphism applied to the LIFT of the unrotated digits (integer vectors, signs and all), then lifted limb by limb"""
        Lq = self.Lq
        vec = [[(-c if neg else c) for c, neg in self._signed(v, g)] for v in self.digit_vectors(ct[1])]
        sig = [[self.evaluate(self.automorphism(self.interpolate(ct[comp][l], l), g, self.q[l]), l) for l in range(Lq)] for comp in range(2)]
        ks = self.keyswitch(sig[1], gk, t_plain, vectors=vec)
        return [[[(x + y) % self.q[l] for x, y in zip(sig[0][l], ks[0][l])] for l in range(Lq)], ks[1]]

    def _signed(self, v, g):
        """X -> X^g on an integer coefficient vector: [(coefficient, negated?)] in the new order"""
        out = [None] * self.N
        for k, c in enumerate(v):
            e = k * g % (2 * self.N)
            out[e % self.N] = (c, e >= self.N)
        return out


@pytest.mark.parametrize("log_n,L,K", [(4, 2, 1), (4, 4, 2), (5, 5, 2), (4, 7, 3), (3, 8, 4)])
def test_special_prime_family_from_the_definitions(oracle_mod, log_n, L, K):
    o = oracle_mod.Oracle(log_n, L)
    ring = SpecialPrimeRing(log_n, o.moduli, o.psi, K)
    Lq, N, t = L - K, o.N, 65537
    oq = oracle_mod.Oracle(log_n, Lq, o.moduli[:Lq])
    a = oq.fill_uniform(21, 2).reshape(2, Lq, N)
    b = oq.fill_uniform(22, 2).reshape(2, Lq, N)
    q = np.array(o.moduli[:Lq], dtype=np.uint64)
    a[1, :, 0] = q - 1
    b[0, :, 1] = 0
    dnum = o.grouped_digits(K)
    assert dnum == ring.dnum
    key = o.fill_uniform(23, 2 * dnum).reshape(dnum, 2, L, N)
    x = o.fill_uniform(24, 1)
    for tp in (0, t):
        assert as_lists(o.mod_down_special(K, x, tp)[0]) == ring.divide_by_P(as_lists(x[0]), tp)
        if K == 1:
            assert as_lists(o.mod_switch_down(x, tp)[0]) == ring.divide_by_P(as_lists(x[0]), tp)
        d = a[1]
        c0, c1 = o.keyswitch_grouped(K, d, key, tp)
        want = ring.keyswitch(as_lists(d), as_lists(key), tp)
        assert as_lists(c0) == want[0] and as_lists(c1) == want[1]
        if K == 1:
            h0, h1 = o.keyswitch_hybrid(d, key, tp)
            assert as_lists(h0) == want[0] and as_lists(h1) == want[1]
    assert as_lists(o.ct_mul_relin_grouped(K, a[None], b[None], key, t)[0]) == ring.ct_mul_relin(as_lists(a), as_lists(b), as_lists(key), t)
    for g in (o.galois_elt(1), o.galois_elt(-1), 2 * N - 1):
        assert as_lists(o.rotate_grouped(K, a[None], g, key, t)[0]) == ring.rotate(as_lists(a), g, as_lists(key), t)
        assert as_lists(o.rotate_hoisted_grouped(K, a[None], [g], key[None], t)[0, 0]) == ring.rotate_hoisted(as_lists(a), g, as_lists(key), t)
