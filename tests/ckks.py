"""CKKS-style approximate-number helpers for the scheme-level tests of rescale / hybrid key switching
(numpy + Python integers; test infrastructure only).

Slot j of a polynomial m is m(zeta_j) with zeta_j = exp(i*pi*5^j / N), j < N/2; the other N/2 evaluation points
are the complex conjugates, so real-coefficient polynomials carry N/2 complex slots and multiply slot-wise."""
import numpy as np


def slot_roots(n, slots):
    e = np.array([pow(5, int(j), 2 * n) for j in slots], dtype=np.float64)
    return np.exp(1j * np.pi * e / n)


def encode(z, slots, n, scale):
    """integer coefficients (Python ints) of the real polynomial with m(zeta_j) = scale * z_j on `slots`, 0 elsewhere"""
    roots = slot_roots(n, slots)
    k = np.arange(n)
    powers = np.conj(roots)[:, None] ** k[None, :]                     # conj(zeta_j)^k
    coeffs = (2.0 / n) * np.real((np.asarray(z)[:, None] * powers).sum(axis=0))
    return [int(round(float(c) * scale)) for c in coeffs]


def decode(coeffs, slots, n, scale):
    roots = slot_roots(n, slots)
    k = np.arange(n)
    c = np.array([float(v) for v in coeffs])
    return (roots[:, None] ** k[None, :] @ c) / scale


def to_rns_eval(o, coeffs):
    """signed integer coefficients -> [L][N] evaluation form under oracle context o"""
    res = np.array([[c % q for c in coeffs] for q in o.moduli], dtype=np.uint64)
    return o.ntt_fwd(res[None])[0]


def encrypt(o, s, coeffs, seed):
    """(-a*s + e + m, a): the oracle's encryption with t = 1 (unscaled noise) of zero, plus the encoded message"""
    ct = o.encrypt(seed, 1, s, np.zeros(o.N, dtype=np.uint64))
    ct[0] = o.poly_add(ct[0][None], to_rns_eval(o, coeffs)[None])[0]
    return ct


def decrypt_coeffs(o, s, ct):
    """centred integer coefficients of c0 + c1*s"""
    ph = o.phase(s, ct)
    Q = 1
    for q in o.moduli:
        Q *= q
    coef = [(Q // q) * pow(Q // q, -1, q) for q in o.moduli]
    out = []
    for n in range(o.N):
        v = sum(int(ph[l][n]) * coef[l] for l in range(o.L)) % Q
        out.append(v - Q if v > Q // 2 else v)
    return out
