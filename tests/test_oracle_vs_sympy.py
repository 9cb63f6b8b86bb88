"""Independent pin of the oracle against sympy (a third-party implementation shipped in this image, unrelated to
both the reference and this repo): primality / maximality of the derived moduli, the order of psi, and the
negacyclic product (NTT -> pointwise -> INTT) against sympy's own NTT-based convolution.  The reference itself offers
nothing to pin against for this path, so this is the strongest external check available."""
import numpy as np
import pytest

sympy = pytest.importorskip("sympy")
from sympy.discrete.convolutions import convolution_ntt  # noqa: E402
from sympy.ntheory import isprime, n_order  # noqa: E402


def test_moduli_are_the_largest_ntt_primes(oracle_mod):
    """the default basis is the L largest primes below 2^60 of the form k * 2^32 + 1 (so 1 mod 2N for every N)"""
    for log_n, L in ((12, 3), (13, 4), (14, 8)):
        o = oracle_mod.Oracle(log_n, L)
        found, cand = [], (1 << 60) + 1
        while len(found) < L:
            cand -= 1 << 32
            if isprime(cand):
                found.append(cand)
        assert found == o.moduli
        assert all(q % (2 << log_n) == 1 for q in o.moduli)


def test_psi_has_order_2n_and_is_minimal(oracle_mod):
    o = oracle_mod.Oracle(10, 2)
    two_n = 2 << 10
    for q, psi in zip(o.moduli, o.psi):
        assert n_order(psi, q) == two_n
        # every primitive 2N-th root is psi^k for odd k: psi is the smallest of them
        roots = [pow(psi, k, q) for k in range(1, two_n, 2)]
        assert psi == min(roots)


@pytest.mark.parametrize("log_n", [6, 10])
def test_negacyclic_product_matches_sympy_convolution(oracle_mod, log_n):
    o = oracle_mod.Oracle(log_n, 2)
    n = o.N
    a = o.fill_uniform(21, 1)[0]
    b = o.fill_uniform(22, 1)[0]
    prod = o.ntt_inv(o.poly_mul_pointwise(o.ntt_fwd(a), o.ntt_fwd(b)))
    for l, q in enumerate(o.moduli):
        lin = convolution_ntt([int(v) for v in a[l]], [int(v) for v in b[l]], prime=q)   # linear convolution mod q
        lin = list(lin) + [0] * (2 * n - len(lin))
        neg = [(lin[k] - lin[k + n]) % q for k in range(n)]
        assert [int(v) for v in prod[l]] == neg
