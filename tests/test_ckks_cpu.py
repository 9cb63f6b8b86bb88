"""CKKS-style use of the row f-2 ops on the CPU oracle: multiply with hybrid key switching (plain rounding, t = 0),
then rescale = mod_switch_down(t = 0), and decode the approximate slot-wise product.  This is the scheme-level meaning of
"rescale" that DESIGN.md 2.9 claims for t = 0 (the BGV meaning, t > 0, is covered in test_oracle_kat.py)."""
import numpy as np

import ckks


def test_ckks_multiply_rescale_rotate(oracle_mod):
    n, logn = 256, 8
    o4 = oracle_mod.Oracle(logn, 4)                       # three ciphertext limbs + special prime
    o3 = oracle_mod.Oracle(logn, 3, o4.moduli[:3])
    o2 = oracle_mod.Oracle(logn, 2, o4.moduli[:2])
    s4 = o4.keygen_secret(5)
    s3, s2 = np.ascontiguousarray(s4[:3]), np.ascontiguousarray(s4[:2])
    rng = np.random.default_rng(11)
    slots = np.arange(n // 2)
    z1 = rng.uniform(-1, 1, n // 2) + 1j * rng.uniform(-1, 1, n // 2)
    z2 = rng.uniform(-1, 1, n // 2) + 1j * rng.uniform(-1, 1, n // 2)
    scale = float(2 ** 50)
    c1 = ckks.encrypt(o3, s3, ckks.encode(z1, slots, n, scale), 21)
    c2 = ckks.encrypt(o3, s3, ckks.encode(z2, slots, n, scale), 22)
    assert np.allclose(ckks.decode(ckks.decrypt_coeffs(o3, s3, c1), slots, n, scale), z1, atol=1e-9)
    evk = o4.keygen_relin_hybrid(23, 1, s4)
    prod = o4.ct_mul_relin_hybrid(c1[None], c2[None], evk, 0)[0]                 # scale^2, three limbs
    got = ckks.decode(ckks.decrypt_coeffs(o3, s3, prod), slots, n, scale * scale)
    assert np.allclose(got, z1 * z2, atol=1e-7)
    low = o3.mod_switch_down(prod, 0)                                             # rescale: divide by q_2
    got = ckks.decode(ckks.decrypt_coeffs(o2, s2, low), slots, n, scale * scale / o3.moduli[2])
    assert np.allclose(got, z1 * z2, atol=1e-7)
    # rotation by one slot with a hybrid Galois key: slot j of the result is slot j+1 of the input
    g = o4.galois_elt(1)
    rot = o4.rotate_hybrid(c1[None], g, o4.keygen_galois_hybrid(24, 1, s4, g), 0)[0]
    got = ckks.decode(ckks.decrypt_coeffs(o3, s3, rot), slots, n, scale)
    assert np.allclose(got, np.roll(z1, -1), atol=1e-7)
