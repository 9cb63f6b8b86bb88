"""A two-level BGV computation on the CPU oracle: multiply (hybrid key switching) -> modulus switch -> multiply again at
the lower level with that level's own context and keys -> decrypt.  This is how the pieces of SURVEY.md section 8 row f-2 compose
into a modulus chain: a level is a context over a prefix of the ciphertext moduli plus the special prime."""
import numpy as np

from test_oracle_kat import negacyclic_mod_t


def build_chain(oracle_mod, logn=10):
    top = oracle_mod.Oracle(logn, 4)                                   # q0 q1 q2 | p
    q0, q1, q2, p = top.moduli
    return {
        "top": top,                                                    # hybrid context of level 3
        "l3": oracle_mod.Oracle(logn, 3, [q0, q1, q2]),                # ciphertext moduli of level 3
        "low": oracle_mod.Oracle(logn, 3, [q0, q1, p]),                # hybrid context of level 2
        "l2": oracle_mod.Oracle(logn, 2, [q0, q1]),                    # ciphertext moduli of level 2
    }


def test_two_level_pipeline(oracle_mod):
    ch = build_chain(oracle_mod)
    top, l3, low, l2 = ch["top"], ch["l3"], ch["low"], ch["l2"]
    t = 65537
    rng = np.random.default_rng(21)
    s_top = top.keygen_secret(31)
    s3, s2 = l3.keygen_secret(31), l2.keygen_secret(31)              # the same ternary secret under each modulus set
    s_low = low.keygen_secret(31)
    assert np.array_equal(s3, s_top[:3]) and np.array_equal(s2, s_top[:2]) and np.array_equal(s_low[2], s_top[3])
    m1, m2, m3 = (rng.integers(0, t, top.N).astype(np.uint64) for _ in range(3))
    c1, c2 = l3.encrypt(41, t, s3, m1), l3.encrypt(42, t, s3, m2)
    prod = top.ct_mul_relin_hybrid(c1[None], c2[None], top.keygen_relin_hybrid(43, t, s_top), t)[0]    # level 3
    down = l3.mod_switch_down(prod, t)                                                                  # level 2, message * q2^-1
    c3 = l2.encrypt(44, t, s2, m3)
    prod2 = low.ct_mul_relin_hybrid(down[None], c3[None], low.keygen_relin_hybrid(45, t, s_low), t)[0]  # level 2
    scale = pow(top.moduli[2], -1, t)
    m12 = (negacyclic_mod_t(m1, m2, t).astype(object) * scale % t).astype(np.uint64)
    assert np.array_equal(l2.decrypt(s2, down.reshape(2, 2, top.N), t), m12)
    assert np.array_equal(l2.decrypt(s2, prod2, t), negacyclic_mod_t(m12, m3, t))
