"""CPU check of the slot encoder used by the config-4 test and of the rotation semantics it relies on:
Dec(rotate(Enc(slots), 5^k)) == slots rolled left by k in both rows."""
import numpy as np

from slots import SlotEncoder


def test_slot_roundtrip_and_rotation(oracle_mod):
    n, t = 1 << 10, 167772161
    o = oracle_mod.Oracle(10, 3)
    enc = SlotEncoder(n, t)
    rng = np.random.default_rng(1)
    sl = rng.integers(-100, 100, (2, n // 2))
    m = enc.encode(sl)
    assert np.array_equal(enc.decode(m).astype(np.int64), sl % t)
    s = o.keygen_secret(1)
    ct = o.encrypt(3, t, s, m)
    for k in (1, 3, -2):
        g = o.galois_elt(k)
        gk = o.keygen_galois(2, t, s, g)
        dec = enc.decode(o.decrypt(s, o.rotate(ct[None], g, gk)[0], t)).astype(np.int64)
        assert np.array_equal(dec, np.roll(sl % t, -k, axis=1))
    # slot-wise product: ct x pt multiplies slots
    w = rng.integers(-50, 50, (2, n // 2))
    pc = enc.encode(w).astype(np.int64)
    pc = np.where(pc > t // 2, pc - t, pc)
    pt = o.ntt_fwd(np.stack([(pc % q).astype(np.uint64) for q in o.moduli])[None])[0]
    dec = enc.decode(o.decrypt(s, o.ct_mul_plain(ct[None], pt)[0], t)).astype(np.int64)
    assert np.array_equal(dec, (sl * w) % t)
