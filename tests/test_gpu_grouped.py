"""Grouped hybrid key switching on the GPU (digits of K limbs, K special primes; DESIGN.md section 2.11, SURVEY.md section 8
row f-2 widening): ks_grouped_kernel through the C ABI, bit-exact against oracle/dpfhe_oracle.c, plus the scheme-level check
that the GPU result decrypts to the product with far less noise than per-limb digits."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from test_gpu_parity import ctxs, dev, dp, host  # noqa: E402,F401  (ctxs and dp are fixtures)
from test_grouped_cpu import crt  # noqa: E402
from test_oracle_kat import negacyclic_mod_t  # noqa: E402


def grouped_inputs(o, K, batch, seed):
    """[batch][2][L-K][N] residues (with edge rows) under the first L-K moduli, and a uniform grouped key"""
    Lq = o.L - K
    x = o.fill_uniform(seed, 2 * batch)[:, :Lq].reshape(batch, 2, Lq, o.N).copy()
    q = np.array(o.moduli[:Lq], dtype=np.uint64)
    x[0, 0] = (q - 1)[:, None]
    x[0, 1, :, ::2] = 0
    dnum = o.grouped_digits(K)
    key = o.fill_uniform(seed + 1, 2 * dnum).reshape(dnum, 2, o.L, o.N)
    return x, key


@pytest.mark.parametrize("log_n,L,K,batch,t", [(12, 4, 2, 3, 65537), (12, 6, 2, 7, 65537), (13, 6, 2, 4, 65537), (13, 6, 2, 131, 0),
                                               (13, 5, 2, 5, 65537), (14, 6, 2, 3, 65537), (12, 10, 3, 2, 65537), (12, 12, 4, 2, 0),
                                               (12, 16, 4, 2, 65537), (13, 4, 1, 3, 65537)])
def test_grouped_keyswitch_family(ctxs, log_n, L, K, batch, t):
    c, o = ctxs(log_n, L)
    Lq = L - K
    assert c.grouped_digits(K) == o.grouped_digits(K) == -(-Lq // K)
    a, key = grouped_inputs(o, K, batch, 81)
    b, _ = grouped_inputs(o, K, batch, 83)
    out = torch.full((batch, 2, Lq, o.N), -1, dtype=torch.int64, device="cuda")
    c.ct_mul_relin_grouped(K, dev(a), dev(b), dev(key), out, batch, t)
    assert np.array_equal(host(out).reshape(a.shape), o.ct_mul_relin_grouped(K, a, b, key, t))
    if K == 1:   # one special prime: the same bits as the hybrid kernel
        ref = torch.empty_like(out)
        c.ct_mul_relin_hybrid(dev(a), dev(b), dev(key), ref, batch, t)
        assert torch.equal(out, ref)
    if batch > 16:
        return   # the large ragged batch exercises scheduling; one mode is enough
    g = o.galois_elt(-2)
    c.rotate_grouped(K, dev(a), g, dev(key), out, batch, t)
    assert np.array_equal(host(out).reshape(a.shape), o.rotate_grouped(K, a, g, key, t))
    d = np.ascontiguousarray(a[:, 1])
    c.keyswitch_grouped(K, dev(d), dev(key), out, batch, t)
    got = host(out).reshape(a.shape)
    for k in range(batch):
        c0, c1 = o.keyswitch_grouped(K, d[k], key, t)
        assert np.array_equal(got[k, 0], c0) and np.array_equal(got[k, 1], c1)


def test_grouped_repeated_calls_and_host_entry_points(ctxs):
    """epochs / flags survive back-to-back launches of different kernels; the host-buffer forms run behind the staging pipeline"""
    c, o = ctxs(12, 6)
    K, batch = 2, 300
    a, key = grouped_inputs(o, K, batch, 91)
    b, _ = grouped_inputs(o, K, batch, 93)
    exp = o.ct_mul_relin_grouped(K, a, b, key, 65537)
    out = np.zeros_like(a)
    c.ct_mul_relin_grouped_host(K, a, b, key, out, 65537)
    assert np.array_equal(out, exp)
    da, db, dk = dev(a), dev(b), dev(key)
    dout = torch.zeros((batch, 2, 4, o.N), dtype=torch.int64, device="cuda")
    for _ in range(3):
        c.ct_mul_relin_grouped(K, da, db, dk, dout, batch, 65537)
    assert np.array_equal(host(dout).reshape(a.shape), exp)
    g = o.galois_elt(7)
    c.rotate_grouped_host(K, a[:5], g, key, out[:5], 65537)
    assert np.array_equal(out[:5], o.rotate_grouped(K, a[:5], g, key, 65537))


def test_grouped_semantics_and_noise(ctxs, oracle_mod):
    """Dec(GPU grouped ct x ct) == m1*m2 with far less noise than per-limb digits without special primes"""
    L, K = 6, 2
    c, o = ctxs(12, L)
    Lq = L - K
    oq = oracle_mod.Oracle(12, Lq, o.moduli[:Lq])
    t = 65537
    rng = np.random.default_rng(8)
    s = o.keygen_secret(61)
    sq = np.ascontiguousarray(s[:Lq])
    m1, m2 = (rng.integers(0, t, o.N).astype(np.uint64) for _ in range(2))
    c1, c2 = oq.encrypt(63, t, sq, m1), oq.encrypt(64, t, sq, m2)
    out = torch.zeros((1, 2, Lq, o.N), dtype=torch.int64, device="cuda")
    c.ct_mul_relin_grouped(K, dev(c1[None]), dev(c2[None]), dev(o.keygen_relin_grouped(K, 62, t, s)), out, 1, t)
    prod = host(out).reshape(2, Lq, o.N).copy()
    assert np.array_equal(oq.decrypt(sq, prod, t), negacyclic_mod_t(m1, m2, t))

    def noise_bits(ct):
        ph = oq.phase(sq, ct)
        worst = 0
        for n in range(0, o.N, 61):
            v, Q = crt(oq, ph, n, range(Lq))
            worst = max(worst, min(v, Q - v))
        return worst.bit_length()

    bv = oq.ct_mul_relin(c1[None], c2[None], oq.keygen_relin(62, t, sq))[0]
    assert noise_bits(prod) + 30 < noise_bits(bv)


@pytest.mark.parametrize("log_n,L,K,t", [(12, 4, 2, 65537), (13, 6, 2, 0), (14, 3, 2, 65537), (12, 7, 4, 65537), (13, 3, 1, 65537)])
def test_mod_down_special(ctxs, log_n, L, K, t):
    """the division by P on its own (md_tau_kernel + md_limb_kernel), device and host-buffer forms"""
    c, o = ctxs(log_n, L)
    n = 5
    x = o.fill_uniform(41, n)
    x[0, :, ::3] = 0
    x[1] = (np.array(o.moduli, dtype=np.uint64) - 1)[:, None]
    exp = o.mod_down_special(K, x, t)
    out = torch.full((n, L - K, o.N), -1, dtype=torch.int64, device="cuda")
    c.mod_down_special(K, dev(x), out, n, t)
    assert np.array_equal(host(out).reshape(exp.shape), exp)
    hout = np.zeros_like(exp)
    c.mod_down_special_host(K, x, hout, t)
    assert np.array_equal(hout, exp)
    if K == 1:
        c.mod_switch_down(dev(x), out, n, t)
        assert np.array_equal(host(out).reshape(exp.shape), exp)


@pytest.mark.parametrize("log_n,L,K,batch,t", [(12, 6, 2, 5, 65537), (13, 6, 2, 37, 0), (13, 4, 1, 3, 65537), (14, 4, 2, 2, 65537), (12, 9, 3, 2, 65537),
                                               (12, 8, 4, 3, 65537)])
def test_rotate_hoisted_grouped(ctxs, log_n, L, K, batch, t, monkeypatch):
    """rotations sharing the mod-up (ks_hoistg_kernel, rot_apply_grouped_kernel, md kernels) against the oracle's hoisted definition;
    the small scratch cap makes the large batch run in several chunks"""
    c, o = ctxs(log_n, L)
    if batch > 16:
        monkeypatch.setenv("DPFHE_HOIST_CAP_MB", "16")
    ct, _ = grouped_inputs(o, K, batch, 51)
    galois = [o.galois_elt(1), o.galois_elt(-2), 2 * o.N - 1]
    dnum = o.grouped_digits(K)
    keys = [o.fill_uniform(60 + r, 2 * dnum).reshape(dnum, 2, L, o.N) for r in range(len(galois))]
    out = torch.full((len(galois), batch, 2, L - K, o.N), -1, dtype=torch.int64, device="cuda")
    dkeys = [dev(k) for k in keys]
    c.rotate_hoisted_grouped(K, dev(ct), galois, dkeys, out, batch, t)
    exp = o.rotate_hoisted_grouped(K, ct, galois, np.stack(keys), t)
    assert np.array_equal(host(out).reshape(exp.shape), exp)
    c.rotate_hoisted_grouped(K, dev(ct), galois, dkeys, out, batch, t)      # a second call: epochs, flags and round marks carry over
    assert np.array_equal(host(out).reshape(exp.shape), exp)


def test_rotate_hoisted_grouped_semantics(ctxs, oracle_mod):
    """the GPU's hoisted rotations decrypt to the rotated messages"""
    L, K = 6, 2
    c, o = ctxs(12, L)
    Lq = L - K
    oq = oracle_mod.Oracle(12, Lq, o.moduli[:Lq])
    t = 65537
    rng = np.random.default_rng(11)
    s = o.keygen_secret(71)
    sq = np.ascontiguousarray(s[:Lq])
    m = rng.integers(0, t, o.N).astype(np.uint64)
    ct = oq.encrypt(72, t, sq, m)
    galois = [o.galois_elt(1), o.galois_elt(5)]
    gks = [dev(o.keygen_galois_grouped(K, 80 + r, t, s, g)) for r, g in enumerate(galois)]
    out = torch.zeros((2, 1, 2, Lq, o.N), dtype=torch.int64, device="cuda")
    c.rotate_hoisted_grouped(K, dev(ct[None]), galois, gks, out, 1, t)
    got = host(out).reshape(2, 2, Lq, o.N)
    for r, g in enumerate(galois):
        exp = np.zeros(o.N, dtype=np.uint64)
        for k in range(o.N):
            e = (k * g) % (2 * o.N)
            exp[e % o.N] = m[k] if e < o.N else (t - m[k]) % t
        assert np.array_equal(oq.decrypt(sq, got[r].copy(), t), exp)


def test_grouped_errors(ctxs):
    c, o = ctxs(12, 4)
    x = torch.zeros((1, 2, 2, o.N), dtype=torch.int64, device="cuda")
    k = torch.zeros((1, 2, 4, o.N), dtype=torch.int64, device="cuda")
    for bad in (0, 3, 5):
        with pytest.raises(RuntimeError, match="n_special"):
            c.ct_mul_relin_grouped(bad, x, x.clone(), k, x.clone(), 1)
    with pytest.raises(RuntimeError, match="overlap"):
        c.ct_mul_relin_grouped(2, x, x.clone(), k, x, 1)
    with pytest.raises(RuntimeError, match="below the special prime"):
        c.ct_mul_relin_grouped(2, x, x.clone(), k, x.clone(), 1, 1 << 61)


def test_round_numbering_restarts(oracle_mod, monkeypatch):
    """the flag / mailbox tags restart long before 32-bit round numbers could wrap (here: every few launches); every kernel family
    that uses them keeps producing the oracle's bits across restarts"""
    import deeppowers_b200
    monkeypatch.setenv("DPFHE_EPOCH_LIMIT", "40")
    c = deeppowers_b200.Context(12, 6)
    monkeypatch.delenv("DPFHE_EPOCH_LIMIT")
    o = oracle_mod.Oracle(12, 6)
    batch = 9
    x = o.fill_uniform(7, 2 * batch).reshape(batch, 2, 6, o.N)
    y = o.fill_uniform(8, 2 * batch).reshape(batch, 2, 6, o.N)
    evk = o.fill_uniform(9, 12).reshape(6, 2, 6, o.N)
    exp_bv = o.ct_mul_relin(x, y, evk)
    a, key = grouped_inputs(o, 2, batch, 81)
    b, _ = grouped_inputs(o, 2, batch, 83)
    exp_g = o.ct_mul_relin_grouped(2, a, b, key, 65537)
    h, hkey = a[:, :, :], o.fill_uniform(85, 10).reshape(5, 2, 6, o.N)
    a5 = o.fill_uniform(86, 2 * batch)[:, :5].reshape(batch, 2, 5, o.N).copy()
    exp_h = o.ct_mul_relin_hybrid(a5, a5, hkey, 65537)
    g = o.galois_elt(1)
    exp_r = o.rotate_hoisted_grouped(2, a, [g], key[None], 65537)
    out6 = torch.zeros((batch, 2, 6, o.N), dtype=torch.int64, device="cuda")
    out5 = torch.zeros((batch, 2, 5, o.N), dtype=torch.int64, device="cuda")
    out4 = torch.zeros((batch, 2, 4, o.N), dtype=torch.int64, device="cuda")
    outr = torch.zeros((1, batch, 2, 4, o.N), dtype=torch.int64, device="cuda")
    dx, dy, dk, da, db, dkey, da5, dhk = (dev(v) for v in (x, y, evk, a, b, key, a5, hkey))
    for _ in range(4):   # ~10 rounds per launch against a limit of 40: a restart every few launches, at different places in the cycle
        c.ct_mul_relin(dx, dy, dk, out6, batch)
        assert np.array_equal(host(out6).reshape(x.shape), exp_bv)
        c.ct_mul_relin_grouped(2, da, db, dkey, out4, batch, 65537)
        assert np.array_equal(host(out4).reshape(a.shape), exp_g)
        c.ct_mul_relin_hybrid(da5, da5.clone(), dhk, out5, batch, 65537)
        assert np.array_equal(host(out5).reshape(a5.shape), exp_h)
        c.rotate_hoisted_grouped(2, da, [g], [dkey], outr, batch, 65537)
        assert np.array_equal(host(outr).reshape(exp_r.shape), exp_r)
        c.ct_mul_relin(dx[:3], dy[:3], dk, out6[:3], 3)
        assert np.array_equal(host(out6[:3]).reshape(3, 2, 6, o.N), exp_bv[:3])
    c.close()
