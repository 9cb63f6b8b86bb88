"""BASELINE.json config 4 as a parity/semantics case: an encrypted 768x768 int8 linear layer (the GPT-2-small
projection shape, reference src/core/execution/model.hpp:47, gpt_weights.cpp:187-190; symmetric int8 range
per quantization_manager.cpp:272-276) evaluated with the diagonal method on top of the hot-path ops
ct_mul_plain + rotate + add, N=8192, L=4.  Decrypting the GPU result must give W @ x exactly."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from slots import SlotEncoder  # noqa: E402

T_PLAIN = 167772161      # 5 * 2^25 + 1, prime, = 1 mod 2N, > 2 * 768 * 127 * 127
DIM = 768


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


def to_rns_eval(o, coeffs_mod_t):
    """centred lift of a plaintext polynomial to every limb, then forward NTT (oracle) -> [L][N]"""
    c = coeffs_mod_t.astype(np.int64)
    c = np.where(c > T_PLAIN // 2, c - T_PLAIN, c)
    limbs = np.stack([(c % q).astype(np.uint64) for q in o.moduli])
    return o.ntt_fwd(limbs[None])[0]


def test_encrypted_linear_layer_768(oracle_mod):
    import deeppowers_b200 as dp
    log_n, L, B = 13, 4, 2
    o = oracle_mod.Oracle(log_n, L)
    ctx = dp.Context(log_n, L)
    N = o.N
    enc = SlotEncoder(N, T_PLAIN)
    rng = np.random.default_rng(0xD3390004)
    W = rng.integers(-127, 128, (DIM, DIM))
    X = rng.integers(-127, 128, (B, DIM))
    g = o.galois_elt(1)   # rotate every row of slots left by one (tests/test_slots_cpu.py pins that meaning)
    s = o.keygen_secret(1)
    gk = o.keygen_galois(2, T_PLAIN, s, g)
    # inputs: x duplicated so that slot i+d holds x[(i+d) mod 768] for i, d < 768
    cts = []
    for b in range(B):
        slots = np.zeros((2, N // 2), dtype=np.int64)
        slots[0, :DIM] = X[b]
        slots[0, DIM:2 * DIM] = X[b]
        cts.append(o.encrypt(10 + b, T_PLAIN, s, enc.encode(slots)))
    ct = np.stack(cts)
    assert np.array_equal(enc.decode(o.decrypt(s, ct[0], T_PLAIN))[0, :DIM].astype(np.int64), X[0] % T_PLAIN)
    # weight diagonals, pre-NTT'd plaintexts (768 x 256 KiB = 192 MiB)
    diag = torch.empty((DIM, L, N), dtype=torch.int64, device="cuda")
    ar = np.arange(DIM)
    for d in range(DIM):
        slots = np.zeros((2, N // 2), dtype=np.int64)
        slots[0, :DIM] = W[ar, (ar + d) % DIM]
        diag[d] = dev(to_rns_eval(o, enc.encode(slots)))
    d_gk = dev(gk)
    cur, nxt = dev(ct), torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    acc, tmp = torch.empty_like(nxt), torch.empty_like(nxt)
    ctx.ct_mul_plain(cur, diag[0], acc, B)
    first_steps = {}
    for d in range(1, DIM):
        ctx.rotate(cur, g, d_gk, nxt, B)
        cur, nxt = nxt, cur
        ctx.ct_mul_plain(cur, diag[d], tmp, B)
        ctx.poly_add(acc, tmp, acc, 2 * B)
        if d <= 2:
            first_steps[d] = (host(cur).reshape(B, 2, L, N).copy(), host(acc).reshape(B, 2, L, N).copy())
    torch.cuda.synchronize()
    # (1) bit-exact against the oracle running the same first steps
    r = ct
    a = o.ct_mul_plain(ct, host(diag[0]).reshape(L, N))
    for d in (1, 2):
        r = o.rotate(r, g, gk)
        a = o.poly_add(a, o.ct_mul_plain(r, host(diag[d]).reshape(L, N)))
        assert np.array_equal(first_steps[d][0], r)
        assert np.array_equal(first_steps[d][1], a)
    # (2) semantics of the whole chain: Dec(result) slots == W @ x
    res = host(acc).reshape(B, 2, L, N)
    for b in range(B):
        y = enc.decode(o.decrypt(s, res[b], T_PLAIN))[0, :DIM].astype(np.int64)
        y = np.where(y > T_PLAIN // 2, y - T_PLAIN, y)
        assert np.array_equal(y, W @ X[b])
    ctx.close()


def test_encrypted_linear_layer_bsgs(oracle_mod):
    """same layer through Context.linear_bsgs: 31 baby + 23 giant rotations instead of 767, fused multiply-accumulate"""
    import deeppowers_b200 as dp
    log_n, L, B, BABY = 13, 4, 2, 32
    o = oracle_mod.Oracle(log_n, L)
    ctx = dp.Context(log_n, L)
    N = o.N
    enc = SlotEncoder(N, T_PLAIN)
    rng = np.random.default_rng(0xD3390044)
    W = rng.integers(-127, 128, (DIM, DIM))
    X = rng.integers(-127, 128, (B, DIM))
    s = o.keygen_secret(1)
    gk1 = o.keygen_galois(2, T_PLAIN, s, o.galois_elt(1))
    gkb = o.keygen_galois(3, T_PLAIN, s, o.galois_elt(BABY))
    cts = []
    for b in range(B):
        slots = np.zeros((2, N // 2), dtype=np.int64)
        slots[0, :DIM] = X[b]
        slots[0, DIM:2 * DIM] = X[b]
        cts.append(o.encrypt(10 + b, T_PLAIN, s, enc.encode(slots)))
    ct = np.stack(cts)
    diag = torch.empty((DIM, L, N), dtype=torch.int64, device="cuda")
    ar = np.arange(DIM)
    for d in range(DIM):
        g = d // BABY
        slots = np.zeros((2, N // 2), dtype=np.int64)
        slots[0, :DIM] = W[ar, (ar + d) % DIM]
        slots = np.roll(slots, g * BABY, axis=1)          # D_{g,b} = rot_{-g*baby}(diag_d)
        diag[d] = dev(to_rns_eval(o, enc.encode(slots)))
    out = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    n0 = ctx.launch_count()
    ctx.linear_bsgs(dev(ct), diag, dev(gk1), dev(gkb), BABY, out, B)                 # fused inner products (default)
    torch.cuda.synchronize()
    n_rot = (BABY - 1) + (DIM // BABY - 1)
    assert ctx.launch_count() - n0 == 2 * n_rot + 1 + (DIM // BABY - 1)     # rotations (2 launches each), one inner-product launch, adds
    ref = torch.empty_like(out)
    n0 = ctx.launch_count()
    ctx.linear_bsgs(dev(ct), diag, dev(gk1), dev(gkb), BABY, ref, B, fused=False)    # composition of ct_mul_plain(_acc)
    torch.cuda.synchronize()
    assert ctx.launch_count() - n0 == 2 * n_rot + DIM + (DIM // BABY - 1)
    assert torch.equal(out, ref)                                                     # same bits either way
    # baby steps as hoisted rotations of the input (one key per step): different keys, same plaintext result
    baby_keys = [dev(o.keygen_galois(100 + b, T_PLAIN, s, o.galois_elt(b))) for b in range(1, BABY)]
    hoisted = torch.empty_like(out)
    ctx.linear_bsgs(dev(ct), diag, baby_keys, dev(gkb), BABY, hoisted, B)
    res_h = host(hoisted).reshape(B, 2, L, N)
    for b in range(B):
        y = enc.decode(o.decrypt(s, res_h[b], T_PLAIN))[0, :DIM].astype(np.int64)
        assert np.array_equal(np.where(y > T_PLAIN // 2, y - T_PLAIN, y), W @ X[b])
    res = host(out).reshape(B, 2, L, N)
    for b in range(B):
        y = enc.decode(o.decrypt(s, res[b], T_PLAIN))[0, :DIM].astype(np.int64)
        y = np.where(y > T_PLAIN // 2, y - T_PLAIN, y)
        assert np.array_equal(y, W @ X[b])
    # the fused multiply-accumulate itself, bit-exact against the oracle
    pt = host(diag[5]).reshape(L, N)
    acc = o.fill_uniform(77, 2 * B).reshape(B, 2, L, N)
    d_acc = dev(acc)
    ctx.ct_mul_plain_acc(dev(ct), diag[5], d_acc, B)
    assert np.array_equal(host(d_acc).reshape(acc.shape), o.poly_add(acc, o.ct_mul_plain(ct, pt)))
    ctx.close()


@pytest.mark.parametrize("log_n,L,nb,ng,batch", [(12, 2, 5, 3, 3), (12, 1, 33, 29, 2), (13, 4, 32, 24, 2), (14, 2, 16, 9, 1), (12, 3, 128, 5, 1)])
def test_plain_inner_products(oracle_mod, log_n, L, nb, ng, batch):
    """dpfhe_ct_mul_plain_inner against the oracle: ragged sizes, several giant-step blocks, worst-case residues"""
    import deeppowers_b200 as dp
    o = oracle_mod.Oracle(log_n, L)
    ctx = dp.Context(log_n, L)
    steps = o.fill_uniform(51, nb * batch * 2).reshape(nb, batch, 2, L, o.N)
    pts = o.fill_uniform(52, ng * nb).reshape(ng, nb, L, o.N)
    q = np.array(o.moduli, dtype=np.uint64)
    steps[:, 0, 0] = (q - 1)[:, None]
    pts[0] = (q - 1)[:, None]
    pts[-1, :, :, ::3] = 0
    out = torch.full((ng, batch, 2, L, o.N), -1, dtype=torch.int64, device="cuda")
    n0 = ctx.launch_count()
    ctx.ct_mul_plain_inner(dev(steps), dev(pts), out, nb, ng, batch)
    assert np.array_equal(host(out).reshape(ng, batch, 2, L, o.N), o.ct_mul_plain_inner(steps, pts))
    assert ctx.launch_count() - n0 >= 1
    with pytest.raises(RuntimeError, match="n_steps"):
        ctx.ct_mul_plain_inner(dev(steps), dev(pts), out, 129, ng, batch)
    ctx.close()


def test_encrypted_linear_layer_special_prime_keys(oracle_mod):
    """the same 768x768 layer with grouped special-prime Galois keys (two special primes, digits of two limbs): hoisted baby steps
    (one mod-up of the input), fused inner products on the ciphertext moduli, giant steps with rotate_grouped.  Decrypts to W @ x,
    with far less noise than per-limb-digit keys leave."""
    import deeppowers_b200 as dp
    log_n, Lq, K, B, BABY = 13, 4, 2, 2, 32
    L = Lq + K
    o = oracle_mod.Oracle(log_n, L)
    oq = oracle_mod.Oracle(log_n, Lq, o.moduli[:Lq])
    ctx, ctx_q = dp.Context(log_n, L), dp.Context(log_n, Lq, o.moduli[:Lq])
    N = o.N
    enc = SlotEncoder(N, T_PLAIN)
    rng = np.random.default_rng(0xD3390045)
    W = rng.integers(-127, 128, (DIM, DIM))
    X = rng.integers(-127, 128, (B, DIM))
    s = o.keygen_secret(1)
    sq = np.ascontiguousarray(s[:Lq])
    cts = []
    for b in range(B):
        slots = np.zeros((2, N // 2), dtype=np.int64)
        slots[0, :DIM] = X[b]
        slots[0, DIM:2 * DIM] = X[b]
        cts.append(oq.encrypt(10 + b, T_PLAIN, sq, enc.encode(slots)))
    ct = np.stack(cts)
    diag = torch.empty((DIM, Lq, N), dtype=torch.int64, device="cuda")
    ar = np.arange(DIM)
    for d in range(DIM):
        g = d // BABY
        slots = np.zeros((2, N // 2), dtype=np.int64)
        slots[0, :DIM] = W[ar, (ar + d) % DIM]
        slots = np.roll(slots, g * BABY, axis=1)
        diag[d] = dev(to_rns_eval(oq, enc.encode(slots)))
    baby_keys = [dev(o.keygen_galois_grouped(K, 100 + b, T_PLAIN, s, o.galois_elt(b))) for b in range(1, BABY)]
    gkb = dev(o.keygen_galois_grouped(K, 3, T_PLAIN, s, o.galois_elt(BABY)))
    out = torch.empty((B, 2, Lq, N), dtype=torch.int64, device="cuda")
    dp.linear_bsgs_grouped(ctx, ctx_q, K, dev(ct), diag, baby_keys, gkb, BABY, out, B, T_PLAIN)
    res = host(out).reshape(B, 2, Lq, N)

    def noise_bits(o_, s_, c):
        ph = o_.phase(s_, c)
        Q = 1
        for q in o_.moduli:
            Q *= q
        coef = [(Q // q) * pow(Q // q, -1, q) for q in o_.moduli]
        worst = 0
        for n in range(0, N, 257):
            v = sum(int(ph[l][n]) * coef[l] for l in range(len(o_.moduli))) % Q
            worst = max(worst, min(v, Q - v))
        return worst.bit_length()

    for b in range(B):
        y = enc.decode(oq.decrypt(sq, res[b], T_PLAIN))[0, :DIM].astype(np.int64)
        assert np.array_equal(np.where(y > T_PLAIN // 2, y - T_PLAIN, y), W @ X[b])
    # the per-limb-digit layer on the same ciphertext moduli, for the noise comparison
    bv_keys = [dev(oq.keygen_galois(100 + b, T_PLAIN, sq, oq.galois_elt(b))) for b in range(1, BABY)]
    bv_out = torch.empty_like(out)
    ctx_q.linear_bsgs(dev(ct), diag, bv_keys, dev(oq.keygen_galois(3, T_PLAIN, sq, oq.galois_elt(BABY))), BABY, bv_out, B)
    bv = host(bv_out).reshape(B, 2, Lq, N)
    assert noise_bits(oq, sq, res[0]) + 20 < noise_bits(oq, sq, bv[0])
    ctx.close()
    ctx_q.close()
