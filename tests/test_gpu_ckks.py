"""CKKS-style pipeline on the GPU: hybrid ct x ct (plain rounding) -> rescale (mod_switch_down, t = 0) -> hybrid rotate,
decoded on the host to the approximate slot-wise product.  Scheme-level check of SURVEY.md section 8 row f-2 through the C ABI."""
import numpy as np
import pytest

import ckks

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint64)


def test_ckks_multiply_rescale_rotate_on_gpu(oracle_mod):
    import deeppowers_b200 as dp
    logn, n = 12, 4096
    o4 = oracle_mod.Oracle(logn, 4)
    o3 = oracle_mod.Oracle(logn, 3, o4.moduli[:3])
    o2 = oracle_mod.Oracle(logn, 2, o4.moduli[:2])
    c4, c3 = dp.Context(logn, 4), dp.Context(logn, 3, o4.moduli[:3])
    s4 = o4.keygen_secret(5)
    s3, s2 = np.ascontiguousarray(s4[:3]), np.ascontiguousarray(s4[:2])
    rng = np.random.default_rng(12)
    slots = np.array([0, 1, 2, 3, 100, 1000, 2046, 2047])   # a few occupied slots keep the host-side encoding cheap
    z1 = rng.uniform(-1, 1, 8) + 1j * rng.uniform(-1, 1, 8)
    z2 = rng.uniform(-1, 1, 8) + 1j * rng.uniform(-1, 1, 8)
    scale = float(2 ** 50)
    ct1 = ckks.encrypt(o3, s3, ckks.encode(z1, slots, n, scale), 21)
    ct2 = ckks.encrypt(o3, s3, ckks.encode(z2, slots, n, scale), 22)
    prod = torch.zeros((1, 2, 3, n), dtype=torch.int64, device="cuda")
    c4.ct_mul_relin_hybrid(dev(ct1[None]), dev(ct2[None]), dev(o4.keygen_relin_hybrid(23, 1, s4)), prod, 1, 0)
    low = torch.zeros((2, 2, n), dtype=torch.int64, device="cuda")
    c3.mod_switch_down(prod, low, 2, 0)                      # rescale by q_2
    got = ckks.decode(ckks.decrypt_coeffs(o2, s2, host(low).reshape(2, 2, n)), slots, n, scale * scale / o3.moduli[2])
    assert np.allclose(got, z1 * z2, atol=1e-6)
    g = o4.galois_elt(1)
    rot = torch.zeros_like(prod)
    c4.rotate_hybrid(dev(ct1[None]), g, dev(o4.keygen_galois_hybrid(24, 1, s4, g)), rot, 1, 0)
    moved = ckks.decode(ckks.decrypt_coeffs(o3, s3, host(rot).reshape(2, 3, n)), np.array([0, 1, 2, 99, 999, 2046]), n, scale)
    assert np.allclose(moved, z1[[1, 2, 3, 4, 5, 7]], atol=1e-6)   # slot j of the result is slot j+1 of the input
    c4.close()
    c3.close()
