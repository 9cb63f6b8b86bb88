"""Throughput of the other BASELINE.json configs (parity-test shapes, not the driver's bench line):
  cfg3  key-switch rotation sweep, N=16384, L=8, batch=1024, k in {+-2^0..+-2^12} (index-outer, batch-inner)
  cfg4  encrypted 768x768 linear layer, N=8192, L=4, batch=512: 768 ct x pt + 767 rotations + 767 adds per batch
Kernel-only CUDA-event timings on synthetic residues; writes one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import deeppowers_b200 as dp

PEAK = 6564.8
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def cfg3():
    log_n, L, B = 14, 8, 1024
    c = dp.Context(log_n, L)
    N = 1 << log_n
    ct = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    out = torch.empty_like(ct)
    c.fill_uniform(0xD3390003, ct, 2 * B)
    ks = [s * (1 << e) for e in range(13) for s in (1, -1)]
    keys = torch.empty((len(ks), L, 2, L, N), dtype=torch.int64, device="cuda")     # 26 x 16 MiB
    c.fill_uniform(0xD3390103, keys, len(ks) * 2 * L)

    def sweep():
        for i, k in enumerate(ks):
            c.rotate(ct, c.galois_elt(k), keys[i], out, B)

    ms = timed(sweep, 2)
    n = len(ks) * B
    P = L * N * 8
    # the same sweep with the digit decomposition shared by the 26 rotations of each ciphertext (dpfhe_rotate_hoisted),
    # in sub-batches of 512 so that the 26 outputs fit a 26 GiB buffer
    SUB = 512
    outs = torch.empty((len(ks), SUB, 2, L, N), dtype=torch.int64, device="cuda")
    galois = [c.galois_elt(k) for k in ks]
    klist = [keys[i] for i in range(len(ks))]

    def sweep_hoisted():
        for first in range(0, B, SUB):
            c.rotate_hoisted(ct[first:first + SUB], galois, klist, outs, SUB)

    ms_h = timed(sweep_hoisted, 2)
    c.rotate(ct[B - SUB:], galois[3], keys[3], out[:SUB], SUB)      # `outs` holds the last sub-batch of the sweep
    same = bool(torch.equal(outs[3], out[:SUB]))
    c.close()
    # per rotation: read the ciphertext and L(L-1) shared transforms, write the ciphertext
    hbytes = (4 + L - 1) * P
    return {"config": "cfg3 rotation sweep N=16384 L=8 batch=1024 x 26 indices", "ms_per_sweep": ms, "rotations_per_s": n / ms * 1e3,
            "roofline": {"bound": "hbm", "achieved": n * 4 * P / ms / 1e6, "peak": PEAK, "unit": "GB/s", "frac": n * 4 * P / ms / 1e6 / PEAK,
                         "algorithmic_bytes_per_rotation": 4 * P},
            "hoisted": {"ms_per_sweep": ms_h, "rotations_per_s": n / ms_h * 1e3, "matches_independent_rotations": same,
                        "GBps_incl_shared_transforms": n * hbytes / ms_h / 1e6, "frac_hbm": n * hbytes / ms_h / 1e6 / PEAK,
                        "GBps_ciphertext_only": n * 4 * P / ms_h / 1e6,
                        "note": "26 rotations of each ciphertext share one digit decomposition (64 transforms) instead of 26 x 64"}}


def cfg4():
    log_n, L, B, DIM = 13, 4, 512, 768
    c = dp.Context(log_n, L)
    N = 1 << log_n
    cur = torch.empty((B, 2, L, N), dtype=torch.int64, device="cuda")
    nxt, acc, tmp = torch.empty_like(cur), torch.empty_like(cur), torch.empty_like(cur)
    c.fill_uniform(0xD3390004, cur, 2 * B)
    diag = torch.empty((DIM, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(0xD3390104, diag, DIM)
    gk = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(0xD3390204, gk, 2 * L)
    g = c.galois_elt(1)
    P = L * N * 8
    t_pt = timed(lambda: c.ct_mul_plain(cur, diag[1], tmp, B), 20)
    t_rot = timed(lambda: c.rotate(cur, g, gk, nxt, B), 5)
    t_add = timed(lambda: c.poly_add(acc, tmp, acc, 2 * B), 20)

    def layer():
        a, b = cur, nxt
        c.ct_mul_plain(a, diag[0], acc, B)
        for d in range(1, DIM):
            c.rotate(a, g, gk, b, B)
            a, b = b, a
            c.ct_mul_plain(a, diag[d], tmp, B)
            c.poly_add(acc, tmp, acc, 2 * B)

    ms = timed(layer, 1)
    # baby-step/giant-step variant with the fused multiply-accumulate (31 + 23 rotations)
    BABY = 32
    gkb = torch.empty((L, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(0xD3390304, gkb, 2 * L)
    GIANT = DIM // BABY
    scratch = torch.empty((BABY + GIANT + 1, B, 2, L, N), dtype=torch.int64, device="cuda")
    ms_bsgs_unfused = timed(lambda: c.linear_bsgs(cur, diag, gk, gkb, BABY, acc, B, scratch=scratch, fused=False), 2)
    ms_bsgs = timed(lambda: c.linear_bsgs(cur, diag, gk, gkb, BABY, acc, B, scratch=scratch), 2)
    baby_keys = torch.empty((BABY - 1, L, 2, L, N), dtype=torch.int64, device="cuda")
    c.fill_uniform(0xD3390404, baby_keys, (BABY - 1) * 2 * L)
    ms_bsgs_h = timed(lambda: c.linear_bsgs(cur, diag, [baby_keys[i] for i in range(BABY - 1)], gkb, BABY, acc, B, scratch=scratch), 2)
    t_inner = timed(lambda: c.ct_mul_plain_inner(scratch[:BABY], diag, scratch[BABY:BABY + GIANT], BABY, GIANT, B), 3)
    t_fma = timed(lambda: c.ct_mul_plain_acc(cur, diag[1], acc, B), 20)
    # the layer as a library object (dpfhe_linear_*): weights and keys resident, one C call per application; and end to end
    # from HOST buffers (pinned, on the GPU's NUMA node), chunks of the batch pipelined through upload / compute / download
    import time
    import numpy as np
    h = lambda t: t.cpu().numpy().view(np.uint64)
    del scratch
    lay = dp.LinearLayer(c, h(diag), BABY, h(baby_keys), h(gkb))
    out_lib = torch.empty_like(cur)
    ms_lib = timed(lambda: lay.apply(cur, out_lib, B), 3)
    scratch = torch.empty((BABY + GIANT + 1, B, 2, L, N), dtype=torch.int64, device="cuda")
    c.linear_bsgs(cur, diag, [baby_keys[i] for i in range(BABY - 1)], gkb, BABY, acc, B, scratch=scratch)
    torch.cuda.synchronize()
    same_lib = bool(torch.equal(out_lib, acc))      # the Python composition of the same calls
    del scratch
    n_words = B * 2 * L * N
    hin, hout = c.pinned_near(n_words), c.pinned_near(n_words)
    torch.from_numpy(hin.array.view(np.int64)).copy_(cur.view(-1))
    torch.cuda.synchronize()
    lay.apply_host(hin.array, hout.array)
    t0 = time.perf_counter()
    for _ in range(3):
        lay.apply_host(hin.array, hout.array)
    s_host = (time.perf_counter() - t0) / 3
    same_host = bool(torch.equal(torch.from_numpy(hout.array.view(np.int64)).cuda().view_as(cur), out_lib))
    lay.close()
    c.close()
    return {"config": "cfg4 encrypted 768x768 linear layer N=8192 L=4 batch=512 (diagonal method, one Galois key)",
            "ms_per_layer_batch": ms, "prompts_per_s": B / ms * 1e3,
            "bsgs": {"baby": BABY, "rotations": BABY - 1 + DIM // BABY - 1, "ms_per_layer_batch": ms_bsgs, "prompts_per_s": B / ms_bsgs * 1e3,
                     "unfused_ms_per_layer_batch": ms_bsgs_unfused, "unfused_prompts_per_s": B / ms_bsgs_unfused * 1e3,
                     "hoisted_baby_steps_ms_per_layer_batch": ms_bsgs_h, "hoisted_baby_steps_prompts_per_s": B / ms_bsgs_h * 1e3},
            "library_layer": {"api": "dpfhe_linear_apply (device buffers, weights and keys resident)", "ms_per_batch": ms_lib,
                              "prompts_per_s": B / ms_lib * 1e3, "matches_python_composition": same_lib},
            "library_layer_host": {"api": "dpfhe_linear_apply_host (pinned host buffers, chunks of the batch pipelined)", "ms_per_batch": s_host * 1e3,
                                   "prompts_per_s": B / s_host, "matches": same_host, "h2d_bytes": n_words * 8, "d2h_bytes": n_words * 8,
                                   "fraction_of_device_resident_rate": (B / s_host) / (B / ms_lib * 1e3)},
            "ct_mul_plain_inner": {"ms": t_inner, "ct_pt_products_per_s": B * DIM / t_inner * 1e3,
                                   "GBps": B * (BABY + GIANT) * 2 * P / t_inner / 1e6, "frac_hbm": B * (BABY + GIANT) * 2 * P / t_inner / 1e6 / PEAK,
                                   "note": "768 ct x pt products per prompt in one launch; traffic = 32 ciphertext rows in + 24 out per prompt"},
            "ct_mul_plain_acc": {"per_s": B / t_fma * 1e3, "GBps": B * 6 * P / t_fma / 1e6, "frac_hbm": B * 6 * P / t_fma / 1e6 / PEAK},
            "ct_mul_plain": {"per_s": B / t_pt * 1e3, "GBps": B * 4 * P / t_pt / 1e6, "frac_hbm": B * 4 * P / t_pt / 1e6 / PEAK},
            "rotate": {"per_s": B / t_rot * 1e3, "GBps": B * 4 * P / t_rot / 1e6, "frac_hbm": B * 4 * P / t_rot / 1e6 / PEAK},
            "ct_add": {"per_s": B / t_add * 1e3, "GBps": B * 6 * P / t_add / 1e6, "frac_hbm": B * 6 * P / t_add / 1e6 / PEAK}}


if __name__ == "__main__":
    print(json.dumps({"cfg3": cfg3(), "cfg4": cfg4()}, indent=1))
