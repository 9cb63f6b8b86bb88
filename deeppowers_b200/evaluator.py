"""Host-side mirror of the C++ `deeppowers::api::fhe` wrappers (include/deeppowers_fhe.hpp).

Thin, allocation-free calls into the C ABI.  Device buffers are torch tensors (any 8-byte
integer dtype, contiguous, on the context's device) or raw device pointers; host buffers are
C-contiguous numpy uint64 arrays.  Errors surface as RuntimeError carrying dpfhe_last_error(),
mirroring the reference's std::runtime_error convention (src/core/hal/cuda/cuda_device.cpp:9-16).
"""
import ctypes as C

import numpy as np

from . import _lib


class DpfheError(RuntimeError):
    pass


def _ptr(x):
    """device pointer of a torch tensor / int; validates dtype width and contiguity"""
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        if x.element_size() != 8 or not x.is_contiguous():
            raise ValueError("device buffers must be contiguous 8-byte integer tensors")
        if not x.is_cuda:
            raise ValueError("expected a CUDA tensor (use the *_host entry points for host arrays)")
        return C.c_void_p(x.data_ptr())
    raise TypeError("unsupported buffer type %r" % type(x))


def _hptr(a, writable=False):
    if not isinstance(a, np.ndarray) or a.dtype != np.uint64 or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("host buffers must be C-contiguous numpy uint64 arrays")
    if writable and not a.flags["WRITEABLE"]:
        raise ValueError("output array is read-only")
    return C.c_void_p(a.ctypes.data)


_CUDA_STREAM_LEGACY = 1   # cudaStreamLegacy: the ABI reserves NULL for "the context's own stream"


def _stream(s):
    """cudaStream_t for the ABI.  None -> torch's current stream (so torch-side copies/events are ordered
    with our kernels); an int is passed through; objects must expose .cuda_stream."""
    if s is None:
        import torch
        h = torch.cuda.current_stream().cuda_stream
    elif isinstance(s, int):
        h = s
    else:
        h = s.cuda_stream
    return C.c_void_p(h if h else _CUDA_STREAM_LEGACY)


class Context:
    """One parameter set bound to one GPU (dpfhe_ctx).  Not thread-safe; one per GPU/process."""

    def __init__(self, log_n, n_limbs, moduli=None, device=0, _borrowed=None):
        self._l = _lib.load()
        self._h = C.c_void_p()
        self._owned = _borrowed is None
        if _borrowed is not None:     # a context owned by a MultiContext
            self._h = C.c_void_p(_borrowed)
        else:
            arr = None
            if moduli is not None:
                if len(moduli) != n_limbs:
                    raise ValueError("need exactly n_limbs moduli")
                arr = (C.c_uint64 * n_limbs)(*[int(m) for m in moduli])
            p = _lib.dpfhe_params(log_n, n_limbs, arr)
            rc = self._l.dpfhe_context_create(C.byref(p), int(device), C.byref(self._h))
            if rc != 0:
                self._h = C.c_void_p()
                raise DpfheError(self._l.dpfhe_last_error().decode())
        self.log_n, self.L, self.N = log_n, n_limbs, 1 << log_n
        self.P = self.L * self.N
        self.device = device
        self.moduli, self.psi = [], []
        for l in range(n_limbs):
            v = C.c_uint64()
            self._chk(self._l.dpfhe_get_modulus(self._h, l, C.byref(v)))
            self.moduli.append(v.value)
            self._chk(self._l.dpfhe_get_psi(self._h, l, C.byref(v)))
            self.psi.append(v.value)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            if self._owned:
                self._l.dpfhe_context_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def synchronize(self):
        """waits for everything issued through this context, on whatever stream"""
        self._chk(self._l.dpfhe_synchronize(self._h))

    # ---- memory other GPUs / processes can write results into (the overlapped gather)
    def device_alloc(self, n_bytes):
        p = C.c_void_p()
        self._chk(self._l.dpfhe_device_alloc(self._h, C.byref(p), n_bytes))
        return p.value

    def device_free(self, ptr):
        self._chk(self._l.dpfhe_device_free(self._h, C.c_void_p(ptr)))

    def ipc_export(self, ptr):
        buf = C.create_string_buffer(64)
        self._chk(self._l.dpfhe_ipc_export(self._h, C.c_void_p(ptr), buf))
        return bytes(buf.raw)

    def ipc_open(self, handle):
        p = C.c_void_p()
        self._chk(self._l.dpfhe_ipc_open(self._h, C.create_string_buffer(bytes(handle), 64), C.byref(p)))
        return p.value

    def ipc_close(self, ptr):
        self._chk(self._l.dpfhe_ipc_close(self._h, C.c_void_p(ptr)))

    # ---- host placement
    def numa_node(self):
        v = C.c_int(-1)
        self._chk(self._l.dpfhe_device_numa_node(self._h, C.byref(v)))
        return v.value

    def bind_thread_near(self):
        """restricts the calling thread to the CPUs of this GPU's NUMA node; returns the number of CPUs (0: unknown topology)"""
        v = C.c_int(0)
        self._chk(self._l.dpfhe_bind_thread_near(self._h, C.byref(v)))
        return v.value

    def pinned_near(self, n_words):
        """pinned host staging memory on this GPU's NUMA node (PinnedBuffer with .array and .node)"""
        return PinnedBuffer(n_words, near=self)

    def _chk(self, rc):
        if rc != 0:
            raise DpfheError(self._l.dpfhe_last_error().decode())

    # ---- introspection
    def root_powers(self, limb, inverse=False):
        out = np.empty(self.N, dtype=np.uint64)
        self._chk(self._l.dpfhe_get_root_powers(self._h, limb, int(inverse), _hptr(out, True)))
        return out

    def launch_count(self):
        return int(self._l.dpfhe_launch_count(self._h))

    def device_bytes(self):
        return int(self._l.dpfhe_context_device_bytes(self._h))

    def phase_cycles(self):
        """diagnostics: per-phase clock64 totals of the fused kernel (needs DPFHE_KS_PROF at context creation)"""
        out = np.zeros(16, dtype=np.uint64)
        self._chk(self._l.dpfhe_debug_phase_cycles(self._h, _hptr(out, True)))
        return out

    def describe(self):
        buf = C.create_string_buffer(1024)
        self._l.dpfhe_describe(self._h, buf, 1024)
        return buf.value.decode()

    # ---- device-pointer ops (asynchronous on `stream`, default: torch's current stream)
    def ntt_fwd(self, data, n_polys, stream=None):
        self._chk(self._l.dpfhe_ntt_fwd(self._h, _ptr(data), n_polys, _stream(stream)))

    def ntt_inv(self, data, n_polys, stream=None):
        self._chk(self._l.dpfhe_ntt_inv(self._h, _ptr(data), n_polys, _stream(stream)))

    def poly_mul_pointwise(self, a, b, out, n_polys, stream=None):
        self._chk(self._l.dpfhe_poly_mul_pointwise(self._h, _ptr(a), _ptr(b), _ptr(out), n_polys, _stream(stream)))

    def poly_add(self, a, b, out, n_polys, stream=None):
        self._chk(self._l.dpfhe_poly_add(self._h, _ptr(a), _ptr(b), _ptr(out), n_polys, _stream(stream)))

    def ct_tensor(self, a, b, d, batch, stream=None):
        self._chk(self._l.dpfhe_ct_tensor(self._h, _ptr(a), _ptr(b), _ptr(d), batch, _stream(stream)))

    def keyswitch(self, d, key, out, batch, stream=None):
        self._chk(self._l.dpfhe_keyswitch(self._h, _ptr(d), _ptr(key), _ptr(out), batch, _stream(stream)))

    def ct_mul_relin(self, a, b, evk, out, batch, stream=None):
        self._chk(self._l.dpfhe_ct_mul_relin(self._h, _ptr(a), _ptr(b), _ptr(evk), _ptr(out), batch, _stream(stream)))

    def ct_mul_plain(self, ct, pt, out, batch, stream=None):
        self._chk(self._l.dpfhe_ct_mul_plain(self._h, _ptr(ct), _ptr(pt), _ptr(out), batch, _stream(stream)))

    def ct_mul_plain_acc(self, ct, pt, acc, batch, stream=None):
        self._chk(self._l.dpfhe_ct_mul_plain_acc(self._h, _ptr(ct), _ptr(pt), _ptr(acc), batch, _stream(stream)))

    def ct_mul_plain_inner(self, steps, pts, out, n_steps, n_groups, batch, stream=None):
        """out[g][k] = sum_b steps[b][k] o pts[g][b]: steps [n_steps][batch][2][L][N], pts [n_groups][n_steps][L][N],
        out [n_groups][batch][2][L][N]; every ciphertext row is read once (the fused BSGS inner loop)"""
        self._chk(self._l.dpfhe_ct_mul_plain_inner(self._h, _ptr(steps), n_steps, _ptr(pts), n_groups, _ptr(out), batch, _stream(stream)))

    def linear_bsgs(self, ct, diags, gk_baby, gk_giant, baby, out, batch, scratch=None, stream=None, fused=True):
        """Encrypted matrix-vector product by baby-step/giant-step diagonals (row f-4, config 4).

        y = sum_g rot_{g*baby}( sum_b D[g*baby + b] o rot_b(x) ), D pre-rotated by -g*baby (caller encodes them so),
        diags: [n][L][N] plaintexts in evaluation form, n a multiple of `baby`; gk_baby / gk_giant: Galois keys of
        rotations by 1 and by `baby` — or gk_baby = list of the baby-1 keys of the rotations by 1 .. baby-1, in which case
        the baby steps are hoisted (one shared digit decomposition).  Uses (baby-1) + (n/baby-1) rotations instead of n-1.  fused=True computes all
        inner sums with one dpfhe_ct_mul_plain_inner call (scratch: baby + n/baby + 1 ciphertext batches); fused=False
        is the reference composition of ct_mul_plain / ct_mul_plain_acc (scratch: baby + 2).  Same bits either way.
        `out` must not alias `ct`."""
        import torch
        n = diags.shape[0]
        assert n % baby == 0, "number of diagonals must be a multiple of the baby-step count"
        giant = n // baby
        shape = (batch, 2, self.L, self.N)
        need = baby + giant + 1 if fused else baby + 2
        if scratch is None:
            scratch = torch.empty((need,) + shape, dtype=torch.int64, device=ct.device)
        assert scratch.shape[0] >= need, "scratch too small for this schedule"
        steps = scratch[:baby]
        g1, gb = self.galois_elt(1), self.galois_elt(baby)
        steps[0].copy_(ct.view(shape))
        if isinstance(gk_baby, (list, tuple)):
            # keys of the rotations by 1 .. baby-1: all baby steps are rotations of the same input and share its digit
            # decomposition (dpfhe_rotate_hoisted)
            assert len(gk_baby) == baby - 1, "need one Galois key per baby step"
            self.rotate_hoisted(steps[0], [self.galois_elt(b) for b in range(1, baby)], list(gk_baby), steps[1:], batch, stream)
        else:
            for b in range(1, baby):
                self.rotate(steps[b - 1], g1, gk_baby, steps[b], batch, stream)
        acc = out
        if fused:
            inner, tmp = scratch[baby:baby + giant], scratch[baby + giant]
            self.ct_mul_plain_inner(steps, diags, inner, baby, giant, batch, stream)
            acc.view(shape).copy_(inner[giant - 1])
            for g in range(giant - 2, -1, -1):
                self.rotate(acc, gb, gk_giant, tmp, batch, stream)          # Horner step: acc = rot_baby(acc) + inner_g
                self.poly_add(tmp, inner[g], acc, 2 * batch, stream)
            return out
        inner, tmp = scratch[baby], scratch[baby + 1]
        for g in range(giant - 1, -1, -1):
            self.ct_mul_plain(steps[0], diags[g * baby], inner, batch, stream)
            for b in range(1, baby):
                self.ct_mul_plain_acc(steps[b], diags[g * baby + b], inner, batch, stream)
            if g == giant - 1:
                acc.view(shape).copy_(inner)
            else:
                self.rotate(acc, gb, gk_giant, tmp, batch, stream)
                self.poly_add(tmp, inner, acc, 2 * batch, stream)
        return out

    def rotate(self, ct, galois_elt, gk, out, batch, stream=None):
        self._chk(self._l.dpfhe_rotate(self._h, _ptr(ct), int(galois_elt), _ptr(gk), _ptr(out), batch, _stream(stream)))

    def rotate_steps(self, ct, k, gk, out, batch, stream=None):
        """rotation by k slots (the Galois element 5^k mod 2N is derived by the library)"""
        self._chk(self._l.dpfhe_rotate_steps(self._h, _ptr(ct), int(k), _ptr(gk), _ptr(out), batch, _stream(stream)))

    def rotate_hoisted(self, ct, galois_elts, gks, out, batch, stream=None):
        """out[r] = rotate(ct, galois_elts[r], gks[r]) for all r, sharing the digit decomposition (bit-identical to rotate);
        gks: list of device tensors, out: [n_rot][batch][2][L][N]"""
        import ctypes as C
        n = len(galois_elts)
        assert len(gks) == n
        ge = (C.c_uint64 * n)(*[int(g) for g in galois_elts])
        kp = (C.c_void_p * n)(*[_ptr(k) for k in gks])
        self._chk(self._l.dpfhe_rotate_hoisted(self._h, _ptr(ct), n, ge, kp, _ptr(out), batch, _stream(stream)))

    def mod_switch_down(self, polys, out, n_polys, t_plain=0, stream=None):
        """drop the last limb: [n_polys][L][N] -> [n_polys][L-1][N] (BGV correction when t_plain > 0)"""
        self._chk(self._l.dpfhe_mod_switch_down(self._h, _ptr(polys), _ptr(out), n_polys, int(t_plain), _stream(stream)))

    # hybrid key switching: this context's last limb is the special prime; data has L-1 limbs, keys [L-1][2][L][N]
    def keyswitch_hybrid(self, d, key, out, batch, t_plain=0, stream=None):
        self._chk(self._l.dpfhe_keyswitch_hybrid(self._h, _ptr(d), _ptr(key), _ptr(out), batch, int(t_plain), _stream(stream)))

    def ct_mul_relin_hybrid(self, a, b, evk, out, batch, t_plain=0, stream=None):
        self._chk(self._l.dpfhe_ct_mul_relin_hybrid(self._h, _ptr(a), _ptr(b), _ptr(evk), _ptr(out), batch, int(t_plain), _stream(stream)))

    def rotate_hybrid(self, ct, galois_elt, gk, out, batch, t_plain=0, stream=None):
        self._chk(self._l.dpfhe_rotate_hybrid(self._h, _ptr(ct), int(galois_elt), _ptr(gk), _ptr(out), batch, int(t_plain), _stream(stream)))

    def fill_uniform(self, seed, data, n_polys, first_poly=0, stream=None):
        self._chk(self._l.dpfhe_fill_uniform(self._h, int(seed), int(first_poly), _ptr(data), n_polys, _stream(stream)))

    # ---- host-buffer ops (synchronous; H2D/compute/D2H pipelined inside the library)
    def ntt_fwd_host(self, data):
        self._chk(self._l.dpfhe_ntt_fwd_host(self._h, _hptr(data, True), data.size // self.P))

    def ntt_inv_host(self, data):
        self._chk(self._l.dpfhe_ntt_inv_host(self._h, _hptr(data, True), data.size // self.P))

    def ct_mul_relin_host(self, a, b, evk, out):
        self._chk(self._l.dpfhe_ct_mul_relin_host(self._h, _hptr(a), _hptr(b), _hptr(evk), _hptr(out, True), a.size // (2 * self.P)))

    def ct_mul_plain_host(self, ct, pt, out):
        self._chk(self._l.dpfhe_ct_mul_plain_host(self._h, _hptr(ct), _hptr(pt), _hptr(out, True), ct.size // (2 * self.P)))

    def rotate_host(self, ct, galois_elt, gk, out):
        self._chk(self._l.dpfhe_rotate_host(self._h, _hptr(ct), int(galois_elt), _hptr(gk), _hptr(out, True), ct.size // (2 * self.P)))

    def ct_mul_relin_hybrid_host(self, a, b, evk, out, t_plain=0):
        pq = 2 * (self.L - 1) * self.N
        self._chk(self._l.dpfhe_ct_mul_relin_hybrid_host(self._h, _hptr(a), _hptr(b), _hptr(evk), _hptr(out, True), a.size // pq, int(t_plain)))

    def rotate_hybrid_host(self, ct, galois_elt, gk, out, t_plain=0):
        pq = 2 * (self.L - 1) * self.N
        self._chk(self._l.dpfhe_rotate_hybrid_host(self._h, _hptr(ct), int(galois_elt), _hptr(gk), _hptr(out, True), ct.size // pq, int(t_plain)))


    # grouped hybrid key switching: the last n_special limbs are special primes, data has L - n_special
    # limbs in digits of n_special limbs, keys [grouped_digits(n_special)][2][L][N]
    def grouped_digits(self, n_special):
        d = C.c_uint(0)
        self._chk(self._l.dpfhe_grouped_digits(self._h, int(n_special), C.byref(d)))
        return int(d.value)

    def keyswitch_grouped(self, n_special, d, key, out, batch, t_plain=0, stream=None):
        self._chk(self._l.dpfhe_keyswitch_grouped(self._h, int(n_special), _ptr(d), _ptr(key), _ptr(out), batch, int(t_plain), _stream(stream)))

    def ct_mul_relin_grouped(self, n_special, a, b, evk, out, batch, t_plain=0, stream=None):
        self._chk(self._l.dpfhe_ct_mul_relin_grouped(self._h, int(n_special), _ptr(a), _ptr(b), _ptr(evk), _ptr(out), batch, int(t_plain),
                                                     _stream(stream)))

    def rotate_grouped(self, n_special, ct, galois_elt, gk, out, batch, t_plain=0, stream=None):
        self._chk(self._l.dpfhe_rotate_grouped(self._h, int(n_special), _ptr(ct), int(galois_elt), _ptr(gk), _ptr(out), batch, int(t_plain),
                                               _stream(stream)))

    def rotate_hoisted_grouped(self, n_special, ct, galois_elts, gks, out, batch, t_plain=0, stream=None):
        """out[r] = a rotation of ct by galois_elts[r] with the grouped hybrid key gks[r], all rotations sharing the mod-up of ct (same
        plaintexts as rotate_grouped, not the same bits); gks: list of device tensors, out: [n_rot][batch][2][L-n_special][N]"""
        n = len(galois_elts)
        assert len(gks) == n
        ge = (C.c_uint64 * n)(*[int(g) for g in galois_elts])
        kp = (C.c_void_p * n)(*[_ptr(k) for k in gks])
        self._chk(self._l.dpfhe_rotate_hoisted_grouped(self._h, int(n_special), _ptr(ct), n, ge, kp, _ptr(out), batch, int(t_plain), _stream(stream)))

    def mod_down_special(self, n_special, polys, out, n_polys, t_plain=0, stream=None):
        self._chk(self._l.dpfhe_mod_down_special(self._h, int(n_special), _ptr(polys), _ptr(out), n_polys, int(t_plain), _stream(stream)))

    def mod_down_special_host(self, n_special, polys, out, t_plain=0):
        self._chk(self._l.dpfhe_mod_down_special_host(self._h, int(n_special), _hptr(polys), _hptr(out, True), polys.size // self.P, int(t_plain)))

    def ct_mul_relin_grouped_host(self, n_special, a, b, evk, out, t_plain=0):
        pq = 2 * (self.L - n_special) * self.N
        self._chk(self._l.dpfhe_ct_mul_relin_grouped_host(self._h, int(n_special), _hptr(a), _hptr(b), _hptr(evk), _hptr(out, True), a.size // pq,
                                                          int(t_plain)))

    def rotate_grouped_host(self, n_special, ct, galois_elt, gk, out, t_plain=0):
        pq = 2 * (self.L - n_special) * self.N
        self._chk(self._l.dpfhe_rotate_grouped_host(self._h, int(n_special), _hptr(ct), int(galois_elt), _hptr(gk), _hptr(out, True), ct.size // pq,
                                                    int(t_plain)))
    def mod_switch_down_host(self, polys, out, t_plain=0):
        self._chk(self._l.dpfhe_mod_switch_down_host(self._h, _hptr(polys), _hptr(out, True), polys.size // self.P, int(t_plain)))

    def galois_elt(self, k):
        """Galois element 5^k mod 2N of a rotation by k slots (k may be negative)."""
        return pow(5, k % (self.N // 2), 2 * self.N)


class PinnedBuffer:
    """Pinned host memory from dpfhe_host_alloc (or, with near=Context, dpfhe_host_alloc_near: pages on that GPU's NUMA
    node, `.node` = the node or -1), exposed as a numpy uint64 array (`.array`); freed on close()/GC."""

    def __init__(self, n_words, near=None):
        self._l = _lib.load()
        self._p = C.c_void_p()
        self.node = -1
        if near is not None:
            node = C.c_int(-1)
            rc = self._l.dpfhe_host_alloc_near(near._h, C.byref(self._p), n_words * 8, C.byref(node))
            self.node = node.value
        else:
            rc = self._l.dpfhe_host_alloc(C.byref(self._p), n_words * 8)
        if rc != 0:
            raise DpfheError(self._l.dpfhe_last_error().decode())
        self.array = np.ctypeslib.as_array(C.cast(self._p, C.POINTER(C.c_uint64)), shape=(n_words,))

    def close(self):
        if self._p and self._p.value:
            self.array = None
            self._l.dpfhe_host_free(self._p)
            self._p = C.c_void_p()

    __del__ = close


def linear_bsgs_grouped(ctx, ctx_q, n_special, ct, diags, gk_baby, gk_giant, baby, out, batch, t_plain=0, scratch=None, stream=None):
    """The baby-step/giant-step matrix-vector product of Context.linear_bsgs with special-prime keys: `ctx` is the
    key-switching context (ciphertext moduli + n_special special primes), `ctx_q` a context over the ciphertext moduli alone for
    the element-wise steps.  ct / out: [batch][2][Lq][N]; diags: [n][Lq][N]; gk_baby: the baby-1 grouped Galois keys of the
    rotations by 1 .. baby-1 (the baby steps are hoisted: one mod-up of the input); gk_giant: the key of the rotation by `baby`.
    scratch: baby + n/baby + 1 ciphertext batches."""
    import torch
    n = diags.shape[0]
    assert n % baby == 0 and len(gk_baby) == baby - 1
    giant = n // baby
    Lq = ctx.L - n_special
    assert ctx_q.L == Lq and ctx_q.moduli == ctx.moduli[:Lq], "ctx_q must be the context of the ciphertext moduli"
    shape = (batch, 2, Lq, ctx.N)
    need = baby + giant + 1
    if scratch is None:
        scratch = torch.empty((need,) + shape, dtype=torch.int64, device=ct.device)
    assert scratch.shape[0] >= need
    steps, inner, tmp = scratch[:baby], scratch[baby:baby + giant], scratch[baby + giant]
    steps[0].copy_(ct.view(shape))
    ctx.rotate_hoisted_grouped(n_special, steps[0], [ctx.galois_elt(b) for b in range(1, baby)], list(gk_baby), steps[1:], batch, t_plain, stream)
    ctx_q.ct_mul_plain_inner(steps, diags, inner, baby, giant, batch, stream)
    acc = out
    acc.view(shape).copy_(inner[giant - 1])
    gb = ctx.galois_elt(baby)
    for g in range(giant - 2, -1, -1):
        ctx.rotate_grouped(n_special, acc, gb, gk_giant, tmp, batch, t_plain, stream)      # Horner step: acc = rot_baby(acc) + inner_g
        ctx_q.poly_add(tmp, inner[g], acc, 2 * batch, stream)
    return out


class LinearLayer:
    """Encrypted linear layer as a library object (dpfhe_linear_*): diagonal plaintexts and Galois keys live on the device; apply()
    takes device buffers, apply_host() host buffers (pipelined in chunks).  diags [n][L][N], gk_baby [baby-1][L][2][L][N] (keys of the
    rotations by 1 .. baby-1), gk_giant [L][2][L][N] (rotation by `baby`): C-contiguous numpy uint64 arrays."""

    def __init__(self, ctx, diags, baby, gk_baby, gk_giant):
        self._l, self.ctx = ctx._l, ctx
        self._h = C.c_void_p()
        n = diags.shape[0]
        rc = self._l.dpfhe_linear_create(ctx._h, _hptr(diags), n, int(baby), _hptr(gk_baby) if gk_baby is not None else None,
                                         _hptr(gk_giant) if gk_giant is not None else None, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise DpfheError(self._l.dpfhe_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._l.dpfhe_linear_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def apply(self, ct, out, batch, stream=None):
        self.ctx._chk(self._l.dpfhe_linear_apply(self._h, _ptr(ct), _ptr(out), batch, _stream(stream)))

    def apply_host(self, ct, out):
        self.ctx._chk(self._l.dpfhe_linear_apply_host(self._h, _hptr(ct), _hptr(out, True), ct.size // (2 * self.ctx.P)))


class MultiContext:
    """Several GPUs in one process (dpfhe_multi_*): one context per device, contiguous shards of the batch, no collective
    while computing.  `devices`: list of CUDA device ids (a device may be listed twice), None = all visible devices."""

    def __init__(self, log_n, n_limbs, moduli=None, devices=None):
        self._l = _lib.load()
        self._h = C.c_void_p()
        arr = None
        if moduli is not None:
            arr = (C.c_uint64 * n_limbs)(*[int(m) for m in moduli])
        p = _lib.dpfhe_params(log_n, n_limbs, arr)
        ids = (C.c_int * len(devices))(*devices) if devices else None
        rc = self._l.dpfhe_multi_create(C.byref(p), ids, len(devices) if devices else 0, C.byref(self._h))
        if rc != 0:
            self._h = C.c_void_p()
            raise DpfheError(self._l.dpfhe_last_error().decode())
        self.n = int(self._l.dpfhe_multi_device_count(self._h))
        self.contexts = [Context(log_n, n_limbs, _borrowed=self._l.dpfhe_multi_context(self._h, r)) for r in range(self.n)]
        self.devices = [int(self._l.dpfhe_context_device(c._h)) for c in self.contexts]
        self.log_n, self.L, self.N = log_n, n_limbs, 1 << log_n
        self.P = self.L * self.N

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            for c in self.contexts:
                c.close()
            self._l.dpfhe_multi_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def _chk(self, rc):
        if rc != 0:
            raise DpfheError(self._l.dpfhe_last_error().decode())

    def shard(self, batch, index):
        first, count = C.c_size_t(), C.c_size_t()
        self._chk(self._l.dpfhe_multi_shard(self._h, batch, index, C.byref(first), C.byref(count)))
        return first.value, count.value

    def ct_mul_relin_host(self, a, b, evk, out):
        self._chk(self._l.dpfhe_multi_ct_mul_relin_host(self._h, _hptr(a), _hptr(b), _hptr(evk), _hptr(out, True), a.size // (2 * self.P)))

    def ct_mul_relin_grouped_host(self, n_special, a, b, evk, out, t_plain=0):
        """special-prime key switching sharded over the devices; a, b, out: [batch][2][L - n_special][N] host arrays"""
        pq = 2 * (self.L - n_special) * self.N
        self._chk(self._l.dpfhe_multi_ct_mul_relin_grouped_host(self._h, int(n_special), _hptr(a), _hptr(b), _hptr(evk), _hptr(out, True), a.size // pq,
                                                                int(t_plain)))

    def rotate_host(self, ct, galois_elt, gk, out):
        self._chk(self._l.dpfhe_multi_rotate_host(self._h, _hptr(ct), int(galois_elt), _hptr(gk), _hptr(out, True), ct.size // (2 * self.P)))

    def ct_mul_relin_gather(self, a_shards, b_shards, evk_copies, out_root, root, batch):
        """a_shards[r], b_shards[r], evk_copies[r]: device tensors / pointers on device r; out_root: [batch][2][L][N] on the
        device of shard `root`.  Every device writes its rows of out_root directly (peer stores); synchronous."""
        n = self.n
        mk = lambda xs: (C.c_void_p * n)(*[_ptr(x) for x in xs])
        self._chk(self._l.dpfhe_multi_ct_mul_relin_gather(self._h, mk(a_shards), mk(b_shards), mk(evk_copies), _ptr(out_root), int(root), batch))


def pinned_empty(n_words):
    """convenience: a PinnedBuffer (keep the object alive while its `.array` is in use)"""
    return PinnedBuffer(n_words)
