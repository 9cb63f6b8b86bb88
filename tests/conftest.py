import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_available():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _gpu_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")


class Emu:
    """ctypes view of tests/_emu/libdpfhe_emu.so: the kernel bodies compiled for the host (test infra)."""

    def __init__(self, lib, log_n, L, moduli=None):
        self._l = lib
        arr = (C.c_uint64 * L)(*[int(m) for m in moduli]) if moduli is not None else None
        self._h = lib.emu_create(log_n, L, arr)
        if not self._h:
            raise ValueError("emu_create rejected the parameters")
        self.log_n, self.L, self.N = log_n, L, 1 << log_n
        self.moduli = [int(lib.emu_modulus(self._h, l)) for l in range(L)]
        self.psi = [int(lib.emu_psi(self._h, l)) for l in range(L)]

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.emu_destroy(self._h)
            self._h = None

    def root_powers(self, l, inverse=False):
        out = np.empty(self.N, dtype=np.uint64)
        self._l.emu_root_powers(self._h, l, int(inverse), out)
        return out

    def ntt_pair(self, data, inverse=False):
        """N = 16384 only: the CTA-pair bodies (ntt_pair_kernel)"""
        d = np.ascontiguousarray(data, dtype=np.uint64).copy()
        assert self._l.emu_ntt_pair(self._h, d.reshape(-1), d.size // (self.L * self.N), int(inverse)) == 0
        return d

    def ntt(self, data, inverse=False):
        d = np.ascontiguousarray(data, dtype=np.uint64).copy()
        assert self._l.emu_ntt(self._h, d.reshape(-1), d.size // (self.L * self.N), int(inverse)) == 0
        return d

    def ks(self, mode, a, b, key, batch, galois=0, G=None):
        out = np.zeros((batch, 2, self.L, self.N), dtype=np.uint64)
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64)
        assert self._l.emu_ks(self._h, mode, a.reshape(-1), b.reshape(-1), np.ascontiguousarray(key).reshape(-1),
                              out.reshape(-1), batch, int(galois), G or 2 * self.L) == 0
        return out

    def ks_hybrid(self, mode, a, b, key, batch, galois=0, t_plain=0, G=None):
        out = np.zeros((batch, 2, self.L - 1, self.N), dtype=np.uint64)
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64)
        assert self._l.emu_ks_hybrid(self._h, mode, a.reshape(-1), b.reshape(-1), np.ascontiguousarray(key).reshape(-1),
                                     out.reshape(-1), batch, int(galois), int(t_plain), G or 2 * self.L) == 0
        return out


    def ks_grouped(self, K, mode, a, b, key, batch, galois=0, t_plain=0, G=None):
        out = np.zeros((batch, 2, self.L - K, self.N), dtype=np.uint64)
        a = np.ascontiguousarray(a, dtype=np.uint64)
        b = a if b is None else np.ascontiguousarray(b, dtype=np.uint64)
        assert self._l.emu_ks_grouped(self._h, int(K), mode, a.reshape(-1), b.reshape(-1), np.ascontiguousarray(key).reshape(-1),
                                      out.reshape(-1), batch, int(galois), int(t_plain), G or 2 * self.L) == 0
        return out
    def rotate_hoisted(self, ct, galois, keys, G=None):
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        batch = ct.size // (2 * self.L * self.N)
        g = np.ascontiguousarray(galois, dtype=np.uint64)
        out = np.zeros((len(g), batch, 2, self.L, self.N), dtype=np.uint64)
        flagged = C.c_uint(0)
        assert self._l.emu_rotate_hoisted(self._h, ct.reshape(-1), len(g), g, np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1),
                                          out.reshape(-1), batch, G or 2 * self.L, C.byref(flagged)) == 0
        return out, flagged.value

    def pt_inner(self, steps, pts, gmax=0):
        steps = np.ascontiguousarray(steps, dtype=np.uint64)
        pts = np.ascontiguousarray(pts, dtype=np.uint64)
        nb, batch, ng = steps.shape[0], steps.shape[1], pts.shape[0]
        out = np.zeros((ng, batch, 2, self.L, self.N), dtype=np.uint64)
        assert self._l.emu_pt_inner(self._h, steps.reshape(-1), nb, pts.reshape(-1), ng, out.reshape(-1), batch, gmax) == 0
        return out

    def mod_switch(self, polys, t_plain=0):
        x = np.ascontiguousarray(polys, dtype=np.uint64).reshape(-1, self.L, self.N)
        out = np.zeros((x.shape[0], self.L - 1, self.N), dtype=np.uint64)
        assert self._l.emu_mod_switch(self._h, x.reshape(-1), out.reshape(-1), x.shape[0], int(t_plain)) == 0
        return out

    def mod_down_special(self, K, polys, t_plain=0):
        x = np.ascontiguousarray(polys, dtype=np.uint64).reshape(-1, self.L, self.N)
        out = np.zeros((x.shape[0], self.L - K, self.N), dtype=np.uint64)
        assert self._l.emu_mod_down_special(self._h, int(K), x.reshape(-1), out.reshape(-1), x.shape[0], int(t_plain)) == 0
        return out

    def rotate_hoisted_grouped(self, K, ct, galois, keys, t_plain=0):
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        g = np.ascontiguousarray(galois, dtype=np.uint64)
        batch = ct.size // (2 * (self.L - K) * self.N)
        out = np.zeros((len(g), batch, 2, self.L - K, self.N), dtype=np.uint64)
        assert self._l.emu_rotate_hoisted_grouped(self._h, int(K), ct.reshape(-1), len(g), g, np.ascontiguousarray(keys, dtype=np.uint64).reshape(-1),
                                                  out.reshape(-1), batch, int(t_plain)) == 0
        return out

    def scalar(self, name, l, *args):
        return int(getattr(self._l, "emu_" + name)(self._h, l, *[C.c_uint64(int(a)) for a in args]))


def _build_emu(variant):
    """tests/_emu/libdpfhe_emu_<variant>.so: the kernel bodies of one arithmetic variant (csrc/types.hpp) compiled for the host"""
    out_dir = os.path.join(ROOT, "tests", "_emu")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libdpfhe_emu_%s.so" % variant)
    csrc = os.path.join(ROOT, "deeppowers_b200", "csrc")
    srcs = [os.path.join(ROOT, "tests", "emu", "emu.cpp"), os.path.join(csrc, "host_params.cpp")]
    deps = srcs + [os.path.join(csrc, f) for f in ("types.hpp", "modarith.cuh", "ntt_core.cuh", "kernel_bodies.cuh", "host_params.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.check_call([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-DDPFHE_FAST=%d" % (variant == "fast"), "-x", "c++", "-I", csrc] + srcs + ["-o", so])
    lib = C.CDLL(so)
    lib.emu_create.restype = C.c_void_p
    lib.emu_create.argtypes = [C.c_uint, C.c_uint, C.c_void_p]
    lib.emu_destroy.argtypes = [C.c_void_p]
    lib.emu_modulus.restype = C.c_uint64
    lib.emu_modulus.argtypes = [C.c_void_p, C.c_uint]
    lib.emu_psi.restype = C.c_uint64
    lib.emu_psi.argtypes = [C.c_void_p, C.c_uint]
    lib.emu_root_powers.argtypes = [C.c_void_p, C.c_uint, C.c_int, _u64p]
    lib.emu_ntt.argtypes = [C.c_void_p, _u64p, C.c_size_t, C.c_int]
    lib.emu_ntt_pair.argtypes = [C.c_void_p, _u64p, C.c_size_t, C.c_int]
    lib.emu_ks.argtypes = [C.c_void_p, C.c_int, _u64p, _u64p, _u64p, _u64p, C.c_size_t, C.c_uint32, C.c_uint]
    lib.emu_ks_grouped.argtypes = [C.c_void_p, C.c_uint, C.c_int, _u64p, _u64p, _u64p, _u64p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_uint]
    lib.emu_ks_hybrid.argtypes = [C.c_void_p, C.c_int, _u64p, _u64p, _u64p, _u64p, C.c_size_t, C.c_uint32, C.c_uint64, C.c_uint]
    lib.emu_rotate_hoisted.argtypes = [C.c_void_p, _u64p, C.c_size_t, _u64p, _u64p, _u64p, C.c_size_t, C.c_uint, C.POINTER(C.c_uint)]
    lib.emu_pt_inner.argtypes = [C.c_void_p, _u64p, C.c_uint, _u64p, C.c_uint, _u64p, C.c_size_t, C.c_uint]
    lib.emu_mod_down_special.argtypes = [C.c_void_p, C.c_uint, _u64p, _u64p, C.c_size_t, C.c_uint64]
    lib.emu_rotate_hoisted_grouped.argtypes = [C.c_void_p, C.c_uint, _u64p, C.c_size_t, _u64p, _u64p, _u64p, C.c_size_t, C.c_uint64]
    lib.emu_mod_switch.argtypes = [C.c_void_p, _u64p, _u64p, C.c_size_t, C.c_uint64]
    for nm, nargs in (("mulmod", 2), ("word_reduce", 1), ("canon", 1), ("canon_store", 1), ("mulmod_lazy", 2), ("barrett_long", 2), ("pti_fold", 4),
                      ("shoup_lazy", 2), ("shoup_exact", 2)):
        f = getattr(lib, "emu_" + nm)
        f.restype = C.c_uint64
        f.argtypes = [C.c_void_p, C.c_uint] + [C.c_uint64] * nargs
    return lib


@pytest.fixture(scope="session")
def emu_libs():
    return {v: _build_emu(v) for v in ("gen", "fast")}


def is_fast_modulus(q):
    return q & 0xFFFFFFFF == 1


@pytest.fixture(scope="session")
def make_emu(emu_libs):
    """make_emu(log_n, L, moduli=None, variant=None): the variant the product would pick for these moduli ("fast" when all are
    k * 2^32 + 1, which includes the default basis), or an explicit "gen" / "fast"."""
    def make(log_n, L, moduli=None, variant=None):
        if variant is None:
            variant = "fast" if moduli is None or all(is_fast_modulus(int(q)) for q in moduli) else "gen"
        return Emu(emu_libs[variant], log_n, L, moduli)
    return make
