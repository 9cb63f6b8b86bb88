"""Slot (batch) encoding for the BGV test scheme: plaintexts are polynomials mod a prime t = 1 (mod 2N);
slot (r, c) holds m(zeta^(+-5^c)), so the Galois element 5^k rotates each row of N/2 slots by k.
numpy only (test infrastructure)."""
import numpy as np


def _find_root(t, two_n):
    for g in range(2, 1000):
        z = pow(g, (t - 1) // two_n, t)
        if pow(z, two_n // 2, t) == t - 1:
            return z
    raise ValueError("no primitive root found")


class SlotEncoder:
    def __init__(self, n, t):
        assert (t - 1) % (2 * n) == 0
        self.n, self.t = n, t
        self.logn = n.bit_length() - 1
        self.zeta = _find_root(t, 2 * n)
        zp = np.ones(2 * n, dtype=np.uint64)
        for k in range(1, 2 * n):
            zp[k] = (int(zp[k - 1]) * self.zeta) % t
        self.zp = zp                               # zeta^k
        self.rev = np.array([int(format(i, "0%db" % self.logn)[::-1], 2) for i in range(n)])
        # slot (r, c) -> index of the odd exponent e in the natural evaluation order (e = 2*idx + 1)
        idx = np.empty((2, n // 2), dtype=np.int64)
        e = 1
        for c in range(n // 2):
            idx[0, c] = (e - 1) // 2
            idx[1, c] = ((2 * n - e) - 1) // 2
            e = (e * 5) % (2 * n)
        self.slot_index = idx

    def _cyclic(self, a, inverse):
        """iterative radix-2 NTT of length n over Z_t with omega = zeta^2 (or its inverse)"""
        n, t = self.n, self.t
        a = a[self.rev].astype(np.uint64)
        size = 2
        while size <= n:
            half, step = size // 2, (2 * n) // size          # omega_size = zeta^(2n/size)
            exps = (np.arange(half) * step) % (2 * n)
            if inverse:
                exps = (2 * n - exps) % (2 * n)
            w = self.zp[exps]
            a = a.reshape(-1, size)
            u = a[:, :half].copy()
            v = (a[:, half:] * w) % t
            a[:, :half] = (u + v) % t
            a[:, half:] = (u + t - v) % t
            a = a.reshape(-1)
            size *= 2
        return a

    def evaluate(self, m):
        """E[i] = m(zeta^(2i+1))"""
        tw = (np.asarray(m, dtype=np.uint64) % self.t * self.zp[: self.n]) % self.t
        return self._cyclic(tw, inverse=False)

    def interpolate(self, e):
        a = self._cyclic(np.asarray(e, dtype=np.uint64) % self.t, inverse=True)
        ninv = pow(self.n, self.t - 2, self.t)
        inv_tw = self.zp[(2 * self.n - np.arange(self.n)) % (2 * self.n)]
        return (a * ninv % self.t) * inv_tw % self.t

    def encode(self, slots):
        """slots: [2][n/2] integers (any sign) -> plaintext coefficients in [0,t)"""
        e = np.zeros(self.n, dtype=np.uint64)
        s = np.asarray(slots, dtype=np.int64) % self.t
        e[self.slot_index[0]] = s[0].astype(np.uint64)
        e[self.slot_index[1]] = s[1].astype(np.uint64)
        return self.interpolate(e)

    def decode(self, m):
        e = self.evaluate(m)
        return np.stack([e[self.slot_index[0]], e[self.slot_index[1]]])
