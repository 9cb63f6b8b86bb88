"""Python mirror of include/dpfhe_wire.hpp (flat DPFHEv1 files: 160-byte header + raw u64 payload)."""
import struct

import numpy as np

MAGIC = b"DPFHEv1\x00"
CIPHERTEXTS, SWITCH_KEY, PLAINTEXTS, HYBRID_SWITCH_KEY, GROUPED_SWITCH_KEY = 1, 2, 3, 4, 5   # hybrid / grouped: n_limbs includes the special primes;
# a grouped key's `count` is the number of special primes K, its payload [ceil((n_limbs-K)/K)][2][n_limbs][N]
_HDR = struct.Struct("<8sIIIIQ16Q")


def payload_words(log_n, n_limbs, kind, count):
    if not (1 <= log_n <= 17 and 1 <= n_limbs <= 16):
        raise ValueError("bad parameters")
    if kind not in (CIPHERTEXTS, SWITCH_KEY, PLAINTEXTS, HYBRID_SWITCH_KEY, GROUPED_SWITCH_KEY):
        raise ValueError("unknown kind")
    poly = (1 << log_n) * n_limbs
    if kind == GROUPED_SWITCH_KEY:
        if not (1 <= count <= 4 and 2 * count <= n_limbs):
            raise ValueError("bad number of special primes")
        return 2 * (-(-(n_limbs - count) // count)) * poly
    return {CIPHERTEXTS: count * 2 * poly, SWITCH_KEY: 2 * n_limbs * poly, PLAINTEXTS: count * poly,
            HYBRID_SWITCH_KEY: 2 * (n_limbs - 1) * poly}[kind]


def write(path, log_n, n_limbs, kind, count, moduli, payload, form=1):
    payload = np.ascontiguousarray(payload, dtype="<u8").reshape(-1)
    if payload.size != payload_words(log_n, n_limbs, kind, count):
        raise ValueError("payload size does not match the header")
    mods = list(int(m) for m in moduli) + [0] * (16 - len(moduli))
    with open(path, "wb") as f:
        f.write(_HDR.pack(MAGIC, log_n, n_limbs, kind, form, count, *mods))
        f.write(payload.tobytes())


def read(path):
    with open(path, "rb") as f:
        raw = f.read(_HDR.size)
        if len(raw) != _HDR.size:
            raise ValueError("truncated header")
        magic, log_n, n_limbs, kind, form, count, *mods = _HDR.unpack(raw)
        if magic != MAGIC:
            raise ValueError("not a DPFHEv1 file")
        words = payload_words(log_n, n_limbs, kind, count)
        f.seek(0, 2)
        if f.tell() != _HDR.size + 8 * words:      # the header is untrusted: it must describe exactly this file
            raise ValueError("file size does not match the header")
        f.seek(_HDR.size)
        data = np.frombuffer(f.read(words * 8), dtype="<u8")
        if data.size != words:
            raise ValueError("truncated payload")
    return {"log_n": log_n, "n_limbs": n_limbs, "kind": kind, "form": form, "count": count, "moduli": mods[:n_limbs]}, data.astype(np.uint64)
