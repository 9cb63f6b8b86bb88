// modarith.cuh — 64-bit modular arithmetic for moduli 2^33 < q < 2^60.
//
// Everything is __host__ __device__ so the same code is exercised by the host emulator
// in tests/emu (test infrastructure; never part of libdpfhe.so's product path).
//
// Lazy-range conventions ("bound B" means value < B*q; 16q < 2^64 because q < 2^60):
//   shoup_lazy(x, w)      any 64-bit x        -> [0, 2q)
//   word_reduce(x)        any 64-bit x        -> [0, 3q)   (3 integer multiplies)
//   barrett_lazy(a*b)     a*b <= 4 q^2        -> [0, 3q)
//   barrett_lazy_long(z)  z < 2^(2b+4)        -> [0, 15q)  (b = bit length of q; sums of up to 16 products)
//   canon(x)              x < 16q             -> [0, q)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define DPFHE_HD __host__ __device__ __forceinline__
#else
#define DPFHE_HD inline
#endif

namespace dpfhe {

typedef uint64_t u64;
typedef uint32_t u32;

struct alignas(16) U64x2 {
    u64 x, y;
};

// Per-limb constants (host-built in params.cpp, resident in device global memory).
struct alignas(16) LimbParams {
    u64 q;           // modulus
    u64 q2;          // 2q
    u64 q4;          // 4q
    u64 q8;          // 8q  (< 2^63)
    u64 nq;          // 2^64 - q: adding h*nq subtracts h*q without a separate negation
    u64 bar_mu;      // floor(2^(bar_shift+64) / q)
    u64 ninv;        // N^-1 mod q                    } folded into the last inverse stage
    u64 ninv_s;      // Shoup companion of ninv
    u64 wninv;       // psi^-bitrev(1) * N^-1 mod q
    u64 wninv_s;     // Shoup companion of wninv
    u32 bar_shift;   // bitlen(q) - 2
    u32 mu32;        // floor(2^64 / q)  (< 2^31 because q > 2^33)
};

DPFHE_HD u64 umulhi64(u64 a, u64 b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (u64)(((unsigned __int128)a * b) >> 64);
#endif
}

DPFHE_HD u32 umulhi32(u32 a, u32 b) {
    // high half of a 32x32 wide multiply (IMAD.WIDE, not the much slower IMAD.HI)
    return (u32)(((u64)a * b) >> 32);
}

// x >= m ? x - m : x, branch-free.  Correct for every 64-bit x when m <= 2^63.
// Device: subtract with borrow and select on the borrow (IADD3, IADD3.X, 2x SEL).
DPFHE_HD u64 csub(u64 x, u64 m) {
#if defined(__CUDA_ARCH__)
    u32 xl = (u32)x, xh = (u32)(x >> 32), ml = (u32)m, mh = (u32)(m >> 32), tl, th, b;
    asm("sub.cc.u32 %0, %3, %5;\n\tsubc.cc.u32 %1, %4, %6;\n\tsubc.u32 %2, 0, 0;"
        : "=&r"(tl), "=&r"(th), "=&r"(b)
        : "r"(xl), "r"(xh), "r"(ml), "r"(mh));
    const u32 rl = b ? xl : tl, rh = b ? xh : th;
    return ((u64)rh << 32) | rl;
#else
    u64 t = x - m;
    return t < x ? t : x;   // unsigned wrap makes t > x exactly when x < m
#endif
}

// c + a*b mod 2^64 as ONE multiply-add chain (IMAD.WIDE + 2 IMAD, no separate adds)
DPFHE_HD u64 mad_lo64(u64 a, u64 b, u64 c) {
#if defined(__CUDA_ARCH__)
    u64 t;
    asm("{\n\t"
        ".reg .u32 al, ah, bl, bh, t0, t1;\n\t"
        ".reg .u64 T;\n\t"
        "mov.b64 {al, ah}, %1;\n\t"
        "mov.b64 {bl, bh}, %2;\n\t"
        "mad.wide.u32 T, al, bl, %3;\n\t"
        "mov.b64 {t0, t1}, T;\n\t"
        "mad.lo.u32 t1, al, bh, t1;\n\t"
        "mad.lo.u32 t1, ah, bl, t1;\n\t"
        "mov.b64 %0, {t0, t1};\n\t"
        "}"
        : "=l"(t)
        : "l"(a), "l"(b), "l"(c));
    return t;
#else
    return c + a * b;
#endif
}

// Shoup multiplication by a fixed w < q with ws = floor(w * 2^64 / q): valid for ANY 64-bit x.
// r = x*w - floor(x*ws / 2^64) * q  (mod 2^64), r in [0, 2q).
// Device form: the quotient uses ptxas' own mul.hi.u64 expansion (4 IMAD.WIDE with carry predicates); the
// low 64 bits are one explicit chain t = xl*wl + hl*nql (2 IMAD.WIDE), t.hi += xl*wh + xh*wl + hl*nqh + hh*nql
// (4 IMAD) with nq = 2^64 - q, which avoids the negation and the split adds nvcc otherwise emits.
DPFHE_HD u64 shoup_lazy(u64 x, u64 w, u64 ws, u64 q, u64 nq) {
#if defined(__CUDA_ARCH__)
    const u64 h = __umul64hi(x, ws);
    u64 t;
    asm("{\n\t"
        ".reg .u32 xl, xh, wl, wh, hl, hh, nl, nh, t0, t1;\n\t"
        ".reg .u64 T;\n\t"
        "mov.b64 {xl, xh}, %1;\n\t"
        "mov.b64 {wl, wh}, %2;\n\t"
        "mov.b64 {hl, hh}, %3;\n\t"
        "mov.b64 {nl, nh}, %4;\n\t"
        "mul.wide.u32 T, xl, wl;\n\t"
        "mad.wide.u32 T, hl, nl, T;\n\t"
        "mov.b64 {t0, t1}, T;\n\t"
        "mad.lo.u32 t1, xl, wh, t1;\n\t"
        "mad.lo.u32 t1, xh, wl, t1;\n\t"
        "mad.lo.u32 t1, hl, nh, t1;\n\t"
        "mad.lo.u32 t1, hh, nl, t1;\n\t"
        "mov.b64 %0, {t0, t1};\n\t"
        "}"
        : "=l"(t)
        : "l"(x), "l"(w), "l"(h), "l"(nq));
    (void)q;
    return t;
#else
    (void)nq;
    u64 h = umulhi64(x, ws);
    return x * w - h * q;   // in [0, 2q)
#endif
}
// One-word quotient estimate: k = floor(x_hi * mu32 / 2^32) <= floor(x/q), off by at most 2.
DPFHE_HD u64 word_reduce(u64 x, const LimbParams &p) {
    u32 k = umulhi32((u32)(x >> 32), p.mu32);
    return mad_lo64((u64)k, p.nq, x);   // x - k*q, in [0, 3q)
}

DPFHE_HD u64 shoup_lazy(u64 x, u64 w, u64 ws, const LimbParams &p) { return shoup_lazy(x, w, ws, p.q, p.nq); }

// 128-bit product (hi:lo) of two 64-bit words.  Device: four IMAD.WIDE partial products combined once
// (nvcc's separate a*b and __umul64hi(a,b) would recompute the low partial product: 5 IMAD.WIDE + 2 IMAD).
DPFHE_HD void mul128(u64 a, u64 b, u64 &hi, u64 &lo) {
#if defined(__CUDA_ARCH__)
    asm("{\n\t"
        ".reg .u32 al, ah, bl, bh, p0l, p0h, ml, mh, cy;\n\t"
        ".reg .u64 p0, p1, p2, p3, m, t;\n\t"
        "mov.b64 {al, ah}, %2;\n\t"
        "mov.b64 {bl, bh}, %3;\n\t"
        "mul.wide.u32 p0, al, bl;\n\t"
        "mul.wide.u32 p1, al, bh;\n\t"
        "mul.wide.u32 p2, ah, bl;\n\t"
        "mul.wide.u32 p3, ah, bh;\n\t"
        "mov.b64 {p0l, p0h}, p0;\n\t"
        "add.cc.u64 m, p1, p2;\n\t"
        "addc.u32 cy, 0, 0;\n\t"
        "cvt.u64.u32 t, p0h;\n\t"
        "add.cc.u64 m, m, t;\n\t"
        "addc.u32 cy, cy, 0;\n\t"
        "mov.b64 {ml, mh}, m;\n\t"
        "mov.b64 %1, {p0l, ml};\n\t"
        "mov.b64 t, {mh, cy};\n\t"
        "add.u64 %0, p3, t;\n\t"
        "}"
        : "=l"(hi), "=l"(lo)
        : "l"(a), "l"(b));
#else
    unsigned __int128 z = (unsigned __int128)a * b;
    lo = (u64)z;
    hi = (u64)(z >> 64);
#endif
}

// Barrett reduction of z = hi:lo.  Requires z < 2^(s+64), s = bitlen(q) - 2, i.e. z <= 4 q^2 (factor bounds Ba*Bb <= 4:
// the shifted value must fit one word), gives [0, 3q).
// With both factors < q the result is in [0, 2q).
DPFHE_HD u64 barrett_lazy(u64 hi, u64 lo, const LimbParams &p) {
    const u32 s = p.bar_shift;                       // 32 <= s <= 58 because 2^33 < q < 2^60
#if defined(__CUDA_ARCH__)
    // floor(z / 2^s) with two funnel shifts over the three upper words of z
    const u32 w1 = (u32)(lo >> 32), w2 = (u32)hi, w3 = (u32)(hi >> 32);
    const u32 zl = __funnelshift_r(w1, w2, s - 32), zh = __funnelshift_r(w2, w3, s - 32);
    const u64 zt = ((u64)zh << 32) | zl;
#else
    const u64 zt = (hi << (64 - s)) | (lo >> s);     // floor(z / 2^s) < 2^64
#endif
    u64 qh = umulhi64(zt, p.bar_mu);
    return mad_lo64(qh, p.nq, lo);   // lo - qh*q
}

// Barrett reduction of a longer sum z = hi:lo < 2^(2b+4), b = bit length of q (e.g. 16 products of canonical factors):
// the quotient is estimated from z / 2^(s+2) so that the shifted value still fits one word; 4*qh is within 14 of
// floor(z/q), so the result is in [0, 15q) (and 15q < 2^64).
DPFHE_HD u64 barrett_lazy_long(u64 hi, u64 lo, const LimbParams &p) {
    const u32 s = p.bar_shift + 2;                   // 34 <= s <= 60
#if defined(__CUDA_ARCH__)
    const u32 w1 = (u32)(lo >> 32), w2 = (u32)hi, w3 = (u32)(hi >> 32);
    const u32 zl = __funnelshift_r(w1, w2, s - 32), zh = __funnelshift_r(w2, w3, s - 32);
    const u64 zt = ((u64)zh << 32) | zl;
#else
    const u64 zt = (hi << (64 - s)) | (lo >> s);     // floor(z / 2^s) < 2^64
#endif
    const u64 qh = umulhi64(zt, p.bar_mu);
    return mad_lo64(qh << 2, p.nq, lo);              // lo - 4*qh*q
}

DPFHE_HD u64 mulmod_lazy(u64 a, u64 b, const LimbParams &p) {
    u64 hi, lo;
    mul128(a, b, hi, lo);
    return barrett_lazy(hi, lo, p);
}

// x < 16q  ->  [0, q)
DPFHE_HD u64 canon(u64 x, const LimbParams &p) {
    u64 r = word_reduce(x, p);   // < 3q
    r = csub(r, p.q2);
    return csub(r, p.q);
}
// x < 4q -> [0, q)
DPFHE_HD u64 canon4(u64 x, const LimbParams &p) { return csub(csub(x, p.q2), p.q); }

DPFHE_HD u64 mulmod(u64 a, u64 b, const LimbParams &p) { return canon4(mulmod_lazy(a, b, p), p); }

// Modulus switching / special-prime division: constants of one call, built on the host
// (host_params.cpp:build_ms_consts) and passed by value in the kernel parameter block.
struct MsConsts {
    u64 inv[16], inv_s[16];     // q_last^-1 mod q_i and its Shoup companion
    u64 sinv[16], sinv_s[16];   // s * q_last^-1 mod q_i (s = t_plain, or 1 for plain rounding)
    u64 qlm[16], qlm_s[16];     // q_last mod q_i and its Shoup companion (hybrid key switching scales by it)
    u64 tinv, tinv_s;           // t_plain^-1 mod q_last (BGV correction), used when has_t
    u64 half;                   // floor(q_last / 2)
    u32 has_t;
};

// splitmix64 finaliser, the synthetic-data hash of DESIGN.md §5
DPFHE_HD u64 splitmix64(u64 x) {
    u64 z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace dpfhe
