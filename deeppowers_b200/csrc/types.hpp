// types.hpp — plain data types shared by every translation unit of libdpfhe.so (host code, both arithmetic variants).
//
// The device arithmetic exists in two compile-time variants:
//   gen   any modulus 2^33 < q < 2^60, q = 1 (mod 2N)
//   fast  every modulus of the context is q = qh * 2^32 + 1: a multiplication by q costs one 32-bit multiply-add
// kernels.cu is compiled once per variant (-DDPFHE_FAST=0 / 1); functions whose code depends on the variant live in
// namespace dpfhe::gen / dpfhe::fast (DPFHE_VNS), the types below are common to both.
#pragma once
#include <stddef.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define DPFHE_HD __host__ __device__ __forceinline__
#else
#define DPFHE_HD inline
#endif

#ifndef DPFHE_FAST
#define DPFHE_FAST 0
#endif
#if DPFHE_FAST
#define DPFHE_VNS fast
#else
#define DPFHE_VNS gen
#endif
// 0: exact Shoup / Barrett quotients; 2: the quotient estimates drop the low partial product and the carry of the
// middle ones (hi64 - {0,1,2}): one IMAD.WIDE and one add less per product, lazy bounds grow by 4q instead of 2q
#ifndef DPFHE_SHOUP_APPROX
#define DPFHE_SHOUP_APPROX 2
#endif

// 1: the ct x ct tensor product forms a0 b1 + a1 b0 from three 128-bit products (Karatsuba) instead of four
#ifndef DPFHE_TENSOR_KARATSUBA
#define DPFHE_TENSOR_KARATSUBA 1
#endif

namespace dpfhe {

typedef uint64_t u64;
typedef uint32_t u32;

struct alignas(16) U64x2 {
    u64 x, y;
};

// (w, floor(w*2^64/q)) pairs, 16 B each, so one 128-bit load fetches a twiddle.
typedef U64x2 Twiddle;

// bound (in units of q) of a lazy Shoup product, and of a lazy Barrett reduction of a product of canonical factors
constexpr int SB = DPFHE_SHOUP_APPROX == 0 ? 2 : 4;

// Per-limb constants (host-built in host_params.cpp; passed by value in the kernel parameter block).
struct alignas(16) LimbParams {
    u64 q;           // modulus
    u64 q2;          // 2q
    u64 qsb;         // SB * q: keeps the differences of lazy butterflies positive
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
    u32 nqh;         // fast moduli (q = qh * 2^32 + 1): -qh mod 2^32; 0 for any other modulus
    u32 pad_;
};

struct LimbTable {
    LimbParams lp[16];
};

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

// Grouped hybrid key switching: the last K limbs of the context are special primes (P = their product), the
// Lq = L - K ciphertext limbs form dnum = ceil(Lq / K) digits of K consecutive limbs.  Constants of one call, built on the
// host (host_params.cpp:build_group_consts) and passed by value in the kernel parameter block, so that every use is a
// constant-bank operand with a CTA-uniform index.  The MsConsts of such a call describe the division by P (inv = P^-1 ...).
constexpr int KS_MAX_SPECIAL = 4;
struct GroupConsts {
    u32 Lq, K, dnum, pad_;
    // limb parameters whose N^-1 (ninv, wninv) carries the factor that the basis conversion wants on the inverse transform's
    // output: Qhat_j^-1 mod q_j for a ciphertext limb j (Qhat_j = product of the other moduli of its group), and
    // (t * Phat_k)^-1 mod p_k for special limb Lq + k (Phat_k = P / p_k; t = 1 for plain rounding)
    LimbParams lp_up[16];
    u64 up[16][16], up_s[16][16];                                // [j][i]: Qhat_j mod q_i, with its Shoup companion (j < Lq, i < L)
    u64 dn[KS_MAX_SPECIAL][16], dn_s[KS_MAX_SPECIAL][16];      // [k][i]: Phat_k mod q_i (i < Lq)
    u64 neg_p[16];                                               // q_i - (P mod q_i): adding it subtracts P
    u64 half[KS_MAX_SPECIAL];                                    // floor(p_k / 2)
};

// ---- twiddle table layout (host_params.cpp writes it, ntt_core.cuh reads it) -----------------------------
// natural index of the twiddle of group i at stage s is 2^s + i.  Stages of the last
// register pass (s >= LOGN-4) are stored transposed so that lane-consecutive rows read
// consecutive table entries:  i = row * 2^u + j  ->  2^s + j * (N/16) + row,  u = s - (LOGN-4).
template <int LOGN>
DPFHE_HD int tw_pos(int s, int i) {
    if (s < LOGN - 4) return (1 << s) + i;
    const int u = s - (LOGN - 4);
    const int row = i >> u, j = i & ((1 << u) - 1);
    return (1 << s) + j * (1 << (LOGN - 4)) + row;
}

enum KsMode { KS_MUL_RELIN = 0, KS_PLAIN = 1, KS_ROTATE = 2 };
// tau' rows of a hybrid key-switching group are double-buffered by round parity (the division step runs one round late)
constexpr int KS_HYB_ROWS = 6;

// splitmix64 finaliser, the synthetic-data hash of DESIGN.md §5
DPFHE_HD u64 splitmix64(u64 x) {
    u64 z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace dpfhe
