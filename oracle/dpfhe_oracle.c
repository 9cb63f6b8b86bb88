/*
 * dpfhe_oracle.c — CPU oracle (TEST INFRASTRUCTURE ONLY; see dpfhe_oracle.h).
 *
 * PARITY UNPINNED against the reference: /root/reference holds no implementation of
 * this path.  Every function below cites the DESIGN.md section
 * (the frozen spec) it restates instead of a reference file:line, and the nearest
 * reference anchor where one exists.
 *
 * Build: see oracle/Makefile (gcc -O3 -march=x86-64-v3 -fopenmp -shared).
 */
#include "dpfhe_oracle.h"

#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <time.h>

typedef unsigned __int128 u128;

struct dpo_ctx {
    unsigned logn, L;
    size_t N;
    uint64_t q[DPO_MAX_L];
    uint64_t psi[DPO_MAX_L];
    uint64_t inv_n[DPO_MAX_L], inv_n_shoup[DPO_MAX_L];
    uint64_t br0[DPO_MAX_L], br1[DPO_MAX_L];    /* floor(2^128/q) = br1:br0 */
    uint64_t *rp[DPO_MAX_L], *rps[DPO_MAX_L];   /* psi^bitrev(i) and Shoup companions */
    uint64_t *irp[DPO_MAX_L], *irps[DPO_MAX_L]; /* inverse of the above               */
};

/* ------------------------------------------------------------------ scalar arithmetic */

/* reference definition of a*b mod q. */
uint64_t dpo_mulmod_ref(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }

/* Barrett reduction of a 128-bit product with the two-word constant
 * floor(2^128/q); result canonical.  Valid for q < 2^63, a,b < q. */
static inline void barrett_ratio(uint64_t q, uint64_t *r0, uint64_t *r1) {
    u128 two64_div = (((u128)1 << 64) / q);           /* floor(2^64/q)               */
    u128 two64_rem = (((u128)1 << 64) % q);
    *r1 = (uint64_t)two64_div;                        /* high word of floor(2^128/q) */
    *r0 = (uint64_t)((two64_rem << 64) / q);          /* low word                    */
}
static inline uint64_t barrett_mul(uint64_t a, uint64_t b, uint64_t q, uint64_t r0, uint64_t r1) {
    u128 z = (u128)a * b;
    uint64_t z0 = (uint64_t)z, z1 = (uint64_t)(z >> 64);
    /* qhat = floor(z * ratio / 2^128) up to an error of 2; its low 64 bits suffice */
    u128 t = ((u128)z0 * r0) >> 64;
    t += (u128)z0 * r1;
    u128 carry = t >> 64;
    t = (uint64_t)t;
    t += (u128)z1 * r0;
    carry += t >> 64;
    uint64_t qhat = (uint64_t)(carry + (u128)z1 * r1);
    uint64_t r = z0 - qhat * q;
    while (r >= q) r -= q;
    return r;
}
uint64_t dpo_mulmod_barrett(uint64_t a, uint64_t b, uint64_t q) {
    uint64_t r0, r1;
    barrett_ratio(q, &r0, &r1);
    return barrett_mul(a, b, q, r0, r1);
}

uint64_t dpo_shoup_precompute(uint64_t w, uint64_t q) { return (uint64_t)(((u128)w << 64) / q); }

/* Shoup multiplication by a fixed operand w with wp = floor(w*2^64/q). */
static inline uint64_t shoup_lazy(uint64_t x, uint64_t w, uint64_t wp, uint64_t q) { /* [0,2q) */
    uint64_t qhat = (uint64_t)(((u128)x * wp) >> 64);
    return x * w - qhat * q;
}
uint64_t dpo_mulmod_shoup(uint64_t x, uint64_t w, uint64_t wp, uint64_t q) {
    uint64_t r = shoup_lazy(x, w, wp, q);
    return r >= q ? r - q : r;
}

static inline uint64_t addmod(uint64_t a, uint64_t b, uint64_t q) { uint64_t s = a + b; return s >= q ? s - q : s; }
static inline uint64_t submod(uint64_t a, uint64_t b, uint64_t q) { return a >= b ? a - b : a + q - b; }
static inline uint64_t mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }

uint64_t dpo_powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}
uint64_t dpo_invmod(uint64_t a, uint64_t q) { return dpo_powmod(a, q - 2, q); } /* q prime */

/* deterministic Miller-Rabin for 64-bit integers */
int dpo_is_prime(uint64_t n) {
    static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (size_t i = 0; i < 12; i++) {
        if (n == bases[i]) return 1;
        if (n % bases[i] == 0) return 0;
    }
    uint64_t d = n - 1;
    int r = 0;
    while (!(d & 1)) { d >>= 1; r++; }
    for (size_t i = 0; i < 12; i++) {
        uint64_t x = dpo_powmod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int comp = 1;
        for (int k = 1; k < r; k++) {
            x = mulmod(x, x, n);
            if (x == n - 1) { comp = 0; break; }
        }
        if (comp) return 0;
    }
    return 1;
}

static inline uint32_t bitrev(uint32_t x, unsigned bits) {
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* smallest primitive 2N-th root of unity mod q. */
static uint64_t smallest_primitive_root(uint64_t q, uint64_t two_n) {
    uint64_t cof = (q - 1) / two_n, root = 0;
    for (uint64_t g = 2;; g++) {
        uint64_t c = dpo_powmod(g, cof, q);
        if (dpo_powmod(c, two_n / 2, q) == q - 1) { root = c; break; }
    }
    /* all primitive 2N-th roots are root^k, k odd; take the minimum */
    uint64_t sq = mulmod(root, root, q), cur = root, best = root;
    for (uint64_t k = 1; k < two_n; k += 2) {
        if (cur < best) best = cur;
        cur = mulmod(cur, sq, q);
    }
    return best;
}

/* ------------------------------------------------------------------ context */

dpo_ctx *dpo_create(unsigned logn, unsigned L, const uint64_t *moduli) {
    if (logn < 1 || logn > 17 || L < 1 || L > DPO_MAX_L) return NULL;
    dpo_ctx *c = (dpo_ctx *)calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->logn = logn; c->L = L; c->N = (size_t)1 << logn;
    uint64_t two_n = (uint64_t)2 << logn;
    if (moduli) {
        for (unsigned l = 0; l < L; l++) {
            uint64_t q = moduli[l];
            if (q >= ((uint64_t)1 << 60) || q <= ((uint64_t)1 << 33) || (q - 1) % two_n || !dpo_is_prime(q)) { free(c); return NULL; }
            for (unsigned k = 0; k < l; k++) if (c->q[k] == q) { free(c); return NULL; }
            c->q[l] = q;
        }
    } else {
        /* the L largest primes below 2^60 of the form k * 2^32 + 1 (hence 1 mod 2N), descending */
        uint64_t cand = ((uint64_t)1 << 60) + 1;
        unsigned found = 0;
        while (found < L) {
            cand -= (uint64_t)1 << 32;
            if (dpo_is_prime(cand)) c->q[found++] = cand;
        }
    }
    for (unsigned l = 0; l < L; l++) {
        uint64_t q = c->q[l];
        c->psi[l] = smallest_primitive_root(q, two_n);
        c->rp[l] = (uint64_t *)malloc(c->N * 8); c->rps[l] = (uint64_t *)malloc(c->N * 8);
        c->irp[l] = (uint64_t *)malloc(c->N * 8); c->irps[l] = (uint64_t *)malloc(c->N * 8);
        uint64_t *pw = (uint64_t *)malloc(c->N * 8);
        pw[0] = 1;
        for (size_t i = 1; i < c->N; i++) pw[i] = mulmod(pw[i - 1], c->psi[l], q);
        for (size_t i = 0; i < c->N; i++) {
            uint64_t w = pw[bitrev((uint32_t)i, logn)];
            uint64_t wi = dpo_invmod(w, q);
            c->rp[l][i] = w;  c->rps[l][i] = dpo_shoup_precompute(w, q);
            c->irp[l][i] = wi; c->irps[l][i] = dpo_shoup_precompute(wi, q);
        }
        free(pw);
        barrett_ratio(q, &c->br0[l], &c->br1[l]);
        c->inv_n[l] = dpo_invmod((uint64_t)c->N % q, q);
        c->inv_n_shoup[l] = dpo_shoup_precompute(c->inv_n[l], q);
    }
    return c;
}

void dpo_destroy(dpo_ctx *c) {
    if (!c) return;
    for (unsigned l = 0; l < c->L; l++) { free(c->rp[l]); free(c->rps[l]); free(c->irp[l]); free(c->irps[l]); }
    free(c);
}
unsigned dpo_logn(const dpo_ctx *c) { return c->logn; }
unsigned dpo_L(const dpo_ctx *c) { return c->L; }
uint64_t dpo_modulus(const dpo_ctx *c, unsigned l) { return c->q[l]; }
uint64_t dpo_psi(const dpo_ctx *c, unsigned l) { return c->psi[l]; }
const uint64_t *dpo_root_powers(const dpo_ctx *c, unsigned l) { return c->rp[l]; }
const uint64_t *dpo_inv_root_powers(const dpo_ctx *c, unsigned l) { return c->irp[l]; }
uint64_t dpo_inv_n(const dpo_ctx *c, unsigned l) { return c->inv_n[l]; }

/* ------------------------------------------------------------------ transforms */

/* forward negacyclic NTT, Cooley-Tukey butterflies with the twist merged
 * (Harvey 2014), natural order in, bit-reversed order out, lazy [0,4q) inside, canonical out. */
static void ntt_fwd_limb(const dpo_ctx *c, unsigned l, uint64_t *a) {
    const uint64_t q = c->q[l], two_q = 2 * q;
    const uint64_t *rp = c->rp[l], *rps = c->rps[l];
    size_t N = c->N, t = N;
    for (size_t m = 1; m < N; m <<= 1) {
        t >>= 1;
        for (size_t i = 0; i < m; i++) {
            uint64_t w = rp[m + i], wp = rps[m + i];
            uint64_t *x = a + 2 * i * t, *y = x + t;
            for (size_t j = 0; j < t; j++) {
                uint64_t u = x[j]; u = u >= two_q ? u - two_q : u;
                uint64_t v = shoup_lazy(y[j], w, wp, q);
                x[j] = u + v;
                y[j] = u + two_q - v;
            }
        }
    }
    for (size_t j = 0; j < N; j++) {
        uint64_t u = a[j];
        u = u >= two_q ? u - two_q : u;
        a[j] = u >= q ? u - q : u;
    }
}

/* inverse, Gentleman-Sande butterflies, bit-reversed in, natural out, N^-1 folded in. */
static void ntt_inv_limb(const dpo_ctx *c, unsigned l, uint64_t *a) {
    const uint64_t q = c->q[l], two_q = 2 * q;
    const uint64_t *irp = c->irp[l], *irps = c->irps[l];
    size_t N = c->N, t = 1;
    for (size_t m = N; m > 1; m >>= 1) {
        size_t h = m >> 1;
        for (size_t i = 0; i < h; i++) {
            uint64_t w = irp[h + i], wp = irps[h + i];
            uint64_t *x = a + 2 * i * t, *y = x + t;
            for (size_t j = 0; j < t; j++) {
                uint64_t u = x[j], v = y[j];           /* both in [0,2q) */
                uint64_t s = u + v; s = s >= two_q ? s - two_q : s;
                x[j] = s;
                y[j] = shoup_lazy(u + two_q - v, w, wp, q);
            }
        }
        t <<= 1;
    }
    for (size_t j = 0; j < N; j++) a[j] = dpo_mulmod_shoup(a[j], c->inv_n[l], c->inv_n_shoup[l], q);
}

void dpo_ntt_fwd(const dpo_ctx *c, uint64_t *data, size_t n_polys) {
    long total = (long)(n_polys * c->L);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < total; k++) ntt_fwd_limb(c, (unsigned)(k % c->L), data + (size_t)k * c->N);
}
void dpo_ntt_inv(const dpo_ctx *c, uint64_t *data, size_t n_polys) {
    long total = (long)(n_polys * c->L);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < total; k++) ntt_inv_limb(c, (unsigned)(k % c->L), data + (size_t)k * c->N);
}

/* Definition check: out[i] = sum_k a[k] * psi^{k*(2*bitrev(i)+1)}, computed directly with % only. */
void dpo_ntt_fwd_limb_slow(const dpo_ctx *c, unsigned l, uint64_t *a) {
    size_t N = c->N;
    uint64_t q = c->q[l];
    uint64_t *out = (uint64_t *)malloc(N * 8);
    for (size_t i = 0; i < N; i++) {
        uint64_t e = 2 * (uint64_t)bitrev((uint32_t)i, c->logn) + 1;
        uint64_t x = dpo_powmod(c->psi[l], e, q), xp = 1, acc = 0;
        for (size_t k = 0; k < N; k++) {
            acc = addmod(acc, mulmod(a[k] % q, xp, q), q);
            xp = mulmod(xp, x, q);
        }
        out[i] = acc;
    }
    memcpy(a, out, N * 8);
    free(out);
}

/* Definition check: out = a*b mod (X^N + 1, q) by the O(N^2) schoolbook rule. */
void dpo_negacyclic_schoolbook(const dpo_ctx *c, unsigned l, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    size_t N = c->N;
    uint64_t q = c->q[l];
    memset(out, 0, N * 8);
    for (size_t i = 0; i < N; i++) {
        if (!a[i]) continue;
        for (size_t j = 0; j < N; j++) {
            uint64_t p = mulmod(a[i], b[j], q);
            size_t k = i + j;
            if (k < N) out[k] = addmod(out[k], p, q);
            else out[k - N] = submod(out[k - N], p, q);
        }
    }
}

/* ------------------------------------------------------------------ evaluator ops */

/* DESIGN.md §2.4 poly_mul_pointwise */
void dpo_poly_mul_pointwise(const dpo_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n_polys) {
    long total = (long)(n_polys * c->L);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < total; k++) {
        unsigned l = (unsigned)(k % c->L);
        uint64_t q = c->q[l], r0 = c->br0[l], r1 = c->br1[l];
        size_t off = (size_t)k * c->N;
        for (size_t j = 0; j < c->N; j++) out[off + j] = barrett_mul(a[off + j], b[off + j], q, r0, r1);
    }
}

/* DESIGN.md §2.4 poly_add */
void dpo_poly_add(const dpo_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n_polys) {
    long total = (long)(n_polys * c->L);
#pragma omp parallel for schedule(static)
    for (long k = 0; k < total; k++) {
        uint64_t q = c->q[k % c->L];
        size_t off = (size_t)k * c->N;
        for (size_t j = 0; j < c->N; j++) out[off + j] = addmod(a[off + j], b[off + j], q);
    }
}

/* DESIGN.md §2.4 ct_tensor: d0 = a0*b0, d1 = a0*b1 + a1*b0, d2 = a1*b1 (pointwise, per limb). */
static void ct_tensor_limbs(const dpo_ctx *c, unsigned nl, const uint64_t *a, const uint64_t *b, uint64_t *d) {
    size_t P = nl * c->N;
    for (unsigned l = 0; l < nl; l++) {
        uint64_t q = c->q[l], r0 = c->br0[l], r1 = c->br1[l];
        for (size_t j = 0; j < c->N; j++) {
            size_t o = l * c->N + j;
            uint64_t a0 = a[o], a1 = a[P + o], b0 = b[o], b1 = b[P + o];
            d[o] = barrett_mul(a0, b0, q, r0, r1);
            d[P + o] = addmod(barrett_mul(a0, b1, q, r0, r1), barrett_mul(a1, b0, q, r0, r1), q);
            d[2 * P + o] = barrett_mul(a1, b1, q, r0, r1);
        }
    }
}
static void ct_tensor_one(const dpo_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *d) { ct_tensor_limbs(c, c->L, a, b, d); }
void dpo_ct_tensor(const dpo_ctx *c, const uint64_t *a, const uint64_t *b, uint64_t *d, size_t batch) {
    size_t P = c->L * c->N;
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)batch; k++) ct_tensor_one(c, a + 2 * P * k, b + 2 * P * k, d + 3 * P * k);
}

/* DESIGN.md §2.5 keyswitch: per-limb-digit (BV-RNS, dnum = L, no special prime).
 *   for digit j: t = INTT_j(d[j]) (integers in [0,q_j));
 *     for limb i: u = (i==j) ? d[j] : NTT_i(t mod q_i);
 *       c0[i] += u o key[j].b[i];  c1[i] += u o key[j].a[i]. */
void dpo_keyswitch(const dpo_ctx *c, const uint64_t *d, const uint64_t *key, uint64_t *c0, uint64_t *c1) {
    size_t N = c->N, P = c->L * N;
    uint64_t *t = (uint64_t *)malloc(N * 8), *u = (uint64_t *)malloc(N * 8);
    memset(c0, 0, P * 8);
    memset(c1, 0, P * 8);
    for (unsigned j = 0; j < c->L; j++) {
        memcpy(t, d + j * N, N * 8);
        ntt_inv_limb(c, j, t);
        for (unsigned i = 0; i < c->L; i++) {
            uint64_t q = c->q[i], r0 = c->br0[i], r1 = c->br1[i];
            const uint64_t *src;
            if (i == j) src = d + j * N;
            else {
                for (size_t n = 0; n < N; n++) u[n] = t[n] % q;
                ntt_fwd_limb(c, i, u);
                src = u;
            }
            const uint64_t *kb = key + ((size_t)j * 2 + 0) * P + i * N;
            const uint64_t *ka = key + ((size_t)j * 2 + 1) * P + i * N;
            for (size_t n = 0; n < N; n++) {
                c0[i * N + n] = addmod(c0[i * N + n], barrett_mul(src[n], kb[n], q, r0, r1), q);
                c1[i * N + n] = addmod(c1[i * N + n], barrett_mul(src[n], ka[n], q, r0, r1), q);
            }
        }
    }
    free(t);
    free(u);
}

/* DESIGN.md §2.6 ct_mul_relin: tensor, keyswitch(d2) with evk, add into (d0,d1). */
static void ct_mul_relin_one(const dpo_ctx *c, const uint64_t *a, const uint64_t *b, const uint64_t *evk, uint64_t *out) {
    size_t N = c->N, P = c->L * N;
    uint64_t *d = (uint64_t *)malloc(3 * P * 8), *k = (uint64_t *)malloc(2 * P * 8);
    ct_tensor_one(c, a, b, d);
    dpo_keyswitch(c, d + 2 * P, evk, k, k + P);
    for (unsigned l = 0; l < c->L; l++)
        for (size_t n = 0; n < N; n++) {
            size_t o = l * N + n;
            out[o] = addmod(d[o], k[o], c->q[l]);
            out[P + o] = addmod(d[P + o], k[P + o], c->q[l]);
        }
    free(d);
    free(k);
}
void dpo_ct_mul_relin(const dpo_ctx *c, const uint64_t *a, const uint64_t *b, const uint64_t *evk, uint64_t *out, size_t batch) {
    size_t P = c->L * c->N;
#pragma omp parallel for schedule(dynamic, 1)
    for (long k = 0; k < (long)batch; k++) ct_mul_relin_one(c, a + 2 * P * k, b + 2 * P * k, evk, out + 2 * P * k);
}

/* DESIGN.md §2.7 ct_mul_plain */
void dpo_ct_mul_plain(const dpo_ctx *c, const uint64_t *ct, const uint64_t *pt, uint64_t *out, size_t batch) {
    size_t N = c->N, P = c->L * N;
#pragma omp parallel for schedule(static)
    for (long k = 0; k < (long)(batch * 2); k++) {
        const uint64_t *src = ct + P * k;
        uint64_t *dst = out + P * k;
        for (unsigned l = 0; l < c->L; l++)
            for (size_t n = 0; n < N; n++) dst[l * N + n] = barrett_mul(src[l * N + n], pt[l * N + n], c->q[l], c->br0[l], c->br1[l]);
    }
}

/* DESIGN.md §2.7b plaintext inner products (inner loop of a baby-step/giant-step matrix-vector product):
 *   out[g][k] = sum_b steps[b][k] o pts[g][b];  steps [nb][batch][2][L][N], pts [ng][nb][L][N], out [ng][batch][2][L][N] */
void dpo_ct_mul_plain_inner(const dpo_ctx *c, const uint64_t *steps, size_t nb, const uint64_t *pts, size_t ng, uint64_t *out, size_t batch) {
    const size_t N = c->N, P = c->L * N;
#pragma omp parallel for schedule(static)
    for (long w = 0; w < (long)(ng * batch * 2); w++) {
        const size_t g = (size_t)w / (batch * 2), kc = (size_t)w % (batch * 2);   /* kc = k*2 + comp */
        uint64_t *dst = out + (g * batch * 2 + kc) * P;
        for (unsigned l = 0; l < c->L; l++)
            for (size_t n = 0; n < N; n++) {
                uint64_t acc = 0;
                for (size_t b = 0; b < nb; b++)
                    acc = addmod(acc, barrett_mul(steps[(b * batch * 2 + kc) * P + l * N + n], pts[(g * nb + b) * P + l * N + n],
                                                  c->q[l], c->br0[l], c->br1[l]), c->q[l]);
                dst[l * N + n] = acc;
            }
    }
}

/* sigma_g in evaluation form is the index permutation
 *   out[i] = in[pi(i)],  2*br(pi(i)) + 1 = g * (2*br(i) + 1) mod 2N. */
void dpo_galois_perm(const dpo_ctx *c, uint64_t g, uint32_t *perm) {
    uint64_t two_n = (uint64_t)2 << c->logn;
    for (size_t i = 0; i < c->N; i++) {
        uint64_t e = (g * (2 * (uint64_t)bitrev((uint32_t)i, c->logn) + 1)) % two_n;
        perm[i] = bitrev((uint32_t)((e - 1) >> 1), c->logn);
    }
}

void dpo_galois_coeff(const dpo_ctx *c, unsigned l, uint64_t g, const uint64_t *in, uint64_t *out) {
    uint64_t two_n = (uint64_t)2 << c->logn, q = c->q[l];
    for (size_t k = 0; k < c->N; k++) {
        uint64_t e = (k * g) % two_n;
        if (e < c->N) out[e] = in[k];
        else out[e - c->N] = in[k] ? q - in[k] : 0;
    }
}

/* DESIGN.md §2.8 rotate: (sigma(c0) + ks0, ks1) with ks = keyswitch(sigma(c1), gk). */
static void rotate_one(const dpo_ctx *c, const uint64_t *ct, const uint32_t *perm, const uint64_t *gk, uint64_t *out) {
    size_t N = c->N, P = c->L * N;
    uint64_t *p = (uint64_t *)malloc(2 * P * 8), *k = (uint64_t *)malloc(2 * P * 8);
    for (unsigned comp = 0; comp < 2; comp++)
        for (unsigned l = 0; l < c->L; l++)
            for (size_t n = 0; n < N; n++) p[comp * P + l * N + n] = ct[comp * P + l * N + perm[n]];
    dpo_keyswitch(c, p + P, gk, k, k + P);
    for (unsigned l = 0; l < c->L; l++)
        for (size_t n = 0; n < N; n++) {
            size_t o = l * N + n;
            out[o] = addmod(p[o], k[o], c->q[l]);
            out[P + o] = k[P + o];
        }
    free(p);
    free(k);
}
void dpo_rotate(const dpo_ctx *c, const uint64_t *ct, uint64_t g, const uint64_t *gk, uint64_t *out, size_t batch) {
    size_t P = c->L * c->N;
    uint32_t *perm = (uint32_t *)malloc(c->N * 4);
    dpo_galois_perm(c, g, perm);
#pragma omp parallel for schedule(dynamic, 1)
    for (long k = 0; k < (long)batch; k++) rotate_one(c, ct + 2 * P * k, perm, gk, out + 2 * P * k);
    free(perm);
}

/* DESIGN.md §2.9 mod_switch_down: drop the last limb q_last of every polynomial (BGV modulus switch when
 * t_plain > 0, plain rounding "rescale" when t_plain == 0):
 *   tau = INTT_last(c[L-1]);  w = centred(tau * t^-1 mod q_last)   (t = 0: w = centred(tau), s = 1; else s = t)
 *   out[i] = (c[i] - s * NTT_i(w mod q_i)) * q_last^-1  mod q_i,   i < L-1.
 * in: [n_polys][L][N], out: [n_polys][L-1][N], both evaluation form. */
void dpo_mod_switch_down(const dpo_ctx *c, const uint64_t *in, uint64_t t_plain, uint64_t *out, size_t n_polys) {
    const size_t N = c->N;
    const unsigned L = c->L;
    if (L < 2) return;
    const uint64_t ql = c->q[L - 1], half = ql >> 1;
    const uint64_t tinv = t_plain ? dpo_invmod(t_plain % ql, ql) : 1;
#pragma omp parallel for schedule(dynamic, 1)
    for (long pidx = 0; pidx < (long)n_polys; pidx++) {
        uint64_t *tau = (uint64_t *)malloc(N * 8), *u = (uint64_t *)malloc(N * 8);
        const uint64_t *src = in + (size_t)pidx * L * N;
        memcpy(tau, src + (size_t)(L - 1) * N, N * 8);
        ntt_inv_limb(c, L - 1, tau);
        if (t_plain)
            for (size_t n = 0; n < N; n++) tau[n] = mulmod(tau[n], tinv, ql);
        for (unsigned i = 0; i + 1 < L; i++) {
            const uint64_t q = c->q[i];
            const uint64_t inv = dpo_invmod(ql % q, q), sfac = t_plain ? t_plain % q : 1, qlm = ql % q;
            for (size_t n = 0; n < N; n++) {
                uint64_t r = tau[n] % q;
                if (tau[n] > half) r = submod(r, qlm, q);      /* centred lift: tau - q_last */
                u[n] = r;
            }
            ntt_fwd_limb(c, i, u);
            uint64_t *dst = out + ((size_t)pidx * (L - 1) + i) * N;
            for (size_t n = 0; n < N; n++)
                dst[n] = mulmod(submod(src[(size_t)i * N + n], mulmod(u[n], sfac, q), q), inv, q);
        }
        free(tau);
        free(u);
    }
}

/* DESIGN.md §2.10 hybrid key switching with one special prime p = q[L-1] (GHS variant, one digit per ciphertext
 * limb).  Ciphertext polynomials carry Lq = L-1 limbs; the switch key is [Lq digits][2][L][N] over all L limbs and
 * encrypts p * g_j * target (dpo_keygen_switch_hybrid).
 *   acc_c[i] = sum_j u_ji o key[j][c][i]   (i < L; u_ji = d[j] if i == j, else NTT_i(INTT_j(d[j]) mod q_i))
 *   (c0, c1) = mod_switch_down(acc_0, acc_1)   -- divides the switched pair, and its noise, by p. */
void dpo_keyswitch_hybrid(const dpo_ctx *c, const uint64_t *d, const uint64_t *key, uint64_t t_plain, uint64_t *c0, uint64_t *c1) {
    const size_t N = c->N, PK = c->L * N;
    const unsigned Lq = c->L - 1;
    uint64_t *t = (uint64_t *)malloc(N * 8), *u = (uint64_t *)malloc(N * 8);
    uint64_t *acc = (uint64_t *)calloc(2 * PK, 8), *low = (uint64_t *)malloc(2 * (size_t)Lq * N * 8);
    for (unsigned j = 0; j < Lq; j++) {
        memcpy(t, d + j * N, N * 8);
        ntt_inv_limb(c, j, t);
        for (unsigned i = 0; i < c->L; i++) {
            uint64_t q = c->q[i], r0 = c->br0[i], r1 = c->br1[i];
            const uint64_t *src;
            if (i == j) src = d + j * N;
            else {
                for (size_t n = 0; n < N; n++) u[n] = t[n] % q;
                ntt_fwd_limb(c, i, u);
                src = u;
            }
            const uint64_t *kb = key + ((size_t)j * 2 + 0) * PK + i * N;
            const uint64_t *ka = key + ((size_t)j * 2 + 1) * PK + i * N;
            for (size_t n = 0; n < N; n++) {
                acc[i * N + n] = addmod(acc[i * N + n], barrett_mul(src[n], kb[n], q, r0, r1), q);
                acc[PK + i * N + n] = addmod(acc[PK + i * N + n], barrett_mul(src[n], ka[n], q, r0, r1), q);
            }
        }
    }
    dpo_mod_switch_down(c, acc, t_plain, low, 2);
    memcpy(c0, low, (size_t)Lq * N * 8);
    memcpy(c1, low + (size_t)Lq * N, (size_t)Lq * N * 8);
    free(t); free(u); free(acc); free(low);
}

static void ct_mul_relin_hybrid_one(const dpo_ctx *c, const uint64_t *a, const uint64_t *b, const uint64_t *evk, uint64_t t_plain, uint64_t *out) {
    const unsigned Lq = c->L - 1;
    const size_t N = c->N, P = (size_t)Lq * N;
    uint64_t *d = (uint64_t *)malloc(3 * P * 8), *k = (uint64_t *)malloc(2 * P * 8);
    ct_tensor_limbs(c, Lq, a, b, d);
    dpo_keyswitch_hybrid(c, d + 2 * P, evk, t_plain, k, k + P);
    for (unsigned l = 0; l < Lq; l++)
        for (size_t n = 0; n < N; n++) {
            size_t o = l * N + n;
            out[o] = addmod(d[o], k[o], c->q[l]);
            out[P + o] = addmod(d[P + o], k[P + o], c->q[l]);
        }
    free(d); free(k);
}
/* a, b, out: [batch][2][L-1][N] */
void dpo_ct_mul_relin_hybrid(const dpo_ctx *c, const uint64_t *a, const uint64_t *b, const uint64_t *evk, uint64_t t_plain,
                             uint64_t *out, size_t batch) {
    if (c->L < 2) return;
    size_t P = (size_t)(c->L - 1) * c->N;
#pragma omp parallel for schedule(dynamic, 1)
    for (long k = 0; k < (long)batch; k++) ct_mul_relin_hybrid_one(c, a + 2 * P * k, b + 2 * P * k, evk, t_plain, out + 2 * P * k);
}

/* ct, out: [batch][2][L-1][N] */
void dpo_rotate_hybrid(const dpo_ctx *c, const uint64_t *ct, uint64_t g, const uint64_t *gk, uint64_t t_plain, uint64_t *out, size_t batch) {
    if (c->L < 2) return;
    const unsigned Lq = c->L - 1;
    const size_t N = c->N, P = (size_t)Lq * N;
    uint32_t *perm = (uint32_t *)malloc(N * 4);
    dpo_galois_perm(c, g, perm);
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < (long)batch; b++) {
        const uint64_t *src = ct + 2 * P * b;
        uint64_t *dst = out + 2 * P * b;
        uint64_t *p = (uint64_t *)malloc(2 * P * 8), *k = (uint64_t *)malloc(2 * P * 8);
        for (unsigned comp = 0; comp < 2; comp++)
            for (unsigned l = 0; l < Lq; l++)
                for (size_t n = 0; n < N; n++) p[comp * P + l * N + n] = src[comp * P + l * N + perm[n]];
        dpo_keyswitch_hybrid(c, p + P, gk, t_plain, k, k + P);
        for (unsigned l = 0; l < Lq; l++)
            for (size_t n = 0; n < N; n++) {
                size_t o = l * N + n;
                dst[o] = addmod(p[o], k[o], c->q[l]);
                dst[P + o] = k[P + o];
            }
        free(p); free(k);
    }
    free(perm);
}

/* ------------------------------------------------------------------ grouped hybrid key switching (dnum < L)
 * DESIGN.md §2.11.  The last K limbs of the context are special primes p_0..p_{K-1} (P = prod p_k); ciphertext
 * polynomials carry Lq = L - K limbs, grouped into dnum = ceil(Lq / K) digits of alpha = K consecutive limbs (the last
 * one may be shorter).  Q_g = product of the moduli of group g, Qhat_j = Q_g / q_j for a member j.
 *   mod-up   (fast basis conversion, no correction: the lift is t + e*Q_g with 0 <= e < alpha):
 *       y_j = INTT_j(d[j]) * Qhat_j^-1 mod q_j;   u_gi = NTT_i( sum_{j in g} y_j * (Qhat_j mod q_i)  mod q_i )   (i not in g)
 *       u_gi = d[i]                                                                                            (i in g)
 *   acc_c[i] = sum_g u_gi o key[g][c][i]                       key: [dnum][2][L][N], encrypts P * (1 on group g, 0 elsewhere) * target
 *   mod-down (every special residue lifted centred, so delta lies in (-K P / 2, K P / 2)):
 *       y_k = INTT(acc[Lq+k]) * (t * Phat_k)^-1 mod p_k  (t = 0: Phat_k^-1),  Phat_k = P / p_k
 *       w_i = sum_k ( y_k * (Phat_k mod q_i) - [y_k > p_k/2] * (P mod q_i) )  mod q_i
 *       out[i] = (acc[i] - s * NTT_i(w_i)) * P^-1 mod q_i        (s = t, or 1 when t = 0)
 * With K = 1 every formula reduces to dpo_mod_switch_down / dpo_keyswitch_hybrid above (tests check the equality). */
static uint64_t prod_mod(const dpo_ctx *c, unsigned lo, unsigned hi, unsigned skip, uint64_t q) {
    uint64_t r = 1 % q;
    for (unsigned m = lo; m < hi; m++)
        if (m != skip) r = mulmod(r, c->q[m] % q, q);
    return r;
}

/* in: [n_polys][L][N], out: [n_polys][L-K][N], evaluation form */
void dpo_mod_down_special(const dpo_ctx *c, unsigned K, const uint64_t *in, uint64_t t_plain, uint64_t *out, size_t n_polys) {
    const size_t N = c->N;
    const unsigned L = c->L;
    if (K < 1 || K >= L) return;
    const unsigned Lq = L - K;
#pragma omp parallel for schedule(dynamic, 1)
    for (long pidx = 0; pidx < (long)n_polys; pidx++) {
        uint64_t *y = (uint64_t *)malloc((size_t)K * N * 8), *u = (uint64_t *)malloc(N * 8);
        const uint64_t *src = in + (size_t)pidx * L * N;
        for (unsigned k = 0; k < K; k++) {
            const unsigned s = Lq + k;
            const uint64_t p = c->q[s];
            uint64_t f = prod_mod(c, Lq, L, s, p);                       /* Phat_k mod p_k */
            if (t_plain) f = mulmod(f, t_plain % p, p);
            f = dpo_invmod(f, p);
            memcpy(y + (size_t)k * N, src + (size_t)s * N, N * 8);
            ntt_inv_limb(c, s, y + (size_t)k * N);
            for (size_t n = 0; n < N; n++) y[(size_t)k * N + n] = mulmod(y[(size_t)k * N + n], f, p);
        }
        for (unsigned i = 0; i < Lq; i++) {
            const uint64_t q = c->q[i];
            const uint64_t Pm = prod_mod(c, Lq, L, L, q), inv = dpo_invmod(Pm, q), sfac = t_plain ? t_plain % q : 1;
            for (size_t n = 0; n < N; n++) u[n] = 0;
            for (unsigned k = 0; k < K; k++) {
                const uint64_t p = c->q[Lq + k], half = p >> 1, ph = prod_mod(c, Lq, L, Lq + k, q);
                for (size_t n = 0; n < N; n++) {
                    const uint64_t v = y[(size_t)k * N + n];
                    uint64_t r = mulmod(v % q, ph, q);
                    if (v > half) r = submod(r, Pm, q);                  /* centred: (v - p_k) * Phat_k */
                    u[n] = addmod(u[n], r, q);
                }
            }
            ntt_fwd_limb(c, i, u);
            uint64_t *dst = out + ((size_t)pidx * Lq + i) * N;
            for (size_t n = 0; n < N; n++)
                dst[n] = mulmod(submod(src[(size_t)i * N + n], mulmod(u[n], sfac, q), q), inv, q);
        }
        free(y);
        free(u);
    }
}

unsigned dpo_grouped_digits(const dpo_ctx *c, unsigned K) { return K >= 1 && K < c->L ? (c->L - K + K - 1) / K : 0; }

/* mod-up of one polynomial d [Lq][N]: U [dnum][L][N], U[g][i] = the lift of digit g in limb i, evaluation form (d[i] itself for
 * the limbs of the digit) */
static void grouped_mod_up(const dpo_ctx *c, unsigned K, const uint64_t *d, uint64_t *U) {
    const size_t N = c->N;
    const unsigned L = c->L, Lq = L - K, dnum = dpo_grouped_digits(c, K);
    uint64_t *y = (uint64_t *)malloc((size_t)K * N * 8);
    for (unsigned g = 0; g < dnum; g++) {
        const unsigned lo = g * K, hi = lo + K < Lq ? lo + K : Lq;
        for (unsigned j = lo; j < hi; j++) {
            const uint64_t qj = c->q[j], f = dpo_invmod(prod_mod(c, lo, hi, j, qj), qj);
            uint64_t *yj = y + (size_t)(j - lo) * N;
            memcpy(yj, d + (size_t)j * N, N * 8);
            ntt_inv_limb(c, j, yj);
            for (size_t n = 0; n < N; n++) yj[n] = mulmod(yj[n], f, qj);
        }
        for (unsigned i = 0; i < L; i++) {
            const uint64_t q = c->q[i];
            uint64_t *u = U + ((size_t)g * L + i) * N;
            if (i >= lo && i < hi) {
                memcpy(u, d + (size_t)i * N, N * 8);
                continue;
            }
            for (size_t n = 0; n < N; n++) u[n] = 0;
            for (unsigned j = lo; j < hi; j++) {
                const uint64_t qh = prod_mod(c, lo, hi, j, q);
                const uint64_t *yj = y + (size_t)(j - lo) * N;
                for (size_t n = 0; n < N; n++) u[n] = addmod(u[n], mulmod(yj[n] % q, qh, q), q);
            }
            ntt_fwd_limb(c, i, u);
        }
    }
    free(y);
}

/* acc [2][L][N] = sum_g U[g][.][perm] o key[g]  (perm == NULL: identity), then the division by P: c0, c1 [Lq][N] */
static void grouped_mac_and_down(const dpo_ctx *c, unsigned K, const uint64_t *U, const uint32_t *perm, const uint64_t *key, uint64_t t_plain,
                                 uint64_t *c0, uint64_t *c1) {
    const size_t N = c->N, PK = c->L * N;
    const unsigned L = c->L, Lq = L - K, dnum = dpo_grouped_digits(c, K);
    uint64_t *acc = (uint64_t *)calloc(2 * PK, 8), *low = (uint64_t *)malloc(2 * (size_t)Lq * N * 8);
    for (unsigned g = 0; g < dnum; g++)
        for (unsigned i = 0; i < L; i++) {
            const uint64_t q = c->q[i], r0 = c->br0[i], r1 = c->br1[i];
            const uint64_t *src = U + ((size_t)g * L + i) * N;
            const uint64_t *kb = key + ((size_t)g * 2 + 0) * PK + i * N;
            const uint64_t *ka = key + ((size_t)g * 2 + 1) * PK + i * N;
            for (size_t n = 0; n < N; n++) {
                const uint64_t u = src[perm ? perm[n] : n];
                acc[i * N + n] = addmod(acc[i * N + n], barrett_mul(u, kb[n], q, r0, r1), q);
                acc[PK + i * N + n] = addmod(acc[PK + i * N + n], barrett_mul(u, ka[n], q, r0, r1), q);
            }
        }
    dpo_mod_down_special(c, K, acc, t_plain, low, 2);
    memcpy(c0, low, (size_t)Lq * N * 8);
    memcpy(c1, low + (size_t)Lq * N, (size_t)Lq * N * 8);
    free(acc); free(low);
}

/* d: [Lq][N]; key: [dnum][2][L][N]; c0, c1: [Lq][N] */
void dpo_keyswitch_grouped(const dpo_ctx *c, unsigned K, const uint64_t *d, const uint64_t *key, uint64_t t_plain, uint64_t *c0, uint64_t *c1) {
    uint64_t *U = (uint64_t *)malloc((size_t)dpo_grouped_digits(c, K) * c->L * c->N * 8);
    grouped_mod_up(c, K, d, U);
    grouped_mac_and_down(c, K, U, NULL, key, t_plain, c0, c1);
    free(U);
}

/* a, b, out: [batch][2][Lq][N]; evk: [dnum][2][L][N] */
void dpo_ct_mul_relin_grouped(const dpo_ctx *c, unsigned K, const uint64_t *a, const uint64_t *b, const uint64_t *evk, uint64_t t_plain,
                              uint64_t *out, size_t batch) {
    if (K < 1 || K >= c->L) return;
    const unsigned Lq = c->L - K;
    const size_t N = c->N, P = (size_t)Lq * N;
#pragma omp parallel for schedule(dynamic, 1)
    for (long b_i = 0; b_i < (long)batch; b_i++) {
        uint64_t *d = (uint64_t *)malloc(3 * P * 8), *k = (uint64_t *)malloc(2 * P * 8);
        uint64_t *dst = out + 2 * P * b_i;
        ct_tensor_limbs(c, Lq, a + 2 * P * b_i, b + 2 * P * b_i, d);
        dpo_keyswitch_grouped(c, K, d + 2 * P, evk, t_plain, k, k + P);
        for (unsigned l = 0; l < Lq; l++)
            for (size_t n = 0; n < N; n++) {
                size_t o = l * N + n;
                dst[o] = addmod(d[o], k[o], c->q[l]);
                dst[P + o] = addmod(d[P + o], k[P + o], c->q[l]);
            }
        free(d); free(k);
    }
}

/* ct, out: [batch][2][Lq][N]; gk: [dnum][2][L][N] */
void dpo_rotate_grouped(const dpo_ctx *c, unsigned K, const uint64_t *ct, uint64_t g, const uint64_t *gk, uint64_t t_plain, uint64_t *out,
                        size_t batch) {
    if (K < 1 || K >= c->L) return;
    const unsigned Lq = c->L - K;
    const size_t N = c->N, P = (size_t)Lq * N;
    uint32_t *perm = (uint32_t *)malloc(N * 4);
    dpo_galois_perm(c, g, perm);
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < (long)batch; b++) {
        const uint64_t *src = ct + 2 * P * b;
        uint64_t *dst = out + 2 * P * b;
        uint64_t *p = (uint64_t *)malloc(2 * P * 8), *k = (uint64_t *)malloc(2 * P * 8);
        for (unsigned comp = 0; comp < 2; comp++)
            for (unsigned l = 0; l < Lq; l++)
                for (size_t n = 0; n < N; n++) p[comp * P + l * N + n] = src[comp * P + l * N + perm[n]];
        dpo_keyswitch_grouped(c, K, p + P, gk, t_plain, k, k + P);
        for (unsigned l = 0; l < Lq; l++)
            for (size_t n = 0; n < N; n++) {
                size_t o = l * N + n;
                dst[o] = addmod(p[o], k[o], c->q[l]);
                dst[P + o] = k[P + o];
            }
        free(p); free(k);
    }
    free(perm);
}

/* Hoisted rotations with grouped hybrid keys: n_rot rotations of the same ciphertexts share the mod-up of the
 * UNPERMUTED c1; rotation r applies its permutation to the lifted digits:
 *     acc_c[i] = sum_g perm_r(U[g][i]) o gk_r[g][c][i],   out_r = (perm_r(c0) + ks0, ks1),  (ks0, ks1) = acc / P.
 * perm_r(U[g]) is the residue vector of sigma_r(lift of digit g) - a lift of the rotated digit of the same size, but not the
 * canonical one dpo_rotate_grouped would build - so the result encrypts the same plaintext with the same noise bound, and is
 * NOT bit-identical to dpo_rotate_grouped.
 * ct: [batch][2][Lq][N]; gks: [n_rot][dnum][2][L][N]; out: [n_rot][batch][2][Lq][N] */
void dpo_rotate_hoisted_grouped(const dpo_ctx *c, unsigned K, const uint64_t *ct, size_t n_rot, const uint64_t *galois, const uint64_t *gks,
                                uint64_t t_plain, uint64_t *out, size_t batch) {
    if (K < 1 || K >= c->L) return;
    const unsigned L = c->L, Lq = L - K, dnum = dpo_grouped_digits(c, K);
    const size_t N = c->N, P = (size_t)Lq * N, key_words = (size_t)dnum * 2 * L * N;
    uint32_t *perm = (uint32_t *)malloc(n_rot * N * 4);
    for (size_t r = 0; r < n_rot; r++) dpo_galois_perm(c, galois[r], perm + r * N);
#pragma omp parallel for schedule(dynamic, 1)
    for (long b = 0; b < (long)batch; b++) {
        const uint64_t *src = ct + 2 * P * b;
        uint64_t *U = (uint64_t *)malloc((size_t)dnum * L * N * 8), *k = (uint64_t *)malloc(2 * P * 8);
        grouped_mod_up(c, K, src + P, U);
        for (size_t r = 0; r < n_rot; r++) {
            const uint32_t *pr = perm + r * N;
            uint64_t *dst = out + (r * batch + (size_t)b) * 2 * P;
            grouped_mac_and_down(c, K, U, pr, gks + r * key_words, t_plain, k, k + P);
            for (unsigned l = 0; l < Lq; l++)
                for (size_t n = 0; n < N; n++) {
                    size_t o = l * N + n;
                    dst[o] = addmod(src[l * N + pr[n]], k[o], c->q[l]);
                    dst[P + o] = k[P + o];
                }
        }
        free(U); free(k);
    }
    free(perm);
}

/* ------------------------------------------------------------------ synthetic data */

uint64_t dpo_splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* counter-based so the GPU can regenerate the identical stream in parallel. */
void dpo_fill_uniform(const dpo_ctx *c, uint64_t seed, uint64_t first_poly, uint64_t *data, size_t n_polys) {
    size_t N = c->N, P = c->L * N;
#pragma omp parallel for schedule(static)
    for (long p = 0; p < (long)n_polys; p++)
        for (unsigned l = 0; l < c->L; l++)
            for (size_t n = 0; n < N; n++) {
                uint64_t k = (first_poly + (uint64_t)p) * P + l * N + n;
                uint64_t h = dpo_splitmix64(seed + k);
                data[(size_t)p * P + l * N + n] = (uint64_t)(((u128)h * c->q[l]) >> 64);
            }
}

/* ------------------------------------------------------------------ BGV-style scheme (tests only) */

typedef struct { uint64_t s[4]; } xo_t;
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static void xo_seed(xo_t *x, uint64_t seed) {
    for (int i = 0; i < 4; i++) { x->s[i] = dpo_splitmix64(seed); seed += 0x9E3779B97F4A7C15ull; }
}
static uint64_t xo_next(xo_t *x) { /* xoshiro256** */
    uint64_t r = rotl(x->s[1] * 5, 7) * 9, t = x->s[1] << 17;
    x->s[2] ^= x->s[0]; x->s[3] ^= x->s[1]; x->s[1] ^= x->s[2]; x->s[0] ^= x->s[3];
    x->s[2] ^= t; x->s[3] = rotl(x->s[3], 45);
    return r;
}
static uint64_t xo_uniform(xo_t *x, uint64_t q) { /* rejection sampling in [0,q) */
    uint64_t mask = ~(uint64_t)0 >> __builtin_clzll(q), r;
    do r = xo_next(x) & mask; while (r >= q);
    return r;
}
static int64_t xo_cbd(xo_t *x) { /* centred binomial, eta = 21, sigma ~ 3.24 */
    uint64_t r = xo_next(x);
    return (int64_t)__builtin_popcountll(r & 0x1FFFFF) - (int64_t)__builtin_popcountll((r >> 21) & 0x1FFFFF);
}

/* small signed polynomial -> RNS evaluation form */
static void small_to_eval(const dpo_ctx *c, const int64_t *v, uint64_t scale, uint64_t *out) {
    for (unsigned l = 0; l < c->L; l++) {
        uint64_t q = c->q[l];
        for (size_t n = 0; n < c->N; n++) {
            int64_t x = v[n];
            uint64_t r = x >= 0 ? (uint64_t)x % q : q - ((uint64_t)(-x) % q);
            if (r == q) r = 0;
            out[l * c->N + n] = mulmod(r, scale % q, q);
        }
        ntt_fwd_limb(c, l, out + l * c->N);
    }
}

void dpo_keygen_secret(const dpo_ctx *c, uint64_t seed, uint64_t *s_eval) {
    xo_t x; xo_seed(&x, seed);
    int64_t *v = (int64_t *)malloc(c->N * 8);
    for (size_t n = 0; n < c->N; n++) v[n] = (int64_t)(xo_next(&x) % 3) - 1;
    small_to_eval(c, v, 1, s_eval);
    free(v);
}

/* key[j] = (b_j, a_j), b_j = -a_j*s + t*e_j + g_j*target, where g_j = delta_{ij} in limb i */
void dpo_keygen_switch(const dpo_ctx *c, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                       const uint64_t *target_eval, uint64_t *key) {
    size_t N = c->N, P = c->L * N;
    xo_t x; xo_seed(&x, seed);
    int64_t *e = (int64_t *)malloc(N * 8);
    uint64_t *e_eval = (uint64_t *)malloc(P * 8);
    for (unsigned j = 0; j < c->L; j++) {
        uint64_t *kb = key + ((size_t)j * 2 + 0) * P, *ka = key + ((size_t)j * 2 + 1) * P;
        for (unsigned l = 0; l < c->L; l++)
            for (size_t n = 0; n < N; n++) ka[l * N + n] = xo_uniform(&x, c->q[l]);
        for (size_t n = 0; n < N; n++) e[n] = xo_cbd(&x);
        small_to_eval(c, e, t_plain, e_eval);
        for (unsigned l = 0; l < c->L; l++) {
            uint64_t q = c->q[l];
            for (size_t n = 0; n < N; n++) {
                size_t o = l * N + n;
                uint64_t v = submod(e_eval[o], mulmod(ka[o], s_eval[o], q), q);
                if (l == j) v = addmod(v, target_eval[o], q);
                kb[o] = v;
            }
        }
    }
    free(e);
    free(e_eval);
}

/* hybrid key: digits j < L-1, limbs over all L moduli; b_j = -a_j*s + t*e_j + p*g_j*target,
 * i.e. (p mod q_j) * target in limb j and nothing in the other limbs (p*g_j = 0 mod p and mod q_i, i != j). */
void dpo_keygen_switch_hybrid(const dpo_ctx *c, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                              const uint64_t *target_eval, uint64_t *key) {
    size_t N = c->N, P = c->L * N;
    const uint64_t sp = c->q[c->L - 1];
    xo_t x; xo_seed(&x, seed);
    int64_t *e = (int64_t *)malloc(N * 8);
    uint64_t *e_eval = (uint64_t *)malloc(P * 8);
    for (unsigned j = 0; j + 1 < c->L; j++) {
        uint64_t *kb = key + ((size_t)j * 2 + 0) * P, *ka = key + ((size_t)j * 2 + 1) * P;
        for (unsigned l = 0; l < c->L; l++)
            for (size_t n = 0; n < N; n++) ka[l * N + n] = xo_uniform(&x, c->q[l]);
        for (size_t n = 0; n < N; n++) e[n] = xo_cbd(&x);
        small_to_eval(c, e, t_plain, e_eval);
        for (unsigned l = 0; l < c->L; l++) {
            uint64_t q = c->q[l];
            for (size_t n = 0; n < N; n++) {
                size_t o = l * N + n;
                uint64_t v = submod(e_eval[o], mulmod(ka[o], s_eval[o], q), q);
                if (l == j) v = addmod(v, mulmod(target_eval[o], sp % q, q), q);
                kb[o] = v;
            }
        }
    }
    free(e);
    free(e_eval);
}

/* grouped hybrid key: digits g < dnum, limbs over all L moduli; b_g = -a_g*s + t*e_g + P*F_g*target with
 * F_g = 1 on the limbs of group g and 0 on every other limb, i.e. (P mod q_l) * target in the limbs of the group and nothing
 * elsewhere.  K = 1 draws the same random stream as dpo_keygen_switch_hybrid. */
void dpo_keygen_switch_grouped(const dpo_ctx *c, unsigned K, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                               const uint64_t *target_eval, uint64_t *key) {
    size_t N = c->N, P = c->L * N;
    const unsigned L = c->L, Lq = L - K, dnum = dpo_grouped_digits(c, K);
    xo_t x; xo_seed(&x, seed);
    int64_t *e = (int64_t *)malloc(N * 8);
    uint64_t *e_eval = (uint64_t *)malloc(P * 8);
    for (unsigned g = 0; g < dnum; g++) {
        const unsigned lo = g * K, hi = lo + K < Lq ? lo + K : Lq;
        uint64_t *kb = key + ((size_t)g * 2 + 0) * P, *ka = key + ((size_t)g * 2 + 1) * P;
        for (unsigned l = 0; l < L; l++)
            for (size_t n = 0; n < N; n++) ka[l * N + n] = xo_uniform(&x, c->q[l]);
        for (size_t n = 0; n < N; n++) e[n] = xo_cbd(&x);
        small_to_eval(c, e, t_plain, e_eval);
        for (unsigned l = 0; l < L; l++) {
            uint64_t q = c->q[l];
            const uint64_t Pm = prod_mod(c, Lq, L, L, q);
            for (size_t n = 0; n < N; n++) {
                size_t o = l * N + n;
                uint64_t v = submod(e_eval[o], mulmod(ka[o], s_eval[o], q), q);
                if (l >= lo && l < hi) v = addmod(v, mulmod(target_eval[o], Pm, q), q);
                kb[o] = v;
            }
        }
    }
    free(e);
    free(e_eval);
}
void dpo_keygen_relin_grouped(const dpo_ctx *c, unsigned K, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t *evk) {
    size_t P = c->L * c->N;
    uint64_t *s2 = (uint64_t *)malloc(P * 8);
    for (unsigned l = 0; l < c->L; l++)
        for (size_t n = 0; n < c->N; n++) s2[l * c->N + n] = mulmod(s_eval[l * c->N + n], s_eval[l * c->N + n], c->q[l]);
    dpo_keygen_switch_grouped(c, K, seed, t_plain, s_eval, s2, evk);
    free(s2);
}
void dpo_keygen_galois_grouped(const dpo_ctx *c, unsigned K, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t g, uint64_t *gk) {
    size_t N = c->N, P = c->L * N;
    uint32_t *perm = (uint32_t *)malloc(N * 4);
    uint64_t *sg = (uint64_t *)malloc(P * 8);
    dpo_galois_perm(c, g, perm);
    for (unsigned l = 0; l < c->L; l++)
        for (size_t n = 0; n < N; n++) sg[l * N + n] = s_eval[l * N + perm[n]];
    dpo_keygen_switch_grouped(c, K, seed, t_plain, s_eval, sg, gk);
    free(perm);
    free(sg);
}

/* hybrid != 0: special-prime key layout [L-1][2][L][N] */
static void keygen_switch_any(const dpo_ctx *c, int hybrid, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                              const uint64_t *target_eval, uint64_t *key) {
    if (hybrid) dpo_keygen_switch_hybrid(c, seed, t_plain, s_eval, target_eval, key);
    else dpo_keygen_switch(c, seed, t_plain, s_eval, target_eval, key);
}
static void keygen_relin_any(const dpo_ctx *c, int hybrid, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t *evk);
static void keygen_galois_any(const dpo_ctx *c, int hybrid, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t g, uint64_t *gk);
void dpo_keygen_relin_hybrid(const dpo_ctx *c, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t *evk) {
    keygen_relin_any(c, 1, seed, t_plain, s_eval, evk);
}
void dpo_keygen_galois_hybrid(const dpo_ctx *c, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t g, uint64_t *gk) {
    keygen_galois_any(c, 1, seed, t_plain, s_eval, g, gk);
}

void dpo_keygen_relin(const dpo_ctx *c, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t *evk) {
    keygen_relin_any(c, 0, seed, t_plain, s_eval, evk);
}
static void keygen_relin_any(const dpo_ctx *c, int hybrid, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t *evk) {
    size_t P = c->L * c->N;
    uint64_t *s2 = (uint64_t *)malloc(P * 8);
    for (unsigned l = 0; l < c->L; l++)
        for (size_t n = 0; n < c->N; n++) s2[l * c->N + n] = mulmod(s_eval[l * c->N + n], s_eval[l * c->N + n], c->q[l]);
    keygen_switch_any(c, hybrid, seed, t_plain, s_eval, s2, evk);
    free(s2);
}

void dpo_keygen_galois(const dpo_ctx *c, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t g, uint64_t *gk) {
    keygen_galois_any(c, 0, seed, t_plain, s_eval, g, gk);
}
static void keygen_galois_any(const dpo_ctx *c, int hybrid, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t g, uint64_t *gk) {
    size_t N = c->N, P = c->L * N;
    uint32_t *perm = (uint32_t *)malloc(N * 4);
    uint64_t *sg = (uint64_t *)malloc(P * 8);
    dpo_galois_perm(c, g, perm);
    for (unsigned l = 0; l < c->L; l++)
        for (size_t n = 0; n < N; n++) sg[l * N + n] = s_eval[l * N + perm[n]];
    keygen_switch_any(c, hybrid, seed, t_plain, s_eval, sg, gk);
    free(perm);
    free(sg);
}

/* ct = (c0, c1) = (-a*s + t*e + m, a) */
void dpo_encrypt(const dpo_ctx *c, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, const uint64_t *msg, uint64_t *ct) {
    size_t N = c->N, P = c->L * N;
    xo_t x; xo_seed(&x, seed);
    int64_t *e = (int64_t *)malloc(N * 8);
    uint64_t *e_eval = (uint64_t *)malloc(P * 8), *m_eval = (uint64_t *)malloc(P * 8);
    for (unsigned l = 0; l < c->L; l++)
        for (size_t n = 0; n < N; n++) ct[P + l * N + n] = xo_uniform(&x, c->q[l]);
    for (size_t n = 0; n < N; n++) e[n] = xo_cbd(&x);
    small_to_eval(c, e, t_plain, e_eval);
    for (size_t n = 0; n < N; n++) e[n] = (int64_t)(msg[n] % t_plain);
    small_to_eval(c, e, 1, m_eval);
    for (unsigned l = 0; l < c->L; l++) {
        uint64_t q = c->q[l];
        for (size_t n = 0; n < N; n++) {
            size_t o = l * N + n;
            ct[o] = addmod(submod(e_eval[o], mulmod(ct[P + o], s_eval[o], q), q), m_eval[o], q);
        }
    }
    free(e); free(e_eval); free(m_eval);
}

void dpo_phase(const dpo_ctx *c, const uint64_t *s_eval, const uint64_t *ct, unsigned n_comp, uint64_t *phase) {
    size_t N = c->N, P = c->L * N;
    for (unsigned l = 0; l < c->L; l++) {
        uint64_t q = c->q[l];
        for (size_t n = 0; n < N; n++) {
            size_t o = l * N + n;
            uint64_t s = s_eval[o], v = addmod(ct[o], mulmod(ct[P + o], s, q), q);
            if (n_comp == 3) v = addmod(v, mulmod(ct[2 * P + o], mulmod(s, s, q), q), q);
            phase[o] = v;
        }
        ntt_inv_limb(c, l, phase + l * N);
    }
}

/* ------------------------------------------------------------------ timing helpers */

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}
int dpo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* processors OpenMP could use (independent of OMP_NUM_THREADS, which launchers such as torchrun set to 1); 1 without OpenMP */
int dpo_num_procs(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
double dpo_time_ct_mul_relin(const dpo_ctx *c, const uint64_t *a, const uint64_t *b, const uint64_t *evk,
                             uint64_t *out, size_t batch, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    double t0 = now_s();
    dpo_ct_mul_relin(c, a, b, evk, out, batch);
    return now_s() - t0;
}
double dpo_time_ntt_fwd(const dpo_ctx *c, uint64_t *data, size_t n_polys, int threads) {
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    double t0 = now_s();
    dpo_ntt_fwd(c, data, n_polys);
    return now_s() - t0;
}
