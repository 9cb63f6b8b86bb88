/*
 * dpfhe_oracle.h — CPU oracle for the RNS negacyclic NTT / ct-mult hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (deeppowers_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and there only as
 * the checker / the timed CPU baseline.
 *
 * PARITY UNPINNED: the reference tree (/root/reference, deeppowers/deeppowers
 * @1cf6449) contains no Ciphertext / Evaluator / NTT / modular-arithmetic code
 * at all and no golden vectors for this path
 * (SURVEY.md §8c).  This oracle therefore restates the textbook constructions
 * frozen in DESIGN.md §2 (Harvey-style merged negacyclic NTT, Shoup/Barrett
 * modular multiply, per-limb-digit RNS key switching) and is pinned by its own
 * known-answer tests (tests/test_oracle_kat.py, tests/golden/) and cross-checked against sympy's
 * independent number-theory / NTT-convolution code (tests/test_oracle_vs_sympy.py).
 *
 * Layout everywhere: uint64 little-endian, row-major
 *   polynomial  : [L][N]            (limb-major)
 *   ciphertext  : [2][L][N]         (c0 then c1), evaluation (NTT) form
 *   batch       : [batch][2][L][N]
 *   switch key  : [L digits][2 {b,a}][L limbs][N], evaluation form
 * All stored residues are canonical, i.e. in [0, q_l).
 */
#ifndef DPFHE_ORACLE_H
#define DPFHE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPO_MAX_L 16

typedef struct dpo_ctx dpo_ctx;

/* moduli == NULL: derive the L largest primes < 2^60 with q = 1 (mod 2N). */
dpo_ctx *dpo_create(unsigned logn, unsigned L, const uint64_t *moduli);
void dpo_destroy(dpo_ctx *);
unsigned dpo_logn(const dpo_ctx *);
unsigned dpo_L(const dpo_ctx *);
uint64_t dpo_modulus(const dpo_ctx *, unsigned limb);
uint64_t dpo_psi(const dpo_ctx *, unsigned limb);          /* smallest primitive 2N-th root */
/* table accessors (N entries each): psi^bitrev(i), its inverse */
const uint64_t *dpo_root_powers(const dpo_ctx *, unsigned limb);
const uint64_t *dpo_inv_root_powers(const dpo_ctx *, unsigned limb);
uint64_t dpo_inv_n(const dpo_ctx *, unsigned limb);

/* ---- scalar arithmetic (three independent restatements, cross-checked) ---- */
uint64_t dpo_mulmod_ref(uint64_t a, uint64_t b, uint64_t q);      /* (u128)a*b % q       */
uint64_t dpo_mulmod_barrett(uint64_t a, uint64_t b, uint64_t q);  /* 2-word Barrett      */
uint64_t dpo_shoup_precompute(uint64_t w, uint64_t q);            /* floor(w*2^64/q)     */
uint64_t dpo_mulmod_shoup(uint64_t x, uint64_t w, uint64_t wp, uint64_t q); /* canonical */
uint64_t dpo_powmod(uint64_t a, uint64_t e, uint64_t q);
uint64_t dpo_invmod(uint64_t a, uint64_t q);
int dpo_is_prime(uint64_t n);

/* ---- transforms: data is [n_polys][L][N], in place ---- */
void dpo_ntt_fwd(const dpo_ctx *, uint64_t *data, size_t n_polys);   /* natural -> bit-reversed */
void dpo_ntt_inv(const dpo_ctx *, uint64_t *data, size_t n_polys);   /* bit-reversed -> natural */
/* slow, table-free restatements used only to pin the fast ones */
void dpo_ntt_fwd_limb_slow(const dpo_ctx *, unsigned limb, uint64_t *a);  /* direct O(N^2) evaluation */
void dpo_negacyclic_schoolbook(const dpo_ctx *, unsigned limb, const uint64_t *a,
                               const uint64_t *b, uint64_t *out);

/* ---- evaluator ops (evaluation form) ---- */
void dpo_poly_mul_pointwise(const dpo_ctx *, const uint64_t *a, const uint64_t *b,
                            uint64_t *out, size_t n_polys);
void dpo_poly_add(const dpo_ctx *, const uint64_t *a, const uint64_t *b, uint64_t *out, size_t n_polys);
/* a,b: [batch][2][L][N] -> d: [batch][3][L][N] */
void dpo_ct_tensor(const dpo_ctx *, const uint64_t *a, const uint64_t *b, uint64_t *d, size_t batch);
/* d: [L][N] eval form; key [L][2][L][N]; out c0,c1: [L][N] each (overwritten) */
void dpo_keyswitch(const dpo_ctx *, const uint64_t *d, const uint64_t *key, uint64_t *c0, uint64_t *c1);
void dpo_ct_mul_relin(const dpo_ctx *, const uint64_t *a, const uint64_t *b, const uint64_t *evk,
                      uint64_t *out, size_t batch);
/* pt: [L][N] eval form, shared by the batch */
void dpo_ct_mul_plain(const dpo_ctx *, const uint64_t *ct, const uint64_t *pt, uint64_t *out, size_t batch);
/* out[g][k] = sum_b steps[b][k] o pts[g][b]; steps [nb][batch][2][L][N], pts [ng][nb][L][N], out [ng][batch][2][L][N] */
void dpo_ct_mul_plain_inner(const dpo_ctx *, const uint64_t *steps, size_t nb, const uint64_t *pts, size_t ng, uint64_t *out,
                            size_t batch);
/* galois_elt odd in [1, 2N); gk is the switch key for sigma_g(s) */
void dpo_rotate(const dpo_ctx *, const uint64_t *ct, uint64_t galois_elt, const uint64_t *gk,
                uint64_t *out, size_t batch);
/* drop the last limb (BGV modulus switch for t_plain > 0, plain rounding for t_plain == 0);
 * in [n_polys][L][N] -> out [n_polys][L-1][N] */
void dpo_mod_switch_down(const dpo_ctx *, const uint64_t *in, uint64_t t_plain, uint64_t *out, size_t n_polys);
/* hybrid (special-prime) key switching, DESIGN.md §2.10: the context's LAST limb is the special prime p;
 * ciphertext polynomials carry L-1 limbs, keys are [L-1 digits][2][L][N].  t_plain as in dpo_mod_switch_down. */
void dpo_keyswitch_hybrid(const dpo_ctx *, const uint64_t *d, const uint64_t *key, uint64_t t_plain, uint64_t *c0, uint64_t *c1);
void dpo_ct_mul_relin_hybrid(const dpo_ctx *, const uint64_t *a, const uint64_t *b, const uint64_t *evk, uint64_t t_plain,
                             uint64_t *out, size_t batch);
void dpo_rotate_hybrid(const dpo_ctx *, const uint64_t *ct, uint64_t galois_elt, const uint64_t *gk, uint64_t t_plain,
                       uint64_t *out, size_t batch);
/* grouped hybrid key switching (dnum < L), DESIGN.md §2.11: the context's last K limbs are special primes, ciphertext
 * polynomials carry Lq = L-K limbs in digits of K consecutive limbs, keys are [dnum = ceil(Lq/K)][2][L][N].
 * K = 1 is the hybrid variant above, bit for bit. */
unsigned dpo_grouped_digits(const dpo_ctx *, unsigned K);
/* in [n_polys][L][N] -> out [n_polys][L-K][N]: division by the product of the special primes */
void dpo_mod_down_special(const dpo_ctx *, unsigned K, const uint64_t *in, uint64_t t_plain, uint64_t *out, size_t n_polys);
void dpo_keyswitch_grouped(const dpo_ctx *, unsigned K, const uint64_t *d, const uint64_t *key, uint64_t t_plain, uint64_t *c0, uint64_t *c1);
void dpo_ct_mul_relin_grouped(const dpo_ctx *, unsigned K, const uint64_t *a, const uint64_t *b, const uint64_t *evk, uint64_t t_plain,
                              uint64_t *out, size_t batch);
void dpo_rotate_grouped(const dpo_ctx *, unsigned K, const uint64_t *ct, uint64_t galois_elt, const uint64_t *gk, uint64_t t_plain,
                        uint64_t *out, size_t batch);
/* n_rot rotations of the same ciphertexts sharing the mod-up (hoisting); same plaintexts as dpo_rotate_grouped, not the same bits.
 * gks [n_rot][dnum][2][L][N], out [n_rot][batch][2][L-K][N] */
void dpo_rotate_hoisted_grouped(const dpo_ctx *, unsigned K, const uint64_t *ct, size_t n_rot, const uint64_t *galois, const uint64_t *gks,
                                uint64_t t_plain, uint64_t *out, size_t batch);
/* permutation table of sigma_g in evaluation form: out[i] = in[perm[i]] */
void dpo_galois_perm(const dpo_ctx *, uint64_t galois_elt, uint32_t *perm);
/* sigma_g in coefficient form on one limb: out(X) = in(X^g) */
void dpo_galois_coeff(const dpo_ctx *, unsigned limb, uint64_t galois_elt, const uint64_t *in, uint64_t *out);

/* ---- synthetic data: x[k] = mulhi64(splitmix64(seed + k), q_limb) ---- */
uint64_t dpo_splitmix64(uint64_t x);
/* fills [n_polys][L][N]; element counter k = first_poly*L*N + linear offset */
void dpo_fill_uniform(const dpo_ctx *, uint64_t seed, uint64_t first_poly, uint64_t *data, size_t n_polys);

/* ---- BGV-style scheme, test infrastructure for semantic checks ---- */
/* secret: ternary, returned in evaluation form [L][N] */
void dpo_keygen_secret(const dpo_ctx *, uint64_t seed, uint64_t *s_eval);
/* key for target polynomial `target_eval` ([L][N], e.g. s^2 or sigma_g(s)) under secret s */
void dpo_keygen_switch(const dpo_ctx *, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                       const uint64_t *target_eval, uint64_t *key);
void dpo_keygen_relin(const dpo_ctx *, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t *evk);
void dpo_keygen_galois(const dpo_ctx *, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                       uint64_t galois_elt, uint64_t *gk);
void dpo_keygen_switch_hybrid(const dpo_ctx *, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                              const uint64_t *target_eval, uint64_t *key);
void dpo_keygen_relin_hybrid(const dpo_ctx *, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t *evk);
void dpo_keygen_galois_hybrid(const dpo_ctx *, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                              uint64_t galois_elt, uint64_t *gk);
void dpo_keygen_switch_grouped(const dpo_ctx *, unsigned K, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                               const uint64_t *target_eval, uint64_t *key);
void dpo_keygen_relin_grouped(const dpo_ctx *, unsigned K, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval, uint64_t *evk);
void dpo_keygen_galois_grouped(const dpo_ctx *, unsigned K, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                               uint64_t galois_elt, uint64_t *gk);
/* msg: N coefficients in [0,t); ct out [2][L][N] eval form */
void dpo_encrypt(const dpo_ctx *, uint64_t seed, uint64_t t_plain, const uint64_t *s_eval,
                 const uint64_t *msg, uint64_t *ct);
/* phase = c0 + c1*s (+ c2*s^2 if n_comp==3), returned in coefficient form [L][N] (caller does CRT) */
void dpo_phase(const dpo_ctx *, const uint64_t *s_eval, const uint64_t *ct, unsigned n_comp, uint64_t *phase);

/* ---- timing helpers for bench.py (OpenMP over ciphertexts) ---- */
int dpo_max_threads(void);
int dpo_num_procs(void);   /* processors available to OpenMP, ignoring OMP_NUM_THREADS */
/* run ct_mul_relin on `batch` cts with `threads` OpenMP threads; returns seconds */
double dpo_time_ct_mul_relin(const dpo_ctx *, const uint64_t *a, const uint64_t *b, const uint64_t *evk,
                             uint64_t *out, size_t batch, int threads);
double dpo_time_ntt_fwd(const dpo_ctx *, uint64_t *data, size_t n_polys, int threads);

#ifdef __cplusplus
}
#endif
#endif
