/*
 * dpfhe.h — C ABI of the B200-native FHE ciphertext-arithmetic engine (libdpfhe.so).
 *
 * This is the drop-in boundary for the hot path named by BASELINE.json:north_star:
 * RNS negacyclic NTT/INTT, Barrett pointwise multiply, key-switch / relinearise,
 * ct x ct, ct x pt and rotate.  The reference (deeppowers/deeppowers @1cf6449) has NO
 * FFI or operator for this path (SURVEY.md §0, §8a row a-0, §8b "Nothing calls an FHE
 * boundary today"), so each entry point below cites the reference interface whose
 * *conventions* it follows rather than one it replaces:
 *
 *   - device-bound context owning its tables, freed in destroy:
 *       hal::CUDADevice ctor/dtor        src/core/hal/cuda/cuda_device.cpp:18-41
 *   - cudaSetDevice at the top of every call, one default stream per device:
 *       src/core/hal/cuda/cuda_device.cpp:23,64,79
 *   - errors: the reference throws std::runtime_error from CUDA_CHECK
 *       (cuda_device.cpp:9-16); exceptions cannot cross a C ABI, so every call
 *       returns a status and the message is fetched with dpfhe_last_error();
 *       the C++ wrapper (include/deeppowers_fhe.hpp) re-throws std::runtime_error.
 *   - caller-owned data buffers, as hal::Tensor buffers are owned by their creator
 *       src/core/hal/cuda/cuda_tensor.cpp:36-58
 *   - the public API the examples include and that the wrapper attaches to:
 *       src/api/cpp/include/deeppowers.hpp:41-87
 *
 * Layouts (uint64 little-endian, row-major, all residues canonical in [0, q_l)):
 *   polynomial [L][N]; ciphertext [2][L][N] (c0,c1) in evaluation (NTT) form;
 *   batch [batch][2][L][N]; switch key [L digits][2 {b,a}][L limbs][N] evaluation form;
 *   plaintext [L][N] evaluation form.
 * Ring Z_q[X]/(X^N+1); forward NTT natural -> bit-reversed order, inverse the opposite
 * (DESIGN.md §2).  Inputs outside [0,q_l) give unspecified (but memory-safe) results.
 *
 * Threading: a context is bound to one device and is NOT thread-safe; use one context
 * per host thread / GPU (matches the reference's one-default-stream-per-device usage).
 * Different contexts may be used from different threads at the same time (dpfhe_multi_* does).
 * `stream` is a cudaStream_t passed as void* (NULL = the context's own stream).
 * Device-pointer entry points are asynchronous with respect to the host.  All calls on one context share
 * its scratch, so the library orders them itself: a call issued on a different stream than the previous
 * one first waits (on the device) for that previous call.  dpfhe_synchronize() waits for all of them.
 * There is no CPU fallback: without a usable CUDA device dpfhe_context_create fails.
 */
#ifndef DPFHE_H
#define DPFHE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPFHE_MAX_LIMBS 16

typedef struct dpfhe_ctx dpfhe_ctx;

typedef struct dpfhe_params {
    uint32_t log_n;         /* N = 1 << log_n; supported: 12, 13, 14                       */
    uint32_t n_limbs;       /* L in [1, DPFHE_MAX_LIMBS]                                   */
    const uint64_t *moduli; /* L distinct primes, 2^33 < q < 2^60, q = 1 mod 2N; NULL = the */
                            /* default basis: the L largest primes k*2^32+1 below 2^60      */
                            /* (DESIGN.md §2.1).  Bases made only of k*2^32+1 primes run    */
                            /* the faster kernel variant; any other basis the generic one.  */
} dpfhe_params;

enum {
    DPFHE_OK = 0,
    DPFHE_ERR_INVALID = -1,    /* bad argument / unsupported parameter set */
    DPFHE_ERR_CUDA = -2,       /* CUDA runtime error (message has file:line) */
    DPFHE_ERR_NOMEM = -3
};

/* thread-local message of the last failing call on this thread */
const char *dpfhe_last_error(void);
/* library / build identification, e.g. "dpfhe 0.1 sm_100a" */
const char *dpfhe_version(void);

/* ---- context ---- */
int dpfhe_context_create(const dpfhe_params *p, int device_id, dpfhe_ctx **out);
void dpfhe_context_destroy(dpfhe_ctx *ctx);
int dpfhe_get_modulus(const dpfhe_ctx *ctx, uint32_t limb, uint64_t *q);
int dpfhe_get_psi(const dpfhe_ctx *ctx, uint32_t limb, uint64_t *psi);
/* copies psi^bitrev(i), i<N (inverse: psi^-bitrev(i)) in natural table order to a HOST buffer of N words */
int dpfhe_get_root_powers(const dpfhe_ctx *ctx, uint32_t limb, int inverse, uint64_t *h_out);
/* bytes of device scratch the context holds (tables + pipeline scratch), for reporting */
size_t dpfhe_context_device_bytes(const dpfhe_ctx *ctx);
/* releases the scratch that grows with use (up to 4 GiB of hoisted-rotation transforms, host staging, ...) */
int dpfhe_context_trim(dpfhe_ctx *ctx);
/* the CUDA device the context is bound to */
int dpfhe_context_device(const dpfhe_ctx *ctx);
/* number of CUDA devices visible to this process (0 and an error status without a driver) */
int dpfhe_device_count(int *out);
/* waits for everything issued through this context, on whatever stream */
int dpfhe_synchronize(dpfhe_ctx *ctx);

/* ---- transforms: d_data is [n_polys][L][N], in place ---- */
int dpfhe_ntt_fwd(dpfhe_ctx *ctx, uint64_t *d_data, size_t n_polys, void *stream);
int dpfhe_ntt_inv(dpfhe_ctx *ctx, uint64_t *d_data, size_t n_polys, void *stream);

/* ---- pointwise ---- */
/* out[p][l][n] = a*b mod q_l, [n_polys][L][N]; out may alias a or b */
int dpfhe_poly_mul_pointwise(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out,
                             size_t n_polys, void *stream);
/* out = a + b mod q_l, [n_polys][L][N] (a ciphertext is two polynomials); out may alias a or b */
int dpfhe_poly_add(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out,
                   size_t n_polys, void *stream);
/* a,b: [batch][2][L][N] -> d: [batch][3][L][N] (d0,d1,d2) */
int dpfhe_ct_tensor(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_d,
                    size_t batch, void *stream);

/* ---- key switching ---- */
/* d: [batch][L][N] (evaluation form) -> out: [batch][2][L][N] = sum_j NTT(INTT(d[j])) o key[j] */
int dpfhe_keyswitch(dpfhe_ctx *ctx, const uint64_t *d_d, const uint64_t *d_key, uint64_t *d_out,
                    size_t batch, void *stream);
/* out = relinearise(a (x) b); a,b,out: [batch][2][L][N]; out must not alias a or b */
int dpfhe_ct_mul_relin(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, const uint64_t *d_evk,
                       uint64_t *d_out, size_t batch, void *stream);
/* out = (c0 o pt, c1 o pt); pt [L][N] shared by the batch; out may alias ct */
int dpfhe_ct_mul_plain(dpfhe_ctx *ctx, const uint64_t *d_ct, const uint64_t *d_pt, uint64_t *d_out,
                       size_t batch, void *stream);
/* acc += (c0 o pt, c1 o pt): fused multiply-accumulate for diagonal-method linear layers; acc [batch][2][L][N] */
int dpfhe_ct_mul_plain_acc(dpfhe_ctx *ctx, const uint64_t *d_ct, const uint64_t *d_pt, uint64_t *d_acc,
                           size_t batch, void *stream);
/* out = (sigma_g(c0) + ks0, ks1), ks = keyswitch(sigma_g(c1), gk); galois_elt odd in [1,2N);
 * out must not alias ct */
int dpfhe_rotate(dpfhe_ctx *ctx, const uint64_t *d_ct, uint64_t galois_elt, const uint64_t *d_gk,
                 uint64_t *d_out, size_t batch, void *stream);

/* rotation by k slots (k < 0: the other direction): dpfhe_rotate with the Galois element 5^k mod 2N, which
 * dpfhe_galois_element returns (SURVEY.md §8b declared rotate with an `int k`; the Galois-element form above
 * also covers the conjugation 2N-1) */
int dpfhe_galois_element(const dpfhe_ctx *ctx, int k, uint64_t *galois_elt);
int dpfhe_rotate_steps(dpfhe_ctx *ctx, const uint64_t *d_ct, int k, const uint64_t *d_gk, uint64_t *d_out,
                       size_t batch, void *stream);

/* ---- hoisted rotations: n_rot rotations of the SAME batch, out[r] = rotate(ct, galois_elts[r], gks[r]).
 *      galois_elts and d_gks are HOST arrays of n_rot entries (d_gks[r] is a device pointer to a key [L][2][L][N]);
 *      d_out is [n_rot][batch][2][L][N].  Bit-identical to n_rot dpfhe_rotate calls, but the digit decomposition and its
 *      L(L-1) forward transforms are computed once per ciphertext.  The context keeps up to 4 GiB of scratch. ---- */
int dpfhe_rotate_hoisted(dpfhe_ctx *ctx, const uint64_t *d_ct, size_t n_rot, const uint64_t *galois_elts,
                         const uint64_t *const *d_gks, uint64_t *d_out, size_t batch, void *stream);

/* ---- plaintext inner products (the inner loop of a baby-step/giant-step matrix-vector product):
 *      out[g][k] = sum_{b < n_steps} steps[b][k] o pts[g][b]   for g < n_groups, k < batch
 *      steps [n_steps][batch][2][L][N] ciphertext batches, pts [n_groups][n_steps][L][N] plaintexts (evaluation form,
 *      shared by the batch), out [n_groups][batch][2][L][N].  Bit-identical to dpfhe_ct_mul_plain followed by
 *      n_steps-1 dpfhe_ct_mul_plain_acc per group, but every ciphertext row is read once. n_steps <= 128. ---- */
int dpfhe_ct_mul_plain_inner(dpfhe_ctx *ctx, const uint64_t *d_steps, size_t n_steps, const uint64_t *d_pts, size_t n_groups,
                             uint64_t *d_out, size_t batch, void *stream);

/* ---- encrypted linear layer (BASELINE.json config 4): y = W x by baby-step/giant-step diagonals,
 *      y = sum_g rot_{g*baby}( sum_b D[g*baby + b] o rot_b(x) ), a composition of the calls above that runs entirely on
 *      the device.  h_diags [n_diags][L][N]: the diagonal plaintexts in evaluation form, diagonal g*baby + b pre-rotated by
 *      -g*baby (the caller encodes them so); n_diags a multiple of baby (<= 128).  h_gk_baby [baby-1][L][2][L][N]: Galois
 *      keys of the rotations by 1 .. baby-1 slots; h_gk_giant [L][2][L][N]: key of the rotation by `baby` slots.
 *      Weights and keys are uploaded once, at creation.  apply: d_ct, d_out [batch][2][L][N] device buffers (asynchronous);
 *      apply_host: host buffers, the batch pipelined in chunks (upload / compute / download overlapped; synchronous).
 *      Uses (baby-1) + (n_diags/baby - 1) rotations per ciphertext instead of n_diags - 1. ---- */
typedef struct dpfhe_linear dpfhe_linear;
int dpfhe_linear_create(dpfhe_ctx *ctx, const uint64_t *h_diags, size_t n_diags, size_t baby, const uint64_t *h_gk_baby,
                        const uint64_t *h_gk_giant, dpfhe_linear **out);
void dpfhe_linear_destroy(dpfhe_linear *layer);
int dpfhe_linear_apply(dpfhe_linear *layer, const uint64_t *d_ct, uint64_t *d_out, size_t batch, void *stream);
int dpfhe_linear_apply_host(dpfhe_linear *layer, const uint64_t *h_ct, uint64_t *h_out, size_t batch);

/* ---- modulus switching / rescale: drop the last limb of every polynomial.
 *      in [n_polys][L][N] -> out [n_polys][L-1][N] (a ciphertext is two polynomials), evaluation form.
 *      t_plain > 0: BGV modulus switch (the plaintext is scaled by q_last^-1 mod t); t_plain == 0: plain rounding.
 *      The result lives under the first L-1 moduli: evaluate it with a context created for those. ---- */
int dpfhe_mod_switch_down(dpfhe_ctx *ctx, const uint64_t *d_in, uint64_t *d_out, size_t n_polys, uint64_t t_plain,
                          void *stream);

/* ---- hybrid (special-prime) key switching.  The context's LAST limb is the special prime p:
 *      ciphertexts carry L-1 limbs ([batch][2][L-1][N]) and switch keys are [L-1 digits][2][L limbs][N], encrypting
 *      p * g_j * target.  The key-switched pair is accumulated over all L limbs and divided by p (rounding as in
 *      dpfhe_mod_switch_down, t_plain > 0 = BGV correction), which divides the key-switching noise by p.
 *      Same fused persistent kernel family as the calls above; outputs must not alias inputs. ---- */
int dpfhe_keyswitch_hybrid(dpfhe_ctx *ctx, const uint64_t *d_d, const uint64_t *d_key, uint64_t *d_out, size_t batch,
                           uint64_t t_plain, void *stream);
int dpfhe_ct_mul_relin_hybrid(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, const uint64_t *d_evk,
                              uint64_t *d_out, size_t batch, uint64_t t_plain, void *stream);
int dpfhe_rotate_hybrid(dpfhe_ctx *ctx, const uint64_t *d_ct, uint64_t galois_elt, const uint64_t *d_gk,
                        uint64_t *d_out, size_t batch, uint64_t t_plain, void *stream);

/* host-buffer forms of the three calls above (synchronous, pipelined like dpfhe_ct_mul_relin_host) */
int dpfhe_ct_mul_relin_hybrid_host(dpfhe_ctx *ctx, const uint64_t *h_a, const uint64_t *h_b, const uint64_t *h_evk,
                                   uint64_t *h_out, size_t batch, uint64_t t_plain);
int dpfhe_rotate_hybrid_host(dpfhe_ctx *ctx, const uint64_t *h_ct, uint64_t galois_elt, const uint64_t *h_gk,
                             uint64_t *h_out, size_t batch, uint64_t t_plain);
int dpfhe_mod_switch_down_host(dpfhe_ctx *ctx, const uint64_t *h_in, uint64_t *h_out, size_t n_polys, uint64_t t_plain);

/* ---- grouped hybrid key switching: digits of several limbs, dnum < L.  The context's last n_special = K
 *      limbs are special primes (P = their product); ciphertexts carry Lq = L-K limbs ([batch][2][Lq][N]), grouped into
 *      dnum = ceil(Lq / K) digits of K consecutive limbs (the last digit may be shorter).  Switch keys are
 *      [dnum][2][L limbs][N] and encrypt P * F_g * target, F_g = 1 on the limbs of digit g and 0 on the others.  Every digit
 *      is raised to all L limbs by fast basis conversion, accumulated against its key, and the pair is divided by P (rounding
 *      as in dpfhe_mod_switch_down, every special residue lifted centred).  Against one special prime this needs fewer
 *      transforms (24 instead of 30 per ct x ct at Lq = 4, K = 2) and keys of dnum instead of Lq digits.
 *      1 <= K <= 4, 2K <= L; K = 1 is the hybrid variant above, bit for bit.  dpfhe_grouped_digits returns dnum.
 *      The reference has no counterpart (SURVEY.md §8 row f-2 widening). ---- */
int dpfhe_grouped_digits(const dpfhe_ctx *ctx, unsigned n_special, unsigned *digits);
int dpfhe_keyswitch_grouped(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_d, const uint64_t *d_key, uint64_t *d_out,
                            size_t batch, uint64_t t_plain, void *stream);
int dpfhe_ct_mul_relin_grouped(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_a, const uint64_t *d_b,
                               const uint64_t *d_evk, uint64_t *d_out, size_t batch, uint64_t t_plain, void *stream);
int dpfhe_rotate_grouped(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_ct, uint64_t galois_elt, const uint64_t *d_gk,
                         uint64_t *d_out, size_t batch, uint64_t t_plain, void *stream);
/* n_rot rotations of the SAME ciphertexts with grouped hybrid keys (d_gks[r]: [dnum][2][L][N], Galois key of galois_elts[r]),
 * sharing the basis conversion and the forward transforms of c1 ("hoisting"): one mod-up per ciphertext,
 * then per rotation only multiply-accumulates over the L limbs and the division by P (2K + 2Lq transforms instead of all of
 * them).  d_out: [n_rot][batch][2][L-K][N].  A rotation permutes the lifted digits instead of lifting the permuted digits:
 * the results decrypt to the same plaintexts with the same noise bound as dpfhe_rotate_grouped but are not the same bits
 * (the parity tests check them against the oracle's restatement of exactly this definition).  Works for n_special = 1 (hybrid
 * keys) as well. */
int dpfhe_rotate_hoisted_grouped(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_ct, size_t n_rot, const uint64_t *galois_elts,
                                 const uint64_t *const *d_gks, uint64_t *d_out, size_t batch, uint64_t t_plain, void *stream);
/* division by the product of the last n_special limbs alone (the mod-down half of the calls above; n_special = 1 is
 * dpfhe_mod_switch_down): in [n_polys][L][N] -> out [n_polys][L - n_special][N], 1 <= n_special <= 4, n_special < L */
int dpfhe_mod_down_special(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_in, uint64_t *d_out, size_t n_polys,
                           uint64_t t_plain, void *stream);
int dpfhe_mod_down_special_host(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *h_in, uint64_t *h_out, size_t n_polys,
                                uint64_t t_plain);
int dpfhe_ct_mul_relin_grouped_host(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *h_a, const uint64_t *h_b,
                                    const uint64_t *h_evk, uint64_t *h_out, size_t batch, uint64_t t_plain);
int dpfhe_rotate_grouped_host(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *h_ct, uint64_t galois_elt,
                              const uint64_t *h_gk, uint64_t *h_out, size_t batch, uint64_t t_plain);

/* ---- synthetic data: x[k] = mulhi64(splitmix64(seed + k), q_limb),
 *      k = (first_poly + p)*L*N + l*N + n.  Fills [n_polys][L][N]. ---- */
int dpfhe_fill_uniform(dpfhe_ctx *ctx, uint64_t seed, uint64_t first_poly, uint64_t *d_data,
                       size_t n_polys, void *stream);

/* ---- host-buffer entry points (what a non-CUDA caller of the reference API would bind):
 *      H2D, compute and D2H are pipelined in chunks on the context's streams; synchronous. ---- */
int dpfhe_ntt_fwd_host(dpfhe_ctx *ctx, uint64_t *h_data, size_t n_polys);
int dpfhe_ntt_inv_host(dpfhe_ctx *ctx, uint64_t *h_data, size_t n_polys);
int dpfhe_ct_mul_relin_host(dpfhe_ctx *ctx, const uint64_t *h_a, const uint64_t *h_b, const uint64_t *h_evk,
                            uint64_t *h_out, size_t batch);
int dpfhe_ct_mul_plain_host(dpfhe_ctx *ctx, const uint64_t *h_ct, const uint64_t *h_pt, uint64_t *h_out,
                            size_t batch);
int dpfhe_rotate_host(dpfhe_ctx *ctx, const uint64_t *h_ct, uint64_t galois_elt, const uint64_t *h_gk,
                      uint64_t *h_out, size_t batch);
/* pinned host memory helpers so callers can reach full PCIe bandwidth */
int dpfhe_host_alloc(void **out, size_t bytes);
/* the same, with the pages placed on the NUMA node of the context's GPU (what multi-GPU hosts need: DESIGN.md §7);
 * *placed_node (may be NULL) = that node, or -1 if the placement could not be enforced */
int dpfhe_host_alloc_near(const dpfhe_ctx *ctx, void **out, size_t bytes, int *placed_node);
int dpfhe_host_free(void *p);   /* frees memory of either allocator */
/* NUMA node of the context's GPU (-1: unknown), and a helper that restricts the CALLING thread to that node's CPUs */
int dpfhe_device_numa_node(const dpfhe_ctx *ctx, int *node);
int dpfhe_bind_thread_near(const dpfhe_ctx *ctx, int *n_cpus);

/* ---- device buffers other GPUs can write into.  The output of dpfhe_ct_mul_relin / dpfhe_keyswitch / dpfhe_rotate is
 *      written exactly once, by the kernel's final stores, so `d_out` may be memory of ANOTHER GPU: a peer-mapped
 *      buffer of the same process (dpfhe_multi_*), or, with one process per GPU, a buffer exported by the owning
 *      process and opened here.  That is how a multi-GPU job gathers its result while it computes. ---- */
#define DPFHE_IPC_HANDLE_BYTES 64
int dpfhe_device_alloc(dpfhe_ctx *ctx, void **d_out, size_t bytes);   /* a cudaMalloc of its own on the context's device */
int dpfhe_device_free(dpfhe_ctx *ctx, void *d_ptr);
int dpfhe_ipc_export(dpfhe_ctx *ctx, const void *d_ptr, unsigned char handle[DPFHE_IPC_HANDLE_BYTES]);
int dpfhe_ipc_open(dpfhe_ctx *ctx, const unsigned char handle[DPFHE_IPC_HANDLE_BYTES], void **d_out);
int dpfhe_ipc_close(dpfhe_ctx *ctx, void *d_ptr);

/* ---- several GPUs in one process: one context per device, contiguous shards of the batch
 *      (the first batch % n shards hold one ciphertext more), no collective while computing.
 *      device_ids NULL = devices 0..n-1; n_devices <= 0 = all visible devices.  A device may be listed twice
 *      (two logical shards on one GPU).  Contrast with the reference's one-MPI-rank-per-GPU DistributedContext,
 *      src/core/distributed/distributed_context.cpp:242-250. ---- */
typedef struct dpfhe_multi dpfhe_multi;
int dpfhe_multi_create(const dpfhe_params *p, const int *device_ids, int n_devices, dpfhe_multi **out);
void dpfhe_multi_destroy(dpfhe_multi *m);
int dpfhe_multi_device_count(const dpfhe_multi *m);
dpfhe_ctx *dpfhe_multi_context(dpfhe_multi *m, int index);   /* borrowed: shard `index`'s context */
int dpfhe_multi_shard(const dpfhe_multi *m, size_t batch, int index, size_t *first, size_t *count);
/* host buffers [batch][2][L][N]: every device pipelines its own shard (H2D, compute, D2H) — no gather needed */
int dpfhe_multi_ct_mul_relin_host(dpfhe_multi *m, const uint64_t *h_a, const uint64_t *h_b, const uint64_t *h_evk,
                                  uint64_t *h_out, size_t batch);
/* the special-prime form (dpfhe_ct_mul_relin_grouped_host) sharded the same way; ciphertexts carry L - n_special limbs */
int dpfhe_multi_ct_mul_relin_grouped_host(dpfhe_multi *m, unsigned n_special, const uint64_t *h_a, const uint64_t *h_b,
                                          const uint64_t *h_evk, uint64_t *h_out, size_t batch, uint64_t t_plain);
int dpfhe_multi_rotate_host(dpfhe_multi *m, const uint64_t *h_ct, uint64_t galois_elt, const uint64_t *h_gk,
                            uint64_t *h_out, size_t batch);
/* device buffers: d_a[r], d_b[r] = shard r of the operands and d_evk[r] = the key, all on device r; the whole result
 * [batch][2][L][N] is gathered on the device of shard `root` (d_out_root), written there directly by every device's
 * kernel through NVLink while it computes.  Synchronous. */
int dpfhe_multi_ct_mul_relin_gather(dpfhe_multi *m, const uint64_t *const *d_a, const uint64_t *const *d_b,
                                    const uint64_t *const *d_evk, uint64_t *d_out_root, int root, size_t batch);

/* ---- diagnostics ---- */
/* number of kernel launches issued through this context since creation */
uint64_t dpfhe_launch_count(const dpfhe_ctx *ctx);
/* per-phase clock64 totals of the fused key-switch kernel (summed over CTAs, then cleared); needs the
 * context to have been created with DPFHE_KS_PROF set in the environment.  out16: 16 words. */
int dpfhe_debug_phase_cycles(dpfhe_ctx *ctx, uint64_t *out16);
/* name + launch geometry of the kernels behind an op, for bench/DESIGN reporting; returns bytes written */
int dpfhe_describe(const dpfhe_ctx *ctx, char *buf, size_t buf_len);

#ifdef __cplusplus
}
#endif
#endif /* DPFHE_H */
