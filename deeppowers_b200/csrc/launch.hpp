// launch.hpp — interface between the C ABI (abi.cu) and the kernel launchers (kernels.cu, compiled once per arithmetic
// variant: namespace dpfhe::gen for any moduli, dpfhe::fast for moduli of the form k * 2^32 + 1; types.hpp).
#pragma once
#include <cuda_runtime.h>

#include "types.hpp"

namespace dpfhe {

// device-resident state shared by all launches of one context
struct LaunchCtx {
    int device = 0;
    int num_sms = 0;
    u32 log_n = 0, L = 0;
    bool fast = false;                // every modulus is k * 2^32 + 1: launches go to the dpfhe::fast kernels
    bool lift_reduce = true;          // some modulus is at least twice another: digits are word-reduced when they change limb
    const LimbParams *lp = nullptr;   // [L] device copy (element-wise kernels)
    LimbTable lt;                     // host copy, passed by value to the transform kernels
    int rot_cfg = 0;                  // tuning variant of rot_apply_kernel (DPFHE_ROT_CFG)
    int ntt_cfg = 0;                  // tuning variant of the N=8192 transform kernel (DPFHE_NTT_CFG)
    int ntt_tma = 1;                  // inverse transforms (N <= 8192) fetch their limb with the TMA unit (DPFHE_NTT_TMA=0: thread copies)
    const Twiddle *tw = nullptr;      // [L][N] forward twiddles, device layout (ntt_core.cuh:tw_pos)
    const Twiddle *itw = nullptr;     // [L][N] inverse twiddles
    // fused key-switch pipeline
    u64 *ks_scratch = nullptr;        // [ks_slots][2][N] digit exchange buffers
    u64 *ks_acc = nullptr;            // [ks_slots][2][N] lazy accumulator rows of the resident work items (second half of ks_scratch's allocation)
    size_t ks_window_bytes = 0;       // bytes of ks_scratch + ks_acc an L2 access-policy window may cover
    size_t l2_persist_max = 0;        // cudaDevAttrMaxPersistingL2CacheSize
    int l2_persist = 0;               // DPFHE_L2_PERSIST: 1 = mark the fused kernel's scratch as persisting in L2
    u64 *ks_acc_hyb = nullptr;        // hybrid key switching: [ks_slots][2 parities][2][N], allocated at the first hybrid call
    u32 *ks_flags = nullptr;          // [ks_slots] monotonically increasing round counters
    u32 *ks_consumed = nullptr;       // [ks_slots] monotone counters: readers of a single-buffered digit slot that have finished
    int ks_single = 0;                // tuning (DPFHE_KS_SINGLE=1): digit slots single-buffered with consumed counters instead of two per CTA
                                      // (by round parity); measured neutral in time and DRAM traffic (profiles/r02/ks_single.txt)
    u32 *ks_ticket = nullptr;         // next ciphertext index (reset per launch)
    u64 *ks_mail = nullptr;           // [ks_slots] per-group mailbox: (round tag << 32) | ciphertext index
    u64 *ks_key_s = nullptr;          // [L][2][L][N] Shoup companions of the current switch key
    u64 *ks_hyb = nullptr;            // hybrid key switching: [ks_slots / 2 + 1][KS_HYB_ROWS][N], allocated at the first hybrid call
    size_t ks_slots = 0;
    u32 ks_epoch = 0;                 // rounds consumed so far (flag values already used)
    unsigned long long ks_epoch_limit = 1ull << 30;   // the round numbering restarts before it gets here (DPFHE_EPOCH_LIMIT: tests)
    unsigned ks_epoch_restarts = 0;
    int ks_prefetch = 0;              // ciphertexts ahead for the bulk L2 prefetch of inputs (DPFHE_KS_PF), 0 = off
    int ks_occ_cap = 0;               // tuning: cap on resident fused-kernel CTAs per SM (DPFHE_KS_OCC), 0 = no cap
    unsigned long long *ks_prof = nullptr;   // [ks_slots][16] phase cycle counters; non-null selects the profiling build
};

// the launchers, declared once per variant namespace
#define DPFHE_DECLARE_LAUNCHERS \
    cudaError_t launch_ntt(const LaunchCtx &lc, u64 *data, size_t n_polys, bool inverse, cudaStream_t st); \
    cudaError_t launch_ks(LaunchCtx &lc, int mode, const u64 *a, const u64 *b, const u64 *key, u64 *out, size_t batch, \
                          u32 galois, cudaStream_t st, const u32 *only = nullptr, bool key_ready = false, const u64 *key_s = nullptr); \
    cudaError_t launch_hoist(LaunchCtx &lc, const u64 *ct, u64 *U, u32 *zero, size_t batch, cudaStream_t st); \
    cudaError_t launch_rot_prepare(LaunchCtx &lc, const u64 *key, u32 galois, const u64 *delta, u64 *M, u64 *kprime, cudaStream_t st, \
                                   u64 *key_s_out = nullptr); \
    cudaError_t launch_rot_apply(const LaunchCtx &lc, const u64 *ct, const u64 *U, const u64 *key, const u64 *kprime, u32 galois, u64 *out, \
                                 size_t batch, cudaStream_t st, const u64 *key_s = nullptr); \
    cudaError_t launch_ks_hybrid(LaunchCtx &lc, int mode, const u64 *a, const u64 *b, const u64 *key, u64 *out, size_t batch, u32 galois, \
                                 const MsConsts &K, cudaStream_t st); \
    cudaError_t launch_ks_grouped(LaunchCtx &lc, int mode, const u64 *a, const u64 *b, const u64 *key, u64 *out, size_t batch, u32 galois, \
                                  const MsConsts &K, const GroupConsts &G, cudaStream_t st); \
    cudaError_t launch_hoist_grouped(LaunchCtx &lc, const u64 *ct, u64 *U, const GroupConsts &G, size_t batch, cudaStream_t st); \
    cudaError_t launch_rot_apply_grouped(LaunchCtx &lc, const u64 *ct, const u64 *U, const u64 *key, const u64 *key_s, u32 galois, u64 *acc, \
                                         const MsConsts &K, const GroupConsts &G, size_t batch, cudaStream_t st); \
    cudaError_t launch_pt_inner(const LaunchCtx &lc, const u64 *steps, u32 nb, const u64 *pts, u32 ng, u64 *out, size_t batch, cudaStream_t st, \
                                unsigned *launches); \
    cudaError_t launch_pointwise_mul(const LaunchCtx &lc, const u64 *a, const u64 *b, u64 *out, size_t n_polys, cudaStream_t st); \
    cudaError_t launch_mod_switch(const LaunchCtx &lc, const u64 *in, u64 *tau, u64 *out, const MsConsts &K, size_t n_polys, cudaStream_t st); \
    cudaError_t launch_mod_down_special(const LaunchCtx &lc, const u64 *in, u64 *tau, u64 *out, const MsConsts &K, const GroupConsts &G, size_t n_polys, \
                                        cudaStream_t st); \
    cudaError_t launch_poly_add(const LaunchCtx &lc, const u64 *a, const u64 *b, u64 *out, size_t n_polys, cudaStream_t st); \
    cudaError_t launch_ct_mul_plain(const LaunchCtx &lc, const u64 *ct, const u64 *pt, u64 *out, size_t batch, cudaStream_t st); \
    cudaError_t launch_ct_mul_plain_acc(const LaunchCtx &lc, const u64 *ct, const u64 *pt, u64 *acc, size_t batch, cudaStream_t st); \
    cudaError_t launch_ct_tensor(const LaunchCtx &lc, const u64 *a, const u64 *b, u64 *d, size_t batch, cudaStream_t st); \
    cudaError_t launch_fill_uniform(const LaunchCtx &lc, u64 seed, u64 first_poly, u64 *data, size_t n_polys, cudaStream_t st);

namespace gen {
DPFHE_DECLARE_LAUNCHERS
}
namespace fast {
DPFHE_DECLARE_LAUNCHERS
}
#undef DPFHE_DECLARE_LAUNCHERS
// launch_ks: `only` = optional [batch] filter (non-zero = process); key_ready: the key's Shoup companions are already in key_s
// (or, with key_s == nullptr, in lc.ks_key_s).  launch_rot_prepare / launch_rot_apply: key_s(_out) == nullptr means lc.ks_key_s.

}  // namespace dpfhe
