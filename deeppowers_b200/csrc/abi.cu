// abi.cu — the extern "C" boundary declared in include/dpfhe.h.
//
// Conventions follow the reference HAL (see the header's citations): cudaSetDevice at the top
// of every call, the context owns its tables/scratch and frees them in destroy, CUDA errors
// are reported with file:line (as hal::CUDADevice::check_cuda_error does) — but as a status
// code plus thread-local message, because exceptions cannot cross a C ABI.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/dpfhe.h"
#include "ctx.hpp"

using namespace dpfhe;

// sets the thread-local message of dpfhe_last_error() and returns `code` (shared with multi.cu / hostmem.cu: ctx.hpp)
int dpfhe_fail(int code, const char *fmt, ...);

namespace {

thread_local std::string g_err;

#define fail dpfhe_fail

#define CU_TRY(expr)                                                                                          \
    do {                                                                                                      \
        cudaError_t e_ = (expr);                                                                              \
        if (e_ != cudaSuccess)                                                                                \
            return fail(DPFHE_ERR_CUDA, "CUDA error at %s:%d: %s (%s)", __FILE__, __LINE__, cudaGetErrorString(e_), #expr); \
    } while (0)

constexpr int PIPE_DEPTH = DPFHE_PIPE_DEPTH;

// launcher of the context's arithmetic variant (launch.hpp): dpfhe::fast when every modulus is k * 2^32 + 1
#define VCALL(fn, lc, ...) ((lc).fast ? fast::fn((lc), __VA_ARGS__) : gen::fn((lc), __VA_ARGS__))

}  // namespace

namespace {

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// byte ranges [a, a + na) and [b, b + nb) intersect (unified addressing: valid across devices too)
bool overlaps(const void *a, size_t na, const void *b, size_t nb) {
    if (!a || !b || !na || !nb) return false;
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(a), b0 = reinterpret_cast<uintptr_t>(b);
    return a0 < b0 + nb && b0 < a0 + na;
}

int enter(const dpfhe_ctx *ctx) {
    if (!ctx) return fail(DPFHE_ERR_INVALID, "null context");
    CU_TRY(cudaSetDevice(ctx->lc.device));
    return DPFHE_OK;
}

// The stream of this call (NULL = the context's own).  If the previous call ran on a different stream, this one waits for it:
// all calls share the context's scratch (key companions, digit slots, tickets), so they must not overlap.
cudaStream_t pick(dpfhe_ctx *ctx, void *stream) {
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    if (ctx->have_last && ctx->last_stream != st) cudaStreamWaitEvent(st, ctx->ev_last, 0);
    ctx->cur = st;
    return st;
}
// counts `n` launches just issued on the stream chosen by pick() and marks the point later calls have to wait for
void note_launch(dpfhe_ctx *ctx, uint64_t n) {
    ctx->launches += n;
    if (ctx->cur && cudaEventRecord(ctx->ev_last, ctx->cur) == cudaSuccess) {
        ctx->last_stream = ctx->cur;
        ctx->have_last = true;
    }
}

int ensure_staging(dpfhe_ctx *ctx, size_t in_bytes, size_t out_bytes, size_t key_bytes) {
    // a size is only recorded once every buffer of its group exists: a failed cudaMalloc leaves the group marked empty
    if (in_bytes > ctx->stage_in_bytes) {
        ctx->stage_in_bytes = 0;
        for (int k = 0; k < PIPE_DEPTH; ++k) {
            if (ctx->stage_in[k]) cudaFree(ctx->stage_in[k]);
            ctx->stage_in[k] = nullptr;
            CU_TRY(cudaMalloc(&ctx->stage_in[k], in_bytes));
        }
        ctx->stage_in_bytes = in_bytes;
    }
    if (out_bytes > ctx->stage_out_bytes) {
        ctx->stage_out_bytes = 0;
        for (int k = 0; k < PIPE_DEPTH; ++k) {
            if (ctx->stage_out[k]) cudaFree(ctx->stage_out[k]);
            ctx->stage_out[k] = nullptr;
            CU_TRY(cudaMalloc(&ctx->stage_out[k], out_bytes));
        }
        ctx->stage_out_bytes = out_bytes;
    }
    if (key_bytes > ctx->stage_key_bytes) {
        ctx->stage_key_bytes = 0;
        if (ctx->stage_key) cudaFree(ctx->stage_key);
        ctx->stage_key = nullptr;
        CU_TRY(cudaMalloc(&ctx->stage_key, key_bytes));
        ctx->stage_key_bytes = key_bytes;
    }
    return DPFHE_OK;
}

// Generic three-stage pipeline over `n_items` items split into chunks:
//   upload(chunk -> stage_in[slot]) on s_h2d, compute on ctx->stream, download(stage_out[slot]) on s_d2h.
// in_item_bytes / out_item_bytes are per item; h_in may be two arrays (a and b) laid out back to back in the stage.
template <class Compute>
int run_pipeline_body(dpfhe_ctx *ctx, const u64 *h_in0, const u64 *h_in1, u64 *h_out, size_t n_items, size_t in_item_words,
                      size_t out_item_words, size_t chunk_items, Compute compute) {
    const size_t n_in = h_in1 ? 2 : 1;
    int rc = ensure_staging(ctx, n_in * chunk_items * in_item_words * 8, chunk_items * out_item_words * 8, 0);
    if (rc) return rc;
    size_t k = 0;
    for (size_t first = 0; first < n_items; first += chunk_items, ++k) {
        const size_t cnt = n_items - first < chunk_items ? n_items - first : chunk_items;
        const int slot = (int)(k % PIPE_DEPTH);
        u64 *din0 = ctx->stage_in[slot], *din1 = din0 + chunk_items * in_item_words, *dout = ctx->stage_out[slot];
        if (k >= PIPE_DEPTH) CU_TRY(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_comp[slot], 0));   // stage_in[slot] free again
        CU_TRY(cudaMemcpyAsync(din0, h_in0 + first * in_item_words, cnt * in_item_words * 8, cudaMemcpyHostToDevice, ctx->s_h2d));
        if (h_in1)
            CU_TRY(cudaMemcpyAsync(din1, h_in1 + first * in_item_words, cnt * in_item_words * 8, cudaMemcpyHostToDevice, ctx->s_h2d));
        CU_TRY(cudaEventRecord(ctx->ev_h2d[slot], ctx->s_h2d));
        CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[slot], 0));
        if (k >= PIPE_DEPTH) CU_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[slot], 0));   // stage_out[slot] drained
        rc = compute(din0, din1, dout, cnt, ctx->stream);
        if (rc) return rc;
        CU_TRY(cudaEventRecord(ctx->ev_comp[slot], ctx->stream));
        CU_TRY(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_comp[slot], 0));
        CU_TRY(cudaMemcpyAsync(h_out + first * out_item_words, dout, cnt * out_item_words * 8, cudaMemcpyDeviceToHost, ctx->s_d2h));
        CU_TRY(cudaEventRecord(ctx->ev_d2h[slot], ctx->s_d2h));
    }
    CU_TRY(cudaStreamSynchronize(ctx->s_d2h));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    return DPFHE_OK;
}

// runs the pipeline; on failure the three streams are drained before returning, so that no copy is still reading or writing
// the caller's host buffers (or the staging slots) when the error is reported
template <class Compute>
int run_pipeline(dpfhe_ctx *ctx, const u64 *h_in0, const u64 *h_in1, u64 *h_out, size_t n_items, size_t in_item_words,
                 size_t out_item_words, size_t chunk_items, Compute compute) {
    const int rc = run_pipeline_body(ctx, h_in0, h_in1, h_out, n_items, in_item_words, out_item_words, chunk_items, compute);
    if (rc != DPFHE_OK) {
        const std::string why = g_err;   // the drains below must not replace the message of the failure
        cudaStreamSynchronize(ctx->s_h2d);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->s_d2h);
        cudaGetLastError();
        g_err = why;
    }
    return rc;
}

size_t pick_chunk(const dpfhe_ctx *ctx, size_t item_bytes, size_t n_items) {
    // ~64 MiB per staged operand keeps PCIe transfers long and the staging footprint small
    size_t c = ((size_t)64 << 20) / item_bytes;
    if (c < 1) c = 1;
    // keep at least one full wave of work items per launch where the batch allows it
    const size_t wave = (size_t)ctx->lc.num_sms;
    if (c < wave && n_items >= wave) c = wave;
    if (c > n_items) c = n_items;
    return c;
}

}  // namespace

int dpfhe_fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

extern "C" {

const char *dpfhe_last_error(void) { return g_err.c_str(); }
const char *dpfhe_version(void) { return "dpfhe 0.1 sm_100a"; }

int dpfhe_context_create(const dpfhe_params *p, int device_id, dpfhe_ctx **out) {
    if (!p || !out) return fail(DPFHE_ERR_INVALID, "null argument");
    *out = nullptr;
    int n_dev = 0;
    cudaError_t ce = cudaGetDeviceCount(&n_dev);
    if (ce != cudaSuccess || n_dev <= 0)
        return fail(DPFHE_ERR_CUDA, "no usable CUDA device (%s); this library has no CPU fallback",
                    ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce));
    if (device_id < 0 || device_id >= n_dev) return fail(DPFHE_ERR_INVALID, "device_id %d out of range [0,%d)", device_id, n_dev);
    dpfhe_ctx *ctx = new (std::nothrow) dpfhe_ctx();
    if (!ctx) return fail(DPFHE_ERR_NOMEM, "out of host memory");
    std::string msg = build_host_params(p->log_n, p->n_limbs, p->moduli, ctx->hp);
    if (!msg.empty()) {
        delete ctx;
        return fail(DPFHE_ERR_INVALID, "%s", msg.c_str());
    }
    ctx->lc.device = device_id;
#define CTX_TRY(expr)                                                                                              \
    do {                                                                                                           \
        cudaError_t e_ = (expr);                                                                                   \
        if (e_ != cudaSuccess) {                                                                                   \
            int rc_ = fail(DPFHE_ERR_CUDA, "CUDA error at %s:%d: %s (%s)", __FILE__, __LINE__, cudaGetErrorString(e_), #expr); \
            dpfhe_context_destroy(ctx);                                                                            \
            return rc_;                                                                                            \
        }                                                                                                          \
    } while (0)
    CTX_TRY(cudaSetDevice(device_id));
    cudaDeviceProp prop;
    CTX_TRY(cudaGetDeviceProperties(&prop, device_id));
    if (prop.major < 10) {
        dpfhe_context_destroy(ctx);
        return fail(DPFHE_ERR_INVALID, "device %d is sm_%d%d; this build targets sm_100a only", device_id, prop.major, prop.minor);
    }
    CTX_TRY(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CTX_TRY(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    CTX_TRY(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (int k = 0; k < PIPE_DEPTH; ++k) {
        CTX_TRY(cudaEventCreateWithFlags(&ctx->ev_h2d[k], cudaEventDisableTiming));
        CTX_TRY(cudaEventCreateWithFlags(&ctx->ev_comp[k], cudaEventDisableTiming));
        CTX_TRY(cudaEventCreateWithFlags(&ctx->ev_d2h[k], cudaEventDisableTiming));
    }
    CTX_TRY(cudaEventCreateWithFlags(&ctx->ev_last, cudaEventDisableTiming));
    const size_t N = ctx->N(), L = ctx->hp.L;
    CTX_TRY(cudaMalloc(&ctx->d_lp, L * sizeof(LimbParams)));
    CTX_TRY(cudaMalloc(&ctx->d_tw, L * N * sizeof(Twiddle)));
    CTX_TRY(cudaMalloc(&ctx->d_itw, L * N * sizeof(Twiddle)));
    ctx->device_bytes += L * sizeof(LimbParams) + 2 * L * N * sizeof(Twiddle);
    std::vector<LimbParams> lps(L);
    for (size_t l = 0; l < L; ++l) {
        lps[l] = ctx->hp.limbs[l].lp;
        CTX_TRY(cudaMemcpy(ctx->d_tw + l * N, ctx->hp.limbs[l].tw.data(), N * sizeof(Twiddle), cudaMemcpyHostToDevice));
        CTX_TRY(cudaMemcpy(ctx->d_itw + l * N, ctx->hp.limbs[l].itw.data(), N * sizeof(Twiddle), cudaMemcpyHostToDevice));
    }
    CTX_TRY(cudaMemcpy(ctx->d_lp, lps.data(), L * sizeof(LimbParams), cudaMemcpyHostToDevice));
    LaunchCtx &lc = ctx->lc;
    lc.num_sms = prop.multiProcessorCount;
    lc.log_n = ctx->hp.log_n;
    lc.L = ctx->hp.L;
    lc.fast = true;
    for (size_t l = 0; l < L; ++l) lc.fast = lc.fast && lps[l].nqh != 0;
    if (getenv("DPFHE_FORCE_GENERIC")) lc.fast = false;   // diagnostics: run fast-class moduli through the generic kernels
    {
        uint64_t qmin = ~0ull, qmax = 0;
        for (size_t l = 0; l < L; ++l) {
            qmin = lps[l].q < qmin ? lps[l].q : qmin;
            qmax = lps[l].q > qmax ? lps[l].q : qmax;
        }
        lc.lift_reduce = !(qmax < 2 * qmin) || getenv("DPFHE_LIFT_REDUCE") != nullptr;
    }
    lc.lp = ctx->d_lp;
    memset(&lc.lt, 0, sizeof(lc.lt));
    for (size_t l = 0; l < L; ++l) lc.lt.lp[l] = lps[l];
    if (const char *env = getenv("DPFHE_NTT_CFG")) lc.ntt_cfg = atoi(env);
    if (const char *env = getenv("DPFHE_NTT_TMA")) lc.ntt_tma = atoi(env);
    if (const char *env = getenv("DPFHE_EPOCH_LIMIT")) lc.ks_epoch_limit = strtoull(env, nullptr, 10);
    if (const char *env = getenv("DPFHE_ROT_CFG")) lc.rot_cfg = atoi(env);
    if (const char *env = getenv("DPFHE_KS_OCC")) lc.ks_occ_cap = atoi(env);
    if (const char *env = getenv("DPFHE_KS_PF")) lc.ks_prefetch = atoi(env);
    lc.tw = ctx->d_tw;
    lc.itw = ctx->d_itw;
    // digit-exchange scratch for the fused key-switch kernel: one slot per resident CTA, two parities
    lc.ks_slots = (size_t)lc.num_sms * 4;
    // digit slots and accumulator rows: ONE allocation, so that a single L2 access-policy window can cover the whole
    // cross-phase working set of the fused kernel (launch_ks_t)
    CTX_TRY(cudaMalloc(&lc.ks_scratch, 2 * lc.ks_slots * 2 * N * 8));
    lc.ks_acc = lc.ks_scratch + lc.ks_slots * 2 * N;
    lc.ks_window_bytes = 2 * lc.ks_slots * 2 * N * 8;
    {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, device_id);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, device_id);
        lc.l2_persist_max = (size_t)(max_persist > 0 ? max_persist : 0);
        if ((size_t)max_window < lc.ks_window_bytes) lc.ks_window_bytes = (size_t)(max_window > 0 ? max_window : 0);
        if (const char *env = getenv("DPFHE_L2_PERSIST")) lc.l2_persist = atoi(env);
        if (lc.l2_persist && lc.l2_persist_max) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lc.l2_persist_max);
    }
    CTX_TRY(cudaMalloc(&lc.ks_key_s, 2 * L * L * N * 8));
    ctx->device_bytes += 2 * L * L * N * 8;
    CTX_TRY(cudaMalloc(&lc.ks_flags, 2 * lc.ks_slots * sizeof(u32)));   // digit flags, then the "round finished" marks of ks_hoistg_kernel
    CTX_TRY(cudaMemset(lc.ks_flags, 0, 2 * lc.ks_slots * sizeof(u32)));
    CTX_TRY(cudaMalloc(&lc.ks_consumed, lc.ks_slots * sizeof(u32)));
    CTX_TRY(cudaMemset(lc.ks_consumed, 0, lc.ks_slots * sizeof(u32)));
    if (const char *env = getenv("DPFHE_KS_SINGLE")) lc.ks_single = atoi(env);
    CTX_TRY(cudaMalloc(&lc.ks_ticket, 64));
    CTX_TRY(cudaMalloc(&lc.ks_mail, lc.ks_slots * sizeof(u64)));
    CTX_TRY(cudaMemset(lc.ks_mail, 0, lc.ks_slots * sizeof(u64)));
    ctx->device_bytes += 2 * lc.ks_slots * 2 * N * 8 + lc.ks_slots * (2 * sizeof(u32) + sizeof(u64)) + 64;
    if (getenv("DPFHE_KS_PROF")) {   // diagnostics: per-phase cycle counters of the fused kernel
        CTX_TRY(cudaMalloc(&lc.ks_prof, lc.ks_slots * 16 * sizeof(unsigned long long)));
        CTX_TRY(cudaMemset(lc.ks_prof, 0, lc.ks_slots * 16 * sizeof(unsigned long long)));
    }
    CTX_TRY(cudaDeviceSynchronize());
#undef CTX_TRY
    *out = ctx;
    return DPFHE_OK;
}

void dpfhe_context_destroy(dpfhe_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->lc.device);
    cudaDeviceSynchronize();
    cudaFree(ctx->d_lp);
    cudaFree(ctx->d_tw);
    cudaFree(ctx->d_itw);
    cudaFree(ctx->lc.ks_scratch);   // ks_acc is the second half of the same allocation
    cudaFree(ctx->lc.ks_acc_hyb);
    cudaFree(ctx->lc.ks_flags);
    cudaFree(ctx->lc.ks_consumed);
    cudaFree(ctx->lc.ks_key_s);
    cudaFree(ctx->lc.ks_ticket);
    cudaFree(ctx->lc.ks_mail);
    cudaFree(ctx->lc.ks_prof);
    cudaFree(ctx->stage_key);
    cudaFree(ctx->ms_tau);
    cudaFree(ctx->hoist_U);
    cudaFree(ctx->hoist_M);
    cudaFree(ctx->hoist_kprime);
    cudaFree(ctx->hoist_delta);
    cudaFree(ctx->hoist_zero);
    cudaFree(ctx->hoistg_buf);
    cudaFree(ctx->lc.ks_hyb);
    for (int k = 0; k < PIPE_DEPTH; ++k) {
        cudaFree(ctx->stage_in[k]);
        cudaFree(ctx->stage_out[k]);
        if (ctx->ev_h2d[k]) cudaEventDestroy(ctx->ev_h2d[k]);
        if (ctx->ev_comp[k]) cudaEventDestroy(ctx->ev_comp[k]);
        if (ctx->ev_d2h[k]) cudaEventDestroy(ctx->ev_d2h[k]);
    }
    if (ctx->ev_last) cudaEventDestroy(ctx->ev_last);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
}

int dpfhe_get_modulus(const dpfhe_ctx *ctx, uint32_t limb, uint64_t *q) {
    if (!ctx || !q || limb >= ctx->hp.L) return fail(DPFHE_ERR_INVALID, "bad argument");
    *q = ctx->hp.limbs[limb].lp.q;
    return DPFHE_OK;
}
int dpfhe_get_psi(const dpfhe_ctx *ctx, uint32_t limb, uint64_t *psi) {
    if (!ctx || !psi || limb >= ctx->hp.L) return fail(DPFHE_ERR_INVALID, "bad argument");
    *psi = ctx->hp.limbs[limb].psi;
    return DPFHE_OK;
}
int dpfhe_get_root_powers(const dpfhe_ctx *ctx, uint32_t limb, int inverse, uint64_t *h_out) {
    if (!ctx || !h_out || limb >= ctx->hp.L) return fail(DPFHE_ERR_INVALID, "bad argument");
    const auto &v = inverse ? ctx->hp.limbs[limb].inv_root_powers : ctx->hp.limbs[limb].root_powers;
    memcpy(h_out, v.data(), v.size() * sizeof(uint64_t));
    return DPFHE_OK;
}
size_t dpfhe_context_device_bytes(const dpfhe_ctx *ctx) {
    if (!ctx) return 0;
    const size_t P8 = ctx->P() * 8, L = ctx->hp.L;
    size_t n = ctx->device_bytes;                                                    // tables, key companions, digit slots, accumulators, flags (+ hybrid rows)
    n += PIPE_DEPTH * (ctx->stage_in_bytes + ctx->stage_out_bytes) + ctx->stage_key_bytes;   // host-entry staging
    n += ctx->ms_tau_bytes;                                                          // modulus-switch scratch
    n += ctx->hoist_chunk * (L * P8 + sizeof(u32));                                  // hoisted rotations: shared transforms + zero flags
    if (ctx->hoist_M) n += 3 * P8 + L * L * 8;                                       //   per-rotation constants
    n += ctx->hoistg_bytes;                                                          //   grouped hybrid keys: lifted digits + accumulators
    if (ctx->lc.ks_prof) n += ctx->lc.ks_slots * 16 * sizeof(unsigned long long);
    return n;
}

// frees the scratch that grows with use (hoisted-rotation transforms, modulus-switch rows, host staging); it comes back on demand
int dpfhe_context_trim(dpfhe_ctx *ctx) {
    int rc = enter(ctx);
    if (rc) return rc;
    rc = dpfhe_synchronize(ctx);
    if (rc) return rc;
    cudaFree(ctx->hoist_U); cudaFree(ctx->hoist_zero); cudaFree(ctx->ms_tau); cudaFree(ctx->stage_key);
    ctx->hoist_U = nullptr; ctx->hoist_zero = nullptr; ctx->hoist_chunk = 0;
    cudaFree(ctx->hoistg_buf);
    ctx->hoistg_buf = nullptr; ctx->hoistg_bytes = 0;
    ctx->ms_tau = nullptr; ctx->ms_tau_bytes = 0;
    ctx->stage_key = nullptr; ctx->stage_key_bytes = 0;
    for (int k = 0; k < PIPE_DEPTH; ++k) {
        cudaFree(ctx->stage_in[k]); cudaFree(ctx->stage_out[k]);
        ctx->stage_in[k] = ctx->stage_out[k] = nullptr;
    }
    ctx->stage_in_bytes = ctx->stage_out_bytes = 0;
    return DPFHE_OK;
}
uint64_t dpfhe_launch_count(const dpfhe_ctx *ctx) { return ctx ? ctx->launches : 0; }

#define CHECK_PTR(p)                                                                          \
    do {                                                                                      \
        if (!(p)) return fail(DPFHE_ERR_INVALID, "null pointer: %s", #p);                     \
        if (!aligned16(p)) return fail(DPFHE_ERR_INVALID, "%s must be 16-byte aligned", #p);  \
    } while (0)

int dpfhe_ntt_fwd(dpfhe_ctx *ctx, uint64_t *d_data, size_t n_polys, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    CHECK_PTR(d_data);
    CU_TRY(VCALL(launch_ntt, ctx->lc, d_data, n_polys, false, pick(ctx, stream)));
    note_launch(ctx, 1);
    return DPFHE_OK;
}
int dpfhe_ntt_inv(dpfhe_ctx *ctx, uint64_t *d_data, size_t n_polys, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    CHECK_PTR(d_data);
    CU_TRY(VCALL(launch_ntt, ctx->lc, d_data, n_polys, true, pick(ctx, stream)));
    note_launch(ctx, 1);
    return DPFHE_OK;
}

int dpfhe_poly_mul_pointwise(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out, size_t n_polys, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    CHECK_PTR(d_a); CHECK_PTR(d_b); CHECK_PTR(d_out);
    CU_TRY(VCALL(launch_pointwise_mul, ctx->lc, d_a, d_b, d_out, n_polys, pick(ctx, stream)));
    note_launch(ctx, 1);
    return DPFHE_OK;
}

int dpfhe_poly_add(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_out, size_t n_polys, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    CHECK_PTR(d_a); CHECK_PTR(d_b); CHECK_PTR(d_out);
    CU_TRY(VCALL(launch_poly_add, ctx->lc, d_a, d_b, d_out, n_polys, pick(ctx, stream)));
    note_launch(ctx, 1);
    return DPFHE_OK;
}

int dpfhe_ct_tensor(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, uint64_t *d_d, size_t batch, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    CHECK_PTR(d_a); CHECK_PTR(d_b); CHECK_PTR(d_d);
    CU_TRY(VCALL(launch_ct_tensor, ctx->lc, d_a, d_b, d_d, batch, pick(ctx, stream)));
    note_launch(ctx, 1);
    return DPFHE_OK;
}

static int ks_common(dpfhe_ctx *ctx, int mode, const uint64_t *a, const uint64_t *b, const uint64_t *key, uint64_t *out,
                     size_t batch, uint64_t galois, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    CHECK_PTR(a); CHECK_PTR(key); CHECK_PTR(out);
    if (mode == KS_MUL_RELIN) CHECK_PTR(b);
    if (mode == KS_ROTATE) {
        const uint64_t two_n = (uint64_t)2 << ctx->hp.log_n;
        if (!(galois & 1) || galois >= two_n) return fail(DPFHE_ERR_INVALID, "galois element must be odd and < 2N");
    }
    {   // the output rows are written while other work items still read their inputs: no overlap at all, not only out == in
        const size_t ct_bytes = 2 * ctx->P() * 8, in_bytes = batch * (mode == KS_PLAIN ? ct_bytes / 2 : ct_bytes);
        if (overlaps(out, batch * ct_bytes, a, in_bytes) || overlaps(out, batch * ct_bytes, b, in_bytes))
            return fail(DPFHE_ERR_INVALID, "output must not overlap an input");
    }
    CU_TRY(VCALL(launch_ks, ctx->lc, mode, a, b, key, out, batch, (u32)galois, pick(ctx, stream)));
    note_launch(ctx, 2);   // key_prepare_kernel + ks_fused_kernel
    return DPFHE_OK;
}

int dpfhe_keyswitch(dpfhe_ctx *ctx, const uint64_t *d_d, const uint64_t *d_key, uint64_t *d_out, size_t batch, void *stream) {
    return ks_common(ctx, KS_PLAIN, d_d, nullptr, d_key, d_out, batch, 0, stream);
}
int dpfhe_ct_mul_relin(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, const uint64_t *d_evk, uint64_t *d_out,
                       size_t batch, void *stream) {
    return ks_common(ctx, KS_MUL_RELIN, d_a, d_b, d_evk, d_out, batch, 0, stream);
}
int dpfhe_rotate(dpfhe_ctx *ctx, const uint64_t *d_ct, uint64_t galois_elt, const uint64_t *d_gk, uint64_t *d_out, size_t batch,
                 void *stream) {
    return ks_common(ctx, KS_ROTATE, d_ct, nullptr, d_gk, d_out, batch, galois_elt, stream);
}

// hybrid (special-prime) variants: the context's last limb is the special prime, data carries L-1 limbs
// n_special = K: the last K limbs are special primes and the digits are groups of K limbs; K = 1 is §2.10
static int check_special(const dpfhe_ctx *ctx, unsigned n_special) {
    const unsigned L = ctx->hp.L;
    if (L < 2) return fail(DPFHE_ERR_INVALID, "hybrid key switching needs a special prime: create the context with at least two limbs");
    if (n_special < 1 || n_special > (unsigned)KS_MAX_SPECIAL || 2 * n_special > L)
        return fail(DPFHE_ERR_INVALID, "n_special must be between 1 and %d and at most half of the context's %u limbs", KS_MAX_SPECIAL, L);
    return DPFHE_OK;
}

static int ks_hybrid_common(dpfhe_ctx *ctx, int mode, const uint64_t *a, const uint64_t *b, const uint64_t *key, uint64_t *out,
                            size_t batch, uint64_t galois, uint64_t t_plain, void *stream, unsigned n_special = 1) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    CHECK_PTR(a); CHECK_PTR(key); CHECK_PTR(out);
    if (mode == KS_MUL_RELIN) CHECK_PTR(b);
    const unsigned L = ctx->hp.L;
    rc = check_special(ctx, n_special);
    if (rc) return rc;
    for (unsigned k = 0; k < n_special; ++k)
        if (t_plain >= ctx->hp.limbs[L - 1 - k].lp.q) return fail(DPFHE_ERR_INVALID, "plaintext modulus must be below the special prime");
    if (mode == KS_ROTATE) {
        const uint64_t two_n = (uint64_t)2 << ctx->hp.log_n;
        if (!(galois & 1) || galois >= two_n) return fail(DPFHE_ERR_INVALID, "galois element must be odd and < 2N");
    }
    {
        const size_t ct_bytes = 2 * (size_t)(L - n_special) * ctx->N() * 8, in_bytes = batch * (mode == KS_PLAIN ? ct_bytes / 2 : ct_bytes);
        if (overlaps(out, batch * ct_bytes, a, in_bytes) || overlaps(out, batch * ct_bytes, b, in_bytes))
            return fail(DPFHE_ERR_INVALID, "output must not overlap an input");
    }
    if (!ctx->lc.ks_hyb) {
        u64 *hyb = nullptr;
        const size_t hyb_bytes = (ctx->lc.ks_slots / 2 + 1) * KS_HYB_ROWS * ctx->N() * sizeof(u64), acc_bytes = ctx->lc.ks_slots * 4 * ctx->N() * sizeof(u64);
        CU_TRY(cudaMalloc(&hyb, hyb_bytes));
        ctx->lc.ks_hyb = hyb;
        CU_TRY(cudaMalloc(&ctx->lc.ks_acc_hyb, acc_bytes));
        ctx->device_bytes += hyb_bytes + acc_bytes;
    }
    MsConsts K;
    if (n_special == 1) {
        build_ms_consts(ctx->hp, t_plain, K);
        CU_TRY(VCALL(launch_ks_hybrid, ctx->lc, mode, a, b, key, out, batch, (u32)galois, K, pick(ctx, stream)));
    } else {
        GroupConsts G;
        build_group_consts(ctx->hp, n_special, t_plain, G, K);
        CU_TRY(VCALL(launch_ks_grouped, ctx->lc, mode, a, b, key, out, batch, (u32)galois, K, G, pick(ctx, stream)));
    }
    note_launch(ctx, 2);   // key_prepare_kernel + ks_hybrid_kernel / ks_grouped_kernel
    return DPFHE_OK;
}
int dpfhe_keyswitch_grouped(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_d, const uint64_t *d_key, uint64_t *d_out, size_t batch,
                            uint64_t t_plain, void *stream) {
    return ks_hybrid_common(ctx, KS_PLAIN, d_d, nullptr, d_key, d_out, batch, 0, t_plain, stream, n_special);
}
int dpfhe_ct_mul_relin_grouped(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_a, const uint64_t *d_b, const uint64_t *d_evk,
                               uint64_t *d_out, size_t batch, uint64_t t_plain, void *stream) {
    return ks_hybrid_common(ctx, KS_MUL_RELIN, d_a, d_b, d_evk, d_out, batch, 0, t_plain, stream, n_special);
}
int dpfhe_rotate_grouped(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_ct, uint64_t galois_elt, const uint64_t *d_gk, uint64_t *d_out,
                         size_t batch, uint64_t t_plain, void *stream) {
    return ks_hybrid_common(ctx, KS_ROTATE, d_ct, nullptr, d_gk, d_out, batch, galois_elt, t_plain, stream, n_special);
}
int dpfhe_grouped_digits(const dpfhe_ctx *ctx, unsigned n_special, unsigned *digits) {
    if (!ctx || !digits) return fail(DPFHE_ERR_INVALID, "null argument");
    int rc = check_special(ctx, n_special);
    if (rc) return rc;
    *digits = (ctx->hp.L - n_special + n_special - 1) / n_special;
    return DPFHE_OK;
}
int dpfhe_keyswitch_hybrid(dpfhe_ctx *ctx, const uint64_t *d_d, const uint64_t *d_key, uint64_t *d_out, size_t batch, uint64_t t_plain,
                           void *stream) {
    return ks_hybrid_common(ctx, KS_PLAIN, d_d, nullptr, d_key, d_out, batch, 0, t_plain, stream);
}
int dpfhe_ct_mul_relin_hybrid(dpfhe_ctx *ctx, const uint64_t *d_a, const uint64_t *d_b, const uint64_t *d_evk, uint64_t *d_out,
                              size_t batch, uint64_t t_plain, void *stream) {
    return ks_hybrid_common(ctx, KS_MUL_RELIN, d_a, d_b, d_evk, d_out, batch, 0, t_plain, stream);
}
int dpfhe_rotate_hybrid(dpfhe_ctx *ctx, const uint64_t *d_ct, uint64_t galois_elt, const uint64_t *d_gk, uint64_t *d_out, size_t batch,
                        uint64_t t_plain, void *stream) {
    return ks_hybrid_common(ctx, KS_ROTATE, d_ct, nullptr, d_gk, d_out, batch, galois_elt, t_plain, stream);
}

// Hoisted rotations: n_rot rotations of the SAME ciphertexts.  Bit-identical to n_rot calls of dpfhe_rotate; the digit
// decomposition and its L(L-1) forward transforms are done once per ciphertext (kernels.cu: ks_hoist_kernel), each
// rotation is then gathers + multiply-accumulates (rot_apply_kernel).  Ciphertexts whose digit has a zero coefficient
// (where the shared-transform identity does not hold) are recomputed by the ordinary rotate kernel.
// prepared: optional per-rotation constants kept by the caller (a linear layer applies the same rotations to every batch):
// for rotation r, prepared[2r] = the key's Shoup companions [L][2][L][N], prepared[2r+1] = kprime [2][L][N].
static int ensure_hoist_consts(dpfhe_ctx *ctx) {
    const size_t L = ctx->hp.L, P = ctx->P();
    if (ctx->hoist_M) return DPFHE_OK;
    CU_TRY(cudaMalloc(&ctx->hoist_M, P * sizeof(u64)));
    CU_TRY(cudaMalloc(&ctx->hoist_kprime, 2 * P * sizeof(u64)));
    CU_TRY(cudaMalloc(&ctx->hoist_delta, L * L * sizeof(u64)));
    std::vector<u64> delta(L * L);
    for (size_t j = 0; j < L; ++j)
        for (size_t i = 0; i < L; ++i) delta[j * L + i] = ctx->hp.limbs[j].lp.q % ctx->hp.limbs[i].lp.q;
    CU_TRY(cudaMemcpy(ctx->hoist_delta, delta.data(), delta.size() * sizeof(u64), cudaMemcpyHostToDevice));
    return DPFHE_OK;
}

static int rotate_hoisted_impl(dpfhe_ctx *ctx, const uint64_t *d_ct, size_t n_rot, const uint64_t *galois_elts, const uint64_t *const *d_gks,
                               const uint64_t *const *prepared, uint64_t *d_out, size_t batch, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0 || n_rot == 0) return DPFHE_OK;
    CHECK_PTR(d_ct); CHECK_PTR(d_out);
    if (!galois_elts || !d_gks) return fail(DPFHE_ERR_INVALID, "null argument");
    const uint64_t two_n = (uint64_t)2 << ctx->hp.log_n;
    const size_t N = ctx->N(), L = ctx->hp.L, P = ctx->P();
    for (size_t r = 0; r < n_rot; ++r) {
        if (!(galois_elts[r] & 1) || galois_elts[r] >= two_n) return fail(DPFHE_ERR_INVALID, "galois element must be odd and < 2N");
        if (!d_gks[r] || !aligned16(d_gks[r])) return fail(DPFHE_ERR_INVALID, "null or misaligned Galois key");
    }
    if (overlaps(d_out, n_rot * batch * 2 * P * 8, d_ct, batch * 2 * P * 8)) return fail(DPFHE_ERR_INVALID, "output must not overlap the input");
    cudaStream_t st = pick(ctx, stream);
    // scratch: at most ~4 GiB of shared transforms at a time (the batch is processed in chunks of that many ciphertexts)
    const size_t per_ct = L * L * N * sizeof(u64);
    size_t cap = (size_t)4 << 30;
    if (const char *e = getenv("DPFHE_HOIST_CAP_MB")) {   // diagnostics / tests: smaller scratch, more chunks
        const long mb = atol(e);
        if (mb > 0) cap = (size_t)mb << 20;
    }
    size_t chunk = cap / per_ct;
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    if (chunk > ctx->hoist_chunk) {
        CU_TRY(cudaStreamSynchronize(st));
        cudaFree(ctx->hoist_U);
        cudaFree(ctx->hoist_zero);
        ctx->hoist_U = nullptr;
        ctx->hoist_zero = nullptr;
        ctx->hoist_chunk = 0;
        CU_TRY(cudaMalloc(&ctx->hoist_U, chunk * per_ct));
        CU_TRY(cudaMalloc(&ctx->hoist_zero, chunk * sizeof(u32)));
        ctx->hoist_chunk = chunk;
    }
    rc = ensure_hoist_consts(ctx);
    if (rc) return rc;
    for (size_t first = 0; first < batch; first += chunk) {
        const size_t cnt = batch - first < chunk ? batch - first : chunk;
        const u64 *in = d_ct + first * 2 * P;
        CU_TRY(cudaMemsetAsync(ctx->hoist_zero, 0, cnt * sizeof(u32), st));
        if (L > 1) {
            CU_TRY(VCALL(launch_hoist, ctx->lc, in, ctx->hoist_U, ctx->hoist_zero, cnt, st));
            note_launch(ctx, 1);
        }
        for (size_t r = 0; r < n_rot; ++r) {
            u64 *out = d_out + (r * batch + first) * 2 * P;
            const u64 *key_s = prepared ? prepared[2 * r] : nullptr, *kprime = prepared ? prepared[2 * r + 1] : ctx->hoist_kprime;
            if (!prepared) {
                CU_TRY(VCALL(launch_rot_prepare, ctx->lc, d_gks[r], (u32)galois_elts[r], ctx->hoist_delta, ctx->hoist_M, ctx->hoist_kprime, st));
                note_launch(ctx, 4);   // key_prepare, negmask, ntt, kprime
            }
            CU_TRY(VCALL(launch_rot_apply, ctx->lc, in, L > 1 ? ctx->hoist_U : nullptr, d_gks[r], kprime, (u32)galois_elts[r], out, cnt, st, key_s));
            note_launch(ctx, 1);
            if (L > 1) {
                CU_TRY(VCALL(launch_ks, ctx->lc, KS_ROTATE, in, nullptr, d_gks[r], out, cnt, (u32)galois_elts[r], st, ctx->hoist_zero, true, key_s));
                note_launch(ctx, 1);
            }
        }
    }
    return DPFHE_OK;
}

int dpfhe_rotate_hoisted(dpfhe_ctx *ctx, const uint64_t *d_ct, size_t n_rot, const uint64_t *galois_elts, const uint64_t *const *d_gks,
                         uint64_t *d_out, size_t batch, void *stream) {
    return rotate_hoisted_impl(ctx, d_ct, n_rot, galois_elts, d_gks, nullptr, d_out, batch, stream);
}

// Hoisted rotations with grouped hybrid keys: the mod-up of c1 is done once per ciphertext
// (ks_hoistg_kernel), every rotation is then gathers + multiply-accumulates over all L limbs (rot_apply_grouped_kernel) and
// the division by P (md_tau / md_limb kernels).  Same plaintexts as n_rot calls of dpfhe_rotate_grouped, not the same bits.
int dpfhe_rotate_hoisted_grouped(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_ct, size_t n_rot, const uint64_t *galois_elts,
                                 const uint64_t *const *d_gks, uint64_t *d_out, size_t batch, uint64_t t_plain, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0 || n_rot == 0) return DPFHE_OK;
    CHECK_PTR(d_ct); CHECK_PTR(d_out);
    if (!galois_elts || !d_gks) return fail(DPFHE_ERR_INVALID, "null argument");
    rc = check_special(ctx, n_special);
    if (rc) return rc;
    const uint64_t two_n = (uint64_t)2 << ctx->hp.log_n;
    const size_t N = ctx->N(), L = ctx->hp.L, Lq = L - n_special, Pq = Lq * N, dnum = (Lq + n_special - 1) / n_special;
    for (unsigned k = 0; k < n_special; ++k)
        if (t_plain >= ctx->hp.limbs[L - 1 - k].lp.q) return fail(DPFHE_ERR_INVALID, "plaintext modulus must be below the special prime");
    for (size_t r = 0; r < n_rot; ++r) {
        if (!(galois_elts[r] & 1) || galois_elts[r] >= two_n) return fail(DPFHE_ERR_INVALID, "galois element must be odd and < 2N");
        if (!d_gks[r] || !aligned16(d_gks[r])) return fail(DPFHE_ERR_INVALID, "null or misaligned Galois key");
    }
    if (overlaps(d_out, n_rot * batch * 2 * Pq * 8, d_ct, batch * 2 * Pq * 8)) return fail(DPFHE_ERR_INVALID, "output must not overlap the input");
    cudaStream_t st = pick(ctx, stream);
    // scratch per ciphertext: lifted digits [dnum][L][N], accumulators [2][L][N], tau' rows [2][K][N]; at most ~4 GiB at a time
    const size_t u_words = dnum * L * N, acc_words = 2 * L * N, tau_words = 2 * (size_t)n_special * N;
    const size_t per_ct = (u_words + acc_words + tau_words) * sizeof(u64);
    size_t cap = (size_t)4 << 30;
    if (const char *e = getenv("DPFHE_HOIST_CAP_MB")) {
        const long mb = atol(e);
        if (mb > 0) cap = (size_t)mb << 20;
    }
    size_t chunk = cap / per_ct;
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    if (chunk * per_ct > ctx->hoistg_bytes) {
        CU_TRY(cudaStreamSynchronize(st));
        cudaFree(ctx->hoistg_buf);
        ctx->hoistg_buf = nullptr;
        ctx->hoistg_bytes = 0;
        CU_TRY(cudaMalloc(&ctx->hoistg_buf, chunk * per_ct));
        ctx->hoistg_bytes = chunk * per_ct;
    }
    u64 *U = ctx->hoistg_buf, *acc = U + chunk * u_words, *tau = acc + chunk * acc_words;
    MsConsts K;
    GroupConsts G;
    build_group_consts(ctx->hp, n_special, t_plain, G, K);
    for (size_t first = 0; first < batch; first += chunk) {
        const size_t cnt = batch - first < chunk ? batch - first : chunk;
        const u64 *in = d_ct + first * 2 * Pq;
        CU_TRY(VCALL(launch_hoist_grouped, ctx->lc, in, U, G, cnt, st));
        note_launch(ctx, 1);
        for (size_t r = 0; r < n_rot; ++r) {
            u64 *out = d_out + (r * batch + first) * 2 * Pq;
            CU_TRY(VCALL(launch_rot_apply_grouped, ctx->lc, in, U, d_gks[r], nullptr, (u32)galois_elts[r], acc, K, G, cnt, st));
            CU_TRY(VCALL(launch_mod_down_special, ctx->lc, acc, tau, out, K, G, 2 * cnt, st));
            note_launch(ctx, 4);   // key_prepare, rot_apply_grouped, md_tau, md_limb
        }
    }
    return DPFHE_OK;
}

int dpfhe_ct_mul_plain(dpfhe_ctx *ctx, const uint64_t *d_ct, const uint64_t *d_pt, uint64_t *d_out, size_t batch, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    CHECK_PTR(d_ct); CHECK_PTR(d_pt); CHECK_PTR(d_out);
    CU_TRY(VCALL(launch_ct_mul_plain, ctx->lc, d_ct, d_pt, d_out, batch, pick(ctx, stream)));
    note_launch(ctx, 1);
    return DPFHE_OK;
}

int dpfhe_ct_mul_plain_acc(dpfhe_ctx *ctx, const uint64_t *d_ct, const uint64_t *d_pt, uint64_t *d_acc, size_t batch, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    CHECK_PTR(d_ct); CHECK_PTR(d_pt); CHECK_PTR(d_acc);
    CU_TRY(VCALL(launch_ct_mul_plain_acc, ctx->lc, d_ct, d_pt, d_acc, batch, pick(ctx, stream)));
    note_launch(ctx, 1);
    return DPFHE_OK;
}

int dpfhe_ct_mul_plain_inner(dpfhe_ctx *ctx, const uint64_t *d_steps, size_t n_steps, const uint64_t *d_pts, size_t n_groups, uint64_t *d_out,
                             size_t batch, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0 || n_groups == 0) return DPFHE_OK;
    CHECK_PTR(d_steps); CHECK_PTR(d_pts); CHECK_PTR(d_out);
    if (n_steps == 0 || n_steps > 128) return fail(DPFHE_ERR_INVALID, "n_steps must be in [1, 128]");
    if (n_groups > 65535) return fail(DPFHE_ERR_INVALID, "n_groups must be below 65536");
    if (overlaps(d_out, n_groups * batch * 2 * ctx->P() * 8, d_steps, n_steps * batch * 2 * ctx->P() * 8))
        return fail(DPFHE_ERR_INVALID, "output must not overlap an input");
    unsigned launches = 0;
    CU_TRY(VCALL(launch_pt_inner, ctx->lc, d_steps, (u32)n_steps, d_pts, (u32)n_groups, d_out, batch, pick(ctx, stream), &launches));
    note_launch(ctx, launches);
    return DPFHE_OK;
}

int dpfhe_mod_switch_down(dpfhe_ctx *ctx, const uint64_t *d_in, uint64_t *d_out, size_t n_polys, uint64_t t_plain, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    CHECK_PTR(d_in); CHECK_PTR(d_out);
    const unsigned L = ctx->hp.L;
    if (L < 2) return fail(DPFHE_ERR_INVALID, "mod_switch_down needs at least two limbs");
    if (overlaps(d_out, n_polys * (size_t)(L - 1) * ctx->N() * 8, d_in, n_polys * ctx->P() * 8))
        return fail(DPFHE_ERR_INVALID, "output must not overlap the input");
    const uint64_t ql = ctx->hp.limbs[L - 1].lp.q;
    if (t_plain >= ql || (t_plain && t_plain % ql == 0)) return fail(DPFHE_ERR_INVALID, "plaintext modulus must be below the dropped modulus");
    const size_t need = n_polys * ctx->N() * 8;
    if (need > ctx->ms_tau_bytes) {
        CU_TRY(cudaStreamSynchronize(pick(ctx, stream)));   // the old scratch may still be in use on this stream
        if (ctx->ms_tau) cudaFree(ctx->ms_tau);
        ctx->ms_tau = nullptr;
        ctx->ms_tau_bytes = 0;
        CU_TRY(cudaMalloc(&ctx->ms_tau, need));
        ctx->ms_tau_bytes = need;
    }
    MsConsts K;
    build_ms_consts(ctx->hp, t_plain, K);
    CU_TRY(VCALL(launch_mod_switch, ctx->lc, d_in, ctx->ms_tau, d_out, K, n_polys, pick(ctx, stream)));
    note_launch(ctx, 2);
    return DPFHE_OK;
}

// division by the product of the last n_special limbs; n_special = 1 is dpfhe_mod_switch_down
int dpfhe_mod_down_special(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *d_in, uint64_t *d_out, size_t n_polys, uint64_t t_plain,
                           void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    CHECK_PTR(d_in); CHECK_PTR(d_out);
    const unsigned L = ctx->hp.L;
    if (n_special < 1 || n_special > (unsigned)KS_MAX_SPECIAL || n_special >= L)
        return fail(DPFHE_ERR_INVALID, "n_special must be between 1 and %d and below the context's %u limbs", KS_MAX_SPECIAL, L);
    if (overlaps(d_out, n_polys * (size_t)(L - n_special) * ctx->N() * 8, d_in, n_polys * ctx->P() * 8))
        return fail(DPFHE_ERR_INVALID, "output must not overlap the input");
    for (unsigned k = 0; k < n_special; ++k)
        if (t_plain >= ctx->hp.limbs[L - 1 - k].lp.q) return fail(DPFHE_ERR_INVALID, "plaintext modulus must be below the dropped moduli");
    const size_t need = n_polys * n_special * ctx->N() * 8;
    if (need > ctx->ms_tau_bytes) {
        CU_TRY(cudaStreamSynchronize(pick(ctx, stream)));   // the old scratch may still be in use on this stream
        if (ctx->ms_tau) cudaFree(ctx->ms_tau);
        ctx->ms_tau = nullptr;
        ctx->ms_tau_bytes = 0;
        CU_TRY(cudaMalloc(&ctx->ms_tau, need));
        ctx->ms_tau_bytes = need;
    }
    MsConsts K;
    GroupConsts G;
    build_group_consts(ctx->hp, n_special, t_plain, G, K);
    CU_TRY(VCALL(launch_mod_down_special, ctx->lc, d_in, ctx->ms_tau, d_out, K, G, n_polys, pick(ctx, stream)));
    note_launch(ctx, 2);
    return DPFHE_OK;
}

int dpfhe_fill_uniform(dpfhe_ctx *ctx, uint64_t seed, uint64_t first_poly, uint64_t *d_data, size_t n_polys, void *stream) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    CHECK_PTR(d_data);
    CU_TRY(VCALL(launch_fill_uniform, ctx->lc, seed, first_poly, d_data, n_polys, pick(ctx, stream)));
    note_launch(ctx, 1);
    return DPFHE_OK;
}

// ---------------------------------------------------------------- host-buffer entry points

static int ntt_host(dpfhe_ctx *ctx, uint64_t *h_data, size_t n_polys, bool inverse) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    if (!h_data) return fail(DPFHE_ERR_INVALID, "null pointer: h_data");
    const size_t P = ctx->P();
    const size_t chunk = pick_chunk(ctx, P * 8, n_polys);
    return run_pipeline(ctx, h_data, nullptr, h_data, n_polys, P, P, chunk,
                        [&](u64 *din, u64 *, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            CU_TRY(VCALL(launch_ntt, ctx->lc, din, cnt, inverse, st));
                            note_launch(ctx, 1);
                            CU_TRY(cudaMemcpyAsync(dout, din, cnt * P * 8, cudaMemcpyDeviceToDevice, st));
                            return DPFHE_OK;
                        });
}
int dpfhe_ntt_fwd_host(dpfhe_ctx *ctx, uint64_t *h_data, size_t n_polys) { return ntt_host(ctx, h_data, n_polys, false); }
int dpfhe_ntt_inv_host(dpfhe_ctx *ctx, uint64_t *h_data, size_t n_polys) { return ntt_host(ctx, h_data, n_polys, true); }

static int upload_key(dpfhe_ctx *ctx, const uint64_t *h_key, size_t words) {
    int rc = ensure_staging(ctx, 0, 0, words * 8);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(ctx->stage_key, h_key, words * 8, cudaMemcpyHostToDevice, pick(ctx, nullptr)));
    return DPFHE_OK;
}

int dpfhe_ct_mul_relin_host(dpfhe_ctx *ctx, const uint64_t *h_a, const uint64_t *h_b, const uint64_t *h_evk, uint64_t *h_out, size_t batch) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    if (!h_a || !h_b || !h_evk || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    const size_t P = ctx->P();
    rc = upload_key(ctx, h_evk, 2 * ctx->hp.L * P);
    if (rc) return rc;
    const size_t chunk = pick_chunk(ctx, 2 * P * 8, batch);
    return run_pipeline(ctx, h_a, h_b, h_out, batch, 2 * P, 2 * P, chunk,
                        [&](u64 *da, u64 *db, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            return ks_common(ctx, KS_MUL_RELIN, da, db, ctx->stage_key, dout, cnt, 0, st);
                        });
}

int dpfhe_rotate_host(dpfhe_ctx *ctx, const uint64_t *h_ct, uint64_t galois_elt, const uint64_t *h_gk, uint64_t *h_out, size_t batch) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    if (!h_ct || !h_gk || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    const size_t P = ctx->P();
    rc = upload_key(ctx, h_gk, 2 * ctx->hp.L * P);
    if (rc) return rc;
    const size_t chunk = pick_chunk(ctx, 2 * P * 8, batch);
    return run_pipeline(ctx, h_ct, nullptr, h_out, batch, 2 * P, 2 * P, chunk,
                        [&](u64 *dc, u64 *, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            return ks_common(ctx, KS_ROTATE, dc, nullptr, ctx->stage_key, dout, cnt, galois_elt, st);
                        });
}

int dpfhe_ct_mul_relin_hybrid_host(dpfhe_ctx *ctx, const uint64_t *h_a, const uint64_t *h_b, const uint64_t *h_evk, uint64_t *h_out,
                                   size_t batch, uint64_t t_plain) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    if (!h_a || !h_b || !h_evk || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    if (ctx->hp.L < 2) return fail(DPFHE_ERR_INVALID, "hybrid key switching needs a special prime: create the context with at least two limbs");
    const size_t Pq = (ctx->hp.L - 1) * ctx->N();   // words of a ciphertext polynomial (L-1 limbs)
    rc = upload_key(ctx, h_evk, 2 * (ctx->hp.L - 1) * ctx->P());
    if (rc) return rc;
    const size_t chunk = pick_chunk(ctx, 2 * Pq * 8, batch);
    return run_pipeline(ctx, h_a, h_b, h_out, batch, 2 * Pq, 2 * Pq, chunk,
                        [&](u64 *da, u64 *db, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            return ks_hybrid_common(ctx, KS_MUL_RELIN, da, db, ctx->stage_key, dout, cnt, 0, t_plain, st);
                        });
}

int dpfhe_rotate_hybrid_host(dpfhe_ctx *ctx, const uint64_t *h_ct, uint64_t galois_elt, const uint64_t *h_gk, uint64_t *h_out,
                             size_t batch, uint64_t t_plain) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    if (!h_ct || !h_gk || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    if (ctx->hp.L < 2) return fail(DPFHE_ERR_INVALID, "hybrid key switching needs a special prime: create the context with at least two limbs");
    const size_t Pq = (ctx->hp.L - 1) * ctx->N();
    rc = upload_key(ctx, h_gk, 2 * (ctx->hp.L - 1) * ctx->P());
    if (rc) return rc;
    const size_t chunk = pick_chunk(ctx, 2 * Pq * 8, batch);
    return run_pipeline(ctx, h_ct, nullptr, h_out, batch, 2 * Pq, 2 * Pq, chunk,
                        [&](u64 *dc, u64 *, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            return ks_hybrid_common(ctx, KS_ROTATE, dc, nullptr, ctx->stage_key, dout, cnt, galois_elt, t_plain, st);
                        });
}

int dpfhe_ct_mul_relin_grouped_host(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *h_a, const uint64_t *h_b, const uint64_t *h_evk,
                                    uint64_t *h_out, size_t batch, uint64_t t_plain) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    if (!h_a || !h_b || !h_evk || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    rc = check_special(ctx, n_special);
    if (rc) return rc;
    const size_t Lq = ctx->hp.L - n_special, Pq = Lq * ctx->N(), dnum = (Lq + n_special - 1) / n_special;
    rc = upload_key(ctx, h_evk, 2 * dnum * ctx->P());
    if (rc) return rc;
    const size_t chunk = pick_chunk(ctx, 2 * Pq * 8, batch);
    return run_pipeline(ctx, h_a, h_b, h_out, batch, 2 * Pq, 2 * Pq, chunk,
                        [&](u64 *da, u64 *db, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            return ks_hybrid_common(ctx, KS_MUL_RELIN, da, db, ctx->stage_key, dout, cnt, 0, t_plain, st, n_special);
                        });
}

int dpfhe_rotate_grouped_host(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *h_ct, uint64_t galois_elt, const uint64_t *h_gk,
                              uint64_t *h_out, size_t batch, uint64_t t_plain) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    if (!h_ct || !h_gk || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    rc = check_special(ctx, n_special);
    if (rc) return rc;
    const size_t Lq = ctx->hp.L - n_special, Pq = Lq * ctx->N(), dnum = (Lq + n_special - 1) / n_special;
    rc = upload_key(ctx, h_gk, 2 * dnum * ctx->P());
    if (rc) return rc;
    const size_t chunk = pick_chunk(ctx, 2 * Pq * 8, batch);
    return run_pipeline(ctx, h_ct, nullptr, h_out, batch, 2 * Pq, 2 * Pq, chunk,
                        [&](u64 *dc, u64 *, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            return ks_hybrid_common(ctx, KS_ROTATE, dc, nullptr, ctx->stage_key, dout, cnt, galois_elt, t_plain, st, n_special);
                        });
}

int dpfhe_mod_switch_down_host(dpfhe_ctx *ctx, const uint64_t *h_in, uint64_t *h_out, size_t n_polys, uint64_t t_plain) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    if (!h_in || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    if (ctx->hp.L < 2) return fail(DPFHE_ERR_INVALID, "mod_switch_down needs at least two limbs");
    const size_t P = ctx->P(), Pq = (ctx->hp.L - 1) * ctx->N();
    const size_t chunk = pick_chunk(ctx, P * 8, n_polys);
    return run_pipeline(ctx, h_in, nullptr, h_out, n_polys, P, Pq, chunk,
                        [&](u64 *din, u64 *, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            return dpfhe_mod_switch_down(ctx, din, dout, cnt, t_plain, st);
                        });
}

int dpfhe_mod_down_special_host(dpfhe_ctx *ctx, unsigned n_special, const uint64_t *h_in, uint64_t *h_out, size_t n_polys, uint64_t t_plain) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (n_polys == 0) return DPFHE_OK;
    if (!h_in || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    if (n_special < 1 || n_special >= ctx->hp.L) return fail(DPFHE_ERR_INVALID, "n_special must be at least 1 and below the context's limbs");
    const size_t P = ctx->P(), Pq = (ctx->hp.L - n_special) * ctx->N();
    const size_t chunk = pick_chunk(ctx, P * 8, n_polys);
    return run_pipeline(ctx, h_in, nullptr, h_out, n_polys, P, Pq, chunk,
                        [&](u64 *din, u64 *, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            return dpfhe_mod_down_special(ctx, n_special, din, dout, cnt, t_plain, st);
                        });
}

int dpfhe_ct_mul_plain_host(dpfhe_ctx *ctx, const uint64_t *h_ct, const uint64_t *h_pt, uint64_t *h_out, size_t batch) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    if (!h_ct || !h_pt || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    const size_t P = ctx->P();
    rc = upload_key(ctx, h_pt, P);
    if (rc) return rc;
    const size_t chunk = pick_chunk(ctx, 2 * P * 8, batch);
    return run_pipeline(ctx, h_ct, nullptr, h_out, batch, 2 * P, 2 * P, chunk,
                        [&](u64 *dc, u64 *, u64 *dout, size_t cnt, cudaStream_t st) -> int {
                            CU_TRY(VCALL(launch_ct_mul_plain, ctx->lc, dc, ctx->stage_key, dout, cnt, st));
                            note_launch(ctx, 1);
                            return DPFHE_OK;
                        });
}

// waits for everything this context has in flight, whatever stream it was issued on
int dpfhe_synchronize(dpfhe_ctx *ctx) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (ctx->have_last) CU_TRY(cudaEventSynchronize(ctx->ev_last));
    CU_TRY(cudaStreamSynchronize(ctx->stream));
    CU_TRY(cudaStreamSynchronize(ctx->s_h2d));
    CU_TRY(cudaStreamSynchronize(ctx->s_d2h));
    return DPFHE_OK;
}

int dpfhe_device_count(int *out) {
    if (!out) return fail(DPFHE_ERR_INVALID, "null argument");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    *out = e == cudaSuccess ? n : 0;
    if (e != cudaSuccess) return fail(DPFHE_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return DPFHE_OK;
}

int dpfhe_context_device(const dpfhe_ctx *ctx) { return ctx ? ctx->lc.device : -1; }

// rotation by k slots (k may be negative): the Galois element is 5^k mod 2N
int dpfhe_galois_element(const dpfhe_ctx *ctx, int k, uint64_t *galois_elt) {
    if (!ctx || !galois_elt) return fail(DPFHE_ERR_INVALID, "null argument");
    const uint64_t two_n = (uint64_t)2 << ctx->hp.log_n, order = two_n / 4;   // 5 has order N/2 in Z_2N^*
    uint64_t e = (uint64_t)(((long long)k % (long long)order + (long long)order) % (long long)order), g = 1, b = 5;
    for (; e; e >>= 1) {
        if (e & 1) g = g * b % two_n;
        b = b * b % two_n;
    }
    *galois_elt = g;
    return DPFHE_OK;
}
int dpfhe_rotate_steps(dpfhe_ctx *ctx, const uint64_t *d_ct, int k, const uint64_t *d_gk, uint64_t *d_out, size_t batch, void *stream) {
    uint64_t g = 0;
    int rc = dpfhe_galois_element(ctx, k, &g);
    if (rc) return rc;
    return dpfhe_rotate(ctx, d_ct, g, d_gk, d_out, batch, stream);
}

// ---------------------------------------------------------------- device memory that other processes / devices can map
int dpfhe_device_alloc(dpfhe_ctx *ctx, void **d_out, size_t bytes) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (!d_out) return fail(DPFHE_ERR_INVALID, "null argument");
    *d_out = nullptr;
    CU_TRY(cudaMalloc(d_out, bytes));   // a whole allocation of its own, so that an IPC handle maps exactly this buffer
    return DPFHE_OK;
}
int dpfhe_device_free(dpfhe_ctx *ctx, void *d_ptr) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (d_ptr) CU_TRY(cudaFree(d_ptr));
    return DPFHE_OK;
}
int dpfhe_ipc_export(dpfhe_ctx *ctx, const void *d_ptr, unsigned char handle[DPFHE_IPC_HANDLE_BYTES]) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (!d_ptr || !handle) return fail(DPFHE_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == DPFHE_IPC_HANDLE_BYTES, "handle size");
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, const_cast<void *>(d_ptr)));
    memcpy(handle, &h, sizeof(h));
    return DPFHE_OK;
}
int dpfhe_ipc_open(dpfhe_ctx *ctx, const unsigned char handle[DPFHE_IPC_HANDLE_BYTES], void **d_out) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (!handle || !d_out) return fail(DPFHE_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    *d_out = nullptr;
    CU_TRY(cudaIpcOpenMemHandle(d_out, h, cudaIpcMemLazyEnablePeerAccess));   // maps the peer's buffer into this context's device
    return DPFHE_OK;
}
int dpfhe_ipc_close(dpfhe_ctx *ctx, void *d_ptr) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (d_ptr) CU_TRY(cudaIpcCloseMemHandle(d_ptr));
    return DPFHE_OK;
}

// Diagnostics (not part of the drop-in surface): sums the fused kernel's per-phase clock64 counters over
// all CTAs into out[16] and clears them.  Only available when the context was created with DPFHE_KS_PROF set.
int dpfhe_debug_phase_cycles(dpfhe_ctx *ctx, uint64_t *out16) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (!out16 || !ctx->lc.ks_prof) return fail(DPFHE_ERR_INVALID, "phase profiling is not enabled (DPFHE_KS_PROF)");
    std::vector<unsigned long long> h(ctx->lc.ks_slots * 16);
    CU_TRY(cudaDeviceSynchronize());
    CU_TRY(cudaMemcpy(h.data(), ctx->lc.ks_prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU_TRY(cudaMemset(ctx->lc.ks_prof, 0, h.size() * sizeof(unsigned long long)));
    for (int k = 0; k < 16; ++k) out16[k] = 0;
    for (size_t s = 0; s < ctx->lc.ks_slots; ++s)
        for (int k = 0; k < 16; ++k) out16[k] += h[s * 16 + k];
    out16[12] = out16[13] = 0;   // [12] = min, [13] = max CTA lifetime (ns) over the CTAs that ran
    for (size_t s = 0; s < ctx->lc.ks_slots; ++s) {
        const uint64_t ns = h[s * 16 + 14];
        if (!ns) continue;
        if (ns > out16[13]) out16[13] = ns;
        if (!out16[12] || ns < out16[12]) out16[12] = ns;
    }
    return DPFHE_OK;
}

// ---------------------------------------------------------------- encrypted linear layer (BASELINE config 4)
// y = sum_g rot_{g*baby}( sum_b D[g*baby + b] o rot_b(x) ): baby-step/giant-step diagonals on top of the hot-path ops.
// The weights (diagonal plaintexts) and Galois keys are uploaded ONCE when the layer is created; an application is
//   baby-1 hoisted rotations of the input (one shared digit decomposition)  -> dpfhe_rotate_hoisted
//   all giant-step inner sums in one pass over the baby steps               -> dpfhe_ct_mul_plain_inner
//   Horner over the giant steps: acc = rot_baby(acc) + inner[g]              -> dpfhe_rotate + dpfhe_poly_add
// and the host-buffer form pipelines chunks of the batch through it (upload / compute / download overlapped).
struct dpfhe_linear {
    dpfhe_ctx *ctx = nullptr;
    size_t n = 0, baby = 0, giant = 0;
    u64 *d_diags = nullptr;                 // [n][L][N]
    u64 *d_keys = nullptr;                  // [baby-1 + 1][L][2][L][N]: baby-step keys, then the giant-step key
    std::vector<uint64_t> g_baby;           // Galois elements 5^b, b = 1 .. baby-1
    std::vector<const uint64_t *> k_baby;   // device pointers of the baby-step keys
    u64 *d_prep = nullptr;                  // per baby step: Shoup companions of its key [L][2][L][N] + kprime [2][L][N], built once
    std::vector<const uint64_t *> prep;     // {companions, kprime} pointers per baby step (rotate_hoisted_impl)
    uint64_t g_giant = 0;
    u64 *scratch = nullptr;                 // [baby + giant + 1][cap][2][L][N]
    size_t cap = 0;                         // ciphertexts the scratch holds
};

static int linear_reserve(dpfhe_linear *lin, size_t batch) {
    if (batch <= lin->cap) return DPFHE_OK;
    dpfhe_ctx *ctx = lin->ctx;
    int rc = dpfhe_synchronize(ctx);
    if (rc) return rc;
    cudaFree(lin->scratch);
    lin->scratch = nullptr;
    lin->cap = 0;
    CU_TRY(cudaMalloc(&lin->scratch, (lin->baby + lin->giant + 1) * batch * 2 * ctx->P() * 8));
    lin->cap = batch;
    return DPFHE_OK;
}

int dpfhe_linear_create(dpfhe_ctx *ctx, const uint64_t *h_diags, size_t n_diags, size_t baby, const uint64_t *h_gk_baby, const uint64_t *h_gk_giant,
                        dpfhe_linear **out) {
    int rc = enter(ctx);
    if (rc) return rc;
    if (!out || !h_diags) return fail(DPFHE_ERR_INVALID, "null argument");
    *out = nullptr;
    if (baby == 0 || baby > 128 || n_diags == 0 || n_diags % baby) return fail(DPFHE_ERR_INVALID, "need 1 <= baby <= 128 and a multiple of baby diagonals");
    const size_t giant = n_diags / baby;
    if (giant > 65535) return fail(DPFHE_ERR_INVALID, "too many giant steps");
    if ((baby > 1 && !h_gk_baby) || (giant > 1 && !h_gk_giant)) return fail(DPFHE_ERR_INVALID, "missing Galois keys");
    dpfhe_linear *lin = new (std::nothrow) dpfhe_linear();
    if (!lin) return fail(DPFHE_ERR_NOMEM, "out of host memory");
    lin->ctx = ctx; lin->n = n_diags; lin->baby = baby; lin->giant = giant;
    const size_t P8 = ctx->P() * 8, key_bytes = 2 * ctx->hp.L * P8;
    cudaError_t e = cudaMalloc(&lin->d_diags, n_diags * P8);
    if (e == cudaSuccess) e = cudaMalloc(&lin->d_keys, baby * key_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(lin->d_diags, h_diags, n_diags * P8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && baby > 1) e = cudaMemcpy(lin->d_keys, h_gk_baby, (baby - 1) * key_bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && giant > 1) e = cudaMemcpy(lin->d_keys + (baby - 1) * key_bytes / 8, h_gk_giant, key_bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        dpfhe_linear_destroy(lin);
        return fail(DPFHE_ERR_CUDA, "linear layer upload: %s", cudaGetErrorString(e));
    }
    for (size_t b = 1; b < baby; ++b) {
        uint64_t g = 0;
        dpfhe_galois_element(ctx, (int)b, &g);
        lin->g_baby.push_back(g);
        lin->k_baby.push_back(lin->d_keys + (b - 1) * key_bytes / 8);
    }
    dpfhe_galois_element(ctx, (int)baby, &lin->g_giant);
    // the constants of the baby-step rotations do not depend on the data: prepare them once (four small launches per rotation
    // that every hoisted call would otherwise repeat — a tenth of a 31-rotation call at batch 512, more for smaller chunks)
    if (baby > 1) {
        const size_t ks_words = 2 * ctx->hp.L * ctx->P(), kp_words = 2 * ctx->P();
        int rc2 = ensure_hoist_consts(ctx);
        e = rc2 == DPFHE_OK ? cudaMalloc(&lin->d_prep, (baby - 1) * (ks_words + kp_words) * 8) : cudaErrorMemoryAllocation;
        cudaStream_t st = pick(ctx, nullptr);
        for (size_t b = 1; b < baby && e == cudaSuccess; ++b) {
            u64 *ks = lin->d_prep + (b - 1) * (ks_words + kp_words), *kp = ks + ks_words;
            e = VCALL(launch_rot_prepare, ctx->lc, lin->k_baby[b - 1], (u32)lin->g_baby[b - 1], ctx->hoist_delta, ctx->hoist_M, kp, st, ks);
            note_launch(ctx, 4);
            lin->prep.push_back(ks);
            lin->prep.push_back(kp);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            dpfhe_linear_destroy(lin);
            return fail(DPFHE_ERR_CUDA, "linear layer constants: %s", cudaGetErrorString(e));
        }
    }
    *out = lin;
    return DPFHE_OK;
}

void dpfhe_linear_destroy(dpfhe_linear *lin) {
    if (!lin) return;
    if (lin->ctx) {
        cudaSetDevice(lin->ctx->lc.device);
        dpfhe_synchronize(lin->ctx);
    }
    cudaFree(lin->d_diags);
    cudaFree(lin->d_keys);
    cudaFree(lin->d_prep);
    cudaFree(lin->scratch);
    delete lin;
}

static int linear_apply_on(dpfhe_linear *lin, const uint64_t *d_ct, uint64_t *d_out, size_t batch, void *stream) {
    dpfhe_ctx *ctx = lin->ctx;
    const size_t ctb = batch * 2 * ctx->P();                 // words of one ciphertext batch
    u64 *steps = lin->scratch, *inner = steps + lin->baby * ctb, *tmp = inner + lin->giant * ctb;
    cudaStream_t st = pick(ctx, stream);
    CU_TRY(cudaMemcpyAsync(steps, d_ct, ctb * 8, cudaMemcpyDeviceToDevice, st));
    int rc = DPFHE_OK;
    if (lin->baby > 1)
        rc = rotate_hoisted_impl(ctx, steps, lin->baby - 1, lin->g_baby.data(), lin->k_baby.data(), lin->prep.data(), steps + ctb, batch, stream);
    if (rc) return rc;
    rc = dpfhe_ct_mul_plain_inner(ctx, steps, lin->baby, lin->d_diags, lin->giant, inner, batch, stream);
    if (rc) return rc;
    st = pick(ctx, stream);
    CU_TRY(cudaMemcpyAsync(d_out, inner + (lin->giant - 1) * ctb, ctb * 8, cudaMemcpyDeviceToDevice, st));
    const u64 *gk_giant = lin->d_keys + (lin->baby - 1) * 2 * ctx->hp.L * ctx->P();
    for (size_t g = lin->giant - 1; g-- > 0;) {
        rc = dpfhe_rotate(ctx, d_out, lin->g_giant, gk_giant, tmp, batch, stream);           // Horner step: acc = rot_baby(acc) + inner[g]
        if (rc) return rc;
        rc = dpfhe_poly_add(ctx, tmp, inner + g * ctb, d_out, 2 * batch, stream);
        if (rc) return rc;
    }
    return DPFHE_OK;
}

int dpfhe_linear_apply(dpfhe_linear *lin, const uint64_t *d_ct, uint64_t *d_out, size_t batch, void *stream) {
    if (!lin) return fail(DPFHE_ERR_INVALID, "null layer");
    int rc = enter(lin->ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    CHECK_PTR(d_ct); CHECK_PTR(d_out);
    if (overlaps(d_out, batch * 2 * lin->ctx->P() * 8, d_ct, batch * 2 * lin->ctx->P() * 8)) return fail(DPFHE_ERR_INVALID, "output must not overlap the input");
    rc = linear_reserve(lin, batch);
    if (rc) return rc;
    return linear_apply_on(lin, d_ct, d_out, batch, stream);
}

int dpfhe_linear_apply_host(dpfhe_linear *lin, const uint64_t *h_ct, uint64_t *h_out, size_t batch) {
    if (!lin) return fail(DPFHE_ERR_INVALID, "null layer");
    dpfhe_ctx *ctx = lin->ctx;
    int rc = enter(ctx);
    if (rc) return rc;
    if (batch == 0) return DPFHE_OK;
    if (!h_ct || !h_out) return fail(DPFHE_ERR_INVALID, "null host pointer");
    // Chunks of about a fifth of the batch, rounded to whole rounds of the persistent key-switch grid (3 CTAs per SM, L CTAs
    // per ciphertext): a chunk that leaves the grid's last round mostly empty costs more than the transfers it hides.  The
    // first upload and the last download are the only transfers not overlapped with a neighbouring chunk's compute.
    const size_t groups = std::max<size_t>(1, (size_t)ctx->lc.num_sms * 3 / ctx->hp.L);
    size_t rounds = (batch / 5 + groups / 2) / groups;
    if (const char *e = getenv("DPFHE_LINEAR_CHUNK_ROUNDS")) rounds = (size_t)atol(e);   // tuning
    if (rounds < 1) rounds = 1;
    size_t chunk = rounds * groups;
    if (chunk > 512) chunk = std::max<size_t>(groups, 512 / groups * groups);
    if (chunk > batch) chunk = batch;
    rc = linear_reserve(lin, chunk);
    if (rc) return rc;
    const size_t ct_words = 2 * ctx->P();
    return run_pipeline(ctx, h_ct, nullptr, h_out, batch, ct_words, ct_words, chunk,
                        [&](u64 *din, u64 *, u64 *dout, size_t cnt, cudaStream_t st) -> int { return linear_apply_on(lin, din, dout, cnt, st); });
}

int dpfhe_describe(const dpfhe_ctx *ctx, char *buf, size_t buf_len) {
    if (!ctx || !buf || !buf_len) return fail(DPFHE_ERR_INVALID, "bad argument");
    const unsigned nt = ctx->hp.log_n == 12 ? 256 : 512;
    int n = snprintf(buf, buf_len,
                     "{\"log_n\": %u, \"n_limbs\": %u, \"num_sms\": %d, \"ntt_kernel\": {\"threads\": %u, \"smem_bytes\": %zu, "
                     "\"grid\": \"one CTA per limb\"}, \"ks_fused_kernel\": {\"threads\": %u, \"smem_bytes\": %zu, "
                     "\"grid\": \"persistent cooperative, multiple of L, <= %zu slots\"}}",
                     ctx->hp.log_n, ctx->hp.L, ctx->lc.num_sms, nt, ctx->N() * 8, 256u, ctx->hp.log_n <= 13 ? ctx->N() * 8 : ctx->N() * 4, ctx->lc.ks_slots);
    return n;
}

}  // extern "C"
