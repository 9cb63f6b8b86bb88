// multi.cu — one process, several GPUs (include/dpfhe.h: dpfhe_multi_*).
//
// Ciphertexts are independent, so a batch is cut into contiguous shards, one per device; keys and tables are replicated,
// nothing is exchanged while computing.  Each device has its own context (tables, scratch, streams) and is driven by
// its own host thread, bound to the CPUs of the GPU's NUMA node.
//   host buffers   : shard r streams through device r's H2D / compute / D2H pipeline and lands in h_out at its global
//                    position — no collective at all.
//   device buffers : the gathered result lives on ONE device.  The fused kernel writes each finished output row straight
//                    into that device's buffer through the peer mapping (NVLink), so the gather is spread over the whole
//                    compute phase instead of following it.
// Contrast: the reference's DistributedContext boots one MPI rank per GPU and gathers with a temporary + copy after the
// compute (src/core/distributed/distributed_context.cpp:107-119, 242-250).
#include <cuda_runtime.h>

#include <cstdio>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dpfhe.h"
#include "ctx.hpp"

int dpfhe_fail(int code, const char *fmt, ...);

struct dpfhe_multi {
    std::vector<dpfhe_ctx *> ctx;
    std::vector<int> dev;
};

namespace {

void shard_of(size_t batch, size_t n, size_t r, size_t *first, size_t *count) {
    const size_t base = batch / n, extra = batch % n;
    *first = r * base + (r < extra ? r : extra);
    *count = base + (r < extra ? 1 : 0);
}

// runs fn(r) on one thread per device and returns the first failure (its message copied to the caller's thread)
template <class F>
int for_each_device(dpfhe_multi *m, F fn) {
    const size_t n = m->ctx.size();
    std::vector<int> rc(n, DPFHE_OK);
    std::vector<std::string> msg(n);
    auto body = [&](size_t r) {
        int cpus = 0;
        dpfhe_bind_thread_near(m->ctx[r], &cpus);
        rc[r] = fn(r);
        if (rc[r] != DPFHE_OK) msg[r] = dpfhe_last_error();
    };
    if (n == 1) {
        rc[0] = fn(0);   // the caller's thread, unbound
        if (rc[0] != DPFHE_OK) msg[0] = dpfhe_last_error();
    } else {
        std::vector<std::thread> th;
        th.reserve(n);
        for (size_t r = 0; r < n; ++r) th.emplace_back(body, r);
        for (auto &t : th) t.join();
    }
    for (size_t r = 0; r < n; ++r)
        if (rc[r] != DPFHE_OK) return dpfhe_fail(rc[r], "device %d (shard %zu): %s", m->dev[r], r, msg[r].c_str());
    return DPFHE_OK;
}

}  // namespace

extern "C" {

int dpfhe_multi_create(const dpfhe_params *p, const int *device_ids, int n_devices, dpfhe_multi **out) {
    if (!p || !out) return dpfhe_fail(DPFHE_ERR_INVALID, "null argument");
    *out = nullptr;
    int avail = 0;
    if (cudaGetDeviceCount(&avail) != cudaSuccess || avail <= 0) return dpfhe_fail(DPFHE_ERR_CUDA, "no usable CUDA device; this library has no CPU fallback");
    if (n_devices <= 0) n_devices = avail;   // all of them
    if (n_devices > 64) return dpfhe_fail(DPFHE_ERR_INVALID, "at most 64 devices");
    dpfhe_multi *m = new dpfhe_multi();
    for (int r = 0; r < n_devices; ++r) {
        const int d = device_ids ? device_ids[r] : r;
        dpfhe_ctx *c = nullptr;
        const int rc = dpfhe_context_create(p, d, &c);
        if (rc != DPFHE_OK) {
            const std::string why = dpfhe_last_error();
            dpfhe_multi_destroy(m);
            return dpfhe_fail(rc, "context on device %d: %s", d, why.c_str());
        }
        m->ctx.push_back(c);
        m->dev.push_back(d);
    }
    // every device may store into every other one's memory (the overlapped gather); a pair without a peer path is found at
    // the first gather instead, with the CUDA error of the failing store
    for (int a = 0; a < n_devices; ++a)
        for (int b = 0; b < n_devices; ++b) {
            if (m->dev[a] == m->dev[b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, m->dev[a], m->dev[b]) != cudaSuccess || !can) continue;
            cudaSetDevice(m->dev[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(m->dev[b], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        }
    *out = m;
    return DPFHE_OK;
}

void dpfhe_multi_destroy(dpfhe_multi *m) {
    if (!m) return;
    for (dpfhe_ctx *c : m->ctx) dpfhe_context_destroy(c);
    delete m;
}

int dpfhe_multi_device_count(const dpfhe_multi *m) { return m ? (int)m->ctx.size() : 0; }

dpfhe_ctx *dpfhe_multi_context(dpfhe_multi *m, int index) {
    if (!m || index < 0 || (size_t)index >= m->ctx.size()) return nullptr;
    return m->ctx[(size_t)index];
}

int dpfhe_multi_shard(const dpfhe_multi *m, size_t batch, int index, size_t *first, size_t *count) {
    if (!m || !first || !count || index < 0 || (size_t)index >= m->ctx.size()) return dpfhe_fail(DPFHE_ERR_INVALID, "bad argument");
    shard_of(batch, m->ctx.size(), (size_t)index, first, count);
    return DPFHE_OK;
}

int dpfhe_multi_ct_mul_relin_host(dpfhe_multi *m, const uint64_t *h_a, const uint64_t *h_b, const uint64_t *h_evk, uint64_t *h_out, size_t batch) {
    if (!m || m->ctx.empty()) return dpfhe_fail(DPFHE_ERR_INVALID, "null multi-device context");
    if (batch == 0) return DPFHE_OK;
    if (!h_a || !h_b || !h_evk || !h_out) return dpfhe_fail(DPFHE_ERR_INVALID, "null host pointer");
    const size_t ct_words = 2 * m->ctx[0]->P();
    return for_each_device(m, [&](size_t r) {
        size_t first, count;
        shard_of(batch, m->ctx.size(), r, &first, &count);
        return dpfhe_ct_mul_relin_host(m->ctx[r], h_a + first * ct_words, h_b + first * ct_words, h_evk, h_out + first * ct_words, count);
    });
}

// special-prime key switching (dpfhe_ct_mul_relin_grouped_host) sharded the same way: ciphertexts carry L - n_special limbs
int dpfhe_multi_ct_mul_relin_grouped_host(dpfhe_multi *m, unsigned n_special, const uint64_t *h_a, const uint64_t *h_b, const uint64_t *h_evk,
                                          uint64_t *h_out, size_t batch, uint64_t t_plain) {
    if (!m || m->ctx.empty()) return dpfhe_fail(DPFHE_ERR_INVALID, "null multi-device context");
    if (batch == 0) return DPFHE_OK;
    if (!h_a || !h_b || !h_evk || !h_out) return dpfhe_fail(DPFHE_ERR_INVALID, "null host pointer");
    if (n_special < 1 || n_special >= m->ctx[0]->hp.L) return dpfhe_fail(DPFHE_ERR_INVALID, "n_special must be at least 1 and below the context's limbs");
    const size_t ct_words = 2 * (m->ctx[0]->hp.L - n_special) * m->ctx[0]->N();
    return for_each_device(m, [&](size_t r) {
        size_t first, count;
        shard_of(batch, m->ctx.size(), r, &first, &count);
        return dpfhe_ct_mul_relin_grouped_host(m->ctx[r], n_special, h_a + first * ct_words, h_b + first * ct_words, h_evk, h_out + first * ct_words, count,
                                               t_plain);
    });
}

int dpfhe_multi_rotate_host(dpfhe_multi *m, const uint64_t *h_ct, uint64_t galois_elt, const uint64_t *h_gk, uint64_t *h_out, size_t batch) {
    if (!m || m->ctx.empty()) return dpfhe_fail(DPFHE_ERR_INVALID, "null multi-device context");
    if (batch == 0) return DPFHE_OK;
    if (!h_ct || !h_gk || !h_out) return dpfhe_fail(DPFHE_ERR_INVALID, "null host pointer");
    const size_t ct_words = 2 * m->ctx[0]->P();
    return for_each_device(m, [&](size_t r) {
        size_t first, count;
        shard_of(batch, m->ctx.size(), r, &first, &count);
        return dpfhe_rotate_host(m->ctx[r], h_ct + first * ct_words, galois_elt, h_gk, h_out + first * ct_words, count);
    });
}

// d_a[r], d_b[r], d_evk[r]: device r's shard of the operands and its copy of the key (device r memory).
// d_out_root: [batch][2][L][N] on the device of shard `root`.  Synchronous.
int dpfhe_multi_ct_mul_relin_gather(dpfhe_multi *m, const uint64_t *const *d_a, const uint64_t *const *d_b, const uint64_t *const *d_evk,
                                    uint64_t *d_out_root, int root, size_t batch) {
    if (!m || m->ctx.empty()) return dpfhe_fail(DPFHE_ERR_INVALID, "null multi-device context");
    if (root < 0 || (size_t)root >= m->ctx.size()) return dpfhe_fail(DPFHE_ERR_INVALID, "root %d out of range", root);
    if (batch == 0) return DPFHE_OK;
    if (!d_a || !d_b || !d_evk || !d_out_root) return dpfhe_fail(DPFHE_ERR_INVALID, "null argument");
    const size_t ct_words = 2 * m->ctx[0]->P();
    return for_each_device(m, [&](size_t r) {
        size_t first, count;
        shard_of(batch, m->ctx.size(), r, &first, &count);
        if (count == 0) return (int)DPFHE_OK;
        // the output rows of this shard are rows [first, first + count) of the root's buffer: the kernel's final stores go there
        int rc = dpfhe_ct_mul_relin(m->ctx[r], d_a[r], d_b[r], d_evk[r], d_out_root + first * ct_words, count, nullptr);
        if (rc != DPFHE_OK) return rc;
        return dpfhe_synchronize(m->ctx[r]);
    });
}

}  // extern "C"
