// hostmem.cu — pinned host staging memory placed next to a GPU (include/dpfhe.h: dpfhe_host_alloc*).
//
// The host-buffer entry points are PCIe-bound; on a two-socket host the staging pages must sit on the NUMA node the GPU
// hangs off, or every transfer crosses the socket interconnect (round 1: 50 GB/s per GPU alone, 21 GB/s with eight GPUs
// pulling from wherever first-touch had put the pages).  No libnuma in this image: sysfs for the topology, the raw mbind
// system call for the placement, cudaHostRegister for the pinning.
#include <cuda_runtime.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cctype>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

#include "../../include/dpfhe.h"
#include "ctx.hpp"

int dpfhe_fail(int code, const char *fmt, ...);

namespace {

struct Mapping {
    void *base;          // what mmap returned
    size_t map_bytes;    // and its length
    size_t reg_bytes;    // bytes registered with CUDA, starting at the user pointer
};
std::mutex g_mu;
std::map<void *, Mapping> g_mapped;   // regions handed out by dpfhe_host_alloc_near, keyed by the user pointer

// NUMA node of a CUDA device from sysfs, -1 if unknown
int device_numa_node(int device) {
    char bus[32] = {0};
    if (cudaDeviceGetPCIBusId(bus, (int)sizeof(bus), device) != cudaSuccess) return -1;
    for (char *c = bus; *c; ++c) *c = (char)tolower(*c);
    const std::string path = std::string("/sys/bus/pci/devices/") + bus + "/numa_node";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// parses a sysfs cpulist ("0-31,64-95") into a cpu_set_t; returns the number of CPUs
int parse_cpulist(const char *s, cpu_set_t *set) {
    CPU_ZERO(set);
    int n = 0;
    while (*s) {
        char *end = nullptr;
        long a = strtol(s, &end, 10);
        if (end == s) break;
        long b = a;
        if (*end == '-') {
            s = end + 1;
            b = strtol(s, &end, 10);
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; ++c) {
            CPU_SET((int)c, set);
            ++n;
        }
        s = end;
        while (*s == ',' || *s == '\n' || *s == ' ') ++s;
    }
    return n;
}

}  // namespace

extern "C" {

int dpfhe_host_alloc(void **out, size_t bytes) {
    if (!out) return dpfhe_fail(DPFHE_ERR_INVALID, "null argument");
    cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) return dpfhe_fail(DPFHE_ERR_CUDA, "cudaHostAlloc(%zu): %s", bytes, cudaGetErrorString(e));
    return DPFHE_OK;
}

int dpfhe_device_numa_node(const dpfhe_ctx *ctx, int *node) {
    if (!ctx || !node) return dpfhe_fail(DPFHE_ERR_INVALID, "null argument");
    *node = device_numa_node(ctx->lc.device);
    return DPFHE_OK;
}

// Pinned host memory whose pages live on the NUMA node of the context's GPU.  *placed_node receives that node, or -1 when
// the placement could not be enforced (unknown topology, mbind refused): the memory is then ordinary pinned memory.
int dpfhe_host_alloc_near(const dpfhe_ctx *ctx, void **out, size_t bytes, int *placed_node) {
    if (!ctx || !out) return dpfhe_fail(DPFHE_ERR_INVALID, "null argument");
    if (placed_node) *placed_node = -1;
    *out = nullptr;
    if (bytes == 0) return DPFHE_OK;
    cudaError_t ce = cudaSetDevice(ctx->lc.device);
    if (ce != cudaSuccess) return dpfhe_fail(DPFHE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(ce));
    // 2 MiB-aligned and advised for transparent huge pages: fewer, larger translations for the DMA engines when eight GPUs
    // stream from host memory at once (the advice is best effort; plain pages work the same way)
    const size_t huge = (size_t)2 << 20, len = (bytes + huge - 1) / huge * huge, map_bytes = len + huge;
    void *base = mmap(nullptr, map_bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return dpfhe_fail(DPFHE_ERR_NOMEM, "mmap of %zu bytes failed", map_bytes);
    void *p = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(base) + huge - 1) / huge * huge);
    if (!getenv("DPFHE_NO_HUGEPAGES")) madvise(p, len, MADV_HUGEPAGE);
    const int node = device_numa_node(ctx->lc.device);
    bool bound = false;
    if (node >= 0 && node < 1024) {
        unsigned long mask[16] = {0};
        mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
        // MPOL_BIND = 2: every page of the range comes from `node` when it is first touched (below)
        bound = syscall(SYS_mbind, p, len, 2, mask, (unsigned long)(8 * sizeof(mask)), 0u) == 0;
    }
    memset(p, 0, len);   // first touch: allocates the pages under the policy
    ce = cudaHostRegister(p, len, cudaHostRegisterPortable);
    if (ce != cudaSuccess) {
        munmap(base, map_bytes);
        return dpfhe_fail(DPFHE_ERR_CUDA, "cudaHostRegister(%zu): %s", len, cudaGetErrorString(ce));
    }
    {
        std::lock_guard<std::mutex> g(g_mu);
        g_mapped[p] = Mapping{base, map_bytes, len};
    }
    *out = p;
    if (placed_node && bound) *placed_node = node;
    return DPFHE_OK;
}

int dpfhe_host_free(void *p) {
    if (!p) return DPFHE_OK;
    Mapping m = {nullptr, 0, 0};
    {
        std::lock_guard<std::mutex> g(g_mu);
        auto it = g_mapped.find(p);
        if (it != g_mapped.end()) {
            m = it->second;
            g_mapped.erase(it);
        }
    }
    if (m.base) {
        cudaHostUnregister(p);
        munmap(m.base, m.map_bytes);
        return DPFHE_OK;
    }
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) return dpfhe_fail(DPFHE_ERR_CUDA, "cudaFreeHost: %s", cudaGetErrorString(e));
    return DPFHE_OK;
}

// Restricts the CALLING thread to the CPUs of the NUMA node of the context's GPU (so that what it first-touches, and the
// driver work it does, stay on that socket).  *n_cpus receives the size of the set, 0 when the topology is unknown.
int dpfhe_bind_thread_near(const dpfhe_ctx *ctx, int *n_cpus) {
    if (!ctx) return dpfhe_fail(DPFHE_ERR_INVALID, "null argument");
    if (n_cpus) *n_cpus = 0;
    const int node = device_numa_node(ctx->lc.device);
    if (node < 0) return DPFHE_OK;
    char path[128], buf[4096] = {0};
    snprintf(path, sizeof(path), "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return DPFHE_OK;
    const size_t got = fread(buf, 1, sizeof(buf) - 1, f);
    fclose(f);
    buf[got] = 0;
    cpu_set_t set;
    const int n = parse_cpulist(buf, &set);
    if (n > 0 && sched_setaffinity(0, sizeof(set), &set) == 0 && n_cpus) *n_cpus = n;
    return DPFHE_OK;
}

}  // extern "C"
