// ntt_harness.cu — times the library's own forward / inverse limb transforms (kernel_bodies.cuh) for one arithmetic
// variant selected at compile time (-DDPFHE_FAST=0|1 -DDPFHE_SHOUP_APPROX=0|1|2), checked against a plain host NTT.
// Build (from the repo root):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDPFHE_FAST=1 -I deeppowers_b200/csrc \
//        profiles/microbench/ntt_harness.cu deeppowers_b200/csrc/host_params.cpp -o profiles/microbench/ntt_f1_a2
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "host_params.hpp"
#include "kernel_bodies.cuh"
using namespace dpfhe;
using namespace dpfhe::DPFHE_VNS;

template <int NT>
struct DevCta {
    __device__ __forceinline__ void mark(int) {}
    __device__ __forceinline__ void wait_ge(const u32 *, u32) {}
    template <class F> __device__ __forceinline__ void par(F f) { f((int)threadIdx.x); __syncthreads(); }
    template <class F> __device__ __forceinline__ void par_dom(F f) { f((int)threadIdx.x); __syncthreads(); }
    template <class F> __device__ __forceinline__ void par_warp(F f) { f((int)threadIdx.x); __syncwarp(); }
};

template <int LOGN, int NT, int MINB, bool INVERSE>
__global__ void __launch_bounds__(NT, MINB) ntt_kernel(u64 *data, const Twiddle *__restrict__ tables, const __grid_constant__ LimbTable lt, u32 L, size_t n_limbs) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    constexpr size_t N = (size_t)1 << LOGN;
    DevCta<NT> cta;
    for (size_t w = blockIdx.x; w < n_limbs; w += gridDim.x) {
        const u32 l = (u32)(w % L);
        const LimbParams &p = lt.lp[l];
        if (INVERSE) ntt_inv_body<LOGN, NT>(cta, buf, data + w * N, tables + (size_t)l * N, p);
        else ntt_fwd_body<LOGN, NT>(cta, buf, data + w * N, tables + (size_t)l * N, p);
    }
}

static u64 sm64(u64 &s) { s += 0x9E3779B97F4A7C15ull; u64 z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

// plain host transform: Cooley-Tukey, natural in, bit-reversed out, twiddle of group i at stage s = root_powers[2^s + i]
static void host_ntt(std::vector<u64> &a, const HostLimb &hl) {
    const size_t n = a.size();
    const u64 q = hl.lp.q;
    size_t t = n;
    for (size_t m = 1; m < n; m <<= 1) {
        t >>= 1;
        for (size_t i = 0; i < m; ++i) {
            const u64 w = hl.root_powers[m + i];
            for (size_t j = 2 * i * t; j < 2 * i * t + t; ++j) {
                const u64 u = a[j], v = host_mulmod(a[j + t], w, q);
                a[j] = (u + v) % q;
                a[j + t] = (u + q - v) % q;
            }
        }
    }
}

int main() {
    constexpr int LOGN = 13, NT = 256, MINB = 3;
    const unsigned L = 4;
    const size_t N = (size_t)1 << LOGN, n_limbs = 32768;
    HostParams hp;
#if DPFHE_FAST
    std::string msg = build_host_params(LOGN, L, nullptr, hp);
#else
    // the largest primes below 2^60 that are 1 mod 2N (the generic class)
    std::vector<u64> qs;
    for (u64 cand = ((1ull << 60) / (2 * N)) * (2 * N) + 1 - 2 * N; qs.size() < L; cand -= 2 * N)
        if (host_is_prime(cand) && (u32)cand != 1u) qs.push_back(cand);
    std::string msg = build_host_params(LOGN, L, qs.data(), hp);
#endif
    if (!msg.empty()) { printf("params: %s\n", msg.c_str()); return 1; }
    LimbTable lt = {};
    for (unsigned l = 0; l < L; ++l) lt.lp[l] = hp.limbs[l].lp;
    Twiddle *tw, *itw; u64 *data;
    cudaMalloc(&tw, L * N * sizeof(Twiddle)); cudaMalloc(&itw, L * N * sizeof(Twiddle));
    for (unsigned l = 0; l < L; ++l) {
        cudaMemcpy(tw + l * N, hp.limbs[l].tw.data(), N * sizeof(Twiddle), cudaMemcpyHostToDevice);
        cudaMemcpy(itw + l * N, hp.limbs[l].itw.data(), N * sizeof(Twiddle), cudaMemcpyHostToDevice);
    }
    std::vector<u64> h(n_limbs * N);
    u64 seed = 7;
    for (size_t w = 0; w < n_limbs; ++w) {
        const u64 q = hp.limbs[w % L].lp.q;
        for (size_t k = 0; k < N; ++k) h[w * N + k] = (u64)(((unsigned __int128)sm64(seed) * q) >> 64);
    }
    for (size_t k = 0; k < N; ++k) { h[k] = hp.limbs[0].lp.q - 1; h[N + k] = k & 1 ? hp.limbs[1].lp.q - 1 : 0; }   // extreme rows
    cudaMalloc(&data, n_limbs * N * 8);
    cudaMemcpy(data, h.data(), n_limbs * N * 8, cudaMemcpyHostToDevice);
    auto kf = ntt_kernel<LOGN, NT, MINB, false>;
    auto ki = ntt_kernel<LOGN, NT, MINB, true>;
    const int smem = (int)(N * 8);
    cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // correctness: forward result of a few limbs vs the host transform, then the round trip of everything
    kf<<<(unsigned)n_limbs, NT, smem>>>(data, tw, lt, L, n_limbs);
    std::vector<u64> g(8 * N);
    cudaMemcpy(g.data(), data, 8 * N * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int w = 0; w < 8; ++w) {
        std::vector<u64> ref(h.begin() + w * N, h.begin() + (w + 1) * N);
        host_ntt(ref, hp.limbs[w % L]);
        for (size_t k = 0; k < N; ++k) bad += ref[k] != g[w * N + k];
    }
    ki<<<(unsigned)n_limbs, NT, smem>>>(data, itw, lt, L, n_limbs);
    std::vector<u64> back(n_limbs * N);
    cudaMemcpy(back.data(), data, n_limbs * N * 8, cudaMemcpyDeviceToHost);
    size_t rt_bad = 0;
    for (size_t k = 0; k < n_limbs * N; ++k) rt_bad += back[k] != h[k];
    float ms_f = 0, ms_i = 0;
    const int reps = 5;
    for (int w = 0; w < 2; ++w) { kf<<<(unsigned)n_limbs, NT, smem>>>(data, tw, lt, L, n_limbs); ki<<<(unsigned)n_limbs, NT, smem>>>(data, itw, lt, L, n_limbs); }
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) kf<<<(unsigned)n_limbs, NT, smem>>>(data, tw, lt, L, n_limbs);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_f, e0, e1);
    cudaEventRecord(e0);
    for (int r = 0; r < reps; ++r) ki<<<(unsigned)n_limbs, NT, smem>>>(data, itw, lt, L, n_limbs);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms_i, e0, e1);
    printf("FAST=%d APPROX=%d (SB=%d)  fwd wrong %d  round-trip wrong %zu  fwd %.3f ms = %.2f M NTT/s   inv %.3f ms = %.2f M NTT/s   %s\n", DPFHE_FAST, DPFHE_SHOUP_APPROX, SB,
           bad, rt_bad, ms_f / reps, n_limbs / (ms_f / reps) / 1e3, ms_i / reps, n_limbs / (ms_i / reps) / 1e3, cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
