// mont_bfly.cu — SASS / throughput study of butterfly arithmetic for "fast" primes q = qh*2^32 + 1 (2^59 < q < 2^60).
// Variants of t = y*w mod q:  V=0 Shoup (library), V>=1 word-serial Montgomery with the prime's structure.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../deeppowers_b200/csrc -cubin mont_bfly.cu -o mont_bfly.cubin
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "ntt_core.cuh"
using namespace dpfhe;

// V1: two word steps, each: multiply-accumulate, negate low word, fold m*qh with the borrow of the negation as carry-in
__device__ __forceinline__ u64 montmul_v1(u64 y, u64 w, u32 qh) {
    u64 r;
    asm("{\n\t"
        ".reg .u32 yl, yh, wl, wh, s0, c0, p1l, p1h, m0, ul, uh, s1, c1, Wl, Wh, m1, rl, rh;\n\t"
        "mov.b64 {yl, yh}, %1;\n\t"
        "mov.b64 {wl, wh}, %2;\n\t"
        "mul.lo.u32 s0, yl, wl;\n\t"
        "mul.hi.u32 c0, yl, wl;\n\t"
        "mad.lo.cc.u32 p1l, yl, wh, c0;\n\t"
        "madc.hi.u32 p1h, yl, wh, 0;\n\t"
        "sub.cc.u32 m0, 0, s0;\n\t"
        "madc.lo.cc.u32 ul, m0, %3, p1l;\n\t"
        "madc.hi.u32 uh, m0, %3, p1h;\n\t"
        "mad.lo.cc.u32 s1, yh, wl, ul;\n\t"
        "madc.hi.u32 c1, yh, wl, 0;\n\t"
        "mad.lo.cc.u32 Wl, yh, wh, uh;\n\t"
        "madc.hi.u32 Wh, yh, wh, 0;\n\t"
        "add.cc.u32 Wl, Wl, c1;\n\t"
        "addc.u32 Wh, Wh, 0;\n\t"
        "sub.cc.u32 m1, 0, s1;\n\t"
        "madc.lo.cc.u32 rl, m1, %3, Wl;\n\t"
        "madc.hi.u32 rh, m1, %3, Wh;\n\t"
        "mov.b64 %0, {rl, rh};\n\t"
        "}"
        : "=l"(r)
        : "l"(y), "l"(w), "r"(qh));
    return r;
}

// V2: same dataflow written with 64-bit wide multiplies and explicit zero-extended addends
__device__ __forceinline__ u64 montmul_v2(u64 y, u64 w, u32 qh) {
    const u32 yl = (u32)y, yh = (u32)(y >> 32), wl = (u32)w, wh = (u32)(w >> 32);
    const u64 p0 = (u64)yl * wl;
    const u64 p1 = (u64)yl * wh + (p0 >> 32);
    const u32 s0 = (u32)p0, m0 = 0u - s0;
    const u64 U = (u64)m0 * qh + p1 + (s0 != 0 ? 1u : 0u);
    const u64 v0 = (u64)yh * wl + (u32)U;
    const u64 W = (u64)yh * wh + (U >> 32) + (v0 >> 32);
    const u32 s1 = (u32)v0, m1 = 0u - s1;
    return (u64)m1 * qh + W + (s1 != 0 ? 1u : 0u);
}


// approximate high product: drops the low partial product, so the result is hi64(y*ws) or one less
__device__ __forceinline__ u64 mulhi_approx_c(u64 y, u64 ws) {
    const u32 yl = (u32)y, yh = (u32)(y >> 32), sl = (u32)ws, sh = (u32)(ws >> 32);
    const u64 m1 = (u64)yh * sl, m2 = (u64)yl * sh;
    return (u64)yh * sh + (m1 >> 32) + (m2 >> 32);   // hi64(y*ws) - {0,1,2}
}
template <int F>
__device__ __forceinline__ u64 mulhi_approx_ptx(u64 y, u64 ws) {
    u64 h;
    if (F == 1) {
    asm("{\n\t"
        ".reg .u32 yl, yh, sl, sh, m0, m1, m2, h0, h1;\n\t"
        ".reg .u64 A, H, M;\n\t"
        "mov.b64 {yl, yh}, %1;\n\t"
        "mov.b64 {sl, sh}, %2;\n\t"
        "mul.wide.u32 A, yh, sl;\n\t"
        "mov.b64 {m0, m1}, A;\n\t"
        "mad.lo.cc.u32 m0, yl, sh, m0;\n\t"
        "madc.hi.cc.u32 m1, yl, sh, m1;\n\t"
        "addc.u32 m2, 0, 0;\n\t"
        "mov.b64 M, {m1, m2};\n\t"
        "mad.wide.u32 %0, yh, sh, M;\n\t"
        "}"
        : "=l"(h)
        : "l"(y), "l"(ws));
    } else if (F == 2) {
    asm("{\n\t"
        ".reg .u32 yl, yh, sl, sh, m0, m1, m2, h0, h1;\n\t"
        ".reg .u64 A, H;\n\t"
        "mov.b64 {yl, yh}, %1;\n\t"
        "mov.b64 {sl, sh}, %2;\n\t"
        "mul.wide.u32 A, yh, sl;\n\t"
        "mov.b64 {m0, m1}, A;\n\t"
        "mad.lo.cc.u32 m0, yl, sh, m0;\n\t"
        "madc.hi.cc.u32 m1, yl, sh, m1;\n\t"
        "addc.u32 m2, 0, 0;\n\t"
        "mul.wide.u32 H, yh, sh;\n\t"
        "mov.b64 {h0, h1}, H;\n\t"
        "add.cc.u32 h0, h0, m1;\n\t"
        "addc.u32 h1, h1, m2;\n\t"
        "mov.b64 %0, {h0, h1};\n\t"
        "}"
        : "=l"(h)
        : "l"(y), "l"(ws));
    } else {
    // order the chain so that the carry of the middle sum feeds the top product directly
    asm("{\n\t"
        ".reg .u32 yl, yh, sl, sh, m0, m1, h0, h1;\n\t"
        ".reg .u64 A;\n\t"
        "mov.b64 {yl, yh}, %1;\n\t"
        "mov.b64 {sl, sh}, %2;\n\t"
        "mul.wide.u32 A, yh, sl;\n\t"
        "mov.b64 {m0, m1}, A;\n\t"
        "mad.lo.cc.u32 m0, yl, sh, m0;\n\t"
        "madc.hi.cc.u32 m1, yl, sh, m1;\n\t"
        "madc.hi.u32 h1, yh, sh, 0;\n\t"
        "mad.lo.cc.u32 h0, yh, sh, m1;\n\t"
        "addc.u32 h1, h1, 0;\n\t"
        "mov.b64 %0, {h0, h1};\n\t"
        "}"
        : "=l"(h)
        : "l"(y), "l"(ws));
    }
    return h;
}
// V3: generic modulus, approximate quotient: [0, 3q)
template <int PTX>
__device__ __forceinline__ u64 shoup3(u64 x, u64 w, u64 ws, u64 nq) {
    const u64 h = PTX ? mulhi_approx_ptx<PTX>(x, ws) : mulhi_approx_c(x, ws);
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
    return t;
}
// V4: q = qh*2^32 + 1: h*q = h + ((hl*qh) << 32)
template <int PTX>
__device__ __forceinline__ u64 shoup3_fast(u64 x, u64 w, u64 ws, u32 nqh) {
    const u64 h = PTX ? mulhi_approx_ptx<PTX>(x, ws) : mulhi_approx_c(x, ws);
    u64 t;
    asm("{\n\t"
        ".reg .u32 xl, xh, wl, wh, hl, hh, t0, t1;\n\t"
        ".reg .u64 T;\n\t"
        "mov.b64 {xl, xh}, %1;\n\t"
        "mov.b64 {wl, wh}, %2;\n\t"
        "mov.b64 {hl, hh}, %3;\n\t"
        "mul.wide.u32 T, xl, wl;\n\t"
        "mov.b64 {t0, t1}, T;\n\t"
        "mad.lo.u32 t1, xl, wh, t1;\n\t"
        "mad.lo.u32 t1, xh, wl, t1;\n\t"
        "mad.lo.u32 t1, hl, %4, t1;\n\t"
        "mov.b64 T, {t0, t1};\n\t"
        "sub.u64 %0, T, %3;\n\t"
        "}"
        : "=l"(t)
        : "l"(x), "l"(w), "l"(h), "r"(nqh));
    return t;
}

template <int V>
__device__ __forceinline__ void bfly(u64 &x, u64 &y, const Twiddle &w, const LimbParams &p, u32 qh) {
    u64 t;
    if (V == 0) t = shoup_lazy(y, w.x, w.y, p);
    else if (V == 1) t = montmul_v1(y, w.x, qh);
    else if (V == 2) t = montmul_v2(y, w.x, qh);
    else if (V == 3) t = shoup3<0>(y, w.x, w.y, p.nq);
    else if (V == 4) t = shoup3<2>(y, w.x, w.y, p.nq);
    else if (V == 5) t = shoup3_fast<0>(y, w.x, w.y, qh);
    else if (V == 6) t = shoup3_fast<2>(y, w.x, w.y, qh);
    else t = shoup3_fast<3>(y, w.x, w.y, qh);
    const u64 a = x;
    x = a + t;
    y = a + (V >= 3 ? p.q4 - p.q : p.q2) - t;
}

template <int V>
__global__ void __launch_bounds__(256, 3) k(u64 *data, const Twiddle *tw, const __grid_constant__ LimbParams p, u32 qh, int iters) {
    u64 x[16];
    u64 *base = data + (size_t)(blockIdx.x * 256 + threadIdx.x) * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = base[i];
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const Twiddle *t = tw + ((it * 37 + (threadIdx.x >> 5)) & 1023) * 16;   // warp-uniform: broadcast loads, as in passes A and B
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int half = 8 >> u;
#pragma unroll
            for (int j = 0; j < (1 << u); ++j) {
                const Twiddle w = (V == 0 || V >= 3) ? t[(1 << u) + j] : Twiddle{reinterpret_cast<const u64 *>(t)[(1 << u) + j], 0};
#pragma unroll
                for (int i = 0; i < half; ++i) {
                    if (u == 2) x[j * 2 * half + i] = csub(x[j * 2 * half + i], p.q8);
                    bfly<V>(x[j * 2 * half + i], x[j * 2 * half + half + i], w, p, qh);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) base[i] = x[i];
}


// ---- self-check: t = variant(y, w) on random operands, verified on the host with 128-bit arithmetic ----
template <int V>
__global__ void chk(const u64 *y, const u64 *w, const u64 *ws, const u64 *wm, u64 *out, const __grid_constant__ LimbParams p, u32 qh, u32 nqh, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u64 t;
    if (V == 0) t = shoup_lazy(y[i], w[i], ws[i], p);
    else if (V == 1) t = montmul_v1(y[i], wm[i], qh);
    else if (V == 2) t = montmul_v2(y[i], wm[i], qh);
    else if (V == 3) t = shoup3<0>(y[i], w[i], ws[i], p.nq);
    else if (V == 4) t = shoup3<2>(y[i], w[i], ws[i], p.nq);
    else if (V == 5) t = shoup3_fast<0>(y[i], w[i], ws[i], nqh);
    else if (V == 6) t = shoup3_fast<2>(y[i], w[i], ws[i], nqh);
    else t = shoup3_fast<3>(y[i], w[i], ws[i], nqh);
    out[i] = t;
}

typedef unsigned __int128 u128;
static u64 splitmix(u64 &s) { s += 0x9E3779B97F4A7C15ull; u64 z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static u64 powmod(u64 a, u64 e, u64 q) { u64 r = 1; while (e) { if (e & 1) r = (u64)((u128)r * a % q); a = (u64)((u128)a * a % q); e >>= 1; } return r; }

template <int V>
static void run(const char *name, int bound, bool mont, u64 *dy, u64 *dw, u64 *dws, u64 *dwm, u64 *dout, const std::vector<u64> &y, const std::vector<u64> &w,
                LimbParams p, u32 qh, u32 nqh, u64 *data, Twiddle *tw, int sms) {
    const int n = (int)y.size();
    chk<V><<<(n + 255) / 256, 256>>>(dy, dw, dws, dwm, dout, p, qh, nqh, n);
    std::vector<u64> out(n);
    cudaMemcpy(out.data(), dout, n * 8, cudaMemcpyDeviceToHost);
    const u64 q = p.q;
    const u64 rinv = powmod((u64)(((u128)1 << 64) % q), q - 2, q);
    int bad = 0; double maxb = 0;
    for (int i = 0; i < n; ++i) {
        u64 want = (u64)((u128)(y[i] % q) * w[i] % q);
        (void)rinv; (void)mont;
        if (out[i] % q != want) ++bad;
        const double b = (double)out[i] / (double)q;
        if (b > maxb) maxb = b;
    }
    const int iters = 2000, grid = sms * 3;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<grid, 256>>>(data, tw, p, V >= 5 ? nqh : qh, 50);
    cudaEventRecord(e0);
    k<V><<<grid, 256>>>(data, tw, p, V >= 5 ? nqh : qh, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double bf = (double)grid * 256 * 32 * iters;
    printf("%-34s wrong %d/%d  max t/q %.4f (claimed < %d)  %.3f ms  %.2f Gbfly/s  %.1f SM-clk per warp-butterfly-SMSP @1.965GHz\n", name, bad, n, maxb, bound, ms,
           bf / ms / 1e6, ms * 1e-3 * 1.965e9 * sms * 4 / (bf / 32));
}

int main() {
    const u64 q = 1152921092289986561ull;   // 2^60 - 96*2^32 + 1
    LimbParams p = {};
    p.q = q; p.q2 = 2 * q; p.q4 = 4 * q; p.q8 = 8 * q; p.nq = 0 - q;
    const u32 qh = (u32)(q >> 32), nqh = 0u - qh;
    int dev = 0, sms = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n = 1 << 20;
    std::vector<u64> y(n), w(n), ws(n), wm(n);
    u64 seed = 12345;
    for (int i = 0; i < n; ++i) {
        y[i] = splitmix(seed);
        if (i % 7 == 0) y[i] = 16 * q - 1 - (i & 1023);        // top of the lazy range
        if (i % 7 == 1) y[i] = ~0ull - (u64)(i & 1023);         // any 64-bit value
        if (i % 7 == 2) y[i] = (u64)(i & 3);
        if (i % 11 == 3) y[i] &= 0xffffffff00000000ull;          // zero low word
        w[i] = splitmix(seed) % q;
        if (i % 13 == 0) w[i] = q - 1 - (i & 7);
        ws[i] = (u64)((((u128)w[i]) << 64) / q);
        wm[i] = (u64)((((u128)w[i]) << 64) % q);
    }
    u64 *dy, *dw, *dws, *dwm, *dout, *data; Twiddle *tw;
    cudaMalloc(&dy, n * 8); cudaMalloc(&dw, n * 8); cudaMalloc(&dws, n * 8); cudaMalloc(&dwm, n * 8); cudaMalloc(&dout, n * 8);
    cudaMemcpy(dy, y.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dw, w.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dws, ws.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(dwm, wm.data(), n * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&data, (size_t)sms * 3 * 256 * 16 * 8); cudaMemset(data, 1, (size_t)sms * 3 * 256 * 16 * 8);
    std::vector<Twiddle> htw(1024 * 16 + 16);
    for (auto &t : htw) { t.x = splitmix(seed) % q; t.y = (u64)((((u128)t.x) << 64) / q); }
    cudaMalloc(&tw, htw.size() * sizeof(Twiddle)); cudaMemcpy(tw, htw.data(), htw.size() * sizeof(Twiddle), cudaMemcpyHostToDevice);
    run<0>("V0 Shoup exact (library)", 2, false, dy, dw, dws, dwm, dout, y, w, p, qh, nqh, data, tw, sms);
    run<1>("V1 Montgomery fast (carry chain)", 3, true, dy, dw, dws, dwm, dout, y, w, p, qh, nqh, data, tw, sms);
    run<2>("V2 Montgomery fast (C)", 3, true, dy, dw, dws, dwm, dout, y, w, p, qh, nqh, data, tw, sms);
    run<3>("V3 Shoup generic, hi-2 (C)", 4, false, dy, dw, dws, dwm, dout, y, w, p, qh, nqh, data, tw, sms);
    run<4>("V4 Shoup generic, hi-1 (PTX)", 3, false, dy, dw, dws, dwm, dout, y, w, p, qh, nqh, data, tw, sms);
    run<5>("V5 Shoup fast prime, hi-2 (C)", 4, false, dy, dw, dws, dwm, dout, y, w, p, qh, nqh, data, tw, sms);
    run<6>("V6 Shoup fast prime, hi-1 (PTX a)", 3, false, dy, dw, dws, dwm, dout, y, w, p, qh, nqh, data, tw, sms);
    run<7>("V7 Shoup fast prime, hi-1 (PTX b)", 3, false, dy, dw, dws, dwm, dout, y, w, p, qh, nqh, data, tw, sms);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
