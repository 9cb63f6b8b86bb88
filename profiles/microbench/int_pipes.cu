// int_pipes.cu — integer-pipe throughput microbenchmark for sm_100a (B200).
// Measures warp-instruction issue rates that bound the 64-bit modular butterflies:
// IMAD.WIDE.U32, IMAD (lo), IMAD.HI, 64-bit add (IADD3 + IADD3.X), __umul64hi, and the
// library's own shoup_lazy / ct_bfly.  Output: results per SM per clock.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../deeppowers_b200/csrc int_pipes.cu -o int_pipes
#include <cstdio>
#include <cuda_runtime.h>
#include "ntt_core.cuh"
using namespace dpfhe;

constexpr int ITERS = 4096, CH = 8;

template <int OP>
__global__ void __launch_bounds__(1024) bench(unsigned long long *out, long long *cycles, unsigned a0, unsigned b0, LimbParams p) {
    unsigned a = a0 + threadIdx.x, b = b0 ^ threadIdx.x;
    u64 acc[CH];
    unsigned r32[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) { acc[c] = threadIdx.x * 977ull + c; r32[c] = threadIdx.x + c; }
    Twiddle w; w.x = p.ninv; w.y = p.ninv_s;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < ITERS; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (OP == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
            if (OP == 1) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r32[c]) : "r"(a), "r"(b));
            if (OP == 2) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r32[c]) : "r"(a), "r"(b));
            if (OP == 3) asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[c]) : "l"((unsigned long long)a << 20 | b));
            if (OP == 4) acc[c] = __umul64hi(acc[c], p.bar_mu) + c;
            if (OP == 5) acc[c] = shoup_lazy(acc[c], w.x, w.y, p);
            if (OP == 6) acc[c] = csub(acc[c] + p.q, p.q2);
            if (OP == 7) acc[c] = word_reduce(acc[c] * 5 + 1, p);
            if (OP == 8) acc[c] = mulmod_lazy(acc[c], p.wninv + c, p);
        }
        if (OP == 11) {   // pure 32-bit ALU adds (LOP3-free): 8 independent chains
#pragma unroll
            for (int c = 0; c < CH; ++c) asm volatile("add.u32 %0, %0, %1;" : "+r"(r32[c]) : "r"(a));
        }
        if (OP == 12) {   // mixed: 4 IMAD chains + 4 IADD chains per iteration, all independent
#pragma unroll
            for (int c = 0; c < CH / 2; ++c) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r32[c]) : "r"(a), "r"(b));
#pragma unroll
            for (int c = CH / 2; c < CH; ++c) asm volatile("add.u32 %0, %0, %1;" : "+r"(r32[c]) : "r"(a));
        }
        if (OP == 13) {   // mixed: 4 IMAD.WIDE chains + 4 x 64-bit add chains
#pragma unroll
            for (int c = 0; c < CH / 2; ++c) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
#pragma unroll
            for (int c = CH / 2; c < CH; ++c) asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[c]) : "l"((unsigned long long)a << 20 | b));
        }
        if (OP == 14) {   // LOP3 only
#pragma unroll
            for (int c = 0; c < CH; ++c) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r32[c]) : "r"(a), "r"(b));
        }
        if (OP == 15) {   // FFMA only (fp32 pipe) for comparison
#pragma unroll
            for (int c = 0; c < CH; ++c) { float f = __uint_as_float(r32[c]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); r32[c] = __float_as_uint(f); }
        }
        if (OP == 16) {   // mixed IMAD + FFMA
#pragma unroll
            for (int c = 0; c < CH / 2; ++c) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r32[c]) : "r"(a), "r"(b));
#pragma unroll
            for (int c = CH / 2; c < CH; ++c) { float f = __uint_as_float(r32[c]); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f)); r32[c] = __float_as_uint(f); }
        }
        if (OP == 17) {   // IMAD with a uniform (kernel-parameter) multiplier: 2 vector-register reads
#pragma unroll
            for (int c = 0; c < CH; ++c) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r32[c]) : "r"(a0), "r"(b));
        }
        if (OP == 18) {   // IMAD.WIDE reg x uniform + 0-ish: result does not feed itself as addend
#pragma unroll
            for (int c = 0; c < CH; ++c) { unsigned long long t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(r32[c]), "r"(a0)); r32[c] = (unsigned)(t >> 32) ^ (unsigned)t; }
        }
        if (OP == 19) {   // IMAD.WIDE reg x reg, no addend
#pragma unroll
            for (int c = 0; c < CH; ++c) { unsigned long long t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(r32[c]), "r"(b)); r32[c] = (unsigned)(t >> 32) ^ (unsigned)t; }
        }
        if (OP == 20) {   // shoup_lazy with a uniform twiddle (kernel parameter)
#pragma unroll
            for (int c = 0; c < CH; ++c) acc[c] = shoup_lazy(acc[c], p.ninv, p.ninv_s, p);
        }
        if (OP == 21) {   // IMAD.WIDE without addend (RZ), result folded with one IADD3
#pragma unroll
            for (int c = 0; c < CH; ++c) { unsigned lo, hi; asm volatile("{.reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0,%1}, t;}" : "=r"(lo), "=r"(hi) : "r"(r32[c]), "r"(b)); r32[c] = lo + hi; }
        }
        if (OP == 22) {   // IMAD.WIDE with a live 64-bit addend, result folded with one IADD3
#pragma unroll
            for (int c = 0; c < CH; ++c) { unsigned lo, hi; asm volatile("{.reg .u64 t; mad.wide.u32 t, %2, %3, %4; mov.b64 {%0,%1}, t;}" : "=r"(lo), "=r"(hi) : "r"(r32[c]), "r"(b), "l"(acc[c])); r32[c] = lo + hi; }
        }
        if (OP == 23) {   // two IADD3 only (reference for the fold cost)
#pragma unroll
            for (int c = 0; c < CH; ++c) { r32[c] = r32[c] + b; asm volatile("" : "+r"(r32[c])); r32[c] = r32[c] + a; asm volatile("" : "+r"(r32[c])); }
        }
        if (OP == 27) {   // IMAD.WIDE with 64-bit addend, multiplicand taken from the chain itself (OP 0's product is loop-invariant and hoisted)
#pragma unroll
            for (int c = 0; c < CH; ++c) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mad.wide.u32 %0, l, %1, %0;}" : "+l"(acc[c]) : "r"(b));
        }
        if (OP == 28) {   // IMAD.WIDE without addend, same dependence
#pragma unroll
            for (int c = 0; c < CH; ++c) asm volatile("{.reg .u32 l, h; mov.b64 {l, h}, %0; mul.wide.u32 %0, l, %1;}" : "+l"(acc[c]) : "r"(b));
        }
        if (OP == 29) {   // IMAD (32-bit) with the multiplicand from the chain
#pragma unroll
            for (int c = 0; c < CH; ++c) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r32[c]) : "r"(b), "r"(a));
        }
        if (OP == 24) {   // DFMA only (fp64 pipe): could the idle fp64 lanes carry part of the modular arithmetic?
#pragma unroll
            for (int c = 0; c < CH; ++c) { double f = __longlong_as_double((long long)acc[c]); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f) : "d"(1.0000001), "d"(0.5)); acc[c] = (u64)__double_as_longlong(f); }
        }
        if (OP == 25) {   // mixed: 4 IMAD.WIDE chains + 4 DFMA chains
#pragma unroll
            for (int c = 0; c < CH / 2; ++c) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[c]) : "r"(a), "r"(b));
#pragma unroll
            for (int c = CH / 2; c < CH; ++c) { double f = __longlong_as_double((long long)acc[c]); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f) : "d"(1.0000001), "d"(0.5)); acc[c] = (u64)__double_as_longlong(f); }
        }
        if (OP == 26) {   // mixed: 4 IMAD chains + 4 DFMA chains
#pragma unroll
            for (int c = 0; c < CH / 2; ++c) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r32[c]) : "r"(a), "r"(b));
#pragma unroll
            for (int c = CH / 2; c < CH; ++c) { double f = __longlong_as_double((long long)acc[c]); asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f) : "d"(1.0000001), "d"(0.5)); acc[c] = (u64)__double_as_longlong(f); }
        }
        if (OP == 9) {
#pragma unroll
            for (int c = 0; c < CH; c += 2) ct_bfly(acc[c], acc[c + 1], w, p);
#pragma unroll
            for (int c = 0; c < CH; c += 1) acc[c] &= 0x0fffffffffffffffull;   // keep the lazy bound (1 LOP3 per element)
        }
        if (OP == 10) {
#pragma unroll
            for (int c = 0; c < CH; c += 2) gs_bfly(acc[c], acc[c + 1], w, p);
        }
    }
    long long t1 = clock64();
    u64 s = 0;
#pragma unroll
    for (int c = 0; c < CH; ++c) s += acc[c] + r32[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, double per_iter_results, LimbParams p, int threads) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(unsigned long long) * sms * 1024);
    cudaMalloc(&cyc, sizeof(long long) * sms);
    bench<OP><<<sms, threads>>>(out, cyc, 12345u, 67891u, p);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OP><<<sms, threads>>>(out, cyc, 12345u, 67891u, p);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    long long h[256];
    cudaMemcpy(h, cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; ++i) avg += h[i];
    avg /= sms;
    double results = (double)threads * ITERS * per_iter_results;
    printf("%-34s threads/SM=%4d  %8.2f results/clk/SM   %7.3f clk per warp-result/SMSP   (%.3f ms, %.0f MHz eff)\n", name, threads,
           results / avg, avg / (results / 32.0 / 4.0), ms, avg / ms / 1e3);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    LimbParams p;
    p.q = 0xfffffffffffc001ull; p.q2 = 2 * p.q; p.q4 = 4 * p.q; p.q8 = 8 * p.q; p.nq = 0 - p.q;
    p.bar_shift = 58; p.bar_mu = (unsigned long long)(((unsigned __int128)1 << 122) / p.q);
    p.mu32 = (unsigned)(((unsigned __int128)1 << 64) / p.q);
    p.ninv = 0x123456789abcdefull % p.q; p.ninv_s = (unsigned long long)(((unsigned __int128)p.ninv << 64) / p.q);
    p.wninv = 0xfedcba987654321ull % p.q; p.wninv_s = 0;
    for (int threads : {1024}) {
        run<0>("mad.wide.u32 (IMAD.WIDE)", CH, p, threads);
        run<1>("mad.lo.u32 (IMAD)", CH, p, threads);
        run<2>("mad.hi.u32 (IMAD.HI)", CH, p, threads);
        run<3>("add.u64 (IADD3+IADD3.X)", CH, p, threads);
        run<4>("__umul64hi + add", CH, p, threads);
        run<5>("shoup_lazy", CH, p, threads);
        run<6>("csub(x+q, 2q)", CH, p, threads);
        run<7>("word_reduce(5x+1)", CH, p, threads);
        run<8>("mulmod_lazy (128b product+Barrett)", CH, p, threads);
        run<9>("ct_bfly (+1 LOP3/elt)", CH / 2, p, threads);
        run<10>("gs_bfly", CH / 2, p, threads);
        run<17>("IMAD reg*uniform+reg", CH, p, threads);
        run<18>("IMAD.WIDE reg*uniform (+LOP)", CH, p, threads);
        run<19>("IMAD.WIDE reg*reg (+LOP)", CH, p, threads);
        run<20>("shoup_lazy, uniform twiddle", CH, p, threads);
        run<21>("IMAD.WIDE no addend + IADD3", CH, p, threads);
        run<22>("IMAD.WIDE 64-bit addend + IADD3", CH, p, threads);
        run<23>("2 x IADD3", CH, p, threads);
        run<11>("add.u32 (IADD3) x8", CH, p, threads);
        run<12>("4 IMAD + 4 IADD3 mixed", CH, p, threads);
        run<13>("4 IMAD.WIDE + 4 add.u64 mixed", CH, p, threads);
        run<14>("lop3 x8", CH, p, threads);
        run<15>("FFMA x8", CH, p, threads);
        run<16>("4 IMAD + 4 FFMA mixed", CH, p, threads);
        run<27>("IMAD.WIDE + addend (chain-fed)", CH, p, threads);
        run<28>("IMAD.WIDE no addend (chain-fed)", CH, p, threads);
        run<29>("IMAD lo (chain-fed)", CH, p, threads);
        run<24>("DFMA x8", CH, p, threads);
        run<25>("4 IMAD.WIDE + 4 DFMA mixed", CH, p, threads);
        run<26>("4 IMAD + 4 DFMA mixed", CH, p, threads);
        printf("\n");
    }
    return 0;
}
