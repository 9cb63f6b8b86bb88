// kernels.cu — sm_100a kernels and their launchers.
//
// No tensor cores (pure 64-bit modular integer work), no Triton, no CPU fallback.
//   ntt_kernel            one CTA per limb transform, limb resident in swizzled shared memory
//   ks_fused_kernel       persistent cooperative kernel: tensor / permute -> INTT -> publish digit
//                         -> (L-1) x [lift + NTT + multiply-accumulate with the switch key] -> out
//   pointwise kernels     128-bit vectorised grid-stride loops
#include <cooperative_groups.h>
#include <cuda.h>   // CUtensorMap (types only: the driver entry point is fetched at run time, libcuda is not linked)
#include <cuda_runtime.h>

#include <atomic>

#include "kernel_bodies.cuh"
#include "launch.hpp"

// The file is large (every kernel x three ring degrees x three modes); the build compiles it in two parts per variant, in parallel:
// -DDPFHE_PART=1 everything but the special-prime key-switch family, 2 its one-special-prime kernel, 3 the grouped kernels; no flag = all.
#ifndef DPFHE_PART
#define DPFHE_PART 0
#endif
#define DPFHE_PART_MAIN (DPFHE_PART == 0 || DPFHE_PART == 1)
#define DPFHE_PART_HYBRID (DPFHE_PART == 0 || DPFHE_PART == 2)
#define DPFHE_PART_GROUPED (DPFHE_PART == 0 || DPFHE_PART == 3)

// compiled twice: -DDPFHE_FAST=0 -> namespace dpfhe::gen, -DDPFHE_FAST=1 -> namespace dpfhe::fast (types.hpp)
namespace dpfhe {
namespace DPFHE_VNS {

// barrier scopes of the CTA policy (ntt_core.cuh): CTA, 256-thread domain, warp.
// PROF: thread 0 accumulates clock64() deltas per phase id into prof[blockIdx][id] (diagnostics only).
template <int NT, bool PROF = false>
struct DevCta {
    unsigned long long *prof = nullptr;
    long long last = 0;
    __device__ __forceinline__ void mark(int id) {
        if (PROF && threadIdx.x == 0) {
            const long long now = clock64();
            prof[id] += (unsigned long long)(now - last);
            last = now;
        }
    }
    template <class F>
    __device__ __forceinline__ void par(F f) {
        f((int)threadIdx.x);
        __syncthreads();
    }
    // blocks the CTA until the monotone counter *flag has reached `target` (wrap-safe comparison)
    __device__ __forceinline__ void wait_ge(const u32 *flag, u32 target) {
        if (threadIdx.x == 0) {
            u32 v;
            do asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            while ((int)(v - target) < 0);
        }
        __syncthreads();
    }
    template <class F>
    __device__ __forceinline__ void par_dom(F f) {
        f((int)threadIdx.x);
        if (NT <= 256) __syncthreads();
        else asm volatile("bar.sync %0, 256;" ::"r"(1 + ((int)threadIdx.x >> 8)) : "memory");
    }
    template <class F>
    __device__ __forceinline__ void par_warp(F f) {
        f((int)threadIdx.x);
        __syncwarp();
    }
};

#if DPFHE_PART_MAIN
// ------------------------------------------------------------------ standalone transforms
// The per-limb constants travel in the kernel parameter block (constant bank), so q, 2q, 8q ...
// are read through uniform registers / constant operands instead of occupying vector registers.
template <int LOGN, int NT, int MINB, bool INVERSE>
__global__ void __launch_bounds__(NT, MINB) ntt_kernel(u64 *data, const Twiddle *__restrict__ tables,
                                                        const __grid_constant__ LimbTable lt, u32 L, size_t n_limbs) {
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

// Inverse transform whose input copy is done by the TMA unit: the swizzled shared-memory layout of ntt_core.cuh IS the layout a
// 2-D tensor map with CU_TENSOR_MAP_SWIZZLE_128B produces (rows of 128 bytes = 16 coefficients, 16-byte chunk index XOR row & 7),
// so one elected thread issues cp.async.bulk.tensor for the whole limb (boxes of 256 rows = 32 KiB) and everybody waits on the
// mbarrier, instead of 256 threads looping over LDG.128 + STS.128.  One limb per CTA (grid = n_limbs): the barrier is used once.
__device__ __forceinline__ u32 smem_addr(const void *p) { return (u32)__cvta_generic_to_shared(p); }

template <int LOGN, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ntt_inv_tma_kernel(const __grid_constant__ CUtensorMap tm, u64 *data, const Twiddle *__restrict__ itw,
                                                                const __grid_constant__ LimbTable lt, u32 L) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    constexpr size_t N = (size_t)1 << LOGN;
    constexpr int ROWS = (int)(N * 8 / 128), BOX_ROWS = ROWS < 256 ? ROWS : 256, BOXES = ROWS / BOX_ROWS;
    DevCta<NT> cta;
    const size_t w = blockIdx.x;
    const u32 l = (u32)(w % L);
    const u32 bar_a = smem_addr(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((u32)(N * 8)) : "memory");
#pragma unroll
        for (int b = 0; b < BOXES; ++b) {
            const int row = (int)(w * ROWS) + b * BOX_ROWS;
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                         ::"r"(smem_addr(smem_raw + (size_t)b * BOX_ROWS * 128)), "l"(reinterpret_cast<unsigned long long>(&tm)), "r"(0), "r"(row), "r"(bar_a)
                         : "memory");
        }
    }
    {   // every thread waits for the bytes to land (phase 0 of the barrier)
        u32 done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done)
                         : "r"(bar_a)
                         : "memory");
        }
    }
    ntt_inv_resident<LOGN, NT>(cta, buf, data + w * N, itw + (size_t)l * N, lt.lp[l]);
}

// N = 16384: one limb per CLUSTER of two CTAs (64 KiB of shared memory each, so three CTAs still share an SM); see
// kernel_bodies.cuh "transforms by a PAIR of CTAs".  The inverse reads the partner's half through distributed shared memory.
template <int NT, int MINB, bool INVERSE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NT, MINB)
    ntt_pair_kernel(u64 *data, const Twiddle *__restrict__ tables, const __grid_constant__ LimbTable lt, u32 L, size_t n_limbs) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    constexpr size_t N = (size_t)1 << NTT_PAIR_LOGN;
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int h = (int)cluster.block_rank();
    const u64 *peer = cluster.map_shared_rank(buf, h ^ 1);
    DevCta<NT> cta;
    for (size_t w = blockIdx.x / 2; w < n_limbs; w += gridDim.x / 2) {
        const u32 l = (u32)(w % L);
        const LimbParams &p = lt.lp[l];
        const Twiddle *tw = tables + (size_t)l * N;
        if (!INVERSE) {
            ntt_fwd_half_load<NT>(cta, buf, data + w * N, tw, p, h);
            cluster.sync();   // both CTAs have read the whole limb: the in-place stores may begin
            ntt_fwd_half_finish<NT>(cta, buf, data + w * N, tw, p, h);
        } else {
            ntt_inv_half_passes<NT>(cta, buf, data + w * N, tw, p, h);
            cluster.sync();   // the partner's half is complete in its shared memory
            ntt_inv_half_outer<NT>(cta, buf, peer, data + w * N, tw, p, h);
            cluster.sync();   // the partner has finished reading this CTA's shared memory
        }
    }
}

// ------------------------------------------------------------------ modulus switching (two launches)
template <int LOGN, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ms_tau_kernel(const u64 *in, u64 *tau, const Twiddle *__restrict__ itw,
                                                           const __grid_constant__ LimbTable lt, const __grid_constant__ MsConsts K,
                                                           u32 L, size_t n_polys) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    constexpr size_t N = (size_t)1 << LOGN;
    DevCta<NT> cta;
    const LimbParams &p = lt.lp[L - 1];
    for (size_t w = blockIdx.x; w < n_polys; w += gridDim.x)
        ms_tau_body<LOGN, NT>(cta, buf, in + (w * L + (L - 1)) * N, nullptr, itw + (size_t)(L - 1) * N, p, tau + w * N, K);
}

template <int LOGN, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ms_limb_kernel(const u64 *in, const u64 *tau, u64 *out, const Twiddle *__restrict__ tw,
                                                            const __grid_constant__ LimbTable lt, const __grid_constant__ MsConsts K,
                                                            u32 L, size_t n_items) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    constexpr size_t N = (size_t)1 << LOGN;
    DevCta<NT> cta;
    const u32 Lo = L - 1;
    for (size_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const size_t poly = w / Lo;
        const u32 i = (u32)(w % Lo);
        ms_limb_body<LOGN, NT>(cta, buf, tau + poly * N, in + (poly * L + i) * N, out + (poly * Lo + i) * N, tw + (size_t)i * N, lt.lp[i], K, i);
    }
}

// division by the product of the last K limbs: tau' of every special limb, then one item per kept limb
template <int LOGN, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) md_tau_kernel(const u64 *in, u64 *tau, const Twiddle *__restrict__ itw, const __grid_constant__ MsConsts K,
                                                           const __grid_constant__ GroupConsts G, size_t n_items) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    constexpr size_t N = (size_t)1 << LOGN;
    DevCta<NT> cta;
    const u32 L = G.Lq + G.K;
    for (size_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const size_t poly = w / G.K;
        const u32 s = G.Lq + (u32)(w % G.K);
        ms_tau_body<LOGN, NT>(cta, buf, in + (poly * L + s) * N, nullptr, itw + (size_t)s * N, G.lp_up[s], tau + w * N, K);
    }
}

template <int LOGN, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) md_limb_kernel(const u64 *in, const u64 *tau, u64 *out, const Twiddle *__restrict__ tw,
                                                            const __grid_constant__ LimbTable lt, const __grid_constant__ MsConsts K,
                                                            const __grid_constant__ GroupConsts G, size_t n_items) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    constexpr size_t N = (size_t)1 << LOGN;
    DevCta<NT> cta;
    const u32 Lq = G.Lq, L = Lq + G.K;
    for (size_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const size_t poly = w / Lq;
        const u32 i = (u32)(w % Lq);
        ms_limb_group<LOGN, NT, false>(cta, buf, tau + poly * G.K * N, N, in + (poly * L + i) * N, out + (poly * Lq + i) * N, tw + (size_t)i * N,
                                       lt.lp[i], K, G, i);
    }
}

#endif
// ------------------------------------------------------------------ fused key-switch family
__device__ __forceinline__ u32 ld_acquire_u32(const u32 *p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(u32 *p, u32 v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u64 ld_acquire_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(u64 *p, u64 v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// TMA-unit bulk prefetch of a contiguous region into L2 (SASS: UBLKPF.L2); bytes must be a multiple of 16
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

#if DPFHE_PART_MAIN
// grid = G CTAs, G a multiple of L, all co-resident (cooperative launch).  The L CTAs of slots
// [g*L, (g+1)*L) form a group that processes one ciphertext at a time: CTA `slot` owns output limb
// i = slot % L.  The group leader (i == 0) draws the next ciphertext index from a global ticket counter
// and posts it in the group's mailbox (dynamic balancing: groups that run ahead take more work);
// the members exchange their INTT'd digits through `scratch` (double-buffered by round parity,
// L2 resident) under release/acquire flags.
// FILTER (hoisted-rotation fallback): only the ciphertexts flagged in A.only are processed; the digit slots then
// alternate over the rounds that actually run.
template <int LOGN, int NT, int MINB, int MODE, bool PROF, bool FILTER = false>
__global__ void __launch_bounds__(NT, MINB) ks_fused_kernel(KsArgs A, const __grid_constant__ LimbTable lt, size_t batch, u32 *flags, u32 epoch,
                                                            u32 *ticket, u64 *mail, unsigned long long *prof, u32 pf_dist, u32 *consumed) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr size_t N = (size_t)1 << LOGN;
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    DevCta<NT, PROF> cta;
    unsigned long long t_start = 0, c_start = 0;
    if (PROF) {
        cta.prof = prof + (size_t)blockIdx.x * 16;
        cta.last = clock64();
        c_start = (unsigned long long)cta.last;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_start));
    }
    __shared__ u32 s_ct;
    const u32 L = A.L, slot = blockIdx.x, i = slot % L, group = slot / L;
    const LimbParams &p = lt.lp[i];
    u32 executed = 0;
    // Single-buffered digit slots (consumed != nullptr): slot s holds ONE digit; its owner may overwrite it only after the L-1
    // siblings that read the previous digit have signalled.  The counters are monotone across launches: the value found at
    // kernel start is the base (no reader of an earlier launch is still running).  One slot per CTA instead of two keeps
    // 28 MB (N = 8192, 444 CTAs) of the kernel's cross-phase working set out of the L2.
    __shared__ u32 s_base;
    if (consumed && threadIdx.x == 0) s_base = ld_acquire_u32(consumed + slot);
    __syncthreads();
    const u32 consumed_base = consumed ? s_base : 0u;
    u32 published = 0;   // digits this CTA has published in this launch
    for (u32 round = 0;; ++round) {
        if (!FILTER && threadIdx.x == 0) {
            const u32 tag = epoch + round + 1;
            if (i == 0) {
                const u32 t = atomicAdd(ticket, 1u);
                if (L > 1) st_release_u64(mail + group, ((u64)tag << 32) | t);
                s_ct = t;
            } else {
                u64 m;
                do m = ld_acquire_u64(mail + group);
                while ((u32)(m >> 32) != tag);
                s_ct = (u32)m;
            }
        }
        __syncthreads();
        // FILTER: static assignment computed by every thread.  Skipped rounds involve no exchange between the members of
        // a group, so a leader handing out tickets could run ahead and overwrite a mailbox tag (or s_ct) before it was read.
        const size_t ct = FILTER ? (size_t)round * (gridDim.x / L) + group : (size_t)s_ct;
        if (ct >= batch) break;   // every member of the group reads the same ticket, so they leave together
        if (FILTER && A.only[ct] == 0u) continue;
        // Tickets are drawn in order, so ciphertext ct + pf_dist will be started by some group a few microseconds
        // from now: pull this CTA's limb of its inputs from HBM into L2 with the TMA unit's bulk prefetch, so the
        // tensor phase that consumes them is L2- rather than HBM-latency bound.
        if (pf_dist && threadIdx.x == 0 && ct + pf_dist < batch) {
            const size_t nc = ct + pf_dist, P = (size_t)L * N;
            constexpr u32 LB = (u32)(N * 8);
            if (MODE == KS_PLAIN) {
                bulk_prefetch_l2(A.a + nc * P + (size_t)i * N, LB);
            } else {
                bulk_prefetch_l2(A.a + nc * 2 * P + (size_t)i * N, LB);
                bulk_prefetch_l2(A.a + nc * 2 * P + P + (size_t)i * N, LB);
                if (MODE == KS_MUL_RELIN) {
                    bulk_prefetch_l2(A.b + nc * 2 * P + (size_t)i * N, LB);
                    bulk_prefetch_l2(A.b + nc * 2 * P + P + (size_t)i * N, LB);
                }
            }
        }
        const u32 parity = consumed ? 0u : (FILTER ? executed++ : round) & 1u;
        const size_t slot_stride = consumed ? 1 : 2;      // digit slots per CTA
        u64 *acc_rows = A.acc + (size_t)slot * 2 * N;   // this CTA's two lazy accumulator rows (L2 resident, reused every round)
        ks_phase1<LOGN, NT, MODE>(cta, buf, A, p, ct, i, A.scratch + ((size_t)slot * slot_stride + parity) * N, acc_rows, 0, 0,
                                  consumed && L > 1 ? consumed + slot : nullptr, consumed_base + published * (L - 1));
        ++published;
        if (L > 1) {
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) st_release_u32(flags + slot, epoch + round + 1);
            for (u32 jj = 1; jj < L; ++jj) {
                const u32 j = (i + jj) % L, sib = slot - i + j;
                if (threadIdx.x == 0) {
                    while ((int)(ld_acquire_u32(flags + sib) - (epoch + round + 1)) < 0) {
                    }
                }
                __syncthreads();
                cta.mark(3);   // waiting for the sibling's digit
                ks_phase2_digit<LOGN, NT>(cta, buf, A, p, ct, i, j, jj, A.scratch + ((size_t)sib * slot_stride + parity) * N, acc_rows);
                // every thread is past its last read of the sibling's digit (the body ends with a CTA barrier): hand the slot back
                if (consumed && threadIdx.x == 0) {
                    __threadfence();
                    atomicAdd(consumed + sib, 1u);
                }
            }
        }
    }
    if (PROF && threadIdx.x == 0) {   // CTA lifetime in nanoseconds (globaltimer) and in SM cycles
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        cta.prof[14] += t_end - t_start;
        cta.prof[15] += (unsigned long long)clock64() - c_start;
    }
}

// Hoisted rotations, step 1: the same group / ticket / flag machinery as ks_fused_kernel, but the digits
// are the unpermuted c1 limbs and phase 2 stores the lifted transforms U[ct][j][i] instead of multiplying them with a key.
template <int LOGN, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ks_hoist_kernel(HoistArgs A, const __grid_constant__ LimbTable lt, size_t batch, u32 *flags, u32 epoch,
                                                            u32 *ticket, u64 *mail) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr size_t N = (size_t)1 << LOGN;
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    DevCta<NT> cta;
    __shared__ u32 s_ct;
    const u32 L = A.L, slot = blockIdx.x, i = slot % L, group = slot / L;
    const LimbParams &p = lt.lp[i];
    for (u32 round = 0;; ++round) {
        const u32 tag = epoch + round + 1;
        if (threadIdx.x == 0) {
            if (i == 0) {
                const u32 t = atomicAdd(ticket, 1u);
                st_release_u64(mail + group, ((u64)tag << 32) | t);
                s_ct = t;
            } else {
                u64 m;
                do m = ld_acquire_u64(mail + group);
                while ((u32)(m >> 32) != tag);
                s_ct = (u32)m;
            }
        }
        __syncthreads();
        const size_t ct = s_ct;
        if (ct >= batch) break;
        const u32 parity = round & 1u;
        hoist_phase1<LOGN, NT>(cta, buf, A, p, ct, i, A.scratch + ((size_t)slot * 2 + parity) * N);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_u32(flags + slot, tag);
        for (u32 jj = 1; jj < L; ++jj) {
            const u32 j = (i + jj) % L, sib = slot - i + j;
            if (threadIdx.x == 0) {
                while ((int)(ld_acquire_u32(flags + sib) - tag) < 0) {
                }
            }
            __syncthreads();
            hoist_phase2<LOGN, NT>(cta, buf, A, p, ct, i, j, A.scratch + ((size_t)sib * 2 + parity) * N);
        }
    }
}

// step 2: one rotation applied to blocks of ROT_CB ciphertexts x one limb; no transforms, only gathers and multiply-accumulates
template <int LOGN, int NT, int MINB, int ROT_CB, bool PF>
__global__ void __launch_bounds__(NT, MINB) rot_apply_kernel(RotApplyArgs A, const __grid_constant__ LimbTable lt, size_t batch, u32 nseg) {
    DevCta<NT> cta;
    constexpr int NC = 1 << (LOGN - 1);
    const size_t n_blocks = (batch + ROT_CB - 1) / ROT_CB, n_items = n_blocks * A.L * nseg;
    const int seg_chunks = NC / (int)nseg;   // nseg is a power of two <= NC / NT
    for (size_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const u32 seg = (u32)(w % nseg), i = (u32)((w / nseg) % A.L);
        const size_t ct0 = (w / nseg / A.L) * ROT_CB;
        const u32 n_ct = (u32)(batch - ct0 < (size_t)ROT_CB ? batch - ct0 : (size_t)ROT_CB);
        rot_apply_rows<LOGN, NT, ROT_CB, PF>(cta, A, lt.lp[i], ct0, n_ct, i, (int)seg * seg_chunks, ((int)seg + 1) * seg_chunks);
    }
}

// coefficient-form indicator of the positions that sigma_g negates: X^k -> X^(kg mod 2N), negative when kg mod 2N >= N
template <int LOGN>
__global__ void __launch_bounds__(256) negmask_kernel(u64 *mask, u32 g, u32 L) {
    constexpr u32 N = 1u << LOGN;
    for (u32 k = blockIdx.x * blockDim.x + threadIdx.x; k < N; k += gridDim.x * blockDim.x) {
        const u32 e = (k * g) & (2 * N - 1);
        const u64 v = e >= N ? 1ull : 0ull;
        for (u32 l = 0; l < L; ++l) mask[(size_t)l * N + (e & (N - 1))] = v;
    }
}

// kprime[c][i][n] = M[i][n] * sum_{j != i} (q_j mod q_i) * key[j][c][i][n]  mod q_i, canonical
template <int LOGN>
__global__ void __launch_bounds__(256) kprime_kernel(const u64 *__restrict__ key, const u64 *__restrict__ M, const u64 *__restrict__ delta,
                                                     u64 *__restrict__ out, const LimbParams *__restrict__ lps, u32 L) {
    constexpr size_t N = (size_t)1 << LOGN;
    const size_t P = (size_t)L * N, total = 2 * P;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const u32 c = (u32)(e / P), i = (u32)((e % P) / N);
        const size_t n = e % N;
        const LimbParams p = lps[i];
        u64 s = 0;
        for (u32 j = 0; j < L; ++j) {
            if (j == i) continue;
            s = csub(s + mulmod(key[((size_t)j * 2 + c) * P + (size_t)i * N + n], delta[j * L + i], p), p.q);
        }
        out[e] = mulmod(s, M[(size_t)i * N + n], p);
    }
}

#endif
#if DPFHE_PART_HYBRID
// Hybrid (special-prime) variant, DESIGN.md §2.10.  A group is L + 1 CTAs: CTA i < L owns ciphertext limb i,
// CTA L owns the special limb.  Per ciphertext:
//   limb CTA    tensor/permute, p*own terms + first key term, INTT, publish digit        (as above)
//               L-1 x [lift + NTT + MAC]                                                  (as above, nothing final)
//               2 x [centred lift of tau' + NTT], out = (acc - s*u) / p                   (ms_limb_body)
//   special CTA L x [lift + NTT_p + MAC into its scratch rows]; 2 x INTT_p (* t^-1) -> tau', publish
// Both roles run six transforms per ciphertext at L = 4.  The special CTA can only finish after every digit arrived,
// so a limb CTA postpones the division step of ciphertext r until it has done the digit and multiply-accumulate
// work of ciphertext r + 1 (software pipeline of depth one): tau' rows, like the digits and the mailbox, are
// double-buffered by round parity, and the output rows keep the lazy accumulators in between.
template <int LOGN, int NT, int MINB, int MODE>
__global__ void __launch_bounds__(NT, MINB) ks_hybrid_kernel(KsArgs A, const __grid_constant__ LimbTable lt, const __grid_constant__ MsConsts K,
                                                             size_t batch, u32 *flags, u32 epoch, u32 *ticket, u64 *mail) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr size_t N = (size_t)1 << LOGN;
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    DevCta<NT> cta;
    __shared__ u32 s_ct;
    const u32 L = A.L, GS = L + 1, slot = blockIdx.x, i = slot % GS, group = slot / GS, base = slot - i;
    const bool special = i == L;
    const LimbParams &p = lt.lp[i];
    u64 *hyb = A.hyb + (size_t)group * KS_HYB_ROWS * N;
    auto wait_for = [&](u32 sib, u32 tag) {
        if (threadIdx.x == 0) {
            while ((int)(ld_acquire_u32(flags + sib) - tag) < 0) {
            }
        }
        __syncthreads();
    };
    auto publish = [&](u32 tag) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_u32(flags + slot, tag);
    };
    // the postponed division of one ciphertext (limb CTAs): needs tau' of that round
    auto acc_of = [&](u32 parity) { return A.acc + ((size_t)slot * 2 + parity) * 2 * N; };   // accumulator rows, double-buffered by round parity
    auto divide = [&](size_t ct, u32 tag, u32 parity) {
        wait_for(base + L, tag);
        const size_t P = (size_t)L * N;
        for (u32 c = 0; c < 2; ++c) {
            u64 *row = A.out + ct * 2 * P + c * P + (size_t)i * N;   // lazy accumulator -> final value in the output row
            ms_limb_body<LOGN, NT, true>(cta, buf, hyb + ks_hyb_tau_row(parity, c) * N, acc_of(parity) + c * N, row, A.tw + (size_t)i * N, p, K, i);
        }
    };
    bool pending = false;
    size_t prev_ct = 0;
    u32 prev_tag = 0, prev_parity = 0;
    for (u32 round = 0;; ++round) {
        const u32 tag = epoch + round + 1, parity = round & 1u;
        if (threadIdx.x == 0) {
            u64 *box = mail + (size_t)group * 2 + parity;
            if (i == 0) {
                const u32 t = atomicAdd(ticket, 1u);
                st_release_u64(box, ((u64)tag << 32) | t);
                s_ct = t;
            } else {
                u64 m;
                do m = ld_acquire_u64(box);
                while ((u32)(m >> 32) != tag);
                s_ct = (u32)m;
            }
        }
        __syncthreads();
        const size_t ct = s_ct;
        if (ct >= batch) break;
        if (!special) {
            ks_phase1<LOGN, NT, MODE, true>(cta, buf, A, p, ct, i, A.scratch + ((size_t)slot * 2 + parity) * N, acc_of(parity), K.qlm[i], K.qlm_s[i]);
            publish(tag);
            for (u32 jj = 1; jj < L; ++jj) {
                const u32 j = (i + jj) % L;
                wait_for(base + j, tag);
                ks_phase2_digit<LOGN, NT, true, false>(cta, buf, A, p, ct, i, j, jj, A.scratch + ((size_t)(base + j) * 2 + parity) * N, acc_of(parity));
            }
            if (pending) divide(prev_ct, prev_tag, prev_parity);
            pending = true;
            prev_ct = ct;
            prev_tag = tag;
            prev_parity = parity;
        } else {
            for (u32 jj = 0; jj < L; ++jj) {
                const u32 j = (group + jj) % L;   // groups start at different digits: spreads the key-column reads
                wait_for(base + j, tag);
                ks_phase2_digit<LOGN, NT, true, true>(cta, buf, A, p, ct, i, j, jj, A.scratch + ((size_t)(base + j) * 2 + parity) * N, hyb);
            }
            for (u32 c = 0; c < 2; ++c)
                ms_tau_body<LOGN, NT, true>(cta, buf, hyb + c * N, hyb + c * N, A.itw + (size_t)i * N, p, hyb + ks_hyb_tau_row(parity, c) * N, K);
            publish(tag);
        }
    }
    if (pending) divide(prev_ct, prev_tag, prev_parity);   // the group's last ciphertext
}

#endif
#if DPFHE_PART_GROUPED
// Grouped hybrid variant (dnum < L), DESIGN.md §2.11.  A group is Lq + K CTAs: CTA i < Lq owns ciphertext limb i, CTA Lq + k
// special prime k.  The roles are those of ks_hybrid_kernel with digits of K limbs:
//   limb CTA    tensor/permute, P*own terms + the key term of its own digit, INTT (scaled by Qhat^-1), publish
//               (dnum-1) x [basis conversion of a foreign digit + NTT + MAC]
//               2 x [basis conversion of the K tau' rows + NTT], out = (acc - s*u) / P     (one round late, as above)
//   special CTA dnum x [basis conversion + NTT + MAC into its scratch rows]; 2 x INTT (* (t Phat)^-1) -> tau', publish
// With Lq = 4, K = 2 every CTA runs four transforms per ciphertext (24 in all, against 30 for one special prime).
template <int LOGN, int NT, int MINB, int MODE>
__global__ void __launch_bounds__(NT, MINB) ks_grouped_kernel(KsArgs A, const __grid_constant__ LimbTable lt, const __grid_constant__ MsConsts K,
                                                              const __grid_constant__ GroupConsts G, size_t batch, u32 *flags, u32 epoch,
                                                              u32 *ticket, u64 *mail) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr size_t N = (size_t)1 << LOGN;
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    DevCta<NT> cta;
    __shared__ u32 s_ct;
    const u32 Lq = G.Lq, Ks = G.K, dnum = G.dnum, GS = Lq + Ks, slot = blockIdx.x, i = slot % GS, group = slot / GS, base = slot - i;
    const bool special = i >= Lq;
    const LimbParams &p = lt.lp[i];
    // rows of special prime k of this group: accumulators (0, 1) and tau' (double-buffered by round parity)
    auto hyb_of = [&](u32 k) { return A.hyb + ((size_t)group * Ks + k) * KS_HYB_ROWS * N; };
    auto wait_for = [&](u32 first, u32 count, u32 tag) {
        if (threadIdx.x < count) {
            while ((int)(ld_acquire_u32(flags + first + threadIdx.x) - tag) < 0) {
            }
        }
        __syncthreads();
    };
    auto publish = [&](u32 tag) {
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) st_release_u32(flags + slot, tag);
    };
    auto acc_of = [&](u32 parity) { return A.acc + ((size_t)slot * 2 + parity) * 2 * N; };
    auto divide = [&](size_t ct, u32 tag, u32 parity) {
        wait_for(base + Lq, Ks, tag);
        const size_t P = (size_t)Lq * N;
        for (u32 c = 0; c < 2; ++c) {
            u64 *row = A.out + ct * 2 * P + c * P + (size_t)i * N;
            ms_limb_group<LOGN, NT>(cta, buf, hyb_of(0) + ks_hyb_tau_row(parity, c) * N, (size_t)KS_HYB_ROWS * N, acc_of(parity) + c * N, row,
                                    A.tw + (size_t)i * N, p, K, G, i);
        }
    };
    bool pending = false;
    size_t prev_ct = 0;
    u32 prev_tag = 0, prev_parity = 0;
    for (u32 round = 0;; ++round) {
        const u32 tag = epoch + round + 1, parity = round & 1u;
        if (threadIdx.x == 0) {
            u64 *box = mail + (size_t)group * 2 + parity;
            if (i == 0) {
                const u32 t = atomicAdd(ticket, 1u);
                st_release_u64(box, ((u64)tag << 32) | t);
                s_ct = t;
            } else {
                u64 m;
                do m = ld_acquire_u64(box);
                while ((u32)(m >> 32) != tag);
                s_ct = (u32)m;
            }
        }
        __syncthreads();
        const size_t ct = s_ct;
        if (ct >= batch) break;
        const u64 *t_rows = A.scratch + ((size_t)base * 2 + parity) * N;   // row of limb j: + j * 2N
        if (!special) {
            const u32 g_own = i / Ks;
            ks_phase1<LOGN, NT, MODE, true>(cta, buf, A, G.lp_up[i], ct, i, A.scratch + ((size_t)slot * 2 + parity) * N, acc_of(parity), K.qlm[i],
                                            K.qlm_s[i], nullptr, 0, g_own);
            publish(tag);
            for (u32 jj = 1; jj < dnum; ++jj) {
                const u32 g = (g_own + jj) % dnum, lo = g * Ks, cnt = lo + Ks < Lq ? Ks : Lq - lo;
                wait_for(base + lo, cnt, tag);
                ks_phase2_group<LOGN, NT, false>(cta, buf, A, G, p, ct, i, g, jj, t_rows, 2 * N, acc_of(parity));
            }
            if (pending) divide(prev_ct, prev_tag, prev_parity);
            pending = true;
            prev_ct = ct;
            prev_tag = tag;
            prev_parity = parity;
        } else {
            u64 *hyb = hyb_of(i - Lq);
            for (u32 jj = 0; jj < dnum; ++jj) {
                const u32 g = (group + jj) % dnum, lo = g * Ks, cnt = lo + Ks < Lq ? Ks : Lq - lo;
                wait_for(base + lo, cnt, tag);
                ks_phase2_group<LOGN, NT, true>(cta, buf, A, G, p, ct, i, g, jj, t_rows, 2 * N, hyb);
            }
            for (u32 c = 0; c < 2; ++c)
                ms_tau_body<LOGN, NT, true>(cta, buf, hyb + c * N, hyb + c * N, A.itw + (size_t)i * N, G.lp_up[i], hyb + ks_hyb_tau_row(parity, c) * N, K);
            publish(tag);
        }
    }
    if (pending) divide(prev_ct, prev_tag, prev_parity);   // the group's last ciphertext
}

// Hoisted rotations with grouped hybrid keys, step 1: groups of L CTAs (every limb of the context) build
// U[ct][g][i]; ticket / mailbox / flag machinery as in ks_hoist_kernel, the digits converted as in ks_grouped_kernel.
// A limb CTA waits only for the digits of foreign groups and a special CTA publishes no digit, so the rounds of a group are
// not held together by the digit flags alone: every CTA also publishes done[slot] = tag at the end of a round, and enters round
// r only when the whole group has finished round r - 2 — the round whose digit slots and mailbox (both double-buffered by
// round parity) round r overwrites.
template <int LOGN, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) ks_hoistg_kernel(HoistGArgs A, const __grid_constant__ LimbTable lt, const __grid_constant__ GroupConsts G,
                                                             size_t batch, u32 *flags, u32 *done, u32 epoch, u32 *ticket, u64 *mail) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr size_t N = (size_t)1 << LOGN;
    u64 *buf = reinterpret_cast<u64 *>(smem_raw);
    DevCta<NT> cta;
    __shared__ u32 s_ct;
    const u32 Lq = G.Lq, Ks = G.K, dnum = G.dnum, GS = Lq + Ks, slot = blockIdx.x, i = slot % GS, group = slot / GS, base = slot - i;
    const LimbParams &p = lt.lp[i];
    auto wait_for = [&](const u32 *f, u32 first, u32 count, u32 tag) {
        if (threadIdx.x < count) {
            while ((int)(ld_acquire_u32(f + first + threadIdx.x) - tag) < 0) {
            }
        }
        __syncthreads();
    };
    for (u32 round = 0;; ++round) {
        const u32 tag = epoch + round + 1, parity = round & 1u;
        if (round >= 2) wait_for(done, base, GS, tag - 2);
        if (threadIdx.x == 0) {
            u64 *box = mail + (size_t)group * 2 + parity;
            if (i == 0) {
                const u32 t = atomicAdd(ticket, 1u);
                st_release_u64(box, ((u64)tag << 32) | t);
                s_ct = t;
            } else {
                u64 m;
                do m = ld_acquire_u64(box);
                while ((u32)(m >> 32) != tag);
                s_ct = (u32)m;
            }
        }
        __syncthreads();
        const size_t ct = s_ct;
        if (ct >= batch) break;
        const u64 *t_rows = A.scratch + ((size_t)base * 2 + parity) * N;
        u32 g_own = dnum;   // a special limb belongs to no digit
        if (i < Lq) {
            g_own = i / Ks;
            hoistg_phase1<LOGN, NT>(cta, buf, A, G, ct, i, A.scratch + ((size_t)slot * 2 + parity) * N);
            __threadfence();
            __syncthreads();
            if (threadIdx.x == 0) st_release_u32(flags + slot, tag);
        }
        for (u32 jj = 0; jj < dnum; ++jj) {
            const u32 g = (group + i + jj) % dnum;
            if (g == g_own) continue;
            const u32 lo = g * Ks, cnt = lo + Ks < Lq ? Ks : Lq - lo;
            wait_for(flags, base + lo, cnt, tag);
            hoistg_phase2<LOGN, NT>(cta, buf, A, G, p, ct, i, g, t_rows, 2 * N);
        }
        __syncthreads();   // every thread is past its last read of the group's digit slots
        if (threadIdx.x == 0) st_release_u32(done + slot, tag);
    }
}

// step 2: one rotation applied to blocks of ROT_CB ciphertexts x one limb of the context (gathers + multiply-accumulates)
template <int LOGN, int NT, int MINB, int ROT_CB>
__global__ void __launch_bounds__(NT, MINB) rot_apply_grouped_kernel(RotApplyGArgs A, const __grid_constant__ LimbTable lt, const __grid_constant__ MsConsts K,
                                                                     const __grid_constant__ GroupConsts G, size_t batch, u32 nseg) {
    DevCta<NT> cta;
    constexpr int NC = 1 << (LOGN - 1);
    const u32 L = G.Lq + G.K;
    const size_t n_blocks = (batch + ROT_CB - 1) / ROT_CB, n_items = n_blocks * L * nseg;
    const int seg_chunks = NC / (int)nseg;
    for (size_t w = blockIdx.x; w < n_items; w += gridDim.x) {
        const u32 seg = (u32)(w % nseg), i = (u32)((w / nseg) % L);
        const size_t ct0 = (w / nseg / L) * ROT_CB;
        const u32 n_ct = (u32)(batch - ct0 < (size_t)ROT_CB ? batch - ct0 : (size_t)ROT_CB);
        rot_apply_grouped_rows<LOGN, NT, ROT_CB>(cta, A, G, K, lt.lp[i], ct0, n_ct, i, (int)seg * seg_chunks, ((int)seg + 1) * seg_chunks);
    }
}

#endif
#if DPFHE_PART_MAIN
// ------------------------------------------------------------------ plaintext inner products (BSGS inner loop)
template <int LOGN, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) pt_inner_kernel(PtInnerArgs A, const __grid_constant__ LimbTable lt, u32 g0, u32 gcnt) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr u32 TILES_PER_LIMB = (1u << LOGN) / PTI_COEFFS;
    DevCta<NT> cta;
    const u32 limb = blockIdx.x / TILES_PER_LIMB, tile = blockIdx.x % TILES_PER_LIMB;
    pt_inner_tile<LOGN, NT>(cta, reinterpret_cast<u64 *>(smem_raw), A, lt.lp[limb], limb, tile, g0, gcnt);
}

// ------------------------------------------------------------------ element-wise kernels
// all operate on 16-byte chunks; chunk index -> limb = (chunk / (N/2)) % L
template <int LOGN>
__global__ void __launch_bounds__(256) pointwise_mul_kernel(const U64x2 *__restrict__ a, const U64x2 *__restrict__ b,
                                                            U64x2 *__restrict__ out, const LimbParams *__restrict__ lps,
                                                            u32 L, size_t n_chunks) {
    constexpr size_t NC = (size_t)1 << (LOGN - 1);
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (size_t)gridDim.x * blockDim.x) {
        const LimbParams p = lps[(c / NC) % L];
        st_stream(out + c, mul_chunk(ld_stream(a + c), ld_stream(b + c), p));
    }
}

// out = a + b mod q (canonical operands); out may alias either input
template <int LOGN>
__global__ void __launch_bounds__(256) poly_add_kernel(const U64x2 *__restrict__ a, const U64x2 *__restrict__ b, U64x2 *__restrict__ out,
                                                       const LimbParams *__restrict__ lps, u32 L, size_t n_chunks) {
    constexpr size_t NC = (size_t)1 << (LOGN - 1);
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (size_t)gridDim.x * blockDim.x) {
        const u64 q = lps[(c / NC) % L].q;
        const U64x2 x = ld_stream(a + c), y = ld_stream(b + c);
        U64x2 r;
        r.x = csub(x.x + y.x, q);
        r.y = csub(x.y + y.y, q);
        st_stream(out + c, r);
    }
}

// ct [batch][2][L][N] x pt [L][N]
template <int LOGN>
__global__ void __launch_bounds__(256) ct_mul_plain_kernel(const U64x2 *__restrict__ ct, const U64x2 *__restrict__ pt,
                                                           U64x2 *__restrict__ out, const LimbParams *__restrict__ lps,
                                                           u32 L, size_t n_chunks) {
    constexpr size_t NC = (size_t)1 << (LOGN - 1);
    const size_t pc = NC * L;   // chunks per polynomial
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (size_t)gridDim.x * blockDim.x) {
        const size_t in_poly = c % pc;
        const LimbParams p = lps[in_poly / NC];
        st_stream(out + c, mul_chunk(ld_stream(ct + c), ld_keep(pt + in_poly), p));
    }
}

// acc += ct o pt  (fused multiply-accumulate used by diagonal-method linear layers); acc may be lazy-free: all canonical
template <int LOGN>
__global__ void __launch_bounds__(256) ct_mul_plain_acc_kernel(const U64x2 *__restrict__ ct, const U64x2 *__restrict__ pt,
                                                               U64x2 *__restrict__ acc, const LimbParams *__restrict__ lps,
                                                               u32 L, size_t n_chunks) {
    constexpr size_t NC = (size_t)1 << (LOGN - 1);
    const size_t pc = NC * L;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (size_t)gridDim.x * blockDim.x) {
        const size_t in_poly = c % pc;
        const LimbParams p = lps[in_poly / NC];
        const U64x2 x = ld_stream(ct + c), m = ld_keep(pt + in_poly), a = ld_stream(acc + c);
        U64x2 r;
        r.x = csub(a.x + canon4(mulmod_lazy(x.x, m.x, p), p), p.q);
        r.y = csub(a.y + canon4(mulmod_lazy(x.y, m.y, p), p), p.q);
        st_stream(acc + c, r);
    }
}

// a,b [batch][2][L][N] -> d [batch][3][L][N]; one thread per chunk of one polynomial position
template <int LOGN>
__global__ void __launch_bounds__(256) ct_tensor_kernel(const U64x2 *__restrict__ a, const U64x2 *__restrict__ b,
                                                        U64x2 *__restrict__ d, const LimbParams *__restrict__ lps,
                                                        u32 L, size_t batch) {
    constexpr size_t NC = (size_t)1 << (LOGN - 1);
    const size_t pc = NC * L, total = batch * pc;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t ct = t / pc, in_poly = t % pc;
        const LimbParams p = lps[in_poly / NC];
        const U64x2 a0 = ld_stream(a + ct * 2 * pc + in_poly), a1 = ld_stream(a + ct * 2 * pc + pc + in_poly);
        const U64x2 b0 = ld_stream(b + ct * 2 * pc + in_poly), b1 = ld_stream(b + ct * 2 * pc + pc + in_poly);
        U64x2 d0, d1, d2;
        tensor_coeff(a0.x, a1.x, b0.x, b1.x, p, d0.x, d1.x, d2.x);
        tensor_coeff(a0.y, a1.y, b0.y, b1.y, p, d0.y, d1.y, d2.y);
        d0.x = canon4(d0.x, p); d0.y = canon4(d0.y, p);
        d1.x = canon4(d1.x, p); d1.y = canon4(d1.y, p);
        d2.x = canon4(d2.x, p); d2.y = canon4(d2.y, p);
        st_stream(d + ct * 3 * pc + in_poly, d0);
        st_stream(d + ct * 3 * pc + pc + in_poly, d1);
        st_stream(d + ct * 3 * pc + 2 * pc + in_poly, d2);
    }
}

// synthetic residues: x[k] = mulhi64(splitmix64(seed + k), q_limb)
template <int LOGN>
__global__ void __launch_bounds__(256) fill_uniform_kernel(U64x2 *__restrict__ out, const LimbParams *__restrict__ lps, u32 L,
                                                           u64 seed, u64 first_elem, size_t n_chunks) {
    constexpr size_t NC = (size_t)1 << (LOGN - 1);
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (size_t)gridDim.x * blockDim.x) {
        const u64 q = lps[(c / NC) % L].q;
        U64x2 v;
        v.x = __umul64hi(splitmix64(seed + first_elem + 2 * c), q);
        v.y = __umul64hi(splitmix64(seed + first_elem + 2 * c + 1), q);
        st_stream(out + c, v);
    }
}

#endif
// Shoup companions of a switch key: ks[e] = floor(key[e] * 2^64 / q_limb(e)); layout [L][2][L][N]
template <int LOGN>
__global__ void __launch_bounds__(256) key_prepare_kernel(const u64 *__restrict__ key, u64 *__restrict__ key_s,
                                                          const LimbParams *__restrict__ lps, u32 L, size_t n) {
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const u64 q = lps[(e >> LOGN) % L].q;
        key_s[e] = (u64)((((unsigned __int128)key[e]) << 64) / q);
    }
}

// ------------------------------------------------------------------ launchers
template <int LOGN>
struct Geometry {
    static constexpr int NT = LOGN == 12 ? 256 : 512;
    static constexpr size_t LIMB_BYTES = (size_t)8 << LOGN;
};

// "already configured on this device" bits of one kernel (a function-local static per launcher instantiation).  Several
// host threads may drive different devices at once (dpfhe_multi_*): the attribute call is idempotent, the bit set atomic.
struct ConfiguredMask {
    std::atomic<unsigned long long> bits{0};
    bool has(int device) const { return (bits.load(std::memory_order_acquire) >> (device & 63)) & 1ull; }
    void set(int device) { bits.fetch_or(1ull << (device & 63), std::memory_order_release); }
};

static unsigned ew_grid(const LaunchCtx &lc, size_t work_items) {
    size_t blocks = (work_items + 255) / 256;
    const size_t cap = (size_t)lc.num_sms * 32;   // 8 resident CTAs of 256 threads per SM, 4 waves
    if (blocks > cap) blocks = cap;
    return (unsigned)(blocks ? blocks : 1);
}

#if DPFHE_PART_MAIN
template <int LOGN, int NT, int MINB, bool INV>
static cudaError_t launch_ntt_t(const LaunchCtx &lc, u64 *data, size_t n_limbs, cudaStream_t st) {
    auto kern = ntt_kernel<LOGN, NT, MINB, INV>;
    const size_t smem = Geometry<LOGN>::LIMB_BYTES;
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    // one CTA per limb transform; the grid-stride loop only matters beyond 2^31-1 limbs
    const size_t grid = n_limbs < 0x7fffffffull ? n_limbs : 0x7fffffffull;
    kern<<<(unsigned)grid, NT, smem, st>>>(data, INV ? lc.itw : lc.tw, lc.lt, lc.L, n_limbs);
    return cudaGetLastError();
}

// tensor map of the caller's array seen as rows of 128 bytes; the encode function comes from the driver through the runtime
// (cudaGetDriverEntryPoint), so libdpfhe.so does not link libcuda
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static std::atomic<EncodeTiledFn> cached{nullptr};
    EncodeTiledFn f = cached.load();
    if (f) return f;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p)
        return nullptr;
    cached.store(reinterpret_cast<EncodeTiledFn>(p));
    return reinterpret_cast<EncodeTiledFn>(p);
}

// returns cudaErrorNotSupported when the TMA path does not apply (no driver entry point, too many rows for 32-bit coordinates):
// the caller then runs the ordinary kernel
template <int LOGN, int NT, int MINB>
static cudaError_t launch_ntt_inv_tma(const LaunchCtx &lc, u64 *data, size_t n_limbs, cudaStream_t st) {
    constexpr size_t ROWS = ((size_t)1 << LOGN) * 8 / 128;
    const size_t rows = n_limbs * ROWS;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || rows >= 0x7fffffffull || n_limbs >= 0x7fffffffull) return cudaErrorNotSupported;
    CUtensorMap tm;
    const cuuint64_t dims[2] = {16, (cuuint64_t)rows}, strides[1] = {128};
    const cuuint32_t box[2] = {16, (cuuint32_t)(ROWS < 256 ? ROWS : 256)}, estr[2] = {1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, data, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return cudaErrorNotSupported;
    auto kern = ntt_inv_tma_kernel<LOGN, NT, MINB>;
    const size_t smem = Geometry<LOGN>::LIMB_BYTES;
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    kern<<<(unsigned)n_limbs, NT, smem, st>>>(tm, data, lc.itw, lc.lt, lc.L);
    return cudaGetLastError();
}

template <int LOGN, int NT, int MINB>
static cudaError_t launch_ntt_dir(const LaunchCtx &lc, u64 *data, size_t n_limbs, bool inverse, cudaStream_t st) {
    if constexpr (LOGN <= 13 && NT == 256) {
        if (inverse && lc.ntt_tma) {
            cudaError_t e = launch_ntt_inv_tma<LOGN, NT, MINB>(lc, data, n_limbs, st);
            if (e != cudaErrorNotSupported) return e;
        }
    }
    return inverse ? launch_ntt_t<LOGN, NT, MINB, true>(lc, data, n_limbs, st) : launch_ntt_t<LOGN, NT, MINB, false>(lc, data, n_limbs, st);
}

template <bool INV>
static cudaError_t launch_ntt_pair_t(const LaunchCtx &lc, u64 *data, size_t n_limbs, cudaStream_t st) {
    constexpr int NT = 256, MINB = 3;
    auto kern = ntt_pair_kernel<NT, MINB, INV>;
    const size_t smem = Geometry<13>::LIMB_BYTES;   // half a limb
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    const size_t pairs = n_limbs < 0x3fffffffull ? n_limbs : 0x3fffffffull;
    kern<<<(unsigned)(2 * pairs), NT, smem, st>>>(data, INV ? lc.itw : lc.tw, lc.lt, lc.L, n_limbs);
    return cudaGetLastError();
}
static cudaError_t launch_ntt_pair(const LaunchCtx &lc, u64 *data, size_t n_limbs, bool inverse, cudaStream_t st) {
    return inverse ? launch_ntt_pair_t<true>(lc, data, n_limbs, st) : launch_ntt_pair_t<false>(lc, data, n_limbs, st);
}

cudaError_t launch_ntt(const LaunchCtx &lc, u64 *data, size_t n_polys, bool inverse, cudaStream_t st) {
    const size_t n_limbs = n_polys * lc.L;
    if (n_limbs == 0) return cudaSuccess;
    switch (lc.log_n) {
        case 12: return launch_ntt_dir<12, 256, 2>(lc, data, n_limbs, inverse, st);
        case 13:
            switch (lc.ntt_cfg) {   // tuning variants (DPFHE_NTT_CFG), default 0
                case 1: return launch_ntt_dir<13, 512, 1>(lc, data, n_limbs, inverse, st);   //  9.9 M NTT/s
                case 2: return launch_ntt_dir<13, 512, 2>(lc, data, n_limbs, inverse, st);   // 12.6 M (64 regs, spills)
                case 3: return launch_ntt_dir<13, 256, 2>(lc, data, n_limbs, inverse, st);   // 12.3 M
                default: return launch_ntt_dir<13, 256, 3>(lc, data, n_limbs, inverse, st);  // 13.2 M: 3 CTAs/SM (smem-limited), 80 regs
            }
        case 14:
            if (lc.ntt_cfg == 1) return launch_ntt_dir<14, 512, 1>(lc, data, n_limbs, inverse, st);   // whole limb per CTA, 1 CTA/SM
            return launch_ntt_pair(lc, data, n_limbs, inverse, st);                                   // CTA pair per limb, 3 CTAs/SM
    }
    return cudaErrorInvalidValue;
}

template <int LOGN, int NT, int MINB>
static cudaError_t launch_ms_t(const LaunchCtx &lc, const u64 *in, u64 *tau, u64 *out, const MsConsts &K, size_t n_polys, cudaStream_t st) {
    auto k1 = ms_tau_kernel<LOGN, NT, MINB>;
    auto k2 = ms_limb_kernel<LOGN, NT, MINB>;
    const size_t smem = Geometry<LOGN>::LIMB_BYTES;
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    const size_t n_items = n_polys * (lc.L - 1);
    k1<<<(unsigned)(n_polys < 0x7fffffffull ? n_polys : 0x7fffffffull), NT, smem, st>>>(in, tau, lc.itw, lc.lt, K, lc.L, n_polys);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k2<<<(unsigned)(n_items < 0x7fffffffull ? n_items : 0x7fffffffull), NT, smem, st>>>(in, tau, out, lc.tw, lc.lt, K, lc.L, n_items);
    return cudaGetLastError();
}

cudaError_t launch_mod_switch(const LaunchCtx &lc, const u64 *in, u64 *tau, u64 *out, const MsConsts &K, size_t n_polys, cudaStream_t st) {
    if (n_polys == 0) return cudaSuccess;
    switch (lc.log_n) {
        case 12: return launch_ms_t<12, 256, 2>(lc, in, tau, out, K, n_polys, st);
        case 13: return launch_ms_t<13, 256, 3>(lc, in, tau, out, K, n_polys, st);
        case 14: return launch_ms_t<14, 512, 1>(lc, in, tau, out, K, n_polys, st);
    }
    return cudaErrorInvalidValue;
}

template <int LOGN, int NT, int MINB>
static cudaError_t launch_md_t(const LaunchCtx &lc, const u64 *in, u64 *tau, u64 *out, const MsConsts &K, const GroupConsts &G, size_t n_polys,
                               cudaStream_t st) {
    auto k1 = md_tau_kernel<LOGN, NT, MINB>;
    auto k2 = md_limb_kernel<LOGN, NT, MINB>;
    const size_t smem = Geometry<LOGN>::LIMB_BYTES;
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    const size_t n_tau = n_polys * G.K, n_items = n_polys * G.Lq;
    k1<<<(unsigned)(n_tau < 0x7fffffffull ? n_tau : 0x7fffffffull), NT, smem, st>>>(in, tau, lc.itw, K, G, n_tau);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k2<<<(unsigned)(n_items < 0x7fffffffull ? n_items : 0x7fffffffull), NT, smem, st>>>(in, tau, out, lc.tw, lc.lt, K, G, n_items);
    return cudaGetLastError();
}

// in [n_polys][L][N] -> out [n_polys][L-K][N]; tau: n_polys * K * N words of scratch
cudaError_t launch_mod_down_special(const LaunchCtx &lc, const u64 *in, u64 *tau, u64 *out, const MsConsts &K, const GroupConsts &G, size_t n_polys,
                                    cudaStream_t st) {
    if (n_polys == 0) return cudaSuccess;
    if (G.Lq + G.K != lc.L) return cudaErrorInvalidValue;
    switch (lc.log_n) {
        case 12: return launch_md_t<12, 256, 2>(lc, in, tau, out, K, G, n_polys, st);
        case 13: return launch_md_t<13, 256, 3>(lc, in, tau, out, K, G, n_polys, st);
        case 14: return launch_md_t<14, 512, 1>(lc, in, tau, out, K, G, n_polys, st);
    }
    return cudaErrorInvalidValue;
}

#endif
// Flag, round-mark and mailbox tags are 32-bit round numbers compared by signed difference, so a slot last written more than 2^31
// rounds ago would look "published" (a context that has multiplied 2^31 ciphertexts: an hour of work).  Long before that the
// numbering restarts: flags, marks, mailboxes and hand-back counters are cleared in stream order - after every earlier launch of
// the context (abi.cu orders its calls) and before this one - and the epoch returns to zero.
static cudaError_t epoch_guard(LaunchCtx &lc, size_t batch, cudaStream_t st) {
    if (batch + 1 >= 0x40000000ull) return cudaErrorInvalidValue;
    if ((unsigned long long)lc.ks_epoch + batch + 1 < lc.ks_epoch_limit) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(lc.ks_flags, 0, 2 * lc.ks_slots * sizeof(u32), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(lc.ks_mail, 0, lc.ks_slots * sizeof(u64), st);
    if (e == cudaSuccess) e = cudaMemsetAsync(lc.ks_consumed, 0, lc.ks_slots * sizeof(u32), st);
    if (e != cudaSuccess) return e;
    lc.ks_epoch = 0;
    ++lc.ks_epoch_restarts;
    return cudaSuccess;
}

#if DPFHE_PART_MAIN
template <int LOGN, int MODE>
static cudaError_t launch_ks_t(LaunchCtx &lc, const KsArgs &A, size_t batch, cudaStream_t st) {
    // at most 64 KiB of shared memory per CTA (N = 16384 is processed as two half-limbs) -> three CTAs per SM
    constexpr int NT = 256, MINB = 3;
    const bool filter = A.only != nullptr;
    if (filter && MODE != KS_ROTATE) return cudaErrorInvalidValue;
    auto kern = ks_fused_kernel<LOGN, NT, MINB, MODE, false>;
    if constexpr (MODE == KS_MUL_RELIN) {   // the per-phase clock counters (DPFHE_KS_PROF, tools/phase_prof.py) exist for ct x ct only
        if (lc.ks_prof) kern = ks_fused_kernel<LOGN, NT, MINB, MODE, true>;
    }
    if (filter) kern = ks_fused_kernel<LOGN, NT, MINB, MODE == KS_ROTATE ? MODE : KS_ROTATE, false, MODE == KS_ROTATE>;
    const size_t smem = LOGN <= 13 ? Geometry<LOGN>::LIMB_BYTES : Geometry<13>::LIMB_BYTES;
    static ConfiguredMask configured[3];
    const int variant = filter ? 2 : (lc.ks_prof && MODE == KS_MUL_RELIN ? 1 : 0);
    if (!configured[variant].has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[variant].set(lc.device);
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    if (lc.ks_occ_cap > 0 && occ > lc.ks_occ_cap) occ = lc.ks_occ_cap;
    size_t G = (size_t)lc.num_sms * occ;
    if (G > lc.ks_slots) G = lc.ks_slots;
    G = (G / lc.L) * lc.L;
    const size_t n_work = batch * lc.L;
    if (G > n_work) G = n_work;
    if (G == 0) return cudaErrorInvalidConfiguration;
    // flag / mailbox tags this launch may consume: one per round, and a group runs at most batch + 1 rounds
    const u32 rounds = (u32)(batch + 1);
    cudaError_t em = epoch_guard(lc, batch, st);
    if (em != cudaSuccess) return em;
    em = cudaMemsetAsync(lc.ks_ticket, 0, sizeof(u32), st);
    if (em != cudaSuccess) return em;
    KsArgs args = A;
    size_t batch_arg = batch;
    u32 *flags = lc.ks_flags;
    u32 epoch = lc.ks_epoch;
    LimbTable lt = lc.lt;
    unsigned long long *prof = lc.ks_prof;
    u32 *ticket = lc.ks_ticket;
    u64 *mail = lc.ks_mail;
    u32 pf_dist = (u32)lc.ks_prefetch;
    u32 *consumed = lc.ks_single ? lc.ks_consumed : nullptr;
    void *params[] = {&args, &lt, &batch_arg, &flags, &epoch, &ticket, &mail, &prof, &pf_dist, &consumed};
    if (lc.l2_persist && lc.l2_persist_max && lc.ks_window_bytes) {
        // tuning (DPFHE_L2_PERSIST): the digit slots and accumulator rows are re-read within microseconds, the ciphertext
        // streams never; a persisting access-policy window over the scratch keeps the streams from evicting it
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)G);
        cfg.blockDim = dim3(NT);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeCooperative;
        attr[0].val.cooperative = 1;
        attr[1].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[1].val.accessPolicyWindow.base_ptr = lc.ks_scratch;
        attr[1].val.accessPolicyWindow.num_bytes = lc.ks_window_bytes;
        const double ratio = (double)lc.l2_persist_max / (double)lc.ks_window_bytes;
        attr[1].val.accessPolicyWindow.hitRatio = (float)(ratio > 1.0 ? 1.0 : ratio);
        attr[1].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[1].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        cfg.attrs = attr;
        cfg.numAttrs = 2;
        e = cudaLaunchKernelExC(&cfg, (const void *)kern, params);
    } else {
        e = cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)G), dim3(NT), params, smem, st);
    }
    lc.ks_epoch += rounds;
    return e;
}

#endif
#if DPFHE_PART_HYBRID
template <int LOGN, int MODE>
static cudaError_t launch_ks_hybrid_t(LaunchCtx &lc, const KsArgs &A, const MsConsts &K, size_t batch, cudaStream_t st) {
    constexpr int NT = 256, MINB = 3;
    auto kern = ks_hybrid_kernel<LOGN, NT, MINB, MODE>;
    const size_t smem = LOGN <= 13 ? Geometry<LOGN>::LIMB_BYTES : Geometry<13>::LIMB_BYTES;
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    if (lc.ks_occ_cap > 0 && occ > lc.ks_occ_cap) occ = lc.ks_occ_cap;
    const size_t GS = lc.L;   // group size: L-1 ciphertext limbs + the special limb
    size_t G = (size_t)lc.num_sms * occ;
    if (G > lc.ks_slots) G = lc.ks_slots;
    G = (G / GS) * GS;
    if (G > batch * GS) G = batch * GS;
    if (G == 0) return cudaErrorInvalidConfiguration;
    const u32 rounds = (u32)(batch + 1);
    cudaError_t em = epoch_guard(lc, batch, st);
    if (em != cudaSuccess) return em;
    em = cudaMemsetAsync(lc.ks_ticket, 0, sizeof(u32), st);
    if (em != cudaSuccess) return em;
    KsArgs args = A;
    LimbTable lt = lc.lt;
    MsConsts consts = K;
    size_t batch_arg = batch;
    u32 *flags = lc.ks_flags;
    u32 epoch = lc.ks_epoch;
    u32 *ticket = lc.ks_ticket;
    u64 *mail = lc.ks_mail;
    void *params[] = {&args, &lt, &consts, &batch_arg, &flags, &epoch, &ticket, &mail};
    e = cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)G), dim3(NT), params, smem, st);
    lc.ks_epoch += rounds;
    return e;
}

// data has L-1 limbs, the key [L-1][2][L][N]; lc.ks_hyb must hold (ks_slots / 2 + 1) * KS_HYB_ROWS * N words and
// lc.ks_acc_hyb ks_slots * 2 * 2 * N words
cudaError_t launch_ks_hybrid(LaunchCtx &lc, int mode, const u64 *a, const u64 *b, const u64 *key, u64 *out, size_t batch, u32 galois,
                             const MsConsts &K, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    if (lc.L < 2 || !lc.ks_hyb || !lc.ks_acc_hyb) return cudaErrorInvalidValue;
    {
        const size_t n = (size_t)2 * (lc.L - 1) * lc.L << lc.log_n;
        const unsigned grid = ew_grid(lc, n);
        if (lc.log_n == 12) key_prepare_kernel<12><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        else if (lc.log_n == 13) key_prepare_kernel<13><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        else key_prepare_kernel<14><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    KsArgs A;
    A.a = a; A.b = b; A.key = key; A.key_s = lc.ks_key_s; A.out = out; A.scratch = lc.ks_scratch;
    A.tw = lc.tw; A.itw = lc.itw; A.L = lc.L - 1; A.galois = galois; A.Lk = lc.L; A.hyb = lc.ks_hyb; A.only = nullptr;
    A.acc = lc.ks_acc_hyb; A.acc_par = 2; A.lift_reduce = lc.lift_reduce ? 1u : 0u;
#define KS_HYB_DISPATCH(LOGN)                                                                   \
    switch (mode) {                                                                             \
        case KS_MUL_RELIN: return launch_ks_hybrid_t<LOGN, KS_MUL_RELIN>(lc, A, K, batch, st);   \
        case KS_PLAIN: return launch_ks_hybrid_t<LOGN, KS_PLAIN>(lc, A, K, batch, st);           \
        case KS_ROTATE: return launch_ks_hybrid_t<LOGN, KS_ROTATE>(lc, A, K, batch, st);         \
    }                                                                                           \
    return cudaErrorInvalidValue;
    switch (lc.log_n) {
        case 12: KS_HYB_DISPATCH(12)
        case 13: KS_HYB_DISPATCH(13)
        case 14: KS_HYB_DISPATCH(14)
    }
    return cudaErrorNotSupported;
}

#endif
#if DPFHE_PART_GROUPED
template <int LOGN, int MODE>
static cudaError_t launch_ks_grouped_t(LaunchCtx &lc, const KsArgs &A, const MsConsts &K, const GroupConsts &Gc, size_t batch, cudaStream_t st) {
    constexpr int NT = 256, MINB = 3;
    auto kern = ks_grouped_kernel<LOGN, NT, MINB, MODE>;
    const size_t smem = LOGN <= 13 ? Geometry<LOGN>::LIMB_BYTES : Geometry<13>::LIMB_BYTES;
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    if (lc.ks_occ_cap > 0 && occ > lc.ks_occ_cap) occ = lc.ks_occ_cap;
    const size_t GS = lc.L;   // group size: every limb of the context, ciphertext and special
    size_t G = (size_t)lc.num_sms * occ;
    if (G > lc.ks_slots) G = lc.ks_slots;
    G = (G / GS) * GS;
    if (G > batch * GS) G = batch * GS;
    if (G == 0) return cudaErrorInvalidConfiguration;
    const u32 rounds = (u32)(batch + 1);
    cudaError_t em = epoch_guard(lc, batch, st);
    if (em != cudaSuccess) return em;
    em = cudaMemsetAsync(lc.ks_ticket, 0, sizeof(u32), st);
    if (em != cudaSuccess) return em;
    KsArgs args = A;
    LimbTable lt = lc.lt;
    MsConsts consts = K;
    GroupConsts gc = Gc;
    size_t batch_arg = batch;
    u32 *flags = lc.ks_flags;
    u32 epoch = lc.ks_epoch;
    u32 *ticket = lc.ks_ticket;
    u64 *mail = lc.ks_mail;
    void *params[] = {&args, &lt, &consts, &gc, &batch_arg, &flags, &epoch, &ticket, &mail};
    e = cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)G), dim3(NT), params, smem, st);
    lc.ks_epoch += rounds;
    return e;
}

// data has Lq = L - K limbs, the key [dnum][2][L][N]; scratch requirements as launch_ks_hybrid (K <= Lq keeps the special
// CTAs' rows within lc.ks_hyb)
cudaError_t launch_ks_grouped(LaunchCtx &lc, int mode, const u64 *a, const u64 *b, const u64 *key, u64 *out, size_t batch, u32 galois,
                              const MsConsts &K, const GroupConsts &Gc, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    if (lc.L < 2 || !lc.ks_hyb || !lc.ks_acc_hyb || Gc.Lq + Gc.K != lc.L || Gc.K > Gc.Lq || Gc.K > (u32)KS_MAX_SPECIAL) return cudaErrorInvalidValue;
    {
        const size_t n = (size_t)2 * Gc.dnum * lc.L << lc.log_n;
        const unsigned grid = ew_grid(lc, n);
        if (lc.log_n == 12) key_prepare_kernel<12><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        else if (lc.log_n == 13) key_prepare_kernel<13><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        else key_prepare_kernel<14><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    KsArgs A;
    A.a = a; A.b = b; A.key = key; A.key_s = lc.ks_key_s; A.out = out; A.scratch = lc.ks_scratch;
    A.tw = lc.tw; A.itw = lc.itw; A.L = Gc.Lq; A.galois = galois; A.Lk = lc.L; A.hyb = lc.ks_hyb; A.only = nullptr;
    A.acc = lc.ks_acc_hyb; A.acc_par = 2; A.lift_reduce = 0u;
#define KS_GRP_DISPATCH(LOGN)                                                                        \
    switch (mode) {                                                                                  \
        case KS_MUL_RELIN: return launch_ks_grouped_t<LOGN, KS_MUL_RELIN>(lc, A, K, Gc, batch, st);   \
        case KS_PLAIN: return launch_ks_grouped_t<LOGN, KS_PLAIN>(lc, A, K, Gc, batch, st);           \
        case KS_ROTATE: return launch_ks_grouped_t<LOGN, KS_ROTATE>(lc, A, K, Gc, batch, st);         \
    }                                                                                                \
    return cudaErrorInvalidValue;
    switch (lc.log_n) {
        case 12: KS_GRP_DISPATCH(12)
        case 13: KS_GRP_DISPATCH(13)
        case 14: KS_GRP_DISPATCH(14)
    }
    return cudaErrorNotSupported;
}

#endif
#if DPFHE_PART_MAIN
// shared-memory plan of pt_inner_kernel: two CTAs per SM (113 KiB each); the plaintext tile takes nb * 128 bytes per
// giant step and the two ciphertext-row buffers 2 * nb * 128 bytes each.  Returns giant steps per launch (0: nb too large).
static constexpr size_t PTI_SMEM_BUDGET = (size_t)113 << 10;
static u32 pt_inner_gmax(u32 nb) {
    const size_t row = (size_t)nb * PTI_COEFFS * 8;
    if (5 * row > PTI_SMEM_BUDGET) return 0;
    u32 gmax = (u32)((PTI_SMEM_BUDGET - 4 * row) / row);
    if (gmax >= 8) gmax -= gmax % 8;   // whole rounds of the CTA's eight warps
    return gmax;
}

template <int LOGN>
static cudaError_t launch_pt_inner_t(const LaunchCtx &lc, const PtInnerArgs &A, u32 gmax, cudaStream_t st) {
    constexpr int NT = 256, MINB = 2;
    auto kern = pt_inner_kernel<LOGN, NT, MINB>;
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PTI_SMEM_BUDGET);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    const size_t row = (size_t)A.nb * PTI_COEFFS * 8;
    const unsigned grid = (unsigned)(lc.L * (((size_t)1 << LOGN) / PTI_COEFFS));
    for (u32 g0 = 0; g0 < A.ng; g0 += gmax) {
        const u32 gcnt = A.ng - g0 < gmax ? A.ng - g0 : gmax;
        kern<<<grid, NT, ((size_t)gcnt + 4) * row, st>>>(A, lc.lt, g0, gcnt);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// out[g] = sum_b steps[b] o pts[g][b]; returns the number of kernel launches through *launches
cudaError_t launch_pt_inner(const LaunchCtx &lc, const u64 *steps, u32 nb, const u64 *pts, u32 ng, u64 *out, size_t batch, cudaStream_t st,
                            unsigned *launches) {
    *launches = 0;
    if (!batch || !nb || !ng) return cudaSuccess;
    PtInnerArgs A;
    A.steps = steps; A.pts = pts; A.out = out; A.batch = batch; A.L = lc.L; A.nb = nb; A.ng = ng;
    const u32 gmax = pt_inner_gmax(nb);
    if (gmax == 0) return cudaErrorInvalidValue;
    *launches = (ng + gmax - 1) / gmax;
    switch (lc.log_n) {
        case 12: return launch_pt_inner_t<12>(lc, A, gmax, st);
        case 13: return launch_pt_inner_t<13>(lc, A, gmax, st);
        case 14: return launch_pt_inner_t<14>(lc, A, gmax, st);
    }
    return cudaErrorInvalidValue;
}

template <int LOGN>
static cudaError_t launch_hoist_t(LaunchCtx &lc, const HoistArgs &A, size_t batch, cudaStream_t st) {
    constexpr int NT = 256, MINB = 3;
    auto kern = ks_hoist_kernel<LOGN, NT, MINB>;
    const size_t smem = LOGN <= 13 ? Geometry<LOGN>::LIMB_BYTES : Geometry<13>::LIMB_BYTES;
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    size_t G = (size_t)lc.num_sms * occ;
    if (G > lc.ks_slots) G = lc.ks_slots;
    G = (G / lc.L) * lc.L;
    if (G > batch * lc.L) G = batch * lc.L;
    if (G == 0) return cudaErrorInvalidConfiguration;
    cudaError_t em = epoch_guard(lc, batch, st);
    if (em != cudaSuccess) return em;
    em = cudaMemsetAsync(lc.ks_ticket, 0, sizeof(u32), st);
    if (em != cudaSuccess) return em;
    HoistArgs args = A;
    LimbTable lt = lc.lt;
    size_t batch_arg = batch;
    u32 *flags = lc.ks_flags;
    u32 epoch = lc.ks_epoch;
    u32 *ticket = lc.ks_ticket;
    u64 *mail = lc.ks_mail;
    void *params[] = {&args, &lt, &batch_arg, &flags, &epoch, &ticket, &mail};
    e = cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)G), dim3(NT), params, smem, st);
    lc.ks_epoch += (u32)(batch + 1);
    return e;
}

// hoisted rotations, step 1: U[ct][j][i] and the zero flags of `batch` ciphertexts (L >= 2)
cudaError_t launch_hoist(LaunchCtx &lc, const u64 *ct, u64 *U, u32 *zero, size_t batch, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    HoistArgs A;
    A.ct = ct; A.U = U; A.scratch = lc.ks_scratch; A.zero = zero; A.tw = lc.tw; A.itw = lc.itw; A.L = lc.L;
    switch (lc.log_n) {
        case 12: return launch_hoist_t<12>(lc, A, batch, st);
        case 13: return launch_hoist_t<13>(lc, A, batch, st);
        case 14: return launch_hoist_t<14>(lc, A, batch, st);
    }
    return cudaErrorInvalidValue;
}

#endif
#if DPFHE_PART_GROUPED
template <int LOGN>
static cudaError_t launch_hoistg_t(LaunchCtx &lc, const HoistGArgs &A, const GroupConsts &Gc, size_t batch, cudaStream_t st) {
    constexpr int NT = 256, MINB = 3;
    auto kern = ks_hoistg_kernel<LOGN, NT, MINB>;
    const size_t smem = LOGN <= 13 ? Geometry<LOGN>::LIMB_BYTES : Geometry<13>::LIMB_BYTES;
    static ConfiguredMask configured;
    if (!configured.has(lc.device)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured.set(lc.device);
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (occ < 1) return cudaErrorLaunchOutOfResources;
    size_t G = (size_t)lc.num_sms * occ;
    if (G > lc.ks_slots) G = lc.ks_slots;
    G = (G / lc.L) * lc.L;
    if (G > batch * lc.L) G = batch * lc.L;
    if (G == 0) return cudaErrorInvalidConfiguration;
    cudaError_t em = epoch_guard(lc, batch, st);
    if (em != cudaSuccess) return em;
    em = cudaMemsetAsync(lc.ks_ticket, 0, sizeof(u32), st);
    if (em != cudaSuccess) return em;
    HoistGArgs args = A;
    LimbTable lt = lc.lt;
    GroupConsts gc = Gc;
    size_t batch_arg = batch;
    u32 *flags = lc.ks_flags;
    u32 *done = lc.ks_flags + lc.ks_slots;   // second half of the flag array: "round finished" marks
    u32 epoch = lc.ks_epoch;
    u32 *ticket = lc.ks_ticket;
    u64 *mail = lc.ks_mail;
    void *params[] = {&args, &lt, &gc, &batch_arg, &flags, &done, &epoch, &ticket, &mail};
    e = cudaLaunchCooperativeKernel((void *)kern, dim3((unsigned)G), dim3(NT), params, smem, st);
    lc.ks_epoch += (u32)(batch + 1);
    return e;
}

// hoisted rotations with grouped hybrid keys, step 1: U [batch][dnum][L][N]
cudaError_t launch_hoist_grouped(LaunchCtx &lc, const u64 *ct, u64 *U, const GroupConsts &G, size_t batch, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    if (G.Lq + G.K != lc.L) return cudaErrorInvalidValue;
    HoistGArgs A;
    A.ct = ct; A.U = U; A.scratch = lc.ks_scratch; A.tw = lc.tw; A.itw = lc.itw;
    switch (lc.log_n) {
        case 12: return launch_hoistg_t<12>(lc, A, G, batch, st);
        case 13: return launch_hoistg_t<13>(lc, A, G, batch, st);
        case 14: return launch_hoistg_t<14>(lc, A, G, batch, st);
    }
    return cudaErrorInvalidValue;
}

// step 2: acc [batch][2][L][N] of one rotation (key companions built here into lc.ks_key_s unless the caller supplies them)
cudaError_t launch_rot_apply_grouped(LaunchCtx &lc, const u64 *ct, const u64 *U, const u64 *key, const u64 *key_s, u32 galois, u64 *acc,
                                     const MsConsts &K, const GroupConsts &G, size_t batch, cudaStream_t st) {
    if (batch == 0) return cudaSuccess;
    if (!key_s) {
        const size_t n = (size_t)2 * G.dnum * lc.L << lc.log_n;
        const unsigned grid = ew_grid(lc, n);
        if (lc.log_n == 12) key_prepare_kernel<12><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        else if (lc.log_n == 13) key_prepare_kernel<13><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        else key_prepare_kernel<14><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
        key_s = lc.ks_key_s;
    }
    RotApplyGArgs A;
    A.ct = ct; A.U = U; A.key = key; A.key_s = key_s; A.acc = acc; A.galois = galois;
    constexpr size_t cb = 2;
    const size_t rows = ((batch + cb - 1) / cb) * lc.L, want = (size_t)lc.num_sms * 12;
    u32 nseg = 1;
    const u32 max_seg = (1u << (lc.log_n - 1)) / 256;
    while (nseg < max_seg && rows * nseg < want) nseg *= 2;
    const size_t n_items = rows * nseg, cap = (size_t)lc.num_sms * 8;
    const unsigned grid = (unsigned)(n_items < cap ? n_items : cap);
    switch (lc.log_n) {
        case 12: rot_apply_grouped_kernel<12, 256, 3, 2><<<grid, 256, 0, st>>>(A, lc.lt, K, G, batch, nseg); break;
        case 13: rot_apply_grouped_kernel<13, 256, 3, 2><<<grid, 256, 0, st>>>(A, lc.lt, K, G, batch, nseg); break;
        case 14: rot_apply_grouped_kernel<14, 256, 3, 2><<<grid, 256, 0, st>>>(A, lc.lt, K, G, batch, nseg); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

#endif
#if DPFHE_PART_MAIN
// per-rotation constants: Shoup companions of the key (lc.ks_key_s), M = NTT(negmask_g) (in `M`, [L][N]) and kprime [2][L][N]
cudaError_t launch_rot_prepare(LaunchCtx &lc, const u64 *key, u32 galois, const u64 *delta, u64 *M, u64 *kprime, cudaStream_t st, u64 *key_s_out) {
    u64 *key_s = key_s_out ? key_s_out : lc.ks_key_s;   // a caller that keeps the constants of a rotation supplies its own buffer
    const size_t n = (size_t)2 * lc.L * lc.L << lc.log_n;
    const unsigned grid = ew_grid(lc, n), gsmall = ew_grid(lc, (size_t)2 * lc.L << lc.log_n);
#define ROT_PREP(LOGN)                                                                              \
    key_prepare_kernel<LOGN><<<grid, 256, 0, st>>>(key, key_s, lc.lp, lc.L, n);                      \
    negmask_kernel<LOGN><<<(1u << LOGN) / 256, 256, 0, st>>>(M, galois, lc.L);
    switch (lc.log_n) {
        case 12: ROT_PREP(12) break;
        case 13: ROT_PREP(13) break;
        case 14: ROT_PREP(14) break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    e = launch_ntt(lc, M, 1, false, st);
    if (e != cudaSuccess) return e;
    switch (lc.log_n) {
        case 12: kprime_kernel<12><<<gsmall, 256, 0, st>>>(key, M, delta, kprime, lc.lp, lc.L); break;
        case 13: kprime_kernel<13><<<gsmall, 256, 0, st>>>(key, M, delta, kprime, lc.lp, lc.L); break;
        case 14: kprime_kernel<14><<<gsmall, 256, 0, st>>>(key, M, delta, kprime, lc.lp, lc.L); break;
    }
    return cudaGetLastError();
}

// hoisted rotations, step 2: one rotation of `batch` ciphertexts from the shared transforms (constants from launch_rot_prepare)
cudaError_t launch_rot_apply(const LaunchCtx &lc, const u64 *ct, const u64 *U, const u64 *key, const u64 *kprime, u32 galois, u64 *out,
                             size_t batch, cudaStream_t st, const u64 *key_s) {
    if (batch == 0) return cudaSuccess;
    RotApplyArgs A;
    A.ct = ct; A.U = U; A.key = key; A.key_s = key_s ? key_s : lc.ks_key_s; A.kprime = kprime; A.out = out; A.L = lc.L; A.galois = galois;
    // rows are cut into up to NC / 256 segments so that small batches still fill the machine several times over
    // tuning variant (DPFHE_ROT_CFG): 0 (default) = two ciphertexts share each key chunk, next digit prefetched;
    // 1 = one ciphertext per item, prefetched; 2 = one ciphertext, no prefetch.  Measured 11.1 / 12.1 / 11.9 ms for 31
    // rotations of 512 ciphertexts at N = 8192, L = 4 and 31.9 / 36.5 / 34.8 ms for 26 x 256 at N = 16384, L = 8.
    const int cfg = lc.rot_cfg;
    const size_t cb = cfg == 0 ? 2 : 1;
    const size_t rows = ((batch + cb - 1) / cb) * lc.L, want = (size_t)lc.num_sms * 12;
    u32 nseg = 1;
    const u32 max_seg = (1u << (lc.log_n - 1)) / 256;
    while (nseg < max_seg && rows * nseg < want) nseg *= 2;
    const size_t n_items = rows * nseg, cap = (size_t)lc.num_sms * 8;
    const unsigned grid = (unsigned)(n_items < cap ? n_items : cap);
#define ROT_APPLY(LOGN)                                                                                          \
    if (cfg == 1) rot_apply_kernel<LOGN, 256, 3, 1, true><<<grid, 256, 0, st>>>(A, lc.lt, batch, nseg);           \
    else if (cfg == 2) rot_apply_kernel<LOGN, 256, 3, 1, false><<<grid, 256, 0, st>>>(A, lc.lt, batch, nseg);     \
    else rot_apply_kernel<LOGN, 256, 3, 2, true><<<grid, 256, 0, st>>>(A, lc.lt, batch, nseg);
    switch (lc.log_n) {
        case 12: ROT_APPLY(12) break;
        case 13: ROT_APPLY(13) break;
        case 14: ROT_APPLY(14) break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t launch_ks(LaunchCtx &lc, int mode, const u64 *a, const u64 *b, const u64 *key, u64 *out, size_t batch,
                      u32 galois, cudaStream_t st, const u32 *only, bool key_ready, const u64 *key_s) {
    if (batch == 0) return cudaSuccess;
    // Shoup companions of the key for this launch (2*L*P words, a few microseconds; batch-amortised)
    if (!key_ready) {
        const size_t n = (size_t)2 * lc.L * lc.L << lc.log_n;
        const unsigned grid = ew_grid(lc, n);
        if (lc.log_n == 12) key_prepare_kernel<12><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        else if (lc.log_n == 13) key_prepare_kernel<13><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        else key_prepare_kernel<14><<<grid, 256, 0, st>>>(key, lc.ks_key_s, lc.lp, lc.L, n);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    KsArgs A;
    A.a = a; A.b = b; A.key = key; A.key_s = key_ready && key_s ? key_s : lc.ks_key_s; A.out = out; A.scratch = lc.ks_scratch;
    A.tw = lc.tw; A.itw = lc.itw; A.L = lc.L; A.galois = galois; A.Lk = lc.L; A.hyb = nullptr; A.only = only;
    A.acc = lc.ks_acc; A.acc_par = 1; A.lift_reduce = lc.lift_reduce ? 1u : 0u;
#define KS_DISPATCH(LOGN)                                                              \
    switch (mode) {                                                                    \
        case KS_MUL_RELIN: return launch_ks_t<LOGN, KS_MUL_RELIN>(lc, A, batch, st);   \
        case KS_PLAIN: return launch_ks_t<LOGN, KS_PLAIN>(lc, A, batch, st);           \
        case KS_ROTATE: return launch_ks_t<LOGN, KS_ROTATE>(lc, A, batch, st);         \
    }                                                                                  \
    return cudaErrorInvalidValue;
    switch (lc.log_n) {
        case 12: KS_DISPATCH(12)
        case 13: KS_DISPATCH(13)
        case 14: KS_DISPATCH(14)
    }
    return cudaErrorNotSupported;
}

#define LOGN_SWITCH(call12, call13, call14)  \
    switch (lc.log_n) {                      \
        case 12: call12; break;              \
        case 13: call13; break;              \
        case 14: call14; break;              \
        default: return cudaErrorInvalidValue; \
    }

cudaError_t launch_pointwise_mul(const LaunchCtx &lc, const u64 *a, const u64 *b, u64 *out, size_t n_polys, cudaStream_t st) {
    const size_t n_chunks = n_polys * lc.L * ((size_t)1 << (lc.log_n - 1));
    if (!n_chunks) return cudaSuccess;
    const unsigned grid = ew_grid(lc, n_chunks);
    auto A = reinterpret_cast<const U64x2 *>(a), B = reinterpret_cast<const U64x2 *>(b);
    auto O = reinterpret_cast<U64x2 *>(out);
    LOGN_SWITCH((pointwise_mul_kernel<12><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)),
                (pointwise_mul_kernel<13><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)),
                (pointwise_mul_kernel<14><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)))
    return cudaGetLastError();
}

cudaError_t launch_poly_add(const LaunchCtx &lc, const u64 *a, const u64 *b, u64 *out, size_t n_polys, cudaStream_t st) {
    const size_t n_chunks = n_polys * lc.L * ((size_t)1 << (lc.log_n - 1));
    if (!n_chunks) return cudaSuccess;
    const unsigned grid = ew_grid(lc, n_chunks);
    auto A = reinterpret_cast<const U64x2 *>(a), B = reinterpret_cast<const U64x2 *>(b);
    auto O = reinterpret_cast<U64x2 *>(out);
    LOGN_SWITCH((poly_add_kernel<12><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)),
                (poly_add_kernel<13><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)),
                (poly_add_kernel<14><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)))
    return cudaGetLastError();
}

cudaError_t launch_ct_mul_plain(const LaunchCtx &lc, const u64 *ct, const u64 *pt, u64 *out, size_t batch, cudaStream_t st) {
    const size_t n_chunks = batch * 2 * lc.L * ((size_t)1 << (lc.log_n - 1));
    if (!n_chunks) return cudaSuccess;
    const unsigned grid = ew_grid(lc, n_chunks);
    auto A = reinterpret_cast<const U64x2 *>(ct), B = reinterpret_cast<const U64x2 *>(pt);
    auto O = reinterpret_cast<U64x2 *>(out);
    LOGN_SWITCH((ct_mul_plain_kernel<12><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)),
                (ct_mul_plain_kernel<13><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)),
                (ct_mul_plain_kernel<14><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)))
    return cudaGetLastError();
}

cudaError_t launch_ct_mul_plain_acc(const LaunchCtx &lc, const u64 *ct, const u64 *pt, u64 *acc, size_t batch, cudaStream_t st) {
    const size_t n_chunks = batch * 2 * lc.L * ((size_t)1 << (lc.log_n - 1));
    if (!n_chunks) return cudaSuccess;
    const unsigned grid = ew_grid(lc, n_chunks);
    auto A = reinterpret_cast<const U64x2 *>(ct), B = reinterpret_cast<const U64x2 *>(pt);
    auto O = reinterpret_cast<U64x2 *>(acc);
    LOGN_SWITCH((ct_mul_plain_acc_kernel<12><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)),
                (ct_mul_plain_acc_kernel<13><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)),
                (ct_mul_plain_acc_kernel<14><<<grid, 256, 0, st>>>(A, B, O, lc.lp, lc.L, n_chunks)))
    return cudaGetLastError();
}

cudaError_t launch_ct_tensor(const LaunchCtx &lc, const u64 *a, const u64 *b, u64 *d, size_t batch, cudaStream_t st) {
    const size_t items = batch * lc.L * ((size_t)1 << (lc.log_n - 1));
    if (!items) return cudaSuccess;
    const unsigned grid = ew_grid(lc, items);
    auto A = reinterpret_cast<const U64x2 *>(a), B = reinterpret_cast<const U64x2 *>(b);
    auto D = reinterpret_cast<U64x2 *>(d);
    LOGN_SWITCH((ct_tensor_kernel<12><<<grid, 256, 0, st>>>(A, B, D, lc.lp, lc.L, batch)),
                (ct_tensor_kernel<13><<<grid, 256, 0, st>>>(A, B, D, lc.lp, lc.L, batch)),
                (ct_tensor_kernel<14><<<grid, 256, 0, st>>>(A, B, D, lc.lp, lc.L, batch)))
    return cudaGetLastError();
}

cudaError_t launch_fill_uniform(const LaunchCtx &lc, u64 seed, u64 first_poly, u64 *data, size_t n_polys, cudaStream_t st) {
    const size_t P = (size_t)lc.L << lc.log_n;
    const size_t n_chunks = n_polys * P / 2;
    if (!n_chunks) return cudaSuccess;
    const unsigned grid = ew_grid(lc, n_chunks);
    auto O = reinterpret_cast<U64x2 *>(data);
    const u64 first_elem = first_poly * P;
    LOGN_SWITCH((fill_uniform_kernel<12><<<grid, 256, 0, st>>>(O, lc.lp, lc.L, seed, first_elem, n_chunks)),
                (fill_uniform_kernel<13><<<grid, 256, 0, st>>>(O, lc.lp, lc.L, seed, first_elem, n_chunks)),
                (fill_uniform_kernel<14><<<grid, 256, 0, st>>>(O, lc.lp, lc.L, seed, first_elem, n_chunks)))
    return cudaGetLastError();
}

#endif
}  // namespace DPFHE_VNS
}  // namespace dpfhe
