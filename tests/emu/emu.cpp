// emu.cpp — host emulator of the CTA-level kernel bodies (TEST INFRASTRUCTURE ONLY).
//
// Compiles deeppowers_b200/csrc/kernel_bodies.cuh with a sequential CTA policy so that the
// index algebra (swizzle, pass decomposition, twiddle layout, digit exchange) is checked
// against the oracle on a machine without a GPU.  It is built into tests/_emu/libdpfhe_emu.so
// by tests/conftest.py, is never linked into libdpfhe.so and is not a fallback for anything.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "host_params.hpp"
#include "kernel_bodies.cuh"

using namespace dpfhe;
using namespace dpfhe::DPFHE_VNS;   // built once per arithmetic variant (-DDPFHE_FAST=0 / 1)

namespace {
struct HostCta {
    int nt;
    template <class F>
    void par(F f) {
        for (int t = 0; t < nt; ++t) f(t);
    }
    // the emulator runs whole segments in order, so every barrier scope degenerates to "all threads"
    template <class F>
    void par_dom(F f) { par(f); }
    template <class F>
    void par_warp(F f) { par(f); }
    void mark(int) {}
    void wait_ge(const uint32_t *, uint32_t) {}   // sequential order: the readers of a slot have always finished
};

template <class T>
T *aligned_new(size_t n) {
    void *p = nullptr;
    if (posix_memalign(&p, 128, n * sizeof(T))) return nullptr;
    return (T *)p;
}

struct Emu {
    HostParams hp;
    std::vector<LimbParams> lp;
    uint32_t lift_reduce = 1;   // as abi.cu: 0 when every modulus is below twice every other one
    Twiddle *tw = nullptr, *itw = nullptr;
    ~Emu() {
        free(tw);
        free(itw);
    }
};

template <int LOGN, int NT>
void run_ntt(Emu &e, uint64_t *data, size_t n_polys, bool inverse) {
    const size_t N = (size_t)1 << LOGN;
    uint64_t *buf = aligned_new<uint64_t>(N);
    HostCta cta{NT};
    for (size_t w = 0; w < n_polys * e.hp.L; ++w) {
        const unsigned l = (unsigned)(w % e.hp.L);
        const LimbParams p = e.lp[l];
        if (inverse) ntt_inv_body<LOGN, NT>(cta, buf, data + w * N, e.itw + l * N, p);
        else ntt_fwd_body<LOGN, NT>(cta, buf, data + w * N, e.tw + l * N, p);
    }
    free(buf);
}

// N = 16384 by a pair of CTAs (ntt_pair_kernel): two half buffers; the cluster barriers become the order of the phases
template <int NT>
void run_ntt_pair(Emu &e, uint64_t *data, size_t n_polys, bool inverse) {
    const size_t N = (size_t)1 << NTT_PAIR_LOGN;
    uint64_t *buf[2] = {aligned_new<uint64_t>(N / 2), aligned_new<uint64_t>(N / 2)};
    HostCta cta{NT};
    for (size_t w = 0; w < n_polys * e.hp.L; ++w) {
        const unsigned l = (unsigned)(w % e.hp.L);
        const LimbParams p = e.lp[l];
        uint64_t *limb = data + w * N;
        if (!inverse) {
            for (int h = 0; h < 2; ++h) ntt_fwd_half_load<NT>(cta, buf[h], limb, e.tw + l * N, p, h);
            for (int h = 0; h < 2; ++h) ntt_fwd_half_finish<NT>(cta, buf[h], limb, e.tw + l * N, p, h);
        } else {
            for (int h = 0; h < 2; ++h) ntt_inv_half_passes<NT>(cta, buf[h], limb, e.itw + l * N, p, h);
            for (int h = 0; h < 2; ++h) ntt_inv_half_outer<NT>(cta, buf[h], buf[h ^ 1], limb, e.itw + l * N, p, h);
        }
    }
    free(buf[0]);
    free(buf[1]);
}

// persistent-grid emulation: G slots, rounds of G work items; all phase 1 of a round before its phase 2
template <int LOGN, int NT, int MODE>
void run_ks(Emu &e, const uint64_t *a, const uint64_t *b, const uint64_t *key, uint64_t *out, size_t batch,
            uint32_t galois, unsigned G) {
    const size_t N = (size_t)1 << LOGN;
    const unsigned L = e.hp.L;
    G = (G / L) * L;
    if (G == 0) G = L;
    std::vector<uint64_t *> buf(G);
    for (unsigned s = 0; s < G; ++s) {
        buf[s] = aligned_new<uint64_t>(N);
    }
    uint64_t *scratch = aligned_new<uint64_t>((size_t)G * 2 * N), *acc = aligned_new<uint64_t>((size_t)G * 2 * N);
    // Shoup companions of the key (the device builds them with key_prepare_kernel)
    const size_t key_words = (size_t)2 * L * L * N;
    uint64_t *key_s = aligned_new<uint64_t>(key_words);
    for (size_t k = 0; k < key_words; ++k)
        key_s[k] = (uint64_t)((((unsigned __int128)key[k]) << 64) / e.lp[(k / N) % L].q);
    KsArgs A;
    A.a = a; A.b = b; A.key = key; A.key_s = key_s; A.out = out; A.scratch = scratch;
    A.tw = e.tw; A.itw = e.itw; A.L = L; A.galois = galois; A.Lk = L; A.hyb = nullptr; A.only = nullptr;
    A.acc = acc; A.acc_par = 1; A.lift_reduce = e.lift_reduce;
    HostCta cta{NT};
    const size_t n_work = batch * L;
    for (size_t r = 0; r * G < n_work; ++r) {
        const unsigned par = (unsigned)(r & 1);
        for (unsigned s = 0; s < G; ++s) {
            const size_t w = r * G + s;
            if (w >= n_work) break;
            ks_phase1<LOGN, NT, MODE>(cta, buf[s], A, e.lp[w % L], w / L, (uint32_t)(w % L), scratch + ((size_t)s * 2 + par) * N, acc + (size_t)s * 2 * N);
        }
        for (unsigned s = 0; s < G; ++s) {
            const size_t w = r * G + s;
            if (w >= n_work) break;
            const uint32_t i = (uint32_t)(w % L);
            for (uint32_t jj = 1; jj < L; ++jj) {
                const uint32_t j = (i + jj) % L;
                const unsigned sib = s - i + j;
                ks_phase2_digit<LOGN, NT>(cta, buf[s], A, e.lp[i], w / L, i, j, jj, scratch + ((size_t)sib * 2 + par) * N, acc + (size_t)s * 2 * N);
            }
        }
    }
    for (unsigned s = 0; s < G; ++s) {
        free(buf[s]);
    }
    free(scratch);
    free(acc);
    free(key_s);
}

// hybrid key switching: groups of L+1 slots (L ciphertext limbs + the special limb), the same role programs as
// ks_hybrid_kernel, run in dependency order (all digits, then the multiply-accumulates, tau', the final division)
template <int LOGN, int NT, int MODE>
void run_ks_hybrid(Emu &e, const uint64_t *a, const uint64_t *b, const uint64_t *key, uint64_t *out, size_t batch,
                   uint32_t galois, uint64_t t_plain, unsigned G) {
    const size_t N = (size_t)1 << LOGN;
    const unsigned LK = e.hp.L, L = LK - 1, GS = LK;
    unsigned groups = G / GS;
    if (groups == 0) groups = 1;
    uint64_t *buf = aligned_new<uint64_t>(N);
    uint64_t *scratch = aligned_new<uint64_t>((size_t)groups * GS * 2 * N);
    uint64_t *hyb_all = aligned_new<uint64_t>((size_t)groups * KS_HYB_ROWS * N);
    uint64_t *acc = aligned_new<uint64_t>((size_t)groups * GS * 2 * 2 * N);   // [slot][parity][2][N]
    const size_t key_words = (size_t)2 * L * LK * N;
    uint64_t *key_s = aligned_new<uint64_t>(key_words);
    for (size_t k = 0; k < key_words; ++k)
        key_s[k] = (uint64_t)((((unsigned __int128)key[k]) << 64) / e.lp[(k / N) % LK].q);
    MsConsts K;
    build_ms_consts(e.hp, t_plain, K);
    KsArgs A;
    A.a = a; A.b = b; A.key = key; A.key_s = key_s; A.out = out; A.scratch = scratch;
    A.tw = e.tw; A.itw = e.itw; A.L = L; A.galois = galois; A.Lk = LK; A.hyb = hyb_all; A.only = nullptr;
    A.acc = acc; A.acc_par = 2; A.lift_reduce = e.lift_reduce;
    auto acc_of = [&](unsigned slot, unsigned parity) { return acc + ((size_t)slot * 2 + parity) * 2 * N; };
    HostCta cta{NT};
    for (size_t r = 0; r * groups < batch; ++r) {
        const unsigned par = (unsigned)(r & 1);
        for (unsigned g = 0; g < groups; ++g) {
            const size_t ct = r * groups + g;
            if (ct >= batch) break;
            const unsigned base = g * GS;
            uint64_t *hyb = hyb_all + (size_t)g * KS_HYB_ROWS * N;
            for (unsigned i = 0; i < L; ++i)
                ks_phase1<LOGN, NT, MODE, true>(cta, buf, A, e.lp[i], ct, i, scratch + ((size_t)(base + i) * 2 + par) * N, acc_of(base + i, par), K.qlm[i], K.qlm_s[i]);
            for (unsigned i = 0; i < L; ++i)
                for (uint32_t jj = 1; jj < L; ++jj) {
                    const uint32_t j = (i + jj) % L;
                    ks_phase2_digit<LOGN, NT, true, false>(cta, buf, A, e.lp[i], ct, i, j, jj, scratch + ((size_t)(base + j) * 2 + par) * N, acc_of(base + i, par));
                }
            for (uint32_t jj = 0; jj < L; ++jj) {
                const uint32_t j = (g + jj) % L;
                ks_phase2_digit<LOGN, NT, true, true>(cta, buf, A, e.lp[L], ct, L, j, jj, scratch + ((size_t)(base + j) * 2 + par) * N, hyb);
            }
            for (unsigned c = 0; c < 2; ++c)
                ms_tau_body<LOGN, NT, true>(cta, buf, hyb + c * N, hyb + c * N, A.itw + (size_t)L * N, e.lp[L], hyb + ks_hyb_tau_row(par, c) * N, K);
            const size_t P = (size_t)L * N;
            for (unsigned i = 0; i < L; ++i)
                for (unsigned c = 0; c < 2; ++c) {
                    uint64_t *row = out + ct * 2 * P + c * P + (size_t)i * N;
                    ms_limb_body<LOGN, NT, true>(cta, buf, hyb + ks_hyb_tau_row(par, c) * N, acc_of(base + i, par) + c * N, row, A.tw + (size_t)i * N, e.lp[i], K, i);
                }
        }
    }
    free(buf);
    free(scratch);
    free(hyb_all);
    free(acc);
    free(key_s);
}

// grouped hybrid key switching: groups of Lq + K slots, the role programs of ks_grouped_kernel in dependency order
template <int LOGN, int NT, int MODE>
void run_ks_grouped(Emu &e, unsigned Ks, const uint64_t *a, const uint64_t *b, const uint64_t *key, uint64_t *out, size_t batch,
                    uint32_t galois, uint64_t t_plain, unsigned G) {
    const size_t N = (size_t)1 << LOGN;
    const unsigned LK = e.hp.L, Lq = LK - Ks, GS = LK;
    unsigned groups = G / GS;
    if (groups == 0) groups = 1;
    MsConsts K;
    GroupConsts Gc;
    build_group_consts(e.hp, Ks, t_plain, Gc, K);
    const unsigned dnum = Gc.dnum;
    uint64_t *buf = aligned_new<uint64_t>(N);
    uint64_t *scratch = aligned_new<uint64_t>((size_t)groups * GS * 2 * N);
    uint64_t *hyb_all = aligned_new<uint64_t>((size_t)groups * Ks * KS_HYB_ROWS * N);
    uint64_t *acc = aligned_new<uint64_t>((size_t)groups * GS * 2 * 2 * N);   // [slot][parity][2][N]
    const size_t key_words = (size_t)2 * dnum * LK * N;
    uint64_t *key_s = aligned_new<uint64_t>(key_words);
    for (size_t k = 0; k < key_words; ++k)
        key_s[k] = (uint64_t)((((unsigned __int128)key[k]) << 64) / e.lp[(k / N) % LK].q);
    KsArgs A;
    A.a = a; A.b = b; A.key = key; A.key_s = key_s; A.out = out; A.scratch = scratch;
    A.tw = e.tw; A.itw = e.itw; A.L = Lq; A.galois = galois; A.Lk = LK; A.hyb = hyb_all; A.only = nullptr;
    A.acc = acc; A.acc_par = 2; A.lift_reduce = 0;
    auto acc_of = [&](unsigned slot, unsigned parity) { return acc + ((size_t)slot * 2 + parity) * 2 * N; };
    HostCta cta{NT};
    for (size_t r = 0; r * groups < batch; ++r) {
        const unsigned par = (unsigned)(r & 1);
        for (unsigned g = 0; g < groups; ++g) {
            const size_t ct = r * groups + g;
            if (ct >= batch) break;
            const unsigned base = g * GS;
            auto hyb_of = [&](unsigned k) { return hyb_all + ((size_t)g * Ks + k) * KS_HYB_ROWS * N; };
            const uint64_t *t_rows = scratch + ((size_t)base * 2 + par) * N;
            for (unsigned i = 0; i < Lq; ++i)
                ks_phase1<LOGN, NT, MODE, true>(cta, buf, A, Gc.lp_up[i], ct, i, scratch + ((size_t)(base + i) * 2 + par) * N, acc_of(base + i, par),
                                                K.qlm[i], K.qlm_s[i], nullptr, 0, i / Ks);
            for (unsigned i = 0; i < Lq; ++i)
                for (uint32_t jj = 1; jj < dnum; ++jj)
                    ks_phase2_group<LOGN, NT, false>(cta, buf, A, Gc, e.lp[i], ct, i, (i / Ks + jj) % dnum, jj, t_rows, 2 * N, acc_of(base + i, par));
            for (unsigned k = 0; k < Ks; ++k) {
                const unsigned i = Lq + k;
                uint64_t *hyb = hyb_of(k);
                for (uint32_t jj = 0; jj < dnum; ++jj)
                    ks_phase2_group<LOGN, NT, true>(cta, buf, A, Gc, e.lp[i], ct, i, (g + jj) % dnum, jj, t_rows, 2 * N, hyb);
                for (unsigned c = 0; c < 2; ++c)
                    ms_tau_body<LOGN, NT, true>(cta, buf, hyb + c * N, hyb + c * N, A.itw + (size_t)i * N, Gc.lp_up[i], hyb + ks_hyb_tau_row(par, c) * N, K);
            }
            const size_t P = (size_t)Lq * N;
            for (unsigned i = 0; i < Lq; ++i)
                for (unsigned c = 0; c < 2; ++c) {
                    uint64_t *row = out + ct * 2 * P + c * P + (size_t)i * N;
                    ms_limb_group<LOGN, NT>(cta, buf, hyb_of(0) + ks_hyb_tau_row(par, c) * N, (size_t)KS_HYB_ROWS * N, acc_of(base + i, par) + c * N, row,
                                            A.tw + (size_t)i * N, e.lp[i], K, Gc, i);
                }
        }
    }
    free(buf);
    free(scratch);
    free(hyb_all);
    free(acc);
    free(key_s);
}

// hoisted rotations with grouped hybrid keys: hoistg_phase1/2 in the role order of ks_hoistg_kernel, then per rotation the rows of
// rot_apply_grouped_kernel and the division by P (md_tau / md_limb bodies)
template <int LOGN, int NT, int NT_MD>
void run_rotate_hoisted_grouped(Emu &e, unsigned Ks, const uint64_t *ct, size_t n_rot, const uint64_t *galois, const uint64_t *keys, uint64_t *out,
                                size_t batch, uint64_t t_plain) {
    const size_t N = (size_t)1 << LOGN;
    const unsigned L = e.hp.L, Lq = L - Ks;
    MsConsts K;
    GroupConsts G;
    build_group_consts(e.hp, Ks, t_plain, G, K);
    const unsigned dnum = G.dnum;
    const size_t P = (size_t)L * N, key_words = (size_t)2 * dnum * L * N;
    uint64_t *buf = aligned_new<uint64_t>(N), *scratch = aligned_new<uint64_t>((size_t)L * 2 * N);
    uint64_t *U = aligned_new<uint64_t>(batch * dnum * L * N), *acc = aligned_new<uint64_t>(batch * 2 * P), *tau = aligned_new<uint64_t>((size_t)Ks * N);
    uint64_t *key_s = aligned_new<uint64_t>(key_words);
    HoistGArgs H;
    H.ct = ct; H.U = U; H.scratch = scratch; H.tw = e.tw; H.itw = e.itw;
    HostCta cta{NT}, cta_md{NT_MD};
    for (size_t c = 0; c < batch; ++c) {
        const unsigned par = (unsigned)(c & 1);
        const uint64_t *t_rows = scratch + (size_t)par * N;
        for (unsigned i = 0; i < Lq; ++i) hoistg_phase1<LOGN, NT>(cta, buf, H, G, c, i, scratch + ((size_t)i * 2 + par) * N);
        for (unsigned i = 0; i < L; ++i)
            for (unsigned g = 0; g < dnum; ++g)
                if (i >= Lq || i / Ks != g) hoistg_phase2<LOGN, NT>(cta, buf, H, G, e.lp[i], c, i, g, t_rows, 2 * N);
    }
    for (size_t r = 0; r < n_rot; ++r) {
        const uint64_t *key = keys + r * key_words;
        for (size_t k = 0; k < key_words; ++k) key_s[k] = (uint64_t)((((unsigned __int128)key[k]) << 64) / e.lp[(k / N) % L].q);
        RotApplyGArgs A;
        A.ct = ct; A.U = U; A.key = key; A.key_s = key_s; A.acc = acc; A.galois = (uint32_t)galois[r];
        for (size_t c0 = 0; c0 < batch; c0 += 2)
            for (unsigned i = 0; i < L; ++i)
                rot_apply_grouped_rows<LOGN, NT, 2>(cta, A, G, K, e.lp[i], c0, (uint32_t)(batch - c0 < 2 ? batch - c0 : 2), i);
        for (size_t w = 0; w < 2 * batch; ++w) {   // the division by P, one polynomial at a time
            for (unsigned k = 0; k < Ks; ++k)
                ms_tau_body<LOGN, NT_MD>(cta_md, buf, acc + (w * L + Lq + k) * N, nullptr, e.itw + (size_t)(Lq + k) * N, G.lp_up[Lq + k], tau + (size_t)k * N, K);
            for (unsigned i = 0; i < Lq; ++i)
                ms_limb_group<LOGN, NT_MD, false>(cta_md, buf, tau, N, acc + (w * L + i) * N, out + ((r * batch * 2 + w) * Lq + i) * N, e.tw + (size_t)i * N,
                                               e.lp[i], K, G, i);
        }
    }
    free(buf); free(scratch); free(U); free(acc); free(tau); free(key_s);
}

// hoisted rotations: the device bodies (hoist_phase1/2, rot_apply_row) in kernel order, the per-rotation constants
// computed the way launch_rot_prepare does (negmask -> NTT -> kprime), flagged ciphertexts through the ordinary rotate
template <int LOGN, int NT>
void run_rotate_hoisted(Emu &e, const uint64_t *ct, size_t n_rot, const uint64_t *galois, const uint64_t *keys, uint64_t *out,
                        size_t batch, unsigned G, unsigned *n_flagged) {
    const size_t N = (size_t)1 << LOGN;
    const unsigned L = e.hp.L;
    const size_t P = (size_t)L * N, key_words = (size_t)2 * L * L * N;
    G = (G / L) * L;
    if (G == 0) G = L;
    uint64_t *buf = aligned_new<uint64_t>(N);
    uint64_t *scratch = aligned_new<uint64_t>((size_t)G * 2 * N);
    uint64_t *U = aligned_new<uint64_t>(batch * L * L * N);
    std::vector<uint32_t> zero(batch, 0);
    HostCta cta{NT};
    HoistArgs H;
    H.ct = ct; H.U = U; H.scratch = scratch; H.zero = zero.data(); H.tw = e.tw; H.itw = e.itw; H.L = L;
    if (L > 1) {
        const size_t n_work = batch * L;
        for (size_t r = 0; r * G < n_work; ++r) {
            const unsigned par = (unsigned)(r & 1);
            for (unsigned s = 0; s < G && r * G + s < n_work; ++s) {
                const size_t w = r * G + s;
                hoist_phase1<LOGN, NT>(cta, buf, H, e.lp[w % L], w / L, (uint32_t)(w % L), scratch + ((size_t)s * 2 + par) * N);
            }
            for (unsigned s = 0; s < G && r * G + s < n_work; ++s) {
                const size_t w = r * G + s;
                const uint32_t i = (uint32_t)(w % L);
                for (uint32_t jj = 1; jj < L; ++jj) {
                    const uint32_t j = (i + jj) % L;
                    hoist_phase2<LOGN, NT>(cta, buf, H, e.lp[i], w / L, i, j, scratch + ((size_t)(s - i + j) * 2 + par) * N);
                }
            }
        }
    }
    *n_flagged = 0;
    for (size_t k = 0; k < batch; ++k) *n_flagged += zero[k] ? 1u : 0u;
    uint64_t *key_s = aligned_new<uint64_t>(key_words), *M = aligned_new<uint64_t>(P), *kprime = aligned_new<uint64_t>(2 * P);
    for (size_t r = 0; r < n_rot; ++r) {
        const uint64_t *key = keys + r * key_words;
        const uint32_t g = (uint32_t)galois[r];
        for (size_t k = 0; k < key_words; ++k) key_s[k] = (uint64_t)((((unsigned __int128)key[k]) << 64) / e.lp[(k / N) % L].q);
        for (uint32_t k = 0; k < N; ++k) {
            const uint32_t ex = (k * g) & (uint32_t)(2 * N - 1);
            for (unsigned l = 0; l < L; ++l) M[(size_t)l * N + (ex & (N - 1))] = ex >= N ? 1 : 0;
        }
        for (unsigned l = 0; l < L; ++l) ntt_fwd_body<LOGN, NT>(cta, buf, M + (size_t)l * N, e.tw + (size_t)l * N, e.lp[l]);
        for (unsigned c = 0; c < 2; ++c)
            for (unsigned i = 0; i < L; ++i)
                for (size_t n = 0; n < N; ++n) {
                    const uint64_t q = e.lp[i].q;
                    uint64_t s = 0;
                    for (unsigned j = 0; j < L; ++j)
                        if (j != i) s = (s + host_mulmod(key[((size_t)j * 2 + c) * P + (size_t)i * N + n], e.lp[j].q % q, q)) % q;
                    kprime[(size_t)c * P + (size_t)i * N + n] = host_mulmod(s, M[(size_t)i * N + n], q);
                }
        RotApplyArgs R;
        R.ct = ct; R.U = L > 1 ? U : nullptr; R.key = key; R.key_s = key_s; R.kprime = kprime; R.out = out + r * batch * 2 * P; R.L = L; R.galois = g;
        for (size_t k = 0; k < batch; k += 2)   // blocks of two ciphertexts, the last one ragged; odd rotations use the prefetching form
            for (unsigned i = 0; i < L; ++i) {
                const unsigned n_ct = (unsigned)(batch - k < 2 ? batch - k : 2);
                if (r & 1) rot_apply_rows<LOGN, NT, 2, true>(cta, R, e.lp[i], k, n_ct, i);
                else rot_apply_rows<LOGN, NT, 2, false>(cta, R, e.lp[i], k, n_ct, i, 0, 1 << (LOGN - 2)),
                     rot_apply_rows<LOGN, NT, 2, false>(cta, R, e.lp[i], k, n_ct, i, 1 << (LOGN - 2), 1 << (LOGN - 1));
            }
        for (size_t k = 0; k < batch; ++k)   // flagged ciphertexts: ordinary rotate, as the device does with its filter
            if (zero[k]) run_ks<LOGN, NT, KS_ROTATE>(e, ct + k * 2 * P, ct + k * 2 * P, key, out + (r * batch + k) * 2 * P, 1, g, L);
    }
    free(buf); free(scratch); free(U); free(key_s); free(M); free(kprime);
}
}  // namespace

extern "C" {

void *emu_create(unsigned log_n, unsigned L, const uint64_t *moduli) {
    Emu *e = new Emu();
    if (!build_host_params(log_n, L, moduli, e->hp).empty()) {
        delete e;
        return nullptr;
    }
#if DPFHE_FAST
    for (unsigned l = 0; l < L; ++l)   // the fast bodies are only valid for moduli k * 2^32 + 1
        if (e->hp.limbs[l].lp.nqh == 0) {
            delete e;
            return nullptr;
        }
#endif
    const size_t N = (size_t)1 << log_n;
    e->tw = aligned_new<Twiddle>(N * L);
    e->itw = aligned_new<Twiddle>(N * L);
    for (unsigned l = 0; l < L; ++l) {
        e->lp.push_back(e->hp.limbs[l].lp);
        memcpy(e->tw + l * N, e->hp.limbs[l].tw.data(), N * sizeof(Twiddle));
        memcpy(e->itw + l * N, e->hp.limbs[l].itw.data(), N * sizeof(Twiddle));
    }
    uint64_t qmin = ~0ull, qmax = 0;
    for (unsigned l = 0; l < L; ++l) {
        qmin = e->lp[l].q < qmin ? e->lp[l].q : qmin;
        qmax = e->lp[l].q > qmax ? e->lp[l].q : qmax;
    }
    e->lift_reduce = qmax < 2 * qmin ? 0u : 1u;
    return e;
}
void emu_destroy(void *h) { delete (Emu *)h; }
uint64_t emu_modulus(void *h, unsigned l) { return ((Emu *)h)->hp.limbs[l].lp.q; }
uint64_t emu_psi(void *h, unsigned l) { return ((Emu *)h)->hp.limbs[l].psi; }
void emu_root_powers(void *h, unsigned l, int inverse, uint64_t *out) {
    Emu *e = (Emu *)h;
    const auto &v = inverse ? e->hp.limbs[l].inv_root_powers : e->hp.limbs[l].root_powers;
    memcpy(out, v.data(), v.size() * 8);
}

int emu_ntt(void *h, uint64_t *data, size_t n_polys, int inverse) {
    Emu *e = (Emu *)h;
    switch (e->hp.log_n) {
        case 12: run_ntt<12, 256>(*e, data, n_polys, inverse != 0); return 0;
        case 13: run_ntt<13, 512>(*e, data, n_polys, inverse != 0); return 0;
        case 14: run_ntt<14, 512>(*e, data, n_polys, inverse != 0); return 0;
    }
    return -1;
}

// the CTA-pair form of the N = 16384 transforms
int emu_ntt_pair(void *h, uint64_t *data, size_t n_polys, int inverse) {
    Emu *e = (Emu *)h;
    if (e->hp.log_n != 14) return -1;
    run_ntt_pair<256>(*e, data, n_polys, inverse != 0);
    return 0;
}

// mode: 0 ct_mul_relin (a,b), 1 keyswitch (a = d), 2 rotate (a = ct, galois)
int emu_ks(void *h, int mode, const uint64_t *a, const uint64_t *b, const uint64_t *key, uint64_t *out, size_t batch,
           uint32_t galois, unsigned G) {
    Emu *e = (Emu *)h;
#define DISPATCH(LOGN, NT)                                                                       \
    if (mode == 0) run_ks<LOGN, NT, KS_MUL_RELIN>(*e, a, b, key, out, batch, galois, G);          \
    else if (mode == 1) run_ks<LOGN, NT, KS_PLAIN>(*e, a, b, key, out, batch, galois, G);         \
    else run_ks<LOGN, NT, KS_ROTATE>(*e, a, b, key, out, batch, galois, G);                       \
    return 0;
    switch (e->hp.log_n) {
        case 12: DISPATCH(12, 256)
        case 13: DISPATCH(13, 256)
        case 14: DISPATCH(14, 256)
    }
    return -1;
}

// hybrid variants of emu_ks: the context's last limb is the special prime, data has L-1 limbs
int emu_ks_hybrid(void *h, int mode, const uint64_t *a, const uint64_t *b, const uint64_t *key, uint64_t *out, size_t batch,
                  uint32_t galois, uint64_t t_plain, unsigned G) {
    Emu *e = (Emu *)h;
    if (e->hp.L < 2) return -1;
#define DISPATCH_H(LOGN, NT)                                                                               \
    if (mode == 0) run_ks_hybrid<LOGN, NT, KS_MUL_RELIN>(*e, a, b, key, out, batch, galois, t_plain, G);    \
    else if (mode == 1) run_ks_hybrid<LOGN, NT, KS_PLAIN>(*e, a, b, key, out, batch, galois, t_plain, G);   \
    else run_ks_hybrid<LOGN, NT, KS_ROTATE>(*e, a, b, key, out, batch, galois, t_plain, G);                 \
    return 0;
    switch (e->hp.log_n) {
        case 12: DISPATCH_H(12, 256)
        case 13: DISPATCH_H(13, 256)
        case 14: DISPATCH_H(14, 256)
    }
    return -1;
}

// grouped hybrid variants: the context's last K limbs are special primes, data has L-K limbs, keys ceil((L-K)/K) digits
int emu_ks_grouped(void *h, unsigned K, int mode, const uint64_t *a, const uint64_t *b, const uint64_t *key, uint64_t *out, size_t batch,
                   uint32_t galois, uint64_t t_plain, unsigned G) {
    Emu *e = (Emu *)h;
    if (K < 1 || K > (unsigned)KS_MAX_SPECIAL || 2 * K > e->hp.L) return -1;
#define DISPATCH_G(LOGN, NT)                                                                                   \
    if (mode == 0) run_ks_grouped<LOGN, NT, KS_MUL_RELIN>(*e, K, a, b, key, out, batch, galois, t_plain, G);    \
    else if (mode == 1) run_ks_grouped<LOGN, NT, KS_PLAIN>(*e, K, a, b, key, out, batch, galois, t_plain, G);   \
    else run_ks_grouped<LOGN, NT, KS_ROTATE>(*e, K, a, b, key, out, batch, galois, t_plain, G);                 \
    return 0;
    switch (e->hp.log_n) {
        case 12: DISPATCH_G(12, 256)
        case 13: DISPATCH_G(13, 256)
        case 14: DISPATCH_G(14, 256)
    }
    return -1;
}

int emu_rotate_hoisted_grouped(void *h, unsigned K, const uint64_t *ct, size_t n_rot, const uint64_t *galois, const uint64_t *keys, uint64_t *out,
                               size_t batch, uint64_t t_plain) {
    Emu *e = (Emu *)h;
    if (K < 1 || K > (unsigned)KS_MAX_SPECIAL || 2 * K > e->hp.L) return -1;
    switch (e->hp.log_n) {
        case 12: run_rotate_hoisted_grouped<12, 256, 256>(*e, K, ct, n_rot, galois, keys, out, batch, t_plain); return 0;
        case 13: run_rotate_hoisted_grouped<13, 256, 256>(*e, K, ct, n_rot, galois, keys, out, batch, t_plain); return 0;
        case 14: run_rotate_hoisted_grouped<14, 256, 512>(*e, K, ct, n_rot, galois, keys, out, batch, t_plain); return 0;
    }
    return -1;
}

int emu_rotate_hoisted(void *h, const uint64_t *ct, size_t n_rot, const uint64_t *galois, const uint64_t *keys, uint64_t *out, size_t batch,
                       unsigned G, unsigned *n_flagged) {
    Emu *e = (Emu *)h;
    switch (e->hp.log_n) {
        case 12: run_rotate_hoisted<12, 256>(*e, ct, n_rot, galois, keys, out, batch, G, n_flagged); return 0;
        case 13: run_rotate_hoisted<13, 256>(*e, ct, n_rot, galois, keys, out, batch, G, n_flagged); return 0;
        case 14: run_rotate_hoisted<14, 256>(*e, ct, n_rot, galois, keys, out, batch, G, n_flagged); return 0;
    }
    return -1;
}

// plaintext inner products: the device tile body over all (limb, tile) pairs, g-blocks of `gmax` giant steps
int emu_pt_inner(void *h, const uint64_t *steps, unsigned nb, const uint64_t *pts, unsigned ng, uint64_t *out, size_t batch, unsigned gmax) {
    Emu *e = (Emu *)h;
    PtInnerArgs A;
    A.steps = steps; A.pts = pts; A.out = out; A.batch = batch; A.L = e->hp.L; A.nb = nb; A.ng = ng;
    if (gmax == 0) gmax = ng;
    uint64_t *smem = aligned_new<uint64_t>(((size_t)gmax + 4) * nb * PTI_COEFFS);
    auto run = [&](auto logn_tag) {
        constexpr int LOGN = decltype(logn_tag)::value;
        HostCta cta{256};
        const unsigned tiles = (1u << LOGN) / PTI_COEFFS;
        for (unsigned g0 = 0; g0 < ng; g0 += gmax)
            for (unsigned l = 0; l < e->hp.L; ++l)
                for (unsigned t = 0; t < tiles; ++t)
                    pt_inner_tile<LOGN, 256>(cta, smem, A, e->lp[l], l, t, g0, ng - g0 < gmax ? ng - g0 : gmax);
    };
    int rc = 0;
    switch (e->hp.log_n) {
        case 12: run(std::integral_constant<int, 12>{}); break;
        case 13: run(std::integral_constant<int, 13>{}); break;
        case 14: run(std::integral_constant<int, 14>{}); break;
        default: rc = -1;
    }
    free(smem);
    return rc;
}

// division by the product of the last Ks limbs through the bodies of md_tau_kernel / md_limb_kernel
int emu_mod_down_special(void *h, unsigned Ks, const uint64_t *in, uint64_t *out, size_t n_polys, uint64_t t_plain) {
    Emu *e = (Emu *)h;
    const unsigned L = e->hp.L;
    const size_t N = (size_t)1 << e->hp.log_n;
    if (Ks < 1 || Ks > (unsigned)KS_MAX_SPECIAL || Ks >= L) return -1;
    const unsigned Lq = L - Ks;
    MsConsts K;
    GroupConsts G;
    build_group_consts(e->hp, Ks, t_plain, G, K);
    uint64_t *buf = aligned_new<uint64_t>(N), *tau = aligned_new<uint64_t>((size_t)Ks * N);
    auto run = [&](auto logn_tag, auto nt_tag) {
        constexpr int LOGN = decltype(logn_tag)::value, NT = decltype(nt_tag)::value;
        HostCta cta{NT};
        for (size_t w = 0; w < n_polys; ++w) {
            for (unsigned k = 0; k < Ks; ++k)
                ms_tau_body<LOGN, NT>(cta, buf, in + (w * L + Lq + k) * N, nullptr, e->itw + (size_t)(Lq + k) * N, G.lp_up[Lq + k], tau + (size_t)k * N, K);
            for (unsigned i = 0; i < Lq; ++i)
                ms_limb_group<LOGN, NT, false>(cta, buf, tau, N, in + (w * L + i) * N, out + (w * Lq + i) * N, e->tw + (size_t)i * N, e->lp[i], K, G, i);
        }
    };
    int rc = 0;
    switch (e->hp.log_n) {
        case 12: run(std::integral_constant<int, 12>{}, std::integral_constant<int, 256>{}); break;
        case 13: run(std::integral_constant<int, 13>{}, std::integral_constant<int, 256>{}); break;
        case 14: run(std::integral_constant<int, 14>{}, std::integral_constant<int, 512>{}); break;
        default: rc = -1;
    }
    free(buf);
    free(tau);
    return rc;
}

// modulus switching through the same bodies the device runs
int emu_mod_switch(void *h, const uint64_t *in, uint64_t *out, size_t n_polys, uint64_t t_plain) {
    Emu *e = (Emu *)h;
    const unsigned L = e->hp.L;
    const size_t N = (size_t)1 << e->hp.log_n;
    if (L < 2) return -1;
    MsConsts K;
    build_ms_consts(e->hp, t_plain, K);
    uint64_t *buf = aligned_new<uint64_t>(N), *tau = aligned_new<uint64_t>(N);
    auto run = [&](auto logn_tag, auto nt_tag) {
        constexpr int LOGN = decltype(logn_tag)::value, NT = decltype(nt_tag)::value;
        HostCta cta{NT};
        for (size_t w = 0; w < n_polys; ++w) {
            ms_tau_body<LOGN, NT>(cta, buf, in + (w * L + (L - 1)) * N, nullptr, e->itw + (size_t)(L - 1) * N, e->lp[L - 1], tau, K);
            for (unsigned i = 0; i + 1 < L; ++i)
                ms_limb_body<LOGN, NT>(cta, buf, tau, in + (w * L + i) * N, out + (w * (L - 1) + i) * N, e->tw + (size_t)i * N, e->lp[i], K, i);
        }
    };
    int rc = 0;
    switch (e->hp.log_n) {
        case 12: run(std::integral_constant<int, 12>{}, std::integral_constant<int, 256>{}); break;
        case 13: run(std::integral_constant<int, 13>{}, std::integral_constant<int, 256>{}); break;
        case 14: run(std::integral_constant<int, 14>{}, std::integral_constant<int, 512>{}); break;
        default: rc = -1;
    }
    free(buf);
    free(tau);
    return rc;
}

// scalar checks
uint64_t emu_mulmod(void *h, unsigned l, uint64_t a, uint64_t b) { return mulmod(a, b, ((Emu *)h)->lp[l]); }
uint64_t emu_word_reduce(void *h, unsigned l, uint64_t x) { return word_reduce(x, ((Emu *)h)->lp[l]); }
uint64_t emu_canon(void *h, unsigned l, uint64_t x) { return canon(x, ((Emu *)h)->lp[l]); }
// the one-subtraction form where it applies (moduli above 2^64 / 17), else the general one: what ntt_fwd_body's store loop runs
uint64_t emu_canon_store(void *h, unsigned l, uint64_t x) { return canon_store(x, ((Emu *)h)->lp[l]); }
uint64_t emu_mulmod_lazy(void *h, unsigned l, uint64_t a, uint64_t b) { return mulmod_lazy(a, b, ((Emu *)h)->lp[l]); }
// x * w through the Shoup forms (w < q; the companion is derived here)
uint64_t emu_shoup_lazy(void *h, unsigned l, uint64_t x, uint64_t w) {
    const LimbParams &p = ((Emu *)h)->lp[l];
    return shoup_lazy(x, w, (uint64_t)((((unsigned __int128)w) << 64) / p.q), p);
}
uint64_t emu_shoup_exact(void *h, unsigned l, uint64_t x, uint64_t w) {
    const LimbParams &p = ((Emu *)h)->lp[l];
    return shoup_exact(x, w, (uint64_t)((((unsigned __int128)w) << 64) / p.q), p);
}
// reductions of a two-word value z = hi:lo
uint64_t emu_barrett_long(void *h, unsigned l, uint64_t hi, uint64_t lo) { return barrett_lazy_long(hi, lo, ((Emu *)h)->lp[l]); }
uint64_t emu_pti_fold(void *h, unsigned l, uint64_t a0, uint64_t a1a, uint64_t a1b, uint64_t a2) { return pti_fold(a0, a1a, a1b, a2, ((Emu *)h)->lp[l]); }
}
