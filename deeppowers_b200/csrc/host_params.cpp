// host_params.cpp — see host_params.hpp.  Plain C++17, compiled into libdpfhe.so.
#include "host_params.hpp"

namespace dpfhe {

typedef unsigned __int128 u128;

uint64_t host_mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)((u128)a * b % q); }

uint64_t host_powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q, base = a % q;
    for (; e; e >>= 1) {
        if (e & 1) r = host_mulmod(r, base, q);
        base = host_mulmod(base, base, q);
    }
    return r;
}

// Miller-Rabin, exact for n < 2^64 with the first twelve prime witnesses.
bool host_is_prime(uint64_t n) {
    const uint64_t small[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return false;
    for (uint64_t s : small) {
        if (n == s) return true;
        if (n % s == 0) return false;
    }
    uint64_t odd = n - 1;
    int twos = 0;
    while ((odd & 1) == 0) {
        odd >>= 1;
        ++twos;
    }
    for (uint64_t a : small) {
        uint64_t y = host_powmod(a, odd, n);
        if (y == 1 || y == n - 1) continue;
        bool witness = true;
        for (int r = 1; r < twos && witness; ++r) {
            y = host_mulmod(y, y, n);
            if (y == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

static uint32_t rev_bits(uint32_t v, unsigned bits) {
    uint32_t r = 0;
    for (unsigned k = 0; k < bits; ++k) r |= ((v >> k) & 1u) << (bits - 1 - k);
    return r;
}

static uint64_t shoup_of(uint64_t w, uint64_t q) { return (uint64_t)(((u128)w << 64) / q); }

// smallest primitive 2N-th root of unity: find one, then scan its odd powers.
static uint64_t min_primitive_root(uint64_t q, uint64_t two_n) {
    const uint64_t cofactor = (q - 1) / two_n;
    uint64_t any = 0;
    for (uint64_t g = 2; !any; ++g) {
        uint64_t cand = host_powmod(g, cofactor, q);
        if (host_powmod(cand, two_n >> 1, q) == q - 1) any = cand;
    }
    const uint64_t step = host_mulmod(any, any, q);
    uint64_t least = any, walk = any;
    for (uint64_t k = 3; k < two_n; k += 2) {
        walk = host_mulmod(walk, step, q);
        if (walk < least) least = walk;
    }
    return least;
}

template <int LOGN>
static void layout_tables(HostLimb &hl) {
    const size_t N = (size_t)1 << LOGN;
    const uint64_t q = hl.lp.q;
    hl.tw.assign(N, U64x2{0, 0});
    hl.itw.assign(N, U64x2{0, 0});
    for (int s = 0; s < LOGN; ++s)
        for (int i = 0; i < (1 << s); ++i) {
            const size_t nat = ((size_t)1 << s) + i, pos = (size_t)tw_pos<LOGN>(s, i);
            hl.tw[pos] = U64x2{hl.root_powers[nat], shoup_of(hl.root_powers[nat], q)};
            hl.itw[pos] = U64x2{hl.inv_root_powers[nat], shoup_of(hl.inv_root_powers[nat], q)};
        }
}

std::string build_host_params(unsigned log_n, unsigned L, const uint64_t *moduli, HostParams &out) {
    if (log_n < 12 || log_n > 14) return "log_n must be 12, 13 or 14";
    if (L < 1 || L > 16) return "n_limbs must be in [1,16]";
    const uint64_t two_n = (uint64_t)2 << log_n;
    const size_t N = (size_t)1 << log_n;
    std::vector<uint64_t> qs;
    if (moduli) {
        for (unsigned l = 0; l < L; ++l) {
            const uint64_t q = moduli[l];
            if (q >= (1ull << 60) || q <= (1ull << 33)) return "modulus out of range (2^33, 2^60)";
            if ((q - 1) % two_n) return "modulus is not 1 mod 2N";
            if (!host_is_prime(q)) return "modulus is not prime";
            for (uint64_t prev : qs)
                if (prev == q) return "moduli must be distinct";
            qs.push_back(q);
        }
    } else {
        // default basis: the L largest primes below 2^60 of the form k * 2^32 + 1.  They are 1 mod 2N
        // for every supported N, and multiplying by such a modulus costs one 32-bit multiply-add (modarith.cuh, DPFHE_FAST).
        uint64_t cand = (1ull << 60) + 1;
        while (qs.size() < L) {
            cand -= 1ull << 32;
            if (host_is_prime(cand)) qs.push_back(cand);
        }
    }
    out.log_n = log_n;
    out.L = L;
    out.limbs.assign(L, HostLimb());
    for (unsigned l = 0; l < L; ++l) {
        HostLimb &hl = out.limbs[l];
        const uint64_t q = qs[l];
        hl.psi = min_primitive_root(q, two_n);
        std::vector<uint64_t> pw(N);
        pw[0] = 1;
        for (size_t k = 1; k < N; ++k) pw[k] = host_mulmod(pw[k - 1], hl.psi, q);
        hl.root_powers.resize(N);
        hl.inv_root_powers.resize(N);
        for (size_t k = 0; k < N; ++k) {
            const uint64_t w = pw[rev_bits((uint32_t)k, log_n)];
            hl.root_powers[k] = w;
            hl.inv_root_powers[k] = host_powmod(w, q - 2, q);
        }
        LimbParams &lp = hl.lp;
        lp.q = q;
        lp.q2 = 2 * q;
        lp.qsb = (uint64_t)SB * q;
        lp.q4 = 4 * q;
        lp.q8 = 8 * q;
        lp.nq = 0 - q;
        unsigned bits = 64 - (unsigned)__builtin_clzll(q);
        lp.bar_shift = bits - 2;
        lp.bar_mu = (uint64_t)(((u128)1 << (lp.bar_shift + 64)) / q);
        lp.mu32 = (uint32_t)(((u128)1 << 64) / q);
        lp.nqh = (uint32_t)q == 1u ? 0u - (uint32_t)(q >> 32) : 0u;
        lp.pad_ = 0;
        lp.ninv = host_powmod((uint64_t)N % q, q - 2, q);
        lp.ninv_s = shoup_of(lp.ninv, q);
        lp.wninv = host_mulmod(hl.inv_root_powers[1], lp.ninv, q);
        lp.wninv_s = shoup_of(lp.wninv, q);
        switch (log_n) {
            case 12: layout_tables<12>(hl); break;
            case 13: layout_tables<13>(hl); break;
            default: layout_tables<14>(hl); break;
        }
    }
    return "";
}

void build_ms_consts(const HostParams &hp, uint64_t t_plain, MsConsts &K) {
    typedef unsigned __int128 u128;
    auto shoup = [](uint64_t w, uint64_t q) { return (uint64_t)((((u128)w) << 64) / q); };
    const unsigned L = hp.L;
    const uint64_t ql = hp.limbs[L - 1].lp.q;
    K = MsConsts();
    K.half = ql >> 1;
    K.has_t = t_plain ? 1u : 0u;
    K.tinv = t_plain ? host_powmod(t_plain % ql, ql - 2, ql) : 1;
    K.tinv_s = shoup(K.tinv, ql);
    for (unsigned i = 0; i + 1 < L; ++i) {
        const uint64_t q = hp.limbs[i].lp.q;
        K.qlm[i] = ql % q;
        K.qlm_s[i] = shoup(K.qlm[i], q);
        K.inv[i] = host_powmod(K.qlm[i], q - 2, q);
        K.inv_s[i] = shoup(K.inv[i], q);
        K.sinv[i] = t_plain ? host_mulmod(t_plain % q, K.inv[i], q) : K.inv[i];
        K.sinv_s[i] = shoup(K.sinv[i], q);
    }
}

void build_group_consts(const HostParams &hp, unsigned K, uint64_t t_plain, GroupConsts &G, MsConsts &Km) {
    typedef unsigned __int128 u128;
    auto shoup = [](uint64_t w, uint64_t q) { return (uint64_t)((((u128)w) << 64) / q); };
    const unsigned L = hp.L, Lq = L - K;
    auto q_of = [&](unsigned l) { return hp.limbs[l].lp.q; };
    // product of the moduli lo .. hi-1 except `skip`, modulo q
    auto prod_mod = [&](unsigned lo, unsigned hi, unsigned skip, uint64_t q) {
        uint64_t r = 1 % q;
        for (unsigned m = lo; m < hi; ++m)
            if (m != skip) r = host_mulmod(r, q_of(m) % q, q);
        return r;
    };
    G = GroupConsts();
    G.Lq = Lq;
    G.K = K;
    G.dnum = (Lq + K - 1) / K;
    auto fold = [&](unsigned l, uint64_t f) {   // limb l with N^-1 replaced by N^-1 * f
        LimbParams p = hp.limbs[l].lp;
        p.ninv = host_mulmod(p.ninv, f, p.q);
        p.ninv_s = shoup(p.ninv, p.q);
        p.wninv = host_mulmod(p.wninv, f, p.q);
        p.wninv_s = shoup(p.wninv, p.q);
        return p;
    };
    for (unsigned j = 0; j < Lq; ++j) {
        const unsigned lo = j / K * K, hi = lo + K < Lq ? lo + K : Lq;
        const uint64_t qj = q_of(j);
        G.lp_up[j] = fold(j, host_powmod(prod_mod(lo, hi, j, qj), qj - 2, qj));
        for (unsigned i = 0; i < L; ++i) {
            G.up[j][i] = prod_mod(lo, hi, j, q_of(i));
            G.up_s[j][i] = shoup(G.up[j][i], q_of(i));
        }
    }
    for (unsigned k = 0; k < K; ++k) {
        const unsigned s = Lq + k;
        const uint64_t p = q_of(s);
        uint64_t f = prod_mod(Lq, L, s, p);
        if (t_plain) f = host_mulmod(f, t_plain % p, p);
        G.lp_up[s] = fold(s, host_powmod(f, p - 2, p));
        G.half[k] = p >> 1;
        for (unsigned i = 0; i < Lq; ++i) {
            G.dn[k][i] = prod_mod(Lq, L, s, q_of(i));
            G.dn_s[k][i] = shoup(G.dn[k][i], q_of(i));
        }
    }
    Km = MsConsts();
    Km.has_t = 0;   // t^-1 is folded into lp_up of the special limbs
    Km.tinv = 1;
    for (unsigned i = 0; i < Lq; ++i) {
        const uint64_t q = q_of(i), Pm = prod_mod(Lq, L, L, q);
        G.neg_p[i] = q - Pm;
        Km.qlm[i] = Pm;
        Km.qlm_s[i] = shoup(Pm, q);
        Km.inv[i] = host_powmod(Pm, q - 2, q);
        Km.inv_s[i] = shoup(Km.inv[i], q);
        Km.sinv[i] = t_plain ? host_mulmod(t_plain % q, Km.inv[i], q) : Km.inv[i];
        Km.sinv_s[i] = shoup(Km.sinv[i], q);
    }
}

}  // namespace dpfhe
