// host_params.hpp — host-side derivation of the parameter set:
// NTT-friendly primes, smallest primitive 2N-th roots, twiddle tables in the device layout.
// Independent of oracle/ (the product never links the oracle); tests compare the two.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "types.hpp"

namespace dpfhe {

struct HostLimb {
    LimbParams lp;
    uint64_t psi;
    std::vector<uint64_t> root_powers;      // psi^bitrev(i), natural table order
    std::vector<uint64_t> inv_root_powers;  // psi^-bitrev(i)
    std::vector<U64x2> tw;                  // device layout (tw_pos), with Shoup companions
    std::vector<U64x2> itw;
};

struct HostParams {
    unsigned log_n = 0, L = 0;
    std::vector<HostLimb> limbs;
};

// returns "" on success, else an error message
std::string build_host_params(unsigned log_n, unsigned L, const uint64_t *moduli, HostParams &out);

// constants for dropping the last limb of `hp` (t_plain = 0: plain rounding); requires hp.L >= 2 and t_plain < q_last
void build_ms_consts(const HostParams &hp, uint64_t t_plain, MsConsts &K);

// constants of grouped hybrid key switching with the last K limbs of `hp` as special primes (types.hpp: GroupConsts); Km
// receives the constants of the division by P.  Requires 1 <= K <= KS_MAX_SPECIAL, K < hp.L, t_plain < every special prime.
void build_group_consts(const HostParams &hp, unsigned K, uint64_t t_plain, GroupConsts &G, MsConsts &Km);

uint64_t host_mulmod(uint64_t a, uint64_t b, uint64_t q);
uint64_t host_powmod(uint64_t a, uint64_t e, uint64_t q);
bool host_is_prime(uint64_t n);

}  // namespace dpfhe
