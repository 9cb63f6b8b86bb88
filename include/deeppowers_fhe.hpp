// deeppowers_fhe.hpp — C++ host API of the encrypted hot path, in the reference's house style.
//
// The reference's public header (src/api/cpp/include/deeppowers.hpp:41-87) exposes
// deeppowers::api::{Model, GenerationConfig, load_model, ...} and has no Ciphertext/Evaluator types
// (SURVEY.md §0); this header adds them next to it, as namespace deeppowers::api::fhe, following
// the same conventions:
//   - errors are C++ exceptions (std::runtime_error), as CUDA_CHECK does in
//     src/core/hal/cuda/cuda_device.cpp:9-16 — every non-zero status of the C ABI is re-thrown
//     with dpfhe_last_error() as the message;
//   - resource-owning classes are non-copyable RAII handles (compare Model's pimpl,
//     deeppowers.hpp:73-75, and CUDADevice's dtor, cuda_device.cpp:26-41);
//   - one Evaluator is bound to one device (compare hal::CUDADevice(0), src/api/cpp/src/deeppowers.cpp:15).
// Header-only over the extern "C" library (include/dpfhe.h, libdpfhe.so); no CUDA headers needed.
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "dpfhe.h"

namespace deeppowers {
namespace api {
namespace fhe {

// Parameters of the ring Z_q[X]/(X^N+1) in RNS form.
struct EncryptionParameters {
    unsigned log_n = 13;                   // N = 8192
    unsigned n_limbs = 4;                  // L
    std::vector<std::uint64_t> moduli;     // empty: the L largest primes k*2^32+1 below 2^60 (the default basis)
};

// Non-owning view of `count` ciphertexts [count][2][L][N] (evaluation form) in host or device memory.
struct CiphertextBatch {
    std::uint64_t *data = nullptr;
    std::size_t count = 0;
};
struct ConstCiphertextBatch {
    const std::uint64_t *data = nullptr;
    std::size_t count = 0;
    ConstCiphertextBatch() = default;
    ConstCiphertextBatch(const std::uint64_t *d, std::size_t c) : data(d), count(c) {}
    ConstCiphertextBatch(const CiphertextBatch &b) : data(b.data), count(b.count) {}
};

class Evaluator {
public:
    explicit Evaluator(const EncryptionParameters &parms, int device_id = 0) : log_n_(parms.log_n), limbs_(parms.n_limbs) {
        dpfhe_params p;
        p.log_n = parms.log_n;
        p.n_limbs = parms.n_limbs;
        p.moduli = parms.moduli.empty() ? nullptr : parms.moduli.data();
        if (!parms.moduli.empty() && parms.moduli.size() != parms.n_limbs)
            throw std::runtime_error("EncryptionParameters: moduli.size() must equal n_limbs");
        check(dpfhe_context_create(&p, device_id, &ctx_));
    }
    ~Evaluator() { dpfhe_context_destroy(ctx_); }
    Evaluator(const Evaluator &) = delete;
    Evaluator &operator=(const Evaluator &) = delete;

    std::size_t poly_degree() const { return std::size_t(1) << log_n_; }
    unsigned limbs() const { return limbs_; }
    std::size_t poly_words() const { return poly_degree() * limbs_; }           // [L][N]
    std::size_t ciphertext_words() const { return 2 * poly_words(); }           // [2][L][N]
    std::size_t switch_key_words() const { return 2 * limbs_ * poly_words(); }  // [L][2][L][N]
    std::uint64_t modulus(unsigned limb) const {
        std::uint64_t q = 0;
        check(dpfhe_get_modulus(ctx_, limb, &q));
        return q;
    }
    // Galois element of a rotation by `steps` slots: 5^steps mod 2N
    std::uint64_t galois_element(long steps) const {
        const std::uint64_t two_n = std::uint64_t(2) << log_n_, half = poly_degree() / 2;
        std::uint64_t e = ((steps % (long)half) + (long)half) % (long)half, g = 1, b = 5;
        for (; e; e >>= 1, b = b * b % two_n)
            if (e & 1) g = g * b % two_n;
        return g;
    }

    // ---- host-buffer calls: synchronous; H2D / compute / D2H are pipelined inside the library ----
    void multiply_relin(ConstCiphertextBatch a, ConstCiphertextBatch b, const std::uint64_t *relin_key, CiphertextBatch out) {
        same(a.count, b.count, out.count);
        check(dpfhe_ct_mul_relin_host(ctx_, a.data, b.data, relin_key, out.data, a.count));
    }
    void multiply_plain(ConstCiphertextBatch ct, const std::uint64_t *plain_eval, CiphertextBatch out) {
        same(ct.count, ct.count, out.count);
        check(dpfhe_ct_mul_plain_host(ctx_, ct.data, plain_eval, out.data, ct.count));
    }
    void rotate(ConstCiphertextBatch ct, long steps, const std::uint64_t *galois_key, CiphertextBatch out) {
        same(ct.count, ct.count, out.count);
        check(dpfhe_rotate_host(ctx_, ct.data, galois_element(steps), galois_key, out.data, ct.count));
    }
    // special-prime (hybrid) forms: this evaluator's last limb is the special prime; batches hold limbs()-1 limbs
    void multiply_relin_hybrid(ConstCiphertextBatch a, ConstCiphertextBatch b, const std::uint64_t *relin_key, CiphertextBatch out,
                               std::uint64_t plain_modulus = 0) {
        same(a.count, b.count, out.count);
        check(dpfhe_ct_mul_relin_hybrid_host(ctx_, a.data, b.data, relin_key, out.data, a.count, plain_modulus));
    }
    void rotate_hybrid(ConstCiphertextBatch ct, long steps, const std::uint64_t *galois_key, CiphertextBatch out,
                       std::uint64_t plain_modulus = 0) {
        same(ct.count, ct.count, out.count);
        check(dpfhe_rotate_hybrid_host(ctx_, ct.data, galois_element(steps), galois_key, out.data, ct.count, plain_modulus));
    }
    // grouped hybrid forms (digits of `special` limbs, `special` special primes at the end of the basis; dpfhe.h): batches hold
    // limbs()-special limbs, keys grouped_digits(special) digits
    unsigned grouped_digits(unsigned special) const {
        unsigned d = 0;
        check(dpfhe_grouped_digits(ctx_, special, &d));
        return d;
    }
    void multiply_relin_grouped(unsigned special, ConstCiphertextBatch a, ConstCiphertextBatch b, const std::uint64_t *relin_key, CiphertextBatch out,
                                std::uint64_t plain_modulus = 0) {
        same(a.count, b.count, out.count);
        check(dpfhe_ct_mul_relin_grouped_host(ctx_, special, a.data, b.data, relin_key, out.data, a.count, plain_modulus));
    }
    void rotate_grouped(unsigned special, ConstCiphertextBatch ct, long steps, const std::uint64_t *galois_key, CiphertextBatch out,
                        std::uint64_t plain_modulus = 0) {
        same(ct.count, ct.count, out.count);
        check(dpfhe_rotate_grouped_host(ctx_, special, ct.data, galois_element(steps), galois_key, out.data, ct.count, plain_modulus));
    }
    void multiply_relin_grouped_device(unsigned special, const std::uint64_t *a, const std::uint64_t *b, const std::uint64_t *relin_key,
                                       std::uint64_t *out, std::size_t count, std::uint64_t plain_modulus = 0, void *stream = nullptr) {
        check(dpfhe_ct_mul_relin_grouped(ctx_, special, a, b, relin_key, out, count, plain_modulus, stream));
    }
    void rotate_grouped_device(unsigned special, const std::uint64_t *ct, long steps, const std::uint64_t *galois_key, std::uint64_t *out,
                               std::size_t count, std::uint64_t plain_modulus = 0, void *stream = nullptr) {
        check(dpfhe_rotate_grouped(ctx_, special, ct, galois_element(steps), galois_key, out, count, plain_modulus, stream));
    }
    // out[r] = ct rotated by steps[r] with galois_keys[r], the rotations sharing the basis conversion of ct (same plaintexts as
    // rotate_grouped_device, not the same bits); out holds n_rot * count ciphertexts
    void rotate_hoisted_grouped_device(unsigned special, const std::uint64_t *ct, const std::vector<long> &steps,
                                       const std::vector<const std::uint64_t *> &galois_keys, std::uint64_t *out, std::size_t count,
                                       std::uint64_t plain_modulus = 0, void *stream = nullptr) {
        if (steps.size() != galois_keys.size()) throw std::invalid_argument("one Galois key per rotation");
        std::vector<std::uint64_t> elts(steps.size());
        for (std::size_t r = 0; r < steps.size(); ++r) elts[r] = galois_element(steps[r]);
        check(dpfhe_rotate_hoisted_grouped(ctx_, special, ct, steps.size(), elts.data(), galois_keys.data(), out, count, plain_modulus, stream));
    }
    // divide by the product of the last `special` limbs: in holds limbs() limbs per polynomial, out limbs()-special
    void mod_down_special_device(unsigned special, const std::uint64_t *ct, std::uint64_t *out, std::size_t count, std::uint64_t plain_modulus = 0,
                                 void *stream = nullptr) {
        check(dpfhe_mod_down_special(ctx_, special, ct, out, 2 * count, plain_modulus, stream));
    }
    // drop the last limb: in holds limbs() limbs per polynomial, out limbs()-1
    void mod_switch_to_next(ConstCiphertextBatch in, CiphertextBatch out, std::uint64_t plain_modulus = 0) {
        same(in.count, in.count, out.count);
        check(dpfhe_mod_switch_down_host(ctx_, in.data, out.data, 2 * in.count, plain_modulus));
    }
    void transform_to_ntt(std::uint64_t *polys, std::size_t n_polys) { check(dpfhe_ntt_fwd_host(ctx_, polys, n_polys)); }
    void transform_from_ntt(std::uint64_t *polys, std::size_t n_polys) { check(dpfhe_ntt_inv_host(ctx_, polys, n_polys)); }

    // ---- device-pointer calls: asynchronous on `stream` (a cudaStream_t; nullptr = the evaluator's own) ----
    void multiply_relin_device(const std::uint64_t *a, const std::uint64_t *b, const std::uint64_t *relin_key, std::uint64_t *out,
                               std::size_t count, void *stream = nullptr) {
        check(dpfhe_ct_mul_relin(ctx_, a, b, relin_key, out, count, stream));
    }
    void multiply_plain_device(const std::uint64_t *ct, const std::uint64_t *plain_eval, std::uint64_t *out, std::size_t count,
                               void *stream = nullptr) {
        check(dpfhe_ct_mul_plain(ctx_, ct, plain_eval, out, count, stream));
    }
    void rotate_device(const std::uint64_t *ct, long steps, const std::uint64_t *galois_key, std::uint64_t *out, std::size_t count,
                       void *stream = nullptr) {
        check(dpfhe_rotate(ctx_, ct, galois_element(steps), galois_key, out, count, stream));
    }
    // several rotations of the same batch, sharing the digit decomposition; out is [steps.size()][count] ciphertexts and
    // galois_keys[r] the device pointer of the key for steps[r].  Same bits as rotate_device called steps.size() times.
    void rotate_many_device(const std::uint64_t *ct, const std::vector<long> &steps, const std::vector<const std::uint64_t *> &galois_keys,
                            std::uint64_t *out, std::size_t count, void *stream = nullptr) {
        if (steps.size() != galois_keys.size()) throw std::runtime_error("one Galois key per rotation step");
        std::vector<std::uint64_t> elts;
        for (long k : steps) elts.push_back(galois_element(k));
        check(dpfhe_rotate_hoisted(ctx_, ct, steps.size(), elts.data(), galois_keys.data(), out, count, stream));
    }
    void add_device(const std::uint64_t *a, const std::uint64_t *b, std::uint64_t *out, std::size_t count, void *stream = nullptr) {
        check(dpfhe_poly_add(ctx_, a, b, out, 2 * count, stream));   // a ciphertext is two polynomials
    }
    void multiply_plain_accumulate_device(const std::uint64_t *ct, const std::uint64_t *plain_eval, std::uint64_t *acc, std::size_t count,
                                          void *stream = nullptr) {
        check(dpfhe_ct_mul_plain_acc(ctx_, ct, plain_eval, acc, count, stream));
    }
    // out[g][k] = sum_b steps[b][k] o plain[g][b]: the fused inner loop of a baby-step/giant-step matrix-vector product
    // (steps [n_steps][count] ciphertexts, plain [n_groups][n_steps] plaintexts in evaluation form, out [n_groups][count])
    void multiply_plain_inner_device(const std::uint64_t *steps, std::size_t n_steps, const std::uint64_t *plain, std::size_t n_groups,
                                     std::uint64_t *out, std::size_t count, void *stream = nullptr) {
        check(dpfhe_ct_mul_plain_inner(ctx_, steps, n_steps, plain, n_groups, out, count, stream));
    }
    // drop the last limb of `count` ciphertexts (2*count polynomials); the result belongs to the first L-1 moduli
    void mod_switch_to_next_device(const std::uint64_t *ct, std::uint64_t *out, std::size_t count, std::uint64_t plain_modulus = 0,
                                   void *stream = nullptr) {
        check(dpfhe_mod_switch_down(ctx_, ct, out, 2 * count, plain_modulus, stream));
    }
    // hybrid (special-prime) key switching: this evaluator's last limb is the special prime, ciphertexts carry
    // limbs()-1 limbs and keys are [limbs()-1][2][limbs()][N]
    void multiply_relin_hybrid_device(const std::uint64_t *a, const std::uint64_t *b, const std::uint64_t *relin_key, std::uint64_t *out,
                                      std::size_t count, std::uint64_t plain_modulus = 0, void *stream = nullptr) {
        check(dpfhe_ct_mul_relin_hybrid(ctx_, a, b, relin_key, out, count, plain_modulus, stream));
    }
    void rotate_hybrid_device(const std::uint64_t *ct, long steps, const std::uint64_t *galois_key, std::uint64_t *out, std::size_t count,
                              std::uint64_t plain_modulus = 0, void *stream = nullptr) {
        check(dpfhe_rotate_hybrid(ctx_, ct, galois_element(steps), galois_key, out, count, plain_modulus, stream));
    }
    void keyswitch_device(const std::uint64_t *digits, const std::uint64_t *key, std::uint64_t *out, std::size_t count, void *stream = nullptr) {
        check(dpfhe_keyswitch(ctx_, digits, key, out, count, stream));
    }
    void transform_to_ntt_device(std::uint64_t *polys, std::size_t n_polys, void *stream = nullptr) {
        check(dpfhe_ntt_fwd(ctx_, polys, n_polys, stream));
    }
    void transform_from_ntt_device(std::uint64_t *polys, std::size_t n_polys, void *stream = nullptr) {
        check(dpfhe_ntt_inv(ctx_, polys, n_polys, stream));
    }

    // waits for every call issued through this evaluator, whatever stream it ran on
    void synchronize() { check(dpfhe_synchronize(ctx_)); }

    dpfhe_ctx *native_handle() { return ctx_; }

private:
    friend class LinearLayer;
    static void check(int status) {
        if (status != DPFHE_OK) throw std::runtime_error(dpfhe_last_error());
    }
    static void same(std::size_t a, std::size_t b, std::size_t c) {
        if (a != b || a != c) throw std::runtime_error("ciphertext batches must have the same count");
    }
    dpfhe_ctx *ctx_ = nullptr;
    unsigned log_n_, limbs_;
};

// An encrypted linear layer y = W x (baby-step/giant-step diagonals) whose weights and Galois keys live on the device.
// diagonals: [n][L][N] plaintexts in evaluation form (diagonal g*baby + b pre-rotated by -g*baby), n a multiple of baby;
// baby_keys: [baby-1][L][2][L][N] keys of the rotations by 1 .. baby-1 slots; giant_key: [L][2][L][N], rotation by `baby`.
class LinearLayer {
public:
    LinearLayer(Evaluator &ev, const std::uint64_t *diagonals, std::size_t n_diagonals, std::size_t baby, const std::uint64_t *baby_keys,
                const std::uint64_t *giant_key) {
        Evaluator::check(dpfhe_linear_create(ev.native_handle(), diagonals, n_diagonals, baby, baby_keys, giant_key, &h_));
    }
    ~LinearLayer() { dpfhe_linear_destroy(h_); }
    LinearLayer(const LinearLayer &) = delete;
    LinearLayer &operator=(const LinearLayer &) = delete;
    void apply(ConstCiphertextBatch in, CiphertextBatch out) {   // host buffers, pipelined
        if (in.count != out.count) throw std::runtime_error("ciphertext batches must have the same count");
        Evaluator::check(dpfhe_linear_apply_host(h_, in.data, out.data, in.count));
    }
    void apply_device(const std::uint64_t *in, std::uint64_t *out, std::size_t count, void *stream = nullptr) {
        Evaluator::check(dpfhe_linear_apply(h_, in, out, count, stream));
    }

private:
    dpfhe_linear *h_ = nullptr;
};

// Several GPUs behind one object (dpfhe_multi_*): contiguous shards of every batch, one context and host thread per device, no
// collective.  The reference's distributed layer is one MPI rank per GPU (src/core/distributed/distributed_context.cpp:242-250).
class MultiEvaluator {
public:
    // devices empty = every visible GPU
    explicit MultiEvaluator(const EncryptionParameters &parms, const std::vector<int> &devices = {}) : log_n_(parms.log_n), limbs_(parms.n_limbs) {
        dpfhe_params p;
        p.log_n = parms.log_n;
        p.n_limbs = parms.n_limbs;
        p.moduli = parms.moduli.empty() ? nullptr : parms.moduli.data();
        check(dpfhe_multi_create(&p, devices.empty() ? nullptr : devices.data(), (int)devices.size(), &m_));
    }
    ~MultiEvaluator() { dpfhe_multi_destroy(m_); }
    MultiEvaluator(const MultiEvaluator &) = delete;
    MultiEvaluator &operator=(const MultiEvaluator &) = delete;
    int device_count() const { return dpfhe_multi_device_count(m_); }
    std::size_t poly_degree() const { return std::size_t(1) << log_n_; }
    unsigned limbs() const { return limbs_; }
    std::uint64_t modulus(unsigned limb) const {
        std::uint64_t q = 0;
        check(dpfhe_get_modulus(dpfhe_multi_context(m_, 0), limb, &q));
        return q;
    }
    // host buffers: every device pipelines its own shard; the result needs no gather
    void multiply_relin(ConstCiphertextBatch a, ConstCiphertextBatch b, const std::uint64_t *relin_key, CiphertextBatch out) {
        if (a.count != b.count || a.count != out.count) throw std::runtime_error("ciphertext batches must have the same count");
        check(dpfhe_multi_ct_mul_relin_host(m_, a.data, b.data, relin_key, out.data, a.count));
    }
    // the same with a special-prime relinearisation key (digits of `special` limbs, `special` special primes; batches hold
    // limbs() - special limbs)
    void multiply_relin_grouped(unsigned special, ConstCiphertextBatch a, ConstCiphertextBatch b, const std::uint64_t *relin_key, CiphertextBatch out,
                                std::uint64_t plain_modulus = 0) {
        if (a.count != b.count || a.count != out.count) throw std::runtime_error("ciphertext batches must have the same count");
        check(dpfhe_multi_ct_mul_relin_grouped_host(m_, special, a.data, b.data, relin_key, out.data, a.count, plain_modulus));
    }
    // device buffers: a[r], b[r], relin_key[r] on device r (shard r of the batch); the whole result on device `root`
    void multiply_relin_gather_device(const std::vector<const std::uint64_t *> &a, const std::vector<const std::uint64_t *> &b,
                                      const std::vector<const std::uint64_t *> &relin_key, std::uint64_t *out_on_root, int root, std::size_t count) {
        if ((int)a.size() != device_count() || a.size() != b.size() || a.size() != relin_key.size())
            throw std::runtime_error("one operand pointer per device");
        check(dpfhe_multi_ct_mul_relin_gather(m_, a.data(), b.data(), relin_key.data(), out_on_root, root, count));
    }
    dpfhe_multi *native_handle() { return m_; }

private:
    static void check(int status) {
        if (status != DPFHE_OK) throw std::runtime_error(dpfhe_last_error());
    }
    dpfhe_multi *m_ = nullptr;
    unsigned log_n_, limbs_;
};

}  // namespace fhe
}  // namespace api
}  // namespace deeppowers
