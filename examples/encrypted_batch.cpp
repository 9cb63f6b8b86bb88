// encrypted_batch.cpp — the encrypted counterpart of the reference's examples/batch_generation.cpp
// (which is a plaintext 4-prompt generate_batch call, examples/batch_generation.cpp:23-37): a batch of
// ciphertext x ciphertext multiplications through deeppowers::api::fhe::Evaluator, same try/catch shape.
//   g++ -std=c++17 -Iinclude examples/encrypted_batch.cpp -Ldeeppowers_b200 -ldpfhe -Wl,-rpath,$PWD/deeppowers_b200 -o encrypted_batch
#include <deeppowers_fhe.hpp>

#include <chrono>
#include <iostream>
#include <vector>

using namespace deeppowers::api::fhe;
using namespace std::chrono;

int main() {
    try {
        EncryptionParameters parms;   // N = 8192, L = 4, default moduli
        Evaluator evaluator(parms);
        const std::size_t batch = 64;
        // synthetic residues (a real caller passes ciphertexts produced by its own encryptor)
        std::vector<std::uint64_t> a(batch * evaluator.ciphertext_words()), b(a.size()), out(a.size());
        std::vector<std::uint64_t> relin_key(evaluator.switch_key_words());
        std::uint64_t s = 1;
        auto next = [&](std::uint64_t q) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (s >> 4) % q; };
        for (std::size_t k = 0; k < a.size(); ++k) {
            const std::uint64_t q = evaluator.modulus((unsigned)((k / evaluator.poly_degree()) % evaluator.limbs()));
            a[k] = next(q);
            b[k] = next(q);
        }
        for (std::size_t k = 0; k < relin_key.size(); ++k)
            relin_key[k] = next(evaluator.modulus((unsigned)((k / evaluator.poly_degree()) % evaluator.limbs())));

        std::cout << "Multiplying " << batch << " ciphertext pairs (N=" << evaluator.poly_degree() << ", L=" << evaluator.limbs() << ")..." << std::endl;
        auto t0 = high_resolution_clock::now();
        evaluator.multiply_relin({a.data(), batch}, {b.data(), batch}, relin_key.data(), {out.data(), batch});
        auto dt = duration_cast<microseconds>(high_resolution_clock::now() - t0).count() / 1e6;
        std::cout << "Total time: " << dt << " seconds" << std::endl;
        std::cout << "Throughput: " << batch / dt << " ct-mults per second (host buffers, H2D+D2H included)" << std::endl;

        // the same batch sharded over every GPU of the box (BASELINE.json config 5's layout): one object, no MPI, no collective —
        // each device pipelines its own contiguous shard and writes its slice of `out`
        MultiEvaluator all(parms);
        std::vector<std::uint64_t> out_multi(a.size());
        all.multiply_relin({a.data(), batch}, {b.data(), batch}, relin_key.data(), {out_multi.data(), batch});
        std::cout << "Sharded over " << all.device_count() << " GPU(s): " << (out_multi == out ? "identical result" : "RESULT DIFFERS") << std::endl;
        return out_multi == out ? 0 : 2;
    } catch (const std::exception &e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
}
