// encrypted_job.cpp — the reference's Model API (through include/shim/deeppowers.hpp) driving the encrypted route:
// set_config("fhe","on") makes generate_batch() multiply DPFHEv1 ciphertext files on the GPU.
//   usage: encrypted_job <a.dpfhe> <b.dpfhe> <relin_key.dpfhe> <out.dpfhe> [log_n n_limbs [all|one [plain_modulus]]]
// n_limbs is the basis of the KEY; a key file of kind 4 / 5 (special primes) makes the ciphertext files carry fewer limbs.
// "all" shards the batch over every visible GPU (BASELINE.json config 5: the encrypted batch path across the GPUs of a box).
#include <deeppowers.hpp>

#include <iostream>

using namespace deeppowers::api;

int main(int argc, char **argv) {
    try {
        if (argc < 5) throw std::runtime_error("usage: encrypted_job <a> <b> <relin_key> <out> [log_n n_limbs [all]]");
        auto model = load_model("gpt2");
        model->set_config("fhe", "on");
        if (argc >= 7) {
            model->set_config("fhe.log_n", argv[5]);
            model->set_config("fhe.n_limbs", argv[6]);
        }
        if (argc >= 8) model->set_config("fhe.devices", argv[7]);
        if (argc >= 9) model->set_config("fhe.plain_modulus", argv[8]);
        GenerationConfig config;
        config.batch_size = 1;
        const std::string job = std::string(argv[1]) + " " + argv[2] + " " + argv[3] + " " + argv[4];
        auto results = model->generate_batch({job}, config);
        std::cout << results[0].texts[0] << " in " << results[0].generation_time << " seconds" << std::endl;
        return 0;
    } catch (const std::exception &e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
}
