// shim/deeppowers.hpp — source-compatibility shim for the reference's public C++ API.
//
// BASELINE.json:north_star wants examples/basic_generation.cpp and examples/quantization_example.cpp to "link
// unchanged".  They do not even compile against the reference's own header (logprobs is used as an
// optional, <cmath>/<numeric>/<csignal> are missing, and the quantization calls are not declared in
// src/api/cpp/include/deeppowers.hpp:41-76).  This header declares the API surface those examples (and
// batch_generation.cpp, stream_generation.cpp) actually use, header-only, on top of libdpfhe.so.
//
// What it is NOT: a language model.  The plaintext GPT engine is out of this tier's scope, as it
// is unimplemented in the reference itself (forward pass is a TODO, src/core/execution/models/gpt_model.cpp:103).
// generate() therefore returns a clearly labelled placeholder continuation.  What it adds is the encrypted
// route: with set_config("fhe", "on"), generate_batch() treats each prompt as a job line
//     "<a.dpfhe> <b.dpfhe> <relin_key.dpfhe> <out.dpfhe>"
// of DPFHEv1 files (include/dpfhe_wire.hpp), multiplies the ciphertext batches on the GPU through
// deeppowers::api::fhe::Evaluator and writes the result — the file-exchange shape of the reference's FHE demo
// (demo/fhe_server.py:113-164) behind the Model API the examples already use.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <csignal>
#include <functional>
#include <map>
#include <memory>
#include <numeric>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../deeppowers_fhe.hpp"
#include "../dpfhe_wire.hpp"

namespace deeppowers {
namespace api {

enum class QuantizationType { NONE, INT8, INT4, MIXED };
enum class QuantizationMethod { POST_TRAINING, DYNAMIC, QUANTIZATION_AWARE };

struct QuantizationConfig {
    QuantizationType type = QuantizationType::NONE;
    QuantizationMethod method = QuantizationMethod::POST_TRAINING;
    bool per_channel = false;
    std::unordered_map<std::string, QuantizationType> layer_precisions;
};

struct GenerationConfig {
    std::string model_type = "gpt";
    size_t max_tokens = 100;
    float temperature = 0.7f;
    float top_p = 1.0f;
    float top_k = 0.0f;
    std::vector<std::string> stop_tokens;
    bool stream = false;
    size_t batch_size = 1;
};

struct GenerationResult {
    std::vector<std::string> texts;
    std::optional<std::vector<float>> logprobs;            // the examples test it like a pointer
    std::vector<std::vector<std::string>> tokens;
    std::optional<std::vector<std::string>> stop_reasons;
    double generation_time = 0.0;
};

using StreamCallback = std::function<bool(const GenerationResult &)>;

class Model {
public:
    explicit Model(const std::string &model_path) : path_(model_path) {
        config_["fhe"] = "off";
        config_["fhe.log_n"] = "13";
        config_["fhe.n_limbs"] = "4";
        config_["fhe.device"] = "0";
        config_["fhe.plain_modulus"] = "0";   // BGV plaintext modulus of special-prime key switching (0: plain rounding)
        config_["fhe.devices"] = "one";   // "all": shard every encrypted batch over all visible GPUs (dpfhe_multi_*)
    }

    GenerationResult generate(const std::string &prompt, const GenerationConfig &config = GenerationConfig()) {
        const auto t0 = std::chrono::steady_clock::now();
        GenerationResult r;
        std::string text = prompt + " [deeppowers-b200 shim: plaintext generation is outside the encrypted hot path]";
        for (const auto &stop : config.stop_tokens) {
            const auto pos = text.find(stop, prompt.size());
            if (!stop.empty() && pos != std::string::npos) text.resize(pos);
        }
        r.texts.push_back(text);
        r.stop_reasons = std::vector<std::string>{"max_tokens"};
        r.generation_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() + 1e-9;
        return r;
    }

    void generate_stream(const std::string &prompt, StreamCallback callback, const GenerationConfig &config = GenerationConfig()) {
        GenerationResult whole = generate(prompt, config);
        std::istringstream words(whole.texts[0]);
        std::string w;
        while (words >> w) {
            GenerationResult chunk;
            chunk.texts.push_back(w + " ");
            if (!callback(chunk)) return;
        }
    }

    std::vector<GenerationResult> generate_batch(const std::vector<std::string> &prompts, const GenerationConfig &config = GenerationConfig()) {
        std::vector<GenerationResult> out;
        for (const auto &p : prompts) out.push_back(get_config("fhe") == "on" ? run_encrypted_job(p) : generate(p, config));
        return out;
    }

    std::string model_type() const { return "gpt"; }
    std::string model_path() const { return path_; }
    size_t vocab_size() const { return 50257; }
    size_t max_sequence_length() const { return 2048; }   // src/core/execution/models/gpt_model.hpp:20

    void to_device(const std::string &device) {
        if (device != "cpu" && device.rfind("cuda", 0) != 0) throw std::runtime_error("unknown device: " + device);
        device_ = device;
    }
    std::string device() const { return device_; }

    void set_config(const std::string &key, const std::string &value) {
        config_[key] = value;
        if (key.rfind("fhe", 0) == 0) evaluator_.reset();   // parameters changed: rebuild lazily
    }
    std::string get_config(const std::string &key) const {
        const auto it = config_.find(key);
        return it == config_.end() ? std::string() : it->second;
    }

    // ---- quantization surface used by examples/quantization_example.cpp:50-124 (bookkeeping only) ----
    size_t get_model_size() const {
        const size_t params = 124439808;   // GPT-2 small
        switch (quant_applied_ ? quant_.type : QuantizationType::NONE) {
            case QuantizationType::INT8: return params;
            case QuantizationType::INT4: return params / 2;
            case QuantizationType::MIXED: return params * 3 / 4;
            default: return params * 4;
        }
    }
    void set_quantization_config(const QuantizationConfig &c) { quant_ = c; }
    void quantize(const std::vector<std::string> & /*calibration_data*/) { quant_applied_ = quant_.type != QuantizationType::NONE; }
    void dequantize() { quant_applied_ = false; }

    // ---- the encrypted route ----
    fhe::Evaluator &fhe_evaluator() {
        if (!evaluator_) evaluator_ = std::make_unique<fhe::Evaluator>(fhe_parameters(), std::stoi(get_config("fhe.device")));
        return *evaluator_;
    }
    // every visible GPU behind one evaluator: set_config("fhe.devices", "all") (BASELINE.json config 5: the batch of
    // examples/batch_generation's encrypted path sharded over the GPUs of the box)
    fhe::MultiEvaluator &fhe_multi_evaluator() {
        if (!multi_) multi_ = std::make_unique<fhe::MultiEvaluator>(fhe_parameters());
        return *multi_;
    }

private:
    fhe::EncryptionParameters fhe_parameters() {
        fhe::EncryptionParameters parms;
        parms.log_n = (unsigned)std::stoul(get_config("fhe.log_n"));
        parms.n_limbs = (unsigned)std::stoul(get_config("fhe.n_limbs"));
        return parms;
    }
    GenerationResult run_encrypted_job(const std::string &line) {
        const auto t0 = std::chrono::steady_clock::now();
        std::istringstream in(line);
        std::string fa, fb, fk, fo;
        if (!(in >> fa >> fb >> fk >> fo)) throw std::runtime_error("fhe job must be '<a> <b> <relin_key> <out>' (DPFHEv1 files)");
        std::vector<std::uint64_t> a, b, k;
        const fhe::WireHeader ha = fhe::read_wire_file(fa, a), hb = fhe::read_wire_file(fb, b), hk = fhe::read_wire_file(fk, k);
        fhe::Evaluator &ev = fhe_evaluator();
        if (hk.kind == (std::uint32_t)fhe::WireKind::HybridSwitchKey || hk.kind == (std::uint32_t)fhe::WireKind::GroupedSwitchKey)
            return run_special_prime_job(ev, ha, hb, hk, a, b, k, fo, t0);
        if (ha.kind != (std::uint32_t)fhe::WireKind::Ciphertexts || hb.kind != ha.kind || hk.kind != (std::uint32_t)fhe::WireKind::SwitchKey)
            throw std::runtime_error("fhe job: wrong file kinds");
        if (ha.log_n != hb.log_n || ha.n_limbs != hb.n_limbs || ha.count != hb.count || ha.n_limbs != ev.limbs() ||
            (std::size_t(1) << ha.log_n) != ev.poly_degree() || hk.log_n != ha.log_n || hk.n_limbs != ha.n_limbs)
            throw std::runtime_error("fhe job: parameter mismatch between files and evaluator");
        for (unsigned l = 0; l < ev.limbs(); ++l)
            if (ha.moduli[l] != ev.modulus(l) || hb.moduli[l] != ev.modulus(l) || hk.moduli[l] != ev.modulus(l))
                throw std::runtime_error("fhe job: moduli differ from the evaluator's");
        const std::size_t ct_words = 2 * ev.limbs() * ev.poly_degree();
        if (a.size() != ha.count * ct_words || b.size() != a.size() || k.size() != std::size_t(2) * ev.limbs() * ev.limbs() * ev.poly_degree())
            throw std::runtime_error("fhe job: payload sizes do not match the headers");
        std::vector<std::uint64_t> out(a.size());
        int devices = 1;
        if (get_config("fhe.devices") == "all") {
            fhe::MultiEvaluator &mev = fhe_multi_evaluator();   // contiguous shards, one GPU each, no collective
            devices = mev.device_count();
            mev.multiply_relin({a.data(), (std::size_t)ha.count}, {b.data(), (std::size_t)ha.count}, k.data(), {out.data(), (std::size_t)ha.count});
        } else {
            ev.multiply_relin({a.data(), (std::size_t)ha.count}, {b.data(), (std::size_t)ha.count}, k.data(), {out.data(), (std::size_t)ha.count});
        }
        fhe::write_wire_file(fo, ha, out.data());
        GenerationResult r;
        r.generation_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::ostringstream msg;
        msg << "wrote " << fo << ": " << ha.count << " ciphertext products (N=" << ev.poly_degree() << ", L=" << ev.limbs() << ", " << devices
            << (devices == 1 ? " GPU)" : " GPUs)");
        r.texts.push_back(msg.str());
        r.stop_reasons = std::vector<std::string>{"fhe_job_done"};
        return r;
    }

    // relinearisation key with special primes (wire kinds 4 and 5): the evaluator's basis is the key's (ciphertext moduli + K special
    // primes), the ciphertext files carry the first limbs() - K moduli; set_config("fhe.plain_modulus", t) selects the BGV rounding
    GenerationResult run_special_prime_job(fhe::Evaluator &ev, const fhe::WireHeader &ha, const fhe::WireHeader &hb, const fhe::WireHeader &hk,
                                           const std::vector<std::uint64_t> &a, const std::vector<std::uint64_t> &b, const std::vector<std::uint64_t> &k,
                                           const std::string &fo, std::chrono::steady_clock::time_point t0) {
        const unsigned special = hk.kind == (std::uint32_t)fhe::WireKind::HybridSwitchKey ? 1u : (unsigned)hk.count;
        if (ha.kind != (std::uint32_t)fhe::WireKind::Ciphertexts || hb.kind != ha.kind) throw std::runtime_error("fhe job: wrong file kinds");
        if (hk.n_limbs != ev.limbs() || (std::size_t(1) << hk.log_n) != ev.poly_degree() || ha.log_n != hk.log_n || hb.log_n != hk.log_n ||
            ha.n_limbs + special != hk.n_limbs || hb.n_limbs != ha.n_limbs || ha.count != hb.count)
            throw std::runtime_error("fhe job: parameter mismatch between files and evaluator");
        for (unsigned l = 0; l < ev.limbs(); ++l)
            if (hk.moduli[l] != ev.modulus(l) || (l < ha.n_limbs && (ha.moduli[l] != ev.modulus(l) || hb.moduli[l] != ev.modulus(l))))
                throw std::runtime_error("fhe job: moduli differ from the evaluator's");
        const std::size_t ct_words = 2 * (std::size_t)ha.n_limbs * ev.poly_degree();
        if (a.size() != ha.count * ct_words || b.size() != a.size() ||
            k.size() != std::size_t(2) * ev.grouped_digits(special) * ev.limbs() * ev.poly_degree())
            throw std::runtime_error("fhe job: payload sizes do not match the headers");
        const std::uint64_t t = std::stoull(get_config("fhe.plain_modulus"));
        std::vector<std::uint64_t> out(a.size());
        ev.multiply_relin_grouped(special, {a.data(), (std::size_t)ha.count}, {b.data(), (std::size_t)ha.count}, k.data(), {out.data(), (std::size_t)ha.count}, t);
        fhe::write_wire_file(fo, ha, out.data());
        GenerationResult r;
        r.generation_time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::ostringstream msg;
        msg << "wrote " << fo << ": " << ha.count << " ciphertext products (N=" << ev.poly_degree() << ", " << ha.n_limbs << " limbs + " << special
            << " special prime" << (special == 1 ? "" : "s") << ", 1 GPU)";
        r.texts.push_back(msg.str());
        r.stop_reasons = std::vector<std::string>{"fhe_job_done"};
        return r;
    }

    std::string path_, device_ = "cuda";
    std::map<std::string, std::string> config_;
    QuantizationConfig quant_;
    bool quant_applied_ = false;
    std::unique_ptr<fhe::Evaluator> evaluator_;
    std::unique_ptr<fhe::MultiEvaluator> multi_;
};

inline std::shared_ptr<Model> load_model(const std::string &model_path) { return std::make_shared<Model>(model_path); }
inline std::vector<std::string> list_available_models() { return {"gpt2"}; }
inline bool is_model_available(const std::string &name) { return name == "gpt2"; }
inline std::string version() { return "0.1.0"; }
inline std::string cuda_version() { return "12.9"; }
inline bool cuda_available() {
    try {
        fhe::EncryptionParameters p;
        p.log_n = 12;
        p.n_limbs = 1;
        fhe::Evaluator probe(p);
        return true;
    } catch (const std::exception &) {
        return false;
    }
}
inline size_t cuda_device_count() {
    int n = 0;
    return dpfhe_device_count(&n) == 0 && n > 0 ? (size_t)n : 0;
}

}  // namespace api
}  // namespace deeppowers
