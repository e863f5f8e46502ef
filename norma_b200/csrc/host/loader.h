// loader.h — the checkpoint side files norma reads before a `Model` exists, in C++ because the Rust
// toolchain is absent.  Mirrors what `Definition::blocking_try_to_model` does AFTER hf-hub has fetched the files
// (/root/reference/src/models/whisper/monolingual.rs:347-430, multilingual.rs:225-300):
//   config.json        -> candle `whisper::Config` (serde, `suppress_tokens` defaults to [])
//   tokenizer.json     -> `tokenizers::Tokenizer::from_file`; only `token_to_id` and `decode(ids, skip_special_tokens)` are used
//                         (mod.rs:86-90, model.rs:147,207); tokenizers 0.20.0 semantics are restated in loader.cc
//   model.safetensors  -> `VarBuilder::from_mmaped_safetensors(.., m::DTYPE = F32, ..)`: every tensor converted to f32 (safetensors 0.4.5 layout)
//   whisper_mel_bytes  -> the Slaney filterbank norma embeds as raw bytes, regenerated (see norma_b200/filters.py)
// The download itself (hf-hub, network) is out of scope.
#pragma once
#include <stdint.h>

#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../../include/norma_b200.h"

namespace nb200host {

// the 99 `Language` variants in declaration order (languages.rs:7-107) as their tokens (languages.rs:120-222)
extern const char *const LANGUAGE_CODES[99];
std::string language_token(size_t i);  // "<|en|>", ...

class Tokenizer {
public:
    // Tokenizer::from_file: returns false and sets *err (Error::LoadTokenizer)
    bool load(const std::string &path, std::string *err);
    bool parse(const char *json, size_t n, std::string *err);
    // added vocabulary first, then the model vocabulary (tokenizers `TokenizerImpl::token_to_id`)
    bool token_to_id(const std::string &token, uint32_t *id) const;
    const std::string *id_to_token(uint32_t id) const;
    bool is_special(const std::string &token) const { return special_.count(token) != 0; }
    // `decode(ids, skip_special_tokens)`: ids without a token are dropped; ByteLevel decoder (or " ".join when there is none)
    std::string decode(const uint32_t *ids, size_t n, bool skip_special) const;
    size_t vocab_size() const { return n_vocab_; }

private:
    std::unordered_map<std::string, uint32_t> model_vocab_, added_vocab_;
    std::unordered_map<uint32_t, std::string> added_by_id_;
    std::vector<std::string> model_by_id_;
    std::vector<uint8_t> model_has_;
    std::unordered_set<std::string> special_;
    bool byte_level_ = false;
    size_t n_vocab_ = 0;
};

// String::from_utf8_lossy: invalid sequences become U+FFFD, one per maximal invalid subpart
std::string utf8_lossy(const std::string &bytes);

struct SafetensorsEntry {
    std::string name;
    nb200_dtype dtype;
    std::vector<int64_t> shape;
    size_t begin, end;  // byte range inside the data section
};

class Safetensors {  // read-only mmap of one .safetensors file
public:
    ~Safetensors();
    bool open(const std::string &path, std::string *err);
    const std::vector<SafetensorsEntry> &entries() const { return entries_; }
    const uint8_t *data(const SafetensorsEntry &e) const { return base_ + data_off_ + e.begin; }

private:
    uint8_t *base_ = nullptr;
    size_t size_ = 0, data_off_ = 0;
    std::vector<SafetensorsEntry> entries_;
};

// model-{ext}-q80.gguf of the `Quantized*` model types (monolingual.rs:198-203, 364-369): GGUF v2 / v3, tensors of type F32, F16 or
// Q8_0 (blocks of 32 weights: f16 scale + 32 int8).  Every tensor is DEQUANTISED to f32 at load and then takes the same bf16 path as a
// safetensors checkpoint.  candle instead keeps the weights in q8_0 and multiplies them with activations quantised to q8 blocks
// (`QMatMul`): a different arithmetic whose own quantisation noise (~0.4 % per activation block) is larger than the bf16 rounding here,
// so parity with it is "within its quantisation error", not bit-level.
struct GgufEntry {
    std::string name;
    uint32_t type;               // 0 F32, 1 F16, 8 Q8_0
    std::vector<int64_t> shape;  // row-major (GGUF stores the dimensions innermost first)
    size_t offset, numel;
};

class Gguf {
public:
    ~Gguf();
    bool open(const std::string &path, std::string *err);
    const std::vector<GgufEntry> &entries() const { return entries_; }
    bool dequantize(const GgufEntry &e, float *out, std::string *err) const;  // numel floats

private:
    uint8_t *base_ = nullptr;
    size_t size_ = 0, data_off_ = 0;
    std::vector<GgufEntry> entries_;
};

bool config_from_json(const char *json, size_t n, nb200_config *cfg, std::vector<uint32_t> *suppress, std::string *err);
bool read_file(const std::string &path, std::string *out, std::string *err);
bool mel_filterbank(int n_mel, std::vector<float> *out);  // [n_mel][201]; false unless n_mel is 80 or 128 (Error::MelBins)

}  // namespace nb200host
