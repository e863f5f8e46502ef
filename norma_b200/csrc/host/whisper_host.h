// whisper_host.h — host side of norma's Whisper `Model` above the C ABI, in C++ because the reference's Rust
// toolchain is absent from this image (INTEGRATION.md shows the Rust it stands in for).  Mirrors, line by line:
//   Model::transcribe            /root/reference/src/models/whisper/model.rs:55-160
//   Model::decode_with_fallback  /root/reference/src/models/whisper/model.rs:164-191
//   Model::detect_language       /root/reference/src/models/whisper/model.rs:194-210
//   LanguageState                /root/reference/src/models/whisper/model.rs:392-440
//   SliceExt::inclusive_boxed_by /root/reference/src/utils.rs:1-76
// The device work (pcm_to_mel, encoder, decode) goes through a `Backend`; `Nb200Backend` calls the C ABI,
// `ScriptedBackend` replays canned DecodingResults so the buffering / seek / fallback logic is testable without a GPU.
#pragma once
#include <stdint.h>

#include <deque>
#include <string>
#include <vector>

#include "../../../include/norma_b200.h"
#include "loader.h"

namespace nb200host {

// candle constants norma uses (model.rs:69,88,95,175-179)
constexpr size_t N_SAMPLES = 480000;
constexpr double NO_SPEECH_THRESHOLD = 0.6;
constexpr double LOGPROB_THRESHOLD = -1.0;
constexpr double COMPRESSION_RATIO_THRESHOLD = 2.4;
constexpr double TEMPERATURES[6] = {0.0, 0.2, 0.4, 0.6, 0.8, 1.0};

struct DecodingResult {  // model.rs:493-499
    std::vector<uint32_t> tokens;
    double avg_logprob = 0.0;
    double no_speech_prob = 0.0;
    double compression_ratio = 0.0;  // the reference never computes it: always NaN (model.rs:313,387)
};

struct Backend {
    virtual ~Backend() {}
    // audio::pcm_to_mel + narrow(min(3000, n_len)) + encoder_forward(mel, flush = true)  (model.rs:74-88, 168)
    virtual int encode(const float *pcm, size_t n) = 0;
    // Model::decode(audio_features, t)  (model.rs:279-390)
    virtual int decode(double t, DecodingResult *out) = 0;
    // Model::detect_language(audio_features): id of the most probable of `lang_tokens`  (model.rs:194-210)
    virtual int detect_language(const std::vector<uint32_t> &lang_tokens, uint32_t *token) = 0;
    // the language token `decode` puts in its prompt (model.rs:286-288); UINT32_MAX = none
    virtual int set_language(uint32_t token) = 0;
    virtual int reset_kv_cache() = 0;
    virtual std::string last_error() = 0;
};

class Nb200Backend : public Backend {
public:
    Nb200Backend(nb200_ctx *ctx, const nb200_special_tokens &tok, uint64_t seed) : ctx_(ctx), tok_(tok), seed_(seed) {}
    int encode(const float *pcm, size_t n) override;
    int decode(double t, DecodingResult *out) override;
    int detect_language(const std::vector<uint32_t> &lang_tokens, uint32_t *token) override;
    int set_language(uint32_t token) override;
    int reset_kv_cache() override;
    std::string last_error() override;

private:
    nb200_ctx *ctx_;
    nb200_special_tokens tok_;
    uint64_t seed_, draws_ = 0;
};

class ScriptedBackend : public Backend {
public:
    int encode(const float *pcm, size_t n) override;
    int decode(double t, DecodingResult *out) override;
    int detect_language(const std::vector<uint32_t> &lang_tokens, uint32_t *token) override;
    int set_language(uint32_t token) override { languages_set.push_back(token); return NB200_OK; }
    int reset_kv_cache() override { ++resets; return NB200_OK; }
    std::string last_error() override { return err_; }
    std::deque<DecodingResult> script;
    std::deque<uint32_t> language_script;   // results of successive detect_language calls
    std::vector<uint32_t> languages_set;    // every set_language call
    std::vector<size_t> encode_lens;   // slice length of every encode call
    std::vector<double> decode_temps;  // temperature of every decode call
    int resets = 0;

private:
    std::string err_;
};

// `tokens.inclusive_boxed_by(pred)` (utils.rs): sub-slices [start, end) bounded inclusively by predicate hits
std::vector<std::pair<size_t, size_t>> inclusive_boxed_by(const std::vector<uint32_t> &v, uint32_t no_timestamps, uint32_t eot);

class WhisperModel {
public:
    WhisperModel(Backend *backend, const nb200_special_tokens &tok, size_t max_chunk_len);
    // Model::transcribe(&mut self, data: &mut Vec<f32>, final_chunk) -> Result<String, _>.  `data` is consumed
    // (the reference swaps / appends it into self.buf).  `segments` (optional) receives every emitted token slice.
    int transcribe(const float *data, size_t n, bool final_chunk, std::string *text, std::vector<std::vector<uint32_t>> *segments);
    int decode_with_fallback(bool *some, DecodingResult *out);
    void set_vocab(uint32_t id, const std::string &bytes);
    void set_tokenizer(const Tokenizer &t) { tokenizer_ = t; has_tokenizer_ = true; }
    // LanguageState::Detect { language_token: None, language_tokens_tensor } (multilingual.rs:319-322); the default is
    // LanguageState::ConstLang(tok.lang) (monolingual.rs:449)
    void set_language_detection(const std::vector<uint32_t> &lang_tokens) { detect_ = true; lang_tokens_ = lang_tokens; lang_set_ = false; }
    bool language(uint32_t *token) const { *token = lang_; return !detect_ || lang_set_; }  // LanguageState::language_token
    size_t n_detects = 0;
    size_t buffered() const { return buf_.size(); }
    std::string last_error() const { return err_; }
    size_t n_encodes = 0, n_decodes = 0;
    size_t n_no_progress = 0;  // windows dropped because their decoding result had no drainable segment (the reference hangs there)

private:
    std::string detokenize(const uint32_t *t, size_t n) const;  // tokenizer.decode(.., skip_special_tokens = true)
    Backend *be_;
    nb200_special_tokens tok_;
    std::vector<float> buf_;
    std::vector<std::string> vocab_;
    Tokenizer tokenizer_;
    bool has_tokenizer_ = false;
    bool detect_ = false, lang_set_ = false;  // LanguageState (model.rs:392-440)
    uint32_t lang_ = UINT32_MAX;
    std::vector<uint32_t> lang_tokens_;
    std::string err_;
};

}  // namespace nb200host

struct nb200_model {
    nb200host::Backend *backend = nullptr;
    nb200host::ScriptedBackend *scripted = nullptr;
    nb200host::WhisperModel *model = nullptr;
    std::string err;
    std::string last_text;             // result of the last transcribe, kept until delivered in full (nb200_model_last_result)
    std::vector<uint32_t> last_flat;
};
