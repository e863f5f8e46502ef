// whisper_host.cc — see whisper_host.h.  Control flow follows /root/reference/src/models/whisper/model.rs:55-191
// statement by statement (comments give the reference lines).
#include "whisper_host.h"

#include <math.h>
#include <string.h>

#include <algorithm>

namespace nb200host {

std::vector<std::pair<size_t, size_t>> inclusive_boxed_by(const std::vector<uint32_t> &v, uint32_t nts, uint32_t eot) {
    // utils.rs:30-54: find the first hit, then the next hit after it; yield [s_idx, e_idx] inclusive and continue
    // AFTER the end element; stop as soon as either search fails.
    auto pred = [&](uint32_t t) { return t > nts || t == eot; };  // model.rs:100-102
    std::vector<std::pair<size_t, size_t>> out;
    size_t base = 0;
    while (base < v.size()) {
        size_t s = base;
        while (s < v.size() && !pred(v[s])) ++s;
        if (s >= v.size()) break;
        size_t e = s + 1;
        while (e < v.size() && !pred(v[e])) ++e;
        if (e >= v.size()) break;
        out.emplace_back(s, e + 1);
        base = e + 1;
    }
    return out;
}

WhisperModel::WhisperModel(Backend *backend, const nb200_special_tokens &tok, size_t max_chunk_len) : be_(backend), tok_(tok) {
    buf_.reserve(max_chunk_len);  // monolingual.rs:434
    lang_ = tok.lang;             // LanguageState::ConstLang(lang_token), monolingual.rs:449
}

void WhisperModel::set_vocab(uint32_t id, const std::string &bytes) {
    if (vocab_.size() <= id) vocab_.resize((size_t)id + 1);
    vocab_[id] = bytes;
}

std::string WhisperModel::detokenize(const uint32_t *t, size_t n) const {
    if (has_tokenizer_) return tokenizer_.decode(t, n, true);  // self.tokenizer.decode(.., true), model.rs:147
    std::string s;
    for (size_t i = 0; i < n; ++i) {
        if (t[i] >= tok_.eot) continue;  // skip_special_tokens: every Whisper special id is >= <|endoftext|>
        if (t[i] < vocab_.size()) s += vocab_[t[i]];
    }
    return s;
}

int WhisperModel::decode_with_fallback(bool *some, DecodingResult *out) {
    // model.rs:168 happened in the caller (encode)
    if (detect_ && !lang_set_) {                                                           // model.rs:170 `self.lang.is_none()`
        uint32_t lang = UINT32_MAX;
        int st = be_->detect_language(lang_tokens_, &lang);                                // model.rs:171
        ++n_detects;
        if (st != NB200_OK) { err_ = be_->last_error(); return st; }
        lang_ = lang;                                                                      // model.rs:172
        lang_set_ = true;
        st = be_->set_language(lang);
        if (st != NB200_OK) { err_ = be_->last_error(); return st; }
    }
    for (double t : TEMPERATURES) {                                                        // model.rs:175
        DecodingResult dr;
        int st = be_->decode(t, &dr);                                                      // model.rs:176
        ++n_decodes;
        if (st != NB200_OK) { err_ = be_->last_error(); return st; }
        const bool needs_fallback = dr.compression_ratio > COMPRESSION_RATIO_THRESHOLD     // NaN > x is false
                                    || dr.avg_logprob < LOGPROB_THRESHOLD;                 // model.rs:177-178
        if (!needs_fallback || dr.no_speech_prob > NO_SPEECH_THRESHOLD) {                  // model.rs:179
            *some = true;
            *out = std::move(dr);
            return NB200_OK;
        }
    }
    *some = false;                                                                         // model.rs:189-190
    return NB200_OK;
}

int WhisperModel::transcribe(const float *data, size_t n, bool final_chunk, std::string *text, std::vector<std::vector<uint32_t>> *segments) {
    buf_.insert(buf_.end(), data, data + n);                                               // model.rs:60-64
    std::string res;
    bool new_chunk_break = false;
    while (!buf_.empty() && !new_chunk_break) {                                            // model.rs:68
        const size_t len_before = buf_.size();
        const size_t slice_len = std::min(buf_.size(), N_SAMPLES);                         // model.rs:69
        int st = be_->encode(buf_.data(), slice_len);                                      // model.rs:74-88 + 168
        ++n_encodes;
        if (st != NB200_OK) { err_ = be_->last_error(); return st; }
        bool some = false;
        DecodingResult dr;
        st = decode_with_fallback(&some, &dr);                                             // model.rs:90
        if (st != NB200_OK) return st;
        if (!some) {                                                                       // model.rs:90-93
            buf_.erase(buf_.begin(), buf_.begin() + slice_len);
            continue;
        }
        if (dr.no_speech_prob > NO_SPEECH_THRESHOLD && dr.avg_logprob < LOGPROB_THRESHOLD) {  // model.rs:95-98
            buf_.erase(buf_.begin(), buf_.begin() + slice_len);
            continue;
        }
        for (auto &se : inclusive_boxed_by(dr.tokens, tok_.no_timestamps, tok_.eot)) {     // model.rs:100-102
            const uint32_t *tokens = dr.tokens.data() + se.first;
            const size_t len = se.second - se.first;
            const uint32_t s_timestamp = tokens[0] - tok_.no_timestamps - 1u;              // model.rs:103 (u32, wraps like release Rust)
            const uint32_t e_timestamp_token = tokens[len - 1];                            // model.rs:105
            if (e_timestamp_token == tok_.eot) {                                           // model.rs:107
                if (s_timestamp == 0 || final_chunk) {                                     // model.rs:108
                    if (slice_len == N_SAMPLES || final_chunk) {                           // model.rs:109
                        buf_.erase(buf_.begin(), buf_.begin() + slice_len);                // model.rs:110
                    } else {
                        new_chunk_break = true;                                            // model.rs:122
                        break;
                    }
                } else {
                    const size_t pre_drain_len = buf_.size();                              // model.rs:125
                    const size_t drain = std::min((size_t)s_timestamp * 320, slice_len);   // model.rs:126-127
                    buf_.erase(buf_.begin(), buf_.begin() + drain);
                    if (pre_drain_len > slice_len) break;                                  // model.rs:129-136: get a new slice
                    new_chunk_break = true;                                                // model.rs:143
                    break;
                }
            }
            if (segments) segments->emplace_back(tokens, tokens + len);
            res += detokenize(tokens + 1, len >= 2 ? len - 2 : 0);                         // model.rs:147-149
        }
        // DEVIATION (documented in DESIGN.md / INTEGRATION.md): a decoding result that yields no drainable segment — the silent window,
        // where decode() returns the prompt alone with avg_logprob = 0 (model.rs:308-315), passes the gate of model.rs:95 because 0 is not
        // < -1 — leaves the reference looping forever on the same slice.  A hang cannot be mirrored; the window is dropped like the
        // no-speech skip of model.rs:95-98 does, so a real-time stream keeps going after silence.
        if (!new_chunk_break && buf_.size() == len_before) {
            ++n_no_progress;
            buf_.erase(buf_.begin(), buf_.begin() + slice_len);
        }
    }
    if (final_chunk) {                                                                     // model.rs:153-156
        if (detect_) { lang_set_ = false; lang_ = UINT32_MAX; }                            // model.rs:154 `self.lang.clear()`
        int st = be_->reset_kv_cache();                                                    // model.rs:155
        if (st != NB200_OK) { err_ = be_->last_error(); return st; }
    }
    if (text) *text = res;
    return NB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
int Nb200Backend::encode(const float *pcm, size_t n) {
    size_t n_len = 0;
    static const float zero = 0.f;
    int st = nb200_pcm_to_mel(ctx_, n ? pcm : &zero, n, nullptr, &n_len);  // mel stays on the device as window 0
    if (st != NB200_OK) return st;
    return nb200_encoder_forward(ctx_, nullptr, 1, nullptr);
}

int Nb200Backend::decode(double t, DecodingResult *out) {
    int64_t maxpos = 0;
    nb200_query(ctx_, NB200_Q_MAX_TARGET_POSITIONS, &maxpos);
    std::vector<uint32_t> toks((size_t)maxpos);
    size_t n = 0;
    double alp = 0, nsp = 0;
    int st = nb200_decode(ctx_, 1, (float)t, seed_ + 0x9E3779B97F4A7C15ull * (++draws_), 0, toks.data(), &n, &alp, &nsp);
    if (st != NB200_OK) return st;
    out->tokens.assign(toks.begin(), toks.begin() + n);
    out->avg_logprob = alp;
    out->no_speech_prob = nsp;
    out->compression_ratio = NAN;
    return NB200_OK;
}

int Nb200Backend::detect_language(const std::vector<uint32_t> &lang_tokens, uint32_t *token) {
    return nb200_detect_language(ctx_, 0, lang_tokens.data(), lang_tokens.size(), token, nullptr);
}

int Nb200Backend::set_language(uint32_t token) {
    tok_.lang = token;
    return nb200_set_tokens(ctx_, &tok_);
}

int Nb200Backend::reset_kv_cache() { return nb200_reset_kv_cache(ctx_); }
std::string Nb200Backend::last_error() { return nb200_last_error(ctx_); }

int ScriptedBackend::encode(const float *, size_t n) {
    encode_lens.push_back(n);
    return NB200_OK;
}

int ScriptedBackend::detect_language(const std::vector<uint32_t> &lang_tokens, uint32_t *token) {
    if (language_script.empty()) {
        err_ = "scripted backend ran out of detected languages";
        return NB200_INVALID_ARG;
    }
    *token = language_script.front();
    language_script.pop_front();
    if (std::find(lang_tokens.begin(), lang_tokens.end(), *token) == lang_tokens.end()) {
        err_ = "scripted language is not one of the language tokens";
        return NB200_INVALID_ARG;
    }
    return NB200_OK;
}

int ScriptedBackend::decode(double t, DecodingResult *out) {
    decode_temps.push_back(t);
    if (script.empty()) {
        err_ = "scripted backend ran out of decoding results";
        return NB200_INVALID_ARG;
    }
    *out = script.front();
    script.pop_front();
    out->compression_ratio = NAN;
    return NB200_OK;
}

}  // namespace nb200host

// ---------------------------------------------------------------------------------------------------------
// C ABI (declared in include/norma_b200.h)
// ---------------------------------------------------------------------------------------------------------
extern "C" {

int nb200_model_create(nb200_ctx *ctx, const nb200_special_tokens *tok, size_t max_chunk_len, uint64_t seed, nb200_model **out) {
    if (!tok || !out) return NB200_INVALID_ARG;
    nb200_model *m = new nb200_model();
    if (ctx) m->backend = new nb200host::Nb200Backend(ctx, *tok, seed);
    else m->backend = m->scripted = new nb200host::ScriptedBackend();
    m->model = new nb200host::WhisperModel(m->backend, *tok, max_chunk_len);
    *out = m;
    return NB200_OK;
}

void nb200_model_destroy(nb200_model *m) {
    if (!m) return;
    delete m->model;
    delete m->backend;
    delete m;
}

const char *nb200_model_last_error(nb200_model *m) {
    if (!m) return "";
    m->err = m->model->last_error();
    return m->err.c_str();
}

int nb200_model_script_push(nb200_model *m, double avg_logprob, double no_speech_prob, const uint32_t *tokens, size_t n) {
    if (!m || !m->scripted || (n && !tokens)) return NB200_INVALID_ARG;
    nb200host::DecodingResult dr;
    dr.tokens.assign(tokens, tokens + n);
    dr.avg_logprob = avg_logprob;
    dr.no_speech_prob = no_speech_prob;
    m->scripted->script.push_back(std::move(dr));
    return NB200_OK;
}

int nb200_model_set_vocab(nb200_model *m, uint32_t id, const char *bytes, size_t n) {
    if (!m || (n && !bytes)) return NB200_INVALID_ARG;
    m->model->set_vocab(id, std::string(bytes, n));
    return NB200_OK;
}

int nb200_model_transcribe(nb200_model *m, const float *data, size_t n, int final_chunk, char *text_out, size_t text_cap, size_t *text_len,
                           uint32_t *seg_out, size_t seg_cap, size_t *seg_len) {
    if (!m || (n && !data)) return NB200_INVALID_ARG;
    std::string text;
    std::vector<std::vector<uint32_t>> segs;
    int st = m->model->transcribe(data, n, final_chunk != 0, &text, &segs);
    if (st != NB200_OK) return st;
    // The audio has been consumed, so the result is kept until it has been delivered in full: a caller whose buffers were too small
    // gets NB200_BUFFER_TOO_SMALL with the needed sizes in *text_len / *seg_len and fetches it with nb200_model_last_result.
    m->last_text = std::move(text);
    m->last_flat.clear();
    m->last_flat.push_back((uint32_t)segs.size());  // flattened segments: [n_segments, len_0, tokens_0..., len_1, tokens_1..., ...]
    for (auto &s : segs) {
        m->last_flat.push_back((uint32_t)s.size());
        m->last_flat.insert(m->last_flat.end(), s.begin(), s.end());
    }
    return nb200_model_last_result(m, text_out, text_cap, text_len, seg_out, seg_cap, seg_len);
}

int nb200_model_last_result(nb200_model *m, char *text_out, size_t text_cap, size_t *text_len, uint32_t *seg_out, size_t seg_cap, size_t *seg_len) {
    if (!m) return NB200_INVALID_ARG;
    if (text_len) *text_len = m->last_text.size();
    if (seg_len) *seg_len = m->last_flat.size();
    const bool text_fits = !text_out || m->last_text.size() + 1 <= text_cap;
    const bool seg_fits = !seg_out || m->last_flat.size() <= seg_cap;
    if (!text_fits || !seg_fits) {
        if (text_out && text_cap) text_out[0] = 0;
        return NB200_BUFFER_TOO_SMALL;  // nothing partial is written: no text cut inside a UTF-8 sequence, no half segment list
    }
    if (text_out) {
        memcpy(text_out, m->last_text.data(), m->last_text.size());
        text_out[m->last_text.size()] = 0;
    }
    if (seg_out) memcpy(seg_out, m->last_flat.data(), m->last_flat.size() * 4);
    return NB200_OK;
}

int nb200_model_state(nb200_model *m, size_t *buffered, size_t *n_encodes, size_t *n_decodes, size_t *n_resets) {
    if (!m) return NB200_INVALID_ARG;
    if (buffered) *buffered = m->model->buffered();
    if (n_encodes) *n_encodes = m->model->n_encodes;
    if (n_decodes) *n_decodes = m->model->n_decodes;
    if (n_resets) *n_resets = m->scripted ? (size_t)m->scripted->resets : 0;
    return NB200_OK;
}

int nb200_model_no_progress_windows(nb200_model *m, size_t *n) {
    if (!m || !n) return NB200_INVALID_ARG;
    *n = m->model->n_no_progress;
    return NB200_OK;
}

int nb200_model_set_language_detection(nb200_model *m, const uint32_t *lang_tokens, size_t n) {
    if (!m || !lang_tokens || n == 0) return NB200_INVALID_ARG;
    m->model->set_language_detection(std::vector<uint32_t>(lang_tokens, lang_tokens + n));
    return m->backend->set_language(UINT32_MAX);
}

int nb200_model_language(nb200_model *m, uint32_t *token, size_t *n_detects) {
    if (!m) return NB200_INVALID_ARG;
    uint32_t t = UINT32_MAX;
    const bool some = m->model->language(&t);
    if (token) *token = some ? t : UINT32_MAX;
    if (n_detects) *n_detects = m->model->n_detects;
    return NB200_OK;
}

int nb200_model_script_push_language(nb200_model *m, uint32_t token) {
    if (!m || !m->scripted) return NB200_INVALID_ARG;
    m->scripted->language_script.push_back(token);
    return NB200_OK;
}

// scripted-backend introspection: slice length of encode call i / temperature of decode call i
int nb200_model_script_log(nb200_model *m, size_t i, size_t *encode_len, double *decode_temp) {
    if (!m || !m->scripted) return NB200_INVALID_ARG;
    if (encode_len) *encode_len = i < m->scripted->encode_lens.size() ? m->scripted->encode_lens[i] : (size_t)-1;
    if (decode_temp) *decode_temp = i < m->scripted->decode_temps.size() ? m->scripted->decode_temps[i] : -1.0;
    return NB200_OK;
}

}  // extern "C"
