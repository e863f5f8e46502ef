// loader.cc — see loader.h.  Host-only code (no kernels); compiled into libnorma_b200.so with the rest.
#include "loader.h"

#include <fcntl.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>

#include "json.h"
#include "whisper_host.h"

int nb200_fail(nb200_ctx *ctx, int code, const char *fmt, ...);  // api.cu: records the message on the ctx (or the thread when ctx == NULL)

namespace nb200host {

const char *const LANGUAGE_CODES[99] = {
    "en", "zh", "de", "es", "ru", "ko", "fr", "ja", "pt", "tr", "pl", "ca", "nl", "ar", "sv", "it", "id", "hi", "fi", "vi",
    "he", "uk", "el", "ms", "cs", "ro", "da", "hu", "ta", "no", "th", "ur", "hr", "bg", "lt", "la", "mi", "ml", "cy", "sk",
    "te", "fa", "lv", "bn", "sr", "az", "sl", "kn", "et", "mk", "br", "eu", "is", "hy", "ne", "mn", "bs", "kk", "sq", "sw",
    "gl", "mr", "pa", "si", "km", "sn", "yo", "so", "af", "oc", "ka", "be", "tg", "sd", "gu", "am", "yi", "lo", "uz", "fo",
    "ht", "ps", "tk", "nn", "mt", "sa", "lb", "my", "bo", "tl", "mg", "as", "tt", "haw", "ln", "ha", "ba", "jw", "su"};

std::string language_token(size_t i) { return std::string("<|") + LANGUAGE_CODES[i] + "|>"; }

bool read_file(const std::string &path, std::string *out, std::string *err) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) { *err = "cannot open '" + path + "': " + strerror(errno); return false; }
    out->clear();
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out->append(buf, n);
    bool ok = !ferror(f);
    fclose(f);
    if (!ok) *err = "read error on '" + path + "'";
    return ok;
}

// ---- config.json ------------------------------------------------------------------------------------------------------
bool config_from_json(const char *json, size_t n, nb200_config *cfg, std::vector<uint32_t> *suppress, std::string *err) {
    nb200json::Value v;
    if (!nb200json::Parser(json, n).parse(&v, err)) return false;
    if (v.kind != nb200json::Value::Object) { *err = "config.json: expected an object"; return false; }
    struct { const char *key; int32_t *dst; } req[] = {
        {"num_mel_bins", &cfg->num_mel_bins}, {"max_source_positions", &cfg->max_source_positions}, {"d_model", &cfg->d_model},
        {"encoder_attention_heads", &cfg->encoder_attention_heads}, {"encoder_layers", &cfg->encoder_layers}, {"vocab_size", &cfg->vocab_size},
        {"max_target_positions", &cfg->max_target_positions}, {"decoder_attention_heads", &cfg->decoder_attention_heads},
        {"decoder_layers", &cfg->decoder_layers}};
    for (auto &r : req) {  // serde: every field but suppress_tokens is required and must be an unsigned integer
        const nb200json::Value *f = v.get(r.key);
        if (!f) { *err = std::string("missing field `") + r.key + "`"; return false; }
        if (f->kind != nb200json::Value::Number || !f->is_int || f->inum < 0 || f->inum > INT32_MAX) {
            *err = std::string("invalid type for `") + r.key + "`, expected usize";
            return false;
        }
        *r.dst = (int32_t)f->inum;
    }
    suppress->clear();
    if (const nb200json::Value *s = v.get("suppress_tokens")) {  // #[serde(default)]
        if (s->kind != nb200json::Value::Array) { *err = "invalid type for `suppress_tokens`, expected a sequence"; return false; }
        for (auto &e : s->arr) {
            if (e.kind != nb200json::Value::Number || !e.is_int || e.inum < 0 || e.inum > (int64_t)UINT32_MAX) {
                *err = "invalid element in `suppress_tokens`, expected u32";
                return false;
            }
            suppress->push_back((uint32_t)e.inum);
        }
    }
    return true;
}

// ---- mel filterbank -----------------------------------------------------------------------------------------------------
bool mel_filterbank(int n_mel, std::vector<float> *out) {
    if (n_mel != 80 && n_mel != 128) return false;  // monolingual.rs:351-355
    const int n_freq = 201;
    const double f_sp = 200.0 / 3.0, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = log(6.4) / 27.0;
    auto hz_to_mel = [&](double f) { return f >= min_log_hz ? min_log_mel + log(f / min_log_hz) / logstep : f / f_sp; };
    auto mel_to_hz = [&](double m) { return m >= min_log_mel ? min_log_hz * exp(logstep * (m - min_log_mel)) : m * f_sp; };
    std::vector<double> freqs(n_freq), pts(n_mel + 2);
    for (int i = 0; i < n_freq; ++i) freqs[i] = i == n_freq - 1 ? 8000.0 : i * (8000.0 / (n_freq - 1));
    const double m0 = hz_to_mel(0.0), m1 = hz_to_mel(8000.0), step = (m1 - m0) / (n_mel + 1);
    for (int i = 0; i < n_mel + 2; ++i) pts[i] = mel_to_hz(i == n_mel + 1 ? m1 : m0 + i * step);
    out->assign((size_t)n_mel * n_freq, 0.f);
    for (int m = 0; m < n_mel; ++m) {
        const double enorm = 2.0 / (pts[m + 2] - pts[m]);
        for (int k = 0; k < n_freq; ++k) {
            const double lower = -(pts[m] - freqs[k]) / (pts[m + 1] - pts[m]);
            const double upper = (pts[m + 2] - freqs[k]) / (pts[m + 2] - pts[m + 1]);
            (*out)[(size_t)m * n_freq + k] = (float)(std::max(0.0, std::min(lower, upper)) * enorm);
        }
    }
    return true;
}

// ---- UTF-8 ------------------------------------------------------------------------------------------------------------
std::string utf8_lossy(const std::string &s) {
    static const char REPL[] = "\xEF\xBF\xBD";
    std::string out;
    out.reserve(s.size());
    const size_t n = s.size();
    size_t i = 0;
    auto at = [&](size_t k) -> int { return k < n ? (unsigned char)s[k] : -1; };
    auto cont = [](int b) { return b >= 0x80 && b <= 0xBF; };
    while (i < n) {
        const int b0 = (unsigned char)s[i];
        if (b0 < 0x80) { out += (char)b0; ++i; continue; }
        size_t good = 0, need = 0;  // `good` bytes form a valid prefix; `need` is the full width
        if (b0 >= 0xC2 && b0 <= 0xDF) {
            need = 2;
            good = 1 + (cont(at(i + 1)) ? 1 : 0);
        } else if (b0 >= 0xE0 && b0 <= 0xEF) {
            need = 3;
            const int b1 = at(i + 1);
            const bool ok1 = b0 == 0xE0 ? (b1 >= 0xA0 && b1 <= 0xBF) : b0 == 0xED ? (b1 >= 0x80 && b1 <= 0x9F) : cont(b1);
            good = 1;
            if (ok1) { good = 2; if (cont(at(i + 2))) good = 3; }
        } else if (b0 >= 0xF0 && b0 <= 0xF4) {
            need = 4;
            const int b1 = at(i + 1);
            const bool ok1 = b0 == 0xF0 ? (b1 >= 0x90 && b1 <= 0xBF) : b0 == 0xF4 ? (b1 >= 0x80 && b1 <= 0x8F) : cont(b1);
            good = 1;
            if (ok1) { good = 2; if (cont(at(i + 2))) { good = 3; if (cont(at(i + 3))) good = 4; } }
        } else {  // stray continuation byte, 0xC0/0xC1 or > 0xF4
            need = 0;
            good = 1;
        }
        if (need && good == need) out.append(s, i, need);
        else out += REPL;
        i += good;
    }
    return out;
}

// ---- tokenizer.json ---------------------------------------------------------------------------------------------------
namespace {
// GPT-2 `bytes_char`: printable bytes map to themselves, the other 68 to U+0100.. in byte order
struct ByteMap {
    int cp_to_byte[512];
    ByteMap() {
        for (int &x : cp_to_byte) x = -1;
        int extra = 0;
        for (int b = 0; b < 256; ++b) {
            const bool direct = (b >= 33 && b <= 126) || (b >= 161 && b <= 172) || (b >= 174 && b <= 255);
            cp_to_byte[direct ? b : 256 + extra++] = b;
        }
    }
};
const ByteMap &byte_map() {
    static const ByteMap m;
    return m;
}

// decode one UTF-8 code point of a string the JSON reader produced (always valid); returns its width
size_t next_cp(const std::string &s, size_t i, uint32_t *cp) {
    const unsigned char c = (unsigned char)s[i];
    if (c < 0x80) { *cp = c; return 1; }
    if (c < 0xE0) { *cp = ((c & 0x1Fu) << 6) | ((unsigned char)s[i + 1] & 0x3Fu); return 2; }
    if (c < 0xF0) { *cp = ((c & 0x0Fu) << 12) | (((unsigned char)s[i + 1] & 0x3Fu) << 6) | ((unsigned char)s[i + 2] & 0x3Fu); return 3; }
    *cp = ((c & 0x07u) << 18) | (((unsigned char)s[i + 1] & 0x3Fu) << 12) | (((unsigned char)s[i + 2] & 0x3Fu) << 6) | ((unsigned char)s[i + 3] & 0x3Fu);
    return 4;
}
}  // namespace

bool Tokenizer::load(const std::string &path, std::string *err) {
    std::string text;
    if (!read_file(path, &text, err)) return false;
    return parse(text.data(), text.size(), err);
}

bool Tokenizer::parse(const char *json, size_t n, std::string *err) {
    nb200json::Value v;
    if (!nb200json::Parser(json, n).parse(&v, err)) return false;
    const nb200json::Value *model = v.get("model");
    if (!model || model->kind != nb200json::Value::Object) { *err = "tokenizer.json: missing `model`"; return false; }
    const nb200json::Value *vocab = model->get("vocab");
    if (!vocab || vocab->kind != nb200json::Value::Object) { *err = "tokenizer.json: missing `model.vocab`"; return false; }
    model_vocab_.clear(); added_vocab_.clear(); added_by_id_.clear(); model_by_id_.clear(); model_has_.clear(); special_.clear();
    model_vocab_.reserve(vocab->obj.size() * 2);
    size_t max_id = 0;
    for (auto &kv : vocab->obj) {
        if (kv.second.kind != nb200json::Value::Number || !kv.second.is_int || kv.second.inum < 0 || kv.second.inum > (int64_t)UINT32_MAX) {
            *err = "tokenizer.json: non-integer id in `model.vocab`";
            return false;
        }
        model_vocab_[kv.first] = (uint32_t)kv.second.inum;
        max_id = std::max(max_id, (size_t)kv.second.inum);
    }
    model_by_id_.resize(vocab->obj.empty() ? 0 : max_id + 1);
    model_has_.assign(model_by_id_.size(), 0);
    for (auto &kv : vocab->obj) { model_by_id_[(size_t)kv.second.inum] = kv.first; model_has_[(size_t)kv.second.inum] = 1; }
    n_vocab_ = model_by_id_.size();
    if (const nb200json::Value *added = v.get("added_tokens")) {
        if (added->kind == nb200json::Value::Array) {
            for (auto &t : added->arr) {
                const nb200json::Value *id = t.get("id"), *content = t.get("content"), *special = t.get("special");
                if (!id || !content || id->kind != nb200json::Value::Number || !id->is_int || id->inum < 0 || content->kind != nb200json::Value::String) {
                    *err = "tokenizer.json: malformed entry in `added_tokens`";
                    return false;
                }
                added_vocab_[content->str] = (uint32_t)id->inum;
                added_by_id_[(uint32_t)id->inum] = content->str;
                if (special && special->kind == nb200json::Value::Bool && special->b) special_.insert(content->str);
                n_vocab_ = std::max(n_vocab_, (size_t)id->inum + 1);
            }
        } else if (!added->is_null()) { *err = "tokenizer.json: `added_tokens` is not a list"; return false; }
    }
    byte_level_ = false;
    const nb200json::Value *dec = v.get("decoder");
    if (dec && !dec->is_null()) {
        const nb200json::Value *type = dec->get("type");
        if (!type || type->kind != nb200json::Value::String || type->str != "ByteLevel") {
            *err = "tokenizer.json: only the ByteLevel decoder (Whisper's) is supported, got `" + (type && type->kind == nb200json::Value::String ? type->str : "?") + "`";
            return false;
        }
        byte_level_ = true;
    }
    return true;
}

bool Tokenizer::token_to_id(const std::string &token, uint32_t *id) const {
    auto a = added_vocab_.find(token);
    if (a != added_vocab_.end()) { *id = a->second; return true; }
    auto m = model_vocab_.find(token);
    if (m != model_vocab_.end()) { *id = m->second; return true; }
    return false;
}

const std::string *Tokenizer::id_to_token(uint32_t id) const {
    auto a = added_by_id_.find(id);
    if (a != added_by_id_.end()) return &a->second;
    if (id < model_by_id_.size() && model_has_[id]) return &model_by_id_[id];
    return nullptr;
}

std::string Tokenizer::decode(const uint32_t *ids, size_t n, bool skip_special) const {
    const ByteMap &bm = byte_map();
    std::string bytes;
    bool first = true;
    for (size_t i = 0; i < n; ++i) {
        const std::string *tok = id_to_token(ids[i]);
        if (!tok) continue;                               // ids unknown to both vocabularies are dropped
        if (skip_special && is_special(*tok)) continue;
        if (!byte_level_) {                               // no decoder: tokens.join(" ")
            if (!first) bytes += ' ';
            bytes += *tok;
            first = false;
            continue;
        }
        // ByteLevel::decode_chain: a token whose chars are ALL in the byte alphabet becomes those bytes, any other token its own UTF-8
        std::string mapped;
        bool all = true;
        for (size_t p = 0; p < tok->size();) {
            uint32_t cp;
            p += next_cp(*tok, p, &cp);
            const int b = cp < 512 ? bm.cp_to_byte[cp] : -1;
            if (b < 0) { all = false; break; }
            mapped += (char)b;
        }
        bytes += all ? mapped : *tok;
    }
    return byte_level_ ? utf8_lossy(bytes) : bytes;
}

// ---- safetensors ------------------------------------------------------------------------------------------------------
Safetensors::~Safetensors() {
    if (base_) munmap(base_, size_);
}

bool Safetensors::open(const std::string &path, std::string *err) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { *err = "cannot open '" + path + "': " + strerror(errno); return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 8) { ::close(fd); *err = "safetensors: '" + path + "' is shorter than its 8-byte header length"; return false; }
    size_ = (size_t)st.st_size;
    void *p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) { *err = std::string("safetensors: mmap failed: ") + strerror(errno); return false; }
    base_ = (uint8_t *)p;
    uint64_t hlen = 0;
    for (int i = 0; i < 8; ++i) hlen |= (uint64_t)base_[i] << (8 * i);  // little-endian u64
    if (hlen > 100000000ull || 8 + hlen > size_) { *err = "safetensors: header length " + std::to_string(hlen) + " is invalid for this file"; return false; }
    data_off_ = 8 + (size_t)hlen;
    nb200json::Value v;
    std::string jerr;
    if (!nb200json::Parser((const char *)base_ + 8, (size_t)hlen).parse(&v, &jerr)) { *err = "safetensors: invalid header JSON: " + jerr; return false; }
    if (v.kind != nb200json::Value::Object) { *err = "safetensors: header is not an object"; return false; }
    const size_t data_len = size_ - data_off_;
    for (auto &kv : v.obj) {
        if (kv.first == "__metadata__") continue;
        const nb200json::Value *dt = kv.second.get("dtype"), *sh = kv.second.get("shape"), *off = kv.second.get("data_offsets");
        if (!dt || !sh || !off || dt->kind != nb200json::Value::String || sh->kind != nb200json::Value::Array || off->kind != nb200json::Value::Array || off->arr.size() != 2) {
            *err = "safetensors: malformed entry '" + kv.first + "'";
            return false;
        }
        SafetensorsEntry e;
        e.name = kv.first;
        size_t esz;
        if (dt->str == "F32") { e.dtype = NB200_F32; esz = 4; }
        else if (dt->str == "BF16") { e.dtype = NB200_BF16; esz = 2; }
        else if (dt->str == "F16") { e.dtype = NB200_F16; esz = 2; }
        else if (dt->str == "F64") { e.dtype = NB200_F64; esz = 8; }
        else { *err = "safetensors: tensor '" + kv.first + "' has unsupported dtype " + dt->str; return false; }
        size_t count = 1;
        for (auto &d : sh->arr) {
            if (d.kind != nb200json::Value::Number || !d.is_int || d.inum < 0) { *err = "safetensors: bad shape for '" + kv.first + "'"; return false; }
            e.shape.push_back(d.inum);
            // overflow-checked: a crafted shape such as [2^62, 4] must not wrap round to a count that matches an empty byte range
            if (__builtin_mul_overflow(count, (size_t)d.inum, &count)) { *err = "safetensors: shape of '" + kv.first + "' overflows"; return false; }
        }
        size_t want_bytes = 0;
        if (__builtin_mul_overflow(count, esz, &want_bytes)) { *err = "safetensors: size of '" + kv.first + "' overflows"; return false; }
        if (off->arr[0].kind != nb200json::Value::Number || off->arr[1].kind != nb200json::Value::Number || off->arr[0].inum < 0 || off->arr[1].inum < off->arr[0].inum) {
            *err = "safetensors: bad data_offsets for '" + kv.first + "'";
            return false;
        }
        e.begin = (size_t)off->arr[0].inum;
        e.end = (size_t)off->arr[1].inum;
        if (e.end > data_len || e.end - e.begin != want_bytes) {  // safetensors' TensorInvalidInfo / MetadataIncompleteBuffer checks
            *err = "safetensors: tensor '" + kv.first + "' byte range does not match its dtype and shape";
            return false;
        }
        entries_.push_back(std::move(e));
    }
    return true;
}


// ---- GGUF ---------------------------------------------------------------------------------------------------------------
namespace {
struct Cursor {
    const uint8_t *p, *end;
    bool ok = true;
    template <typename T> T get() {
        T v{};
        if ((size_t)(end - p) < sizeof(T)) { ok = false; p = end; return v; }
        memcpy(&v, p, sizeof(T));
        p += sizeof(T);
        return v;
    }
    std::string str() {
        const uint64_t n = get<uint64_t>();
        if (!ok || n > (uint64_t)(end - p)) { ok = false; return std::string(); }
        std::string s((const char *)p, (size_t)n);
        p += n;
        return s;
    }
    void skip(uint64_t n) {
        if (n > (uint64_t)(end - p)) { ok = false; p = end; } else p += n;
    }
};
const int GGUF_SCALAR_SIZE[13] = {1, 1, 2, 2, 4, 4, 4, 1, 0, 0, 8, 8, 8};  // by value type; 8 = string, 9 = array

bool skip_value(Cursor &c, uint32_t type, uint32_t *alignment, bool is_alignment_key) {
    if (type == 8) { c.str(); return c.ok; }
    if (type == 9) {
        const uint32_t et = c.get<uint32_t>();
        const uint64_t n = c.get<uint64_t>();
        if (!c.ok || et > 12 || et == 9) return false;
        if (et == 8) { for (uint64_t i = 0; i < n && c.ok; ++i) c.str(); }
        else c.skip(n * (uint64_t)GGUF_SCALAR_SIZE[et]);
        return c.ok;
    }
    if (type > 12) return false;
    if (is_alignment_key && type == 4) { *alignment = c.get<uint32_t>(); return c.ok; }
    c.skip((uint64_t)GGUF_SCALAR_SIZE[type]);
    return c.ok;
}

float half_to_float(uint16_t h) {
    const uint32_t sign = (h >> 15) & 1u, ex = (h >> 10) & 0x1fu, man = h & 0x3ffu;
    float f;
    if (ex == 0) f = ldexpf((float)man, -24);
    else if (ex == 31) f = man ? NAN : INFINITY;
    else f = ldexpf((float)(man | 0x400u), (int)ex - 25);
    return sign ? -f : f;
}
}  // namespace

Gguf::~Gguf() {
    if (base_) munmap(base_, size_);
}

bool Gguf::open(const std::string &path, std::string *err) {
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { *err = "cannot open '" + path + "': " + strerror(errno); return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size < 24) { ::close(fd); *err = "gguf: '" + path + "' is too short"; return false; }
    size_ = (size_t)st.st_size;
    void *p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) { *err = std::string("gguf: mmap failed: ") + strerror(errno); return false; }
    base_ = (uint8_t *)p;
    Cursor c{base_, base_ + size_};
    if (c.get<uint32_t>() != 0x46554747u) { *err = "gguf: bad magic"; return false; }
    const uint32_t version = c.get<uint32_t>();
    if (version != 2 && version != 3) { *err = "gguf: unsupported version " + std::to_string(version); return false; }
    const uint64_t n_tensors = c.get<uint64_t>(), n_kv = c.get<uint64_t>();
    if (!c.ok || n_tensors > 1000000 || n_kv > 1000000) { *err = "gguf: implausible header counts"; return false; }
    uint32_t alignment = 32;
    for (uint64_t i = 0; i < n_kv; ++i) {
        const std::string key = c.str();
        const uint32_t type = c.get<uint32_t>();
        if (!c.ok || !skip_value(c, type, &alignment, key == "general.alignment")) { *err = "gguf: malformed metadata"; return false; }
    }
    if (alignment == 0 || (alignment & (alignment - 1))) { *err = "gguf: alignment is not a power of two"; return false; }
    for (uint64_t i = 0; i < n_tensors; ++i) {
        GgufEntry e;
        e.name = c.str();
        const uint32_t nd = c.get<uint32_t>();
        if (!c.ok || nd == 0 || nd > 4) { *err = "gguf: bad rank for tensor " + std::to_string(i); return false; }
        std::vector<int64_t> dims(nd);
        e.numel = 1;
        for (uint32_t k = 0; k < nd; ++k) {
            const uint64_t dk = c.get<uint64_t>();
            if (dk == 0 || dk > (1ull << 40)) { *err = "gguf: bad dimension in '" + e.name + "'"; return false; }
            dims[k] = (int64_t)dk;
            if (__builtin_mul_overflow(e.numel, (size_t)dk, &e.numel)) { *err = "gguf: shape of '" + e.name + "' overflows"; return false; }
        }
        e.shape.assign(dims.rbegin(), dims.rend());  // innermost-first -> row-major
        e.type = c.get<uint32_t>();
        e.offset = (size_t)c.get<uint64_t>();
        if (!c.ok) { *err = "gguf: truncated tensor table"; return false; }
        entries_.push_back(std::move(e));
    }
    data_off_ = ((size_t)(c.p - base_) + alignment - 1) / alignment * alignment;
    for (auto &e : entries_) {
        size_t bytes;
        if (e.numel > size_) { *err = "gguf: tensor '" + e.name + "' has more elements than the file has bytes"; return false; }  // also bounds the products below
        if (e.type == 0) bytes = e.numel * 4;
        else if (e.type == 1) bytes = e.numel * 2;
        else if (e.type == 8) {
            if (e.shape.back() % 32) { *err = "gguf: Q8_0 tensor '" + e.name + "' has a row length that is not a multiple of 32"; return false; }
            bytes = e.numel / 32 * 34;
        } else { *err = "gguf: tensor '" + e.name + "' has unsupported type " + std::to_string(e.type) + " (F32, F16 and Q8_0 are read)"; return false; }
        // step by step against what is left of the file: a large u64 offset must not wrap the sum
        if (data_off_ > size_ || e.offset > size_ - data_off_ || bytes > size_ - data_off_ - e.offset) { *err = "gguf: tensor '" + e.name + "' runs past the end of the file"; return false; }
    }
    return true;
}

bool Gguf::dequantize(const GgufEntry &e, float *out, std::string *err) const {
    const uint8_t *src = base_ + data_off_ + e.offset;
    if (e.type == 0) memcpy(out, src, e.numel * 4);
    else if (e.type == 1) {
        for (size_t i = 0; i < e.numel; ++i) { uint16_t h; memcpy(&h, src + 2 * i, 2); out[i] = half_to_float(h); }
    } else if (e.type == 8) {  // block_q8_0 { f16 d; int8 qs[32]; }: w = d * q
        for (size_t b = 0; b < e.numel / 32; ++b) {
            uint16_t h;
            memcpy(&h, src + 34 * b, 2);
            const float d = half_to_float(h);
            const int8_t *q = (const int8_t *)(src + 34 * b + 2);
            for (int i = 0; i < 32; ++i) out[32 * b + i] = d * (float)q[i];
        }
    } else { *err = "gguf: unsupported tensor type"; return false; }
    return true;
}

}  // namespace nb200host

// ---------------------------------------------------------------------------------------------------------
// C ABI (declared in include/norma_b200.h)
// ---------------------------------------------------------------------------------------------------------
struct nb200_tokenizer {
    nb200host::Tokenizer tok;
};

namespace {
// candle_transformers::models::whisper token strings norma looks up (monolingual.rs:376-384, multilingual.rs:239-249)
const char *SOT = "<|startoftranscript|>", *EOT = "<|endoftext|>", *TRANSCRIBE = "<|transcribe|>", *TRANSLATE = "<|translate|>",
           *NO_TIMESTAMPS = "<|notimestamps|>", *NO_SPEECH[2] = {"<|nocaptions|>", "<|nospeech|>"};

int lookup(const nb200host::Tokenizer &t, const char *name, uint32_t *id) {
    if (!t.token_to_id(name, id)) return nb200_fail(nullptr, NB200_NOT_FOUND, "Failed to get the id for token: %s", name);  // whisper::Error::TokenId
    return NB200_OK;
}

int special_tokens(const nb200host::Tokenizer &t, const char *language_token, int task, nb200_special_tokens *out) {
    int st;
    if ((st = lookup(t, NO_TIMESTAMPS, &out->no_timestamps))) return st;
    if ((st = lookup(t, SOT, &out->sot))) return st;
    if ((st = lookup(t, EOT, &out->eot))) return st;
    if (!t.token_to_id(NO_SPEECH[0], &out->no_speech) && !t.token_to_id(NO_SPEECH[1], &out->no_speech))
        return nb200_fail(nullptr, NB200_NOT_FOUND, "Failed to get the id for token: %s nor %s", NO_SPEECH[0], NO_SPEECH[1]);
    if ((st = lookup(t, task == NB200_TASK_TRANSLATE ? TRANSLATE : TRANSCRIBE, &out->task))) return st;
    out->lang = UINT32_MAX;
    if (language_token && (st = lookup(t, language_token, &out->lang))) return st;
    if ((st = lookup(t, "<|0.00|>", &out->ts_zero))) return st;
    if ((st = lookup(t, "<|1.00|>", &out->ts_one))) return st;
    return NB200_OK;
}
}  // namespace

extern "C" {

int nb200_config_from_file(const char *path, nb200_config *cfg, uint32_t *suppress_out, size_t suppress_cap, size_t *n_suppress) {
    if (!path || !cfg) return nb200_fail(nullptr, NB200_INVALID_ARG, "config_from_file: NULL argument");
    std::string text, err;
    if (!nb200host::read_file(path, &text, &err)) return nb200_fail(nullptr, NB200_IO_ERROR, "%s", err.c_str());
    std::vector<uint32_t> sup;
    nb200_config c;
    memset(&c, 0, sizeof c);
    if (!nb200host::config_from_json(text.data(), text.size(), &c, &sup, &err)) return nb200_fail(nullptr, NB200_PARSE_ERROR, "config.json: %s", err.c_str());
    c.max_batch = 1;
    *cfg = c;
    if (n_suppress) *n_suppress = sup.size();
    if (suppress_out) memcpy(suppress_out, sup.data(), std::min(sup.size(), suppress_cap) * 4);
    return NB200_OK;
}

int nb200_mel_filters(int n_mel, float *out) {
    std::vector<float> f;
    if (!out) return nb200_fail(nullptr, NB200_INVALID_ARG, "mel_filters: NULL output");
    if (!nb200host::mel_filterbank(n_mel, &f)) return nb200_fail(nullptr, NB200_UNSUPPORTED_SHAPE, "Unexpected number of mel bins (num_mel_bins), got: %d", n_mel);
    memcpy(out, f.data(), f.size() * 4);
    return NB200_OK;
}

int nb200_tokenizer_from_file(const char *path, nb200_tokenizer **out) {
    if (!path || !out) return nb200_fail(nullptr, NB200_INVALID_ARG, "tokenizer_from_file: NULL argument");
    nb200_tokenizer *t = new nb200_tokenizer();
    std::string text, err;
    if (!nb200host::read_file(path, &text, &err)) { delete t; return nb200_fail(nullptr, NB200_IO_ERROR, "%s", err.c_str()); }
    if (!t->tok.parse(text.data(), text.size(), &err)) { delete t; return nb200_fail(nullptr, NB200_PARSE_ERROR, "Failed to load the tokenizer: %s", err.c_str()); }
    *out = t;
    return NB200_OK;
}

void nb200_tokenizer_destroy(nb200_tokenizer *t) { delete t; }

int nb200_tokenizer_token_to_id(const nb200_tokenizer *t, const char *token, uint32_t *id) {
    if (!t || !token || !id) return nb200_fail(nullptr, NB200_INVALID_ARG, "token_to_id: NULL argument");
    return lookup(t->tok, token, id);
}

int nb200_tokenizer_decode(const nb200_tokenizer *t, const uint32_t *ids, size_t n, int skip_special_tokens, char *out, size_t cap, size_t *len) {
    if (!t || (n && !ids)) return nb200_fail(nullptr, NB200_INVALID_ARG, "tokenizer_decode: NULL argument");
    std::string s = t->tok.decode(ids, n, skip_special_tokens != 0);
    if (len) *len = s.size();
    if (out && cap) {
        size_t c = std::min(s.size(), cap - 1);
        memcpy(out, s.data(), c);
        out[c] = 0;
    }
    return NB200_OK;
}

int nb200_tokenizer_special_tokens(const nb200_tokenizer *t, const char *language_token, int task, nb200_special_tokens *out) {
    if (!t || !out) return nb200_fail(nullptr, NB200_INVALID_ARG, "tokenizer_special_tokens: NULL argument");
    return special_tokens(t->tok, language_token, task, out);
}

int nb200_tokenizer_language_tokens(const nb200_tokenizer *t, uint32_t *out) {
    if (!t || !out) return nb200_fail(nullptr, NB200_INVALID_ARG, "tokenizer_language_tokens: NULL argument");
    for (size_t i = 0; i < 99; ++i) {  // `Language::iter().map(|lang| token_id(&tokenizer, lang.token()))` (multilingual.rs:251-253)
        int st = lookup(t->tok, nb200host::language_token(i).c_str(), out + i);
        if (st) return st;
    }
    return NB200_OK;
}

int nb200_load_safetensors(nb200_ctx *ctx, const char *path, size_t *n_tensors) {
    if (!ctx || !path) return nb200_fail(ctx, NB200_INVALID_ARG, "load_safetensors: NULL argument");
    nb200host::Safetensors st;
    std::string err;
    if (!st.open(path, &err)) return nb200_fail(ctx, err.rfind("cannot open", 0) == 0 ? NB200_IO_ERROR : NB200_PARSE_ERROR, "%s", err.c_str());
    size_t n = 0;
    for (auto &e : st.entries()) {
        if (e.name.rfind("model.", 0) != 0) continue;  // `proj_out.weight` etc.: not read by Whisper::load
        if (e.shape.empty() || e.shape.size() > 4) continue;
        int rc = nb200_load_tensor(ctx, e.name.c_str(), st.data(e), e.dtype, e.shape.data(), (int)e.shape.size());
        if (rc != NB200_OK) return rc;
        ++n;
    }
    if (n_tensors) *n_tensors = n;
    return NB200_OK;
}

int nb200_load_gguf(nb200_ctx *ctx, const char *path, size_t *n_tensors) {
    if (!ctx || !path) return nb200_fail(ctx, NB200_INVALID_ARG, "load_gguf: NULL argument");
    nb200host::Gguf g;
    std::string err;
    if (!g.open(path, &err)) return nb200_fail(ctx, err.rfind("cannot open", 0) == 0 ? NB200_IO_ERROR : NB200_PARSE_ERROR, "%s", err.c_str());
    size_t n = 0;
    std::vector<float> buf;
    for (auto &e : g.entries()) {
        if (e.name.rfind("model.", 0) != 0) continue;
        buf.resize(e.numel);
        if (!g.dequantize(e, buf.data(), &err)) return nb200_fail(ctx, NB200_PARSE_ERROR, "%s", err.c_str());
        int rc = nb200_load_tensor(ctx, e.name.c_str(), buf.data(), NB200_F32, e.shape.data(), (int)e.shape.size());
        if (rc != NB200_OK) return rc;
        ++n;
    }
    if (n_tensors) *n_tensors = n;
    return NB200_OK;
}

int nb200_gguf_read(const char *path, const char *name, float *out, size_t cap, int64_t *shape, int *rank, int *type) {
    if (!path || !name) return nb200_fail(nullptr, NB200_INVALID_ARG, "gguf_read: NULL argument");
    nb200host::Gguf g;
    std::string err;
    if (!g.open(path, &err)) return nb200_fail(nullptr, err.rfind("cannot open", 0) == 0 ? NB200_IO_ERROR : NB200_PARSE_ERROR, "%s", err.c_str());
    for (auto &e : g.entries()) {
        if (e.name != name) continue;
        if (rank) *rank = (int)e.shape.size();
        if (type) *type = (int)e.type;
        for (size_t i = 0; shape && i < e.shape.size(); ++i) shape[i] = e.shape[i];
        if (out) {
            if (cap < e.numel) return nb200_fail(nullptr, NB200_INVALID_ARG, "gguf_read: buffer holds %zu of %zu values", cap, e.numel);
            if (!g.dequantize(e, out, &err)) return nb200_fail(nullptr, NB200_PARSE_ERROR, "%s", err.c_str());
        }
        return NB200_OK;
    }
    return nb200_fail(nullptr, NB200_NOT_FOUND, "gguf: no tensor named '%s'", name);
}

int nb200_safetensors_read(const char *path, const char *name, float *out, size_t cap, int64_t *shape, int *rank) {
    if (!path || !name) return nb200_fail(nullptr, NB200_INVALID_ARG, "safetensors_read: NULL argument");
    nb200host::Safetensors st;
    std::string err;
    if (!st.open(path, &err)) return nb200_fail(nullptr, err.rfind("cannot open", 0) == 0 ? NB200_IO_ERROR : NB200_PARSE_ERROR, "%s", err.c_str());
    for (auto &e : st.entries()) {
        if (e.name != name) continue;
        if (rank) *rank = (int)e.shape.size();
        size_t n = 1;
        for (size_t i = 0; i < e.shape.size(); ++i) {
            if (shape && i < 8) shape[i] = e.shape[i];
            n *= (size_t)e.shape[i];
        }
        if (out) {
            const uint8_t *src = st.data(e);
            n = std::min(n, cap);
            for (size_t i = 0; i < n; ++i) {
                switch (e.dtype) {
                    case NB200_F32: memcpy(out + i, src + 4 * i, 4); break;
                    case NB200_F64: { double d; memcpy(&d, src + 8 * i, 8); out[i] = (float)d; break; }
                    case NB200_BF16: { uint32_t u = (uint32_t)(src[2 * i] | (src[2 * i + 1] << 8)) << 16; memcpy(out + i, &u, 4); break; }
                    default: {  // F16, through the same f32 path the upload takes
                        uint16_t h = (uint16_t)(src[2 * i] | (src[2 * i + 1] << 8));
                        const uint32_t sign = (h >> 15) & 1u, ex = (h >> 10) & 0x1fu, man = h & 0x3ffu;
                        float f;
                        if (ex == 0) f = ldexpf((float)man, -24);
                        else if (ex == 31) f = man ? NAN : INFINITY;
                        else f = ldexpf((float)(man | 0x400u), (int)ex - 25);
                        out[i] = sign ? -f : f;
                    }
                }
            }
        }
        return NB200_OK;
    }
    return nb200_fail(nullptr, NB200_NOT_FOUND, "safetensors: no tensor named '%s'", name);
}

int nb200_model_set_tokenizer(nb200_model *m, const nb200_tokenizer *t) {
    if (!m || !t) return NB200_INVALID_ARG;
    m->model->set_tokenizer(t->tok);
    return NB200_OK;
}

int nb200_model_from_files(int ordinal, const char *config_json, const char *tokenizer_json, const char *safetensors, nb200_dtype compute,
                           const char *language_token, int task, size_t max_chunk_len, uint64_t seed, nb200_ctx **ctx_out, nb200_model **model_out) {
    if (!config_json || !tokenizer_json || !safetensors || !ctx_out || !model_out) return nb200_fail(nullptr, NB200_INVALID_ARG, "model_from_files: NULL argument");
    nb200_config cfg;
    size_t n_sup = 0;
    std::vector<uint32_t> sup(1 << 16);
    int st = nb200_config_from_file(config_json, &cfg, sup.data(), sup.size(), &n_sup);                 // monolingual.rs:347
    if (st) return st;
    nb200_tokenizer *tk = nullptr;
    if ((st = nb200_tokenizer_from_file(tokenizer_json, &tk))) return st;                               // monolingual.rs:348-349
    std::vector<float> filters;
    if (!nb200host::mel_filterbank(cfg.num_mel_bins, &filters)) {                                       // monolingual.rs:351-355
        nb200_tokenizer_destroy(tk);
        return nb200_fail(nullptr, NB200_UNSUPPORTED_SHAPE, "Unexpected number of mel bins (num_mel_bins), got: %d", cfg.num_mel_bins);
    }
    nb200_special_tokens tok;
    uint32_t langs[99];
    st = special_tokens(tk->tok, language_token, task, &tok);                                           // monolingual.rs:376-384,419-420
    if (!st && !language_token) st = nb200_tokenizer_language_tokens(tk, langs);                        // multilingual.rs:251-254
    if (st) { nb200_tokenizer_destroy(tk); return st; }
    nb200_ctx *ctx = nullptr;
    st = nb200_create(ordinal, &cfg, compute, &ctx);
    if (st) { nb200_tokenizer_destroy(tk); return st; }
    auto bail = [&](int code) {
        std::string msg = nb200_last_error(ctx);
        nb200_destroy(ctx);
        nb200_tokenizer_destroy(tk);
        return nb200_fail(nullptr, code, "%s", msg.c_str());
    };
    bool is_gguf = false;  // `Quantized*` model types hand over model-{ext}-q80.gguf instead (monolingual.rs:364-369)
    {
        FILE *f = fopen(safetensors, "rb");
        char magic[4] = {0, 0, 0, 0};
        if (f) { is_gguf = fread(magic, 1, 4, f) == 4 && !memcmp(magic, "GGUF", 4); fclose(f); }
    }
    if ((st = is_gguf ? nb200_load_gguf(ctx, safetensors, nullptr) : nb200_load_safetensors(ctx, safetensors, nullptr))) return bail(st);  // monolingual.rs:364-373
    if ((st = nb200_finalize_weights(ctx))) return bail(st);
    if ((st = nb200_set_mel_filters(ctx, filters.data(), cfg.num_mel_bins))) return bail(st);
    if ((st = nb200_set_tokens(ctx, &tok))) return bail(st);
    if ((st = nb200_set_suppress(ctx, sup.data(), std::min(n_sup, sup.size())))) return bail(st);       // monolingual.rs:386-395
    nb200_model *m = nullptr;
    if ((st = nb200_model_create(ctx, &tok, max_chunk_len, seed, &m))) return bail(st);
    nb200_model_set_tokenizer(m, tk);
    if (!language_token) nb200_model_set_language_detection(m, langs, 99);                              // LanguageState::Detect (multilingual.rs:296-299)
    nb200_tokenizer_destroy(tk);
    *ctx_out = ctx;
    *model_out = m;
    return NB200_OK;
}

}  // extern "C"
