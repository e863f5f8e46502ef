// json.h — a small JSON reader for the three checkpoint side files norma parses with serde_json / tokenizers
// (config.json, tokenizer.json, the safetensors header; /root/reference/src/models/whisper/monolingual.rs:347-349,372).
// Strict RFC 8259 subset: objects keep insertion order, numbers are kept as double AND int64 (ids, offsets), strings are
// decoded to UTF-8 (\uXXXX with surrogate pairs).  Errors carry the byte offset.
#pragma once
#include <stdint.h>
#include <stdlib.h>

#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace nb200json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    int64_t inum = 0;
    bool is_int = false;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    const Value *get(const char *key) const {
        if (kind != Object) return nullptr;
        for (auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_null() const { return kind == Null; }
};

class Parser {
public:
    Parser(const char *p, size_t n) : p_(p), end_(p + n), begin_(p) {}
    bool parse(Value *out, std::string *err) {
        skip_ws();
        if (!value(out, 0)) { *err = err_ + " at byte " + std::to_string((size_t)(p_ - begin_)); return false; }
        skip_ws();
        if (p_ != end_) { *err = "trailing characters at byte " + std::to_string((size_t)(p_ - begin_)); return false; }
        return true;
    }

private:
    const char *p_, *end_, *begin_;
    std::string err_;
    bool fail(const char *m) { err_ = m; return false; }
    void skip_ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\r' || *p_ == '\t')) ++p_;
    }
    bool lit(const char *s) {
        size_t n = strlen_(s);
        if ((size_t)(end_ - p_) < n) return false;
        for (size_t i = 0; i < n; ++i)
            if (p_[i] != s[i]) return false;
        p_ += n;
        return true;
    }
    static size_t strlen_(const char *s) { size_t n = 0; while (s[n]) ++n; return n; }
    static void put_utf8(std::string &s, uint32_t cp) {
        if (cp < 0x80) s += (char)cp;
        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
        else { s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
    }
    bool hex4(uint32_t *v) {
        if (end_ - p_ < 4) return fail("short \\u escape");
        uint32_t x = 0;
        for (int i = 0; i < 4; ++i) {
            char c = p_[i];
            x <<= 4;
            if (c >= '0' && c <= '9') x |= c - '0';
            else if (c >= 'a' && c <= 'f') x |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') x |= c - 'A' + 10;
            else return fail("bad \\u escape");
        }
        p_ += 4;
        *v = x;
        return true;
    }
    bool string(std::string *out) {
        if (p_ >= end_ || *p_ != '"') return fail("expected string");
        ++p_;
        out->clear();
        while (p_ < end_) {
            unsigned char c = (unsigned char)*p_;
            if (c == '"') { ++p_; return true; }
            if (c < 0x20) return fail("control character in string");
            if (c != '\\') {
                const char *q = p_;
                while (q < end_ && *q != '"' && *q != '\\' && (unsigned char)*q >= 0x20) ++q;
                out->append(p_, q - p_);
                p_ = q;
                continue;
            }
            if (++p_ >= end_) break;
            char e = *p_++;
            switch (e) {
                case '"': *out += '"'; break;
                case '\\': *out += '\\'; break;
                case '/': *out += '/'; break;
                case 'b': *out += '\b'; break;
                case 'f': *out += '\f'; break;
                case 'n': *out += '\n'; break;
                case 'r': *out += '\r'; break;
                case 't': *out += '\t'; break;
                case 'u': {
                    uint32_t cp;
                    if (!hex4(&cp)) return false;
                    if (cp >= 0xD800 && cp <= 0xDBFF) {  // high surrogate: must be followed by \uDC00..DFFF
                        uint32_t lo;
                        if (end_ - p_ < 2 || p_[0] != '\\' || p_[1] != 'u') return fail("lone surrogate");
                        p_ += 2;
                        if (!hex4(&lo)) return false;
                        if (lo < 0xDC00 || lo > 0xDFFF) return fail("lone surrogate");
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    } else if (cp >= 0xDC00 && cp <= 0xDFFF) return fail("lone surrogate");
                    put_utf8(*out, cp);
                    break;
                }
                default: return fail("bad escape");
            }
        }
        return fail("unterminated string");
    }
    bool number(Value *out) {
        const char *s = p_;
        if (p_ < end_ && *p_ == '-') ++p_;
        if (p_ >= end_ || *p_ < '0' || *p_ > '9') return fail("bad number");
        bool integral = true;
        while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        if (p_ < end_ && *p_ == '.') { integral = false; ++p_; while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_; }
        if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
            integral = false;
            ++p_;
            if (p_ < end_ && (*p_ == '+' || *p_ == '-')) ++p_;
            while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
        }
        std::string t(s, p_ - s);
        out->kind = Value::Number;
        out->num = strtod(t.c_str(), nullptr);
        out->is_int = integral && t.size() < 19;
        out->inum = out->is_int ? strtoll(t.c_str(), nullptr, 10) : (int64_t)out->num;
        return true;
    }
    bool value(Value *out, int depth) {
        if (depth > 64) return fail("nesting too deep");
        if (p_ >= end_) return fail("unexpected end");
        char c = *p_;
        if (c == '{') {
            ++p_;
            out->kind = Value::Object;
            skip_ws();
            if (p_ < end_ && *p_ == '}') { ++p_; return true; }
            for (;;) {
                skip_ws();
                std::string k;
                if (!string(&k)) return false;
                skip_ws();
                if (p_ >= end_ || *p_ != ':') return fail("expected ':'");
                ++p_;
                skip_ws();
                out->obj.emplace_back(std::move(k), Value());
                if (!value(&out->obj.back().second, depth + 1)) return false;
                skip_ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == '}') { ++p_; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p_;
            out->kind = Value::Array;
            skip_ws();
            if (p_ < end_ && *p_ == ']') { ++p_; return true; }
            for (;;) {
                skip_ws();
                out->arr.emplace_back();
                if (!value(&out->arr.back(), depth + 1)) return false;
                skip_ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == ']') { ++p_; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') { out->kind = Value::String; return string(&out->str); }
        if (c == 't') { if (!lit("true")) return fail("bad literal"); out->kind = Value::Bool; out->b = true; return true; }
        if (c == 'f') { if (!lit("false")) return fail("bad literal"); out->kind = Value::Bool; out->b = false; return true; }
        if (c == 'n') { if (!lit("null")) return fail("bad literal"); out->kind = Value::Null; return true; }
        return number(out);
    }
};

}  // namespace nb200json
