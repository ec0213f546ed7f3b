// Minimal JSON reader for Avro schema documents (host side, cold path).
// Replaces the serde_json front half of apache_avro::Schema::parse_str, which the
// reference calls at ruhvro/src/deserialize.rs:18-20.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace rv {

struct Json {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    std::string str;  // String payload, or the raw text of a Number
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;  // insertion order kept

    const Json* find(const char* key) const {
        if (kind != Object) return nullptr;
        for (auto& kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
    bool is_string() const { return kind == String; }
};

class JsonReader {
  public:
    JsonReader(const char* p, size_t n) : p_(p), end_(p + n) {}
    Json parse_document() {
        Json v = value(0);
        ws();
        if (p_ != end_) fail("trailing characters after JSON document");
        return v;
    }

  private:
    const char* p_;
    const char* end_;
    [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("schema JSON: ") + what); }
    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
    }
    bool lit(const char* s) {
        size_t n = std::char_traits<char>::length(s);
        if (size_t(end_ - p_) >= n && std::char_traits<char>::compare(p_, s, n) == 0) { p_ += n; return true; }
        return false;
    }
    static int hex(char c) {
        if (c >= '0' && c <= '9') return c - '0';
        if (c >= 'a' && c <= 'f') return c - 'a' + 10;
        if (c >= 'A' && c <= 'F') return c - 'A' + 10;
        return -1;
    }
    unsigned hex4() {
        if (end_ - p_ < 4) fail("truncated \\u escape");
        unsigned v = 0;
        for (int i = 0; i < 4; ++i) {
            int h = hex(*p_++);
            if (h < 0) fail("bad \\u escape");
            v = v * 16 + unsigned(h);
        }
        return v;
    }
    static void utf8(std::string& o, unsigned cp) {
        if (cp < 0x80) o.push_back(char(cp));
        else if (cp < 0x800) { o.push_back(char(0xC0 | (cp >> 6))); o.push_back(char(0x80 | (cp & 0x3F))); }
        else if (cp < 0x10000) { o.push_back(char(0xE0 | (cp >> 12))); o.push_back(char(0x80 | ((cp >> 6) & 0x3F))); o.push_back(char(0x80 | (cp & 0x3F))); }
        else { o.push_back(char(0xF0 | (cp >> 18))); o.push_back(char(0x80 | ((cp >> 12) & 0x3F))); o.push_back(char(0x80 | ((cp >> 6) & 0x3F))); o.push_back(char(0x80 | (cp & 0x3F))); }
    }
    std::string string_body() {
        if (p_ >= end_ || *p_ != '"') fail("expected string");
        ++p_;
        std::string out;
        for (;;) {
            if (p_ >= end_) fail("unterminated string");
            char c = *p_++;
            if (c == '"') return out;
            if (c != '\\') {
                if (static_cast<unsigned char>(c) < 0x20) fail("control character in a string");   // (as serde_json: must be escaped)
                out.push_back(c);
                continue;
            }
            if (p_ >= end_) fail("unterminated escape");
            c = *p_++;
            switch (c) {
                case 'n': out.push_back('\n'); break;
                case 't': out.push_back('\t'); break;
                case 'r': out.push_back('\r'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case '"': case '\\': case '/': out.push_back(c); break;
                case 'u': {
                    unsigned cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                        p_ += 2;
                        unsigned lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    utf8(out, cp);
                    break;
                }
                default: fail("bad escape");
            }
        }
    }
    Json value(int depth) {
        if (depth > 128) fail("nesting too deep");
        ws();
        if (p_ >= end_) fail("unexpected end of document");
        Json v;
        char c = *p_;
        if (c == '"') { v.kind = Json::String; v.str = string_body(); return v; }
        if (c == '{') {
            ++p_; v.kind = Json::Object; ws();
            if (p_ < end_ && *p_ == '}') { ++p_; return v; }
            for (;;) {
                ws();
                std::string k = string_body();
                ws();
                if (p_ >= end_ || *p_ != ':') fail("expected ':'");
                ++p_;
                Json item = value(depth + 1);
                // a repeated key replaces the earlier value (serde_json's Map::insert, which apache_avro's
                // Schema::parse_str reads the document into: the last occurrence wins)
                bool replaced = false;
                for (auto& kv : v.obj)
                    if (kv.first == k) { kv.second = std::move(item); replaced = true; break; }
                if (!replaced) v.obj.emplace_back(std::move(k), std::move(item));
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == '}') { ++p_; return v; }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p_; v.kind = Json::Array; ws();
            if (p_ < end_ && *p_ == ']') { ++p_; return v; }
            for (;;) {
                v.arr.push_back(value(depth + 1));
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == ']') { ++p_; return v; }
                fail("expected ',' or ']'");
            }
        }
        if (lit("true")) { v.kind = Json::Bool; v.b = true; return v; }
        if (lit("false")) { v.kind = Json::Bool; return v; }
        if (lit("null")) return v;
        if (c == '-' || (c >= '0' && c <= '9')) {
            const char* s = p_;
            while (p_ < end_ && ((*p_ >= '0' && *p_ <= '9') || *p_ == '-' || *p_ == '+' || *p_ == '.' || *p_ == 'e' || *p_ == 'E')) ++p_;
            v.kind = Json::Number; v.str.assign(s, p_);
            return v;
        }
        fail("unexpected character");
    }
};

}  // namespace rv
