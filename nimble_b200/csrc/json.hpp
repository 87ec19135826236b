// Minimal JSON reader for the nimble library file `[config, data]`
// (written by nimble/__main__.py:64-65 with json.dump(indent=2)).  Values only; no writer.
#pragma once
#include <cstdint>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace nb200 {

struct JValue {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0.0;
    std::string str;
    std::vector<JValue> arr;
    std::vector<std::pair<std::string, JValue>> obj;

    const JValue *get(const std::string &key) const {
        for (auto &kv : obj)
            if (kv.first == key) return &kv.second;
        return nullptr;
    }
};

class JParser {
  public:
    JParser(const char *p, size_t n) : p_(p), end_(p + n) {}
    JValue parse() {
        JValue v = value();
        ws();
        if (p_ != end_) fail("trailing characters");
        return v;
    }

  private:
    const char *p_, *end_;
    [[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string("JSON: ") + what); }
    void ws() {
        while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
    }
    bool lit(const char *s) {
        size_t n = strlen(s);
        if ((size_t)(end_ - p_) >= n && memcmp(p_, s, n) == 0) { p_ += n; return true; }
        return false;
    }
    static void utf8(std::string &o, uint32_t cp) {
        if (cp < 0x80) o += (char)cp;
        else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
        else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 0x3F)); o += (char)(0x80 | ((cp >> 6) & 0x3F)); o += (char)(0x80 | (cp & 0x3F)); }
    }
    uint32_t hex4() {
        if (end_ - p_ < 4) fail("bad \\u escape");
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) {
            char c = *p_++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else fail("bad \\u escape");
        }
        return v;
    }
    std::string string() {
        if (p_ >= end_ || *p_ != '"') fail("expected string");
        ++p_;
        std::string o;
        const char *run = p_;
        while (true) {
            if (p_ >= end_) fail("unterminated string");
            char c = *p_;
            if (c == '"') { o.append(run, p_ - run); ++p_; return o; }
            if (c == '\\') {
                o.append(run, p_ - run);
                ++p_;
                if (p_ >= end_) fail("bad escape");
                char e = *p_++;
                switch (e) {
                case '"': o += '"'; break;
                case '\\': o += '\\'; break;
                case '/': o += '/'; break;
                case 'b': o += '\b'; break;
                case 'f': o += '\f'; break;
                case 'n': o += '\n'; break;
                case 'r': o += '\r'; break;
                case 't': o += '\t'; break;
                case 'u': {
                    uint32_t cp = hex4();
                    if (cp >= 0xD800 && cp < 0xDC00 && end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
                        p_ += 2;
                        uint32_t lo = hex4();
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    utf8(o, cp);
                    break;
                }
                default: fail("bad escape");
                }
                run = p_;
            } else ++p_;
        }
    }
    JValue value() {
        ws();
        if (p_ >= end_) fail("unexpected end");
        JValue v;
        char c = *p_;
        if (c == '{') {
            v.type = JValue::Obj;
            ++p_; ws();
            if (p_ < end_ && *p_ == '}') { ++p_; return v; }
            while (true) {
                ws();
                std::string k = string();
                ws();
                if (p_ >= end_ || *p_ != ':') fail("expected ':'");
                ++p_;
                v.obj.emplace_back(std::move(k), value());
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == '}') { ++p_; return v; }
                fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            v.type = JValue::Arr;
            ++p_; ws();
            if (p_ < end_ && *p_ == ']') { ++p_; return v; }
            while (true) {
                v.arr.push_back(value());
                ws();
                if (p_ < end_ && *p_ == ',') { ++p_; continue; }
                if (p_ < end_ && *p_ == ']') { ++p_; return v; }
                fail("expected ',' or ']'");
            }
        }
        if (c == '"') { v.type = JValue::Str; v.str = string(); return v; }
        if (lit("true")) { v.type = JValue::Bool; v.b = true; return v; }
        if (lit("false")) { v.type = JValue::Bool; v.b = false; return v; }
        if (lit("null")) { v.type = JValue::Null; return v; }
        if (lit("NaN")) { v.type = JValue::Num; v.num = 0.0 / 0.0; return v; }
        {
            char *e = nullptr;
            std::string tmp(p_, (size_t)std::min<ptrdiff_t>(end_ - p_, 64));
            double d = strtod(tmp.c_str(), &e);
            if (e == tmp.c_str()) fail("unexpected token");
            p_ += (e - tmp.c_str());
            v.type = JValue::Num; v.num = d;
            return v;
        }
    }
};

}  // namespace nb200
