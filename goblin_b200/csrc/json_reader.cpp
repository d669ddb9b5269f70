#include "json_reader.h"

#include <cerrno>
#include <cstdlib>
#include <cstring>

namespace gb {
namespace {

struct Parser {
    const char* p;
    const char* end;
    std::string err;

    void skipWs() {
        while (p < end && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    }
    bool fail(const std::string& m) {
        if (err.empty()) err = m;
        return false;
    }
    static void appendUtf8(std::string& s, uint32_t cp) {
        if (cp < 0x80) s += (char)cp;
        else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) {
            s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F));
        } else {
            s += (char)(0xF0 | (cp >> 18)); s += (char)(0x80 | ((cp >> 12) & 0x3F));
            s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F));
        }
    }
    bool hex4(uint32_t* out) {
        if (end - p < 4) return fail("truncated \\u escape");
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) {
            char c = *p++;
            v <<= 4;
            if (c >= '0' && c <= '9') v |= c - '0';
            else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10;
            else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10;
            else return fail("bad \\u escape");
        }
        *out = v;
        return true;
    }
    bool parseString(std::string* out) {
        if (p >= end || *p != '"') return fail("expected string");
        ++p;
        out->clear();
        while (p < end) {
            char c = *p++;
            if (c == '"') return true;
            if (c == '\\') {
                if (p >= end) break;
                char e = *p++;
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
                    if (cp >= 0xD800 && cp <= 0xDBFF && end - p >= 6 && p[0] == '\\' && p[1] == 'u') {
                        p += 2;
                        uint32_t lo;
                        if (!hex4(&lo)) return false;
                        cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    appendUtf8(*out, cp);
                    break;
                }
                default: return fail("bad escape");
                }
            } else {
                *out += c;
            }
        }
        return fail("unterminated string");
    }
    bool parseNumber(JsonValue* v) {
        const char* s = p;
        if (p < end && *p == '-') ++p;
        if (p >= end || !(*p >= '0' && *p <= '9')) return fail("bad number");
        while (p < end && *p >= '0' && *p <= '9') ++p;
        bool isFloat = false;
        if (p < end && *p == '.') {
            isFloat = true;
            ++p;
            if (p >= end || !(*p >= '0' && *p <= '9')) return fail("bad fraction");
            while (p < end && *p >= '0' && *p <= '9') ++p;
        }
        if (p < end && (*p == 'e' || *p == 'E')) {
            isFloat = true;
            ++p;
            if (p < end && (*p == '+' || *p == '-')) ++p;
            if (p >= end || !(*p >= '0' && *p <= '9')) return fail("bad exponent");
            while (p < end && *p >= '0' && *p <= '9') ++p;
        }
        std::string lit(s, p);
        if (!isFloat) {
            errno = 0;
            long long iv = strtoll(lit.c_str(), nullptr, 10);
            if (errno == 0) {
                v->type = JsonValue::Int;
                v->i = iv;
                v->d = (double)iv;
                return true;
            }
            // out of int64 range: nlohmann falls back to floating point
        }
        v->type = JsonValue::Float;
        v->d = strtod(lit.c_str(), nullptr);
        return true;
    }
    bool parseValue(JsonValue* v, int depth) {
        if (depth > 256) return fail("nesting too deep");
        skipWs();
        if (p >= end) return fail("unexpected end of input");
        char c = *p;
        if (c == '{') {
            ++p;
            v->type = JsonValue::Object;
            skipWs();
            if (p < end && *p == '}') { ++p; return true; }
            while (true) {
                skipWs();
                std::string key;
                if (!parseString(&key)) return false;
                skipWs();
                if (p >= end || *p != ':') return fail("expected ':'");
                ++p;
                JsonValue child;
                if (!parseValue(&child, depth + 1)) return false;
                v->obj[key] = std::move(child);
                skipWs();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == '}') { ++p; return true; }
                return fail("expected ',' or '}'");
            }
        }
        if (c == '[') {
            ++p;
            v->type = JsonValue::Array;
            skipWs();
            if (p < end && *p == ']') { ++p; return true; }
            while (true) {
                JsonValue child;
                if (!parseValue(&child, depth + 1)) return false;
                v->arr.push_back(std::move(child));
                skipWs();
                if (p < end && *p == ',') { ++p; continue; }
                if (p < end && *p == ']') { ++p; return true; }
                return fail("expected ',' or ']'");
            }
        }
        if (c == '"') {
            v->type = JsonValue::String;
            return parseString(&v->s);
        }
        if (c == 't' && end - p >= 4 && !strncmp(p, "true", 4)) { p += 4; v->type = JsonValue::Bool; v->b = true; return true; }
        if (c == 'f' && end - p >= 5 && !strncmp(p, "false", 5)) { p += 5; v->type = JsonValue::Bool; v->b = false; return true; }
        if (c == 'n' && end - p >= 4 && !strncmp(p, "null", 4)) { p += 4; v->type = JsonValue::Null; return true; }
        return parseNumber(v);
    }
};

} // namespace

bool parseJson(const std::string& text, JsonValue* out, std::string* error) {
    Parser ps{text.data(), text.data() + text.size(), {}};
    // tolerate a UTF-8 byte order mark
    if (text.size() >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB &&
        (unsigned char)text[2] == 0xBF) {
        ps.p += 3;
    }
    if (!ps.parseValue(out, 0)) {
        if (error) *error = ps.err + " at byte " + std::to_string(ps.p - text.data());
        return false;
    }
    ps.skipWs();
    if (ps.p != ps.end) {
        if (error) *error = "trailing characters at byte " + std::to_string(ps.p - text.data());
        return false;
    }
    return true;
}

} // namespace gb
