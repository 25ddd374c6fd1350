// seaweedfs_b200/csrc/mini_json.h — just enough JSON to read the two things this path needs from a .vif file
// (protobuf-JSON of VolumeInfo, weed/storage/volume_info/volume_info.go:73-95, weed/pb/volume_server.proto:561-577):
// a numeric member of the top-level object (datFileSize, version) and the members of ecShardConfig.  A real
// tokenizer — strings with escapes, nested objects and arrays are skipped properly, members are matched only at the
// level they belong to — so member order, whitespace, escapes or a key name occurring inside some string value
// cannot confuse it.  protojson writes 64-bit integers as strings and accepts both lowerCamel and proto field names.
#pragma once
#include <cctype>
#include <cstdint>
#include <cstdlib>
#include <string>

namespace swec {
namespace mini_json {

struct Cursor {
    const std::string& s;
    size_t i = 0;
    bool ok = true;
    void ws() {
        while (i < s.size() && (s[i] == ' ' || s[i] == '\t' || s[i] == '\n' || s[i] == '\r')) i++;
    }
    bool eat(char c) {
        ws();
        if (i < s.size() && s[i] == c) {
            i++;
            return true;
        }
        return false;
    }
    bool string(std::string* out) {  // out may be null (skip)
        ws();
        if (i >= s.size() || s[i] != '"') return ok = false;
        for (i++; i < s.size(); i++) {
            if (s[i] == '"') {
                i++;
                return true;
            }
            if (s[i] == '\\') {
                if (++i >= s.size()) break;
                if (s[i] == 'u') {
                    i += 4;
                    if (out) out->push_back('?');
                } else if (out) {
                    out->push_back(s[i] == 'n' ? '\n' : s[i] == 't' ? '\t' : s[i]);
                }
            } else if (out) {
                out->push_back(s[i]);
            }
        }
        return ok = false;
    }
    bool skip_value() {
        ws();
        if (i >= s.size()) return ok = false;
        if (s[i] == '"') return string(nullptr);
        if (s[i] == '{' || s[i] == '[') {
            const char close = s[i] == '{' ? '}' : ']';
            const bool obj = s[i] == '{';
            i++;
            if (eat(close)) return true;
            do {
                if (obj && (!string(nullptr) || !eat(':'))) return ok = false;
                if (!skip_value()) return false;
            } while (eat(','));
            return eat(close) ? true : (ok = false);
        }
        const size_t b = i;  // number / true / false / null
        while (i < s.size() && (isalnum((unsigned char)s[i]) || s[i] == '-' || s[i] == '+' || s[i] == '.')) i++;
        return i > b ? true : (ok = false);
    }
};

// Value of member `key` (or `alt`) of the object that starts at cursor c (c at '{'): returns its raw text span
// [*b, *e) and leaves the cursor after the object.  false when absent or malformed.
inline bool member(Cursor& c, const char* key, const char* alt, size_t* b, size_t* e) {
    bool found = false;
    if (!c.eat('{')) return false;
    if (c.eat('}')) return false;
    do {
        std::string k;
        if (!c.string(&k) || !c.eat(':')) return false;
        c.ws();
        const size_t vb = c.i;
        if (!c.skip_value()) return false;
        if (!found && (k == key || (alt && k == alt))) {
            *b = vb;
            *e = c.i;
            found = true;
        }
    } while (c.eat(','));
    return c.eat('}') && found;
}

// integer written as 123 or "123"
inline bool to_int(const std::string& s, size_t b, size_t e, int64_t* out) {
    while (b < e && (s[b] == '"' || isspace((unsigned char)s[b]))) b++;
    if (b >= e || !(isdigit((unsigned char)s[b]) || s[b] == '-')) return false;
    *out = strtoll(s.c_str() + b, nullptr, 10);
    return true;
}

// top-level numeric member
inline bool top_int(const std::string& txt, const char* key, const char* alt, int64_t* out) {
    Cursor c{txt};
    size_t b = 0, e = 0;
    return member(c, key, alt, &b, &e) && to_int(txt, b, e, out);
}

// numeric member of a top-level object member: top.obj.key
inline bool nested_int(const std::string& txt, const char* obj, const char* obj_alt, const char* key, const char* alt,
                       int64_t* out) {
    Cursor c{txt};
    size_t b = 0, e = 0;
    if (!member(c, obj, obj_alt, &b, &e)) return false;
    const std::string sub = txt.substr(b, e - b);
    Cursor c2{sub};
    size_t b2 = 0, e2 = 0;
    return member(c2, key, alt, &b2, &e2) && to_int(sub, b2, e2, out);
}

}  // namespace mini_json
}  // namespace swec
