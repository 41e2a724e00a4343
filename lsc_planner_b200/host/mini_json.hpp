// Minimal JSON reader for the reference's mission files (missions/*.json; the reference uses rapidjson,
// src/mission.cpp:20-319). Objects, arrays, numbers, strings, true/false/null; '#' and '//' comments tolerated.
#pragma once
#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mini_json {

struct Value {
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Value> arr;
    std::map<std::string, Value> obj;

    bool has(const std::string& k) const { return kind == Object && obj.count(k); }
    const Value& operator[](const std::string& k) const {
        auto it = obj.find(k);
        if (kind != Object || it == obj.end()) throw std::invalid_argument("[Mission] missing key: " + k);
        return it->second;
    }
    const Value& operator[](size_t i) const {
        if (kind != Array || i >= arr.size()) throw std::invalid_argument("[Mission] array index out of range");
        return arr[i];
    }
    size_t size() const { return kind == Array ? arr.size() : obj.size(); }
    double number() const { if (kind != Number) throw std::invalid_argument("[Mission] number expected"); return num; }
    const std::string& string() const { if (kind != String) throw std::invalid_argument("[Mission] string expected"); return str; }
};

class Parser {
public:
    explicit Parser(const std::string& text) : s(text) {}
    Value parse() { Value v = value(); ws(); if (p != s.size()) fail("trailing characters"); return v; }

private:
    const std::string& s;
    size_t p = 0;
    [[noreturn]] void fail(const char* what) const { throw std::invalid_argument(std::string("[Mission] JSON: ") + what + " at offset " + std::to_string(p)); }
    void ws() {
        while (p < s.size()) {
            if (std::isspace((unsigned char)s[p])) p++;
            else if (s[p] == '#' || (s[p] == '/' && p + 1 < s.size() && s[p + 1] == '/')) { while (p < s.size() && s[p] != '\n') p++; }
            else break;
        }
    }
    Value value() {
        ws();
        if (p >= s.size()) fail("unexpected end");
        Value v;
        const char c = s[p];
        if (c == '{') {
            v.kind = Value::Object; p++; ws();
            if (p < s.size() && s[p] == '}') { p++; return v; }
            while (true) {
                ws();
                Value k = value();
                if (k.kind != Value::String) fail("object key must be a string");
                ws();
                if (p >= s.size() || s[p] != ':') fail("':' expected");
                p++;
                v.obj[k.str] = value();
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == '}') { p++; break; }
                fail("',' or '}' expected");
            }
        } else if (c == '[') {
            v.kind = Value::Array; p++; ws();
            if (p < s.size() && s[p] == ']') { p++; return v; }
            while (true) {
                v.arr.push_back(value());
                ws();
                if (p < s.size() && s[p] == ',') { p++; continue; }
                if (p < s.size() && s[p] == ']') { p++; break; }
                fail("',' or ']' expected");
            }
        } else if (c == '"') {
            v.kind = Value::String; p++;
            while (p < s.size() && s[p] != '"') {
                if (s[p] == '\\' && p + 1 < s.size()) { p++; v.str.push_back(s[p] == 'n' ? '\n' : s[p] == 't' ? '\t' : s[p]); }
                else v.str.push_back(s[p]);
                p++;
            }
            if (p >= s.size()) fail("unterminated string");
            p++;
        } else if (s.compare(p, 4, "true") == 0) { v.kind = Value::Bool; v.b = true; p += 4; }
        else if (s.compare(p, 5, "false") == 0) { v.kind = Value::Bool; v.b = false; p += 5; }
        else if (s.compare(p, 4, "null") == 0) { p += 4; }
        else {
            char* end = nullptr;
            v.num = std::strtod(s.c_str() + p, &end);
            if (end == s.c_str() + p) fail("value expected");
            v.kind = Value::Number;
            p = (size_t)(end - s.c_str());
        }
        return v;
    }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace mini_json
