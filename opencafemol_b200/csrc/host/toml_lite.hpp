// toml_lite.hpp — a small TOML 1.0 reader (header-only, C++17, no dependencies).
//
// The reference parses its inputs with the third-party toml11 library (git submodule `toml11`, not vendored
// under the reference tree).  This is an independent reader of the TOML subset its inputs use — and a bit
// more: comments, bare / quoted / dotted keys, [tables], [[arrays of tables]], inline tables (newlines and
// trailing commas inside them are tolerated, as toml11 tolerates them), multi-line arrays, basic / literal /
// multi-line strings, integers (dec/hex/oct/bin, underscores), floats (inf, nan), booleans.  Date-times are
// kept as strings.  The value class carries the accessor names of toml11 (`is_table`, `as_array`, `contains`,
// `at`, ...), which is what the host code and the reference's reader are written against.
//
// tests/test_host_cpp.py checks it against Python's tomllib on every sample input.
#ifndef OCM_TOML_LITE_HPP
#define OCM_TOML_LITE_HPP

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <map>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

namespace tomllite {

enum class value_t { empty, boolean, integer, floating, string, array, table };

inline const char* type_name(value_t t) {
    switch (t) {
        case value_t::empty: return "empty";
        case value_t::boolean: return "boolean";
        case value_t::integer: return "integer";
        case value_t::floating: return "floating";
        case value_t::string: return "string";
        case value_t::array: return "array";
        case value_t::table: return "table";
    }
    return "?";
}
inline std::ostream& operator<<(std::ostream& os, value_t t) { return os << type_name(t); }

struct syntax_error : std::runtime_error { using std::runtime_error::runtime_error; };
struct type_error : std::runtime_error { using std::runtime_error::runtime_error; };

class Value {
  public:
    using array_type = std::vector<Value>;
    using table_type = std::map<std::string, Value>;

    Value() = default;
    Value(bool b) : t_(value_t::boolean), b_(b) {}
    Value(int v) : t_(value_t::integer), i_(v) {}
    Value(long v) : t_(value_t::integer), i_(v) {}
    Value(long long v) : t_(value_t::integer), i_(v) {}
    Value(unsigned long v) : t_(value_t::integer), i_((std::int64_t)v) {}
    Value(double v) : t_(value_t::floating), d_(v) {}
    Value(const char* s) : t_(value_t::string), s_(s) {}
    Value(std::string s) : t_(value_t::string), s_(std::move(s)) {}
    Value(array_type a) : t_(value_t::array), a_(std::move(a)) {}
    Value(table_type m) : t_(value_t::table), m_(std::move(m)) {}

    value_t type() const noexcept { return t_; }
    bool is_empty() const noexcept { return t_ == value_t::empty; }
    bool is_boolean() const noexcept { return t_ == value_t::boolean; }
    bool is_integer() const noexcept { return t_ == value_t::integer; }
    bool is_floating() const noexcept { return t_ == value_t::floating; }
    bool is_string() const noexcept { return t_ == value_t::string; }
    bool is_array() const noexcept { return t_ == value_t::array; }
    bool is_table() const noexcept { return t_ == value_t::table; }
    bool is_array_of_tables() const noexcept {
        if (!is_array() || a_.empty()) return false;
        for (const auto& v : a_) if (!v.is_table()) return false;
        return true;
    }

    bool as_boolean() const { need(value_t::boolean); return b_; }
    std::int64_t as_integer() const { need(value_t::integer); return i_; }
    double as_floating() const { need(value_t::floating); return d_; }
    const std::string& as_string() const { need(value_t::string); return s_; }
    std::string& as_string() { need(value_t::string); return s_; }
    const array_type& as_array() const { need(value_t::array); return a_; }
    array_type& as_array() { need(value_t::array); return a_; }
    const table_type& as_table() const { need(value_t::table); return m_; }
    table_type& as_table() { need(value_t::table); return m_; }

    // a number, whichever way it was written (toml11 would refuse an integer where a float is wanted; the
    // reference's inputs never rely on that refusal)
    double as_number() const {
        if (t_ == value_t::floating) return d_;
        if (t_ == value_t::integer) return (double)i_;
        throw type_error(where() + "expected a number, found " + type_name(t_));
    }

    bool contains(const std::string& k) const { return is_table() && m_.count(k) != 0; }
    std::size_t count(const std::string& k) const { return is_table() ? m_.count(k) : 0; }
    std::size_t size() const {
        if (is_array()) return a_.size();
        if (is_table()) return m_.size();
        if (is_string()) return s_.size();
        throw type_error(where() + "size() of a " + type_name(t_));
    }
    const Value& at(const std::string& k) const {
        need(value_t::table);
        auto it = m_.find(k);
        if (it == m_.end()) throw std::out_of_range(where() + "key \"" + k + "\" not found");
        return it->second;
    }
    Value& at(const std::string& k) {
        need(value_t::table);
        auto it = m_.find(k);
        if (it == m_.end()) throw std::out_of_range(where() + "key \"" + k + "\" not found");
        return it->second;
    }
    const Value& at(std::size_t i) const {
        need(value_t::array);
        if (i >= a_.size()) throw std::out_of_range(where() + "array index " + std::to_string(i) + " out of range");
        return a_[i];
    }
    Value& at(std::size_t i) {
        need(value_t::array);
        if (i >= a_.size()) throw std::out_of_range(where() + "array index " + std::to_string(i) + " out of range");
        return a_[i];
    }
    template <class I, std::enable_if_t<std::is_integral<I>::value, int> = 0>
    const Value& operator[](I i) const { return at((std::size_t)i); }
    template <class I, std::enable_if_t<std::is_integral<I>::value, int> = 0>
    Value& operator[](I i) { return at((std::size_t)i); }
    const Value& operator[](const std::string& k) const { return at(k); }
    Value& operator[](const std::string& k) { need(value_t::table); return m_[k]; }

    int line() const noexcept { return line_; }
    void set_line(int l) noexcept { line_ = l; }
    std::string where() const { return line_ > 0 ? "line " + std::to_string(line_) + ": " : std::string(); }

    // bookkeeping of the reader: a table made by [header] / dotted keys may be re-opened, an inline one may not
    bool sealed = false;
    bool defined = false;

  private:
    void need(value_t t) const {
        if (t_ != t) throw type_error(where() + "expected " + type_name(t) + ", found " + type_name(t_));
    }
    value_t t_ = value_t::empty;
    bool b_ = false;
    std::int64_t i_ = 0;
    double d_ = 0.0;
    std::string s_;
    array_type a_;
    table_type m_;
    int line_ = 0;
};

namespace detail {

class Reader {
  public:
    explicit Reader(const std::string& text, std::string name) : s_(text), name_(std::move(name)) {}

    Value parse() {
        Value root{Value::table_type{}};
        Value* cur = &root;
        skip_ws_nl();
        while (!eof()) {
            if (peek() == '[') {
                cur = header(root);
            } else {
                key_value(*cur);
            }
            skip_ws();
            if (!eof() && peek() == '#') skip_comment();
            if (!eof()) {
                if (peek() == '\n') ++pos_, ++line_;
                else if (peek() == '\r' && peek(1) == '\n') pos_ += 2, ++line_;
                else fail("expected end of line");
            }
            skip_ws_nl();
        }
        return root;
    }

  private:
    const std::string& s_;
    std::string name_;
    std::size_t pos_ = 0;
    int line_ = 1;

    bool eof() const { return pos_ >= s_.size(); }
    char peek(std::size_t o = 0) const { return pos_ + o < s_.size() ? s_[pos_ + o] : '\0'; }
    [[noreturn]] void fail(const std::string& m) const {
        throw syntax_error(name_ + ":" + std::to_string(line_) + ": " + m);
    }
    void skip_ws() { while (!eof() && (peek() == ' ' || peek() == '\t')) ++pos_; }
    void skip_comment() { while (!eof() && peek() != '\n') ++pos_; }
    void skip_ws_nl() {
        for (;;) {
            skip_ws();
            if (eof()) return;
            if (peek() == '#') skip_comment();
            else if (peek() == '\n') ++pos_, ++line_;
            else if (peek() == '\r') ++pos_;
            else return;
        }
    }
    static bool bare(char c) { return std::isalnum((unsigned char)c) || c == '_' || c == '-'; }

    std::string simple_key() {
        if (peek() == '"') return basic_string(false);
        if (peek() == '\'') return literal_string(false);
        std::size_t b = pos_;
        // bare keys are ASCII in TOML 1.0; bytes >= 0x80 are accepted too (the reference's inputs use "NPγT")
        while (!eof() && (bare(peek()) || (unsigned char)peek() >= 0x80)) ++pos_;
        if (pos_ == b) fail("expected a key");
        return s_.substr(b, pos_ - b);
    }
    std::vector<std::string> dotted_key() {
        std::vector<std::string> ks;
        for (;;) {
            skip_ws();
            ks.push_back(simple_key());
            skip_ws();
            if (peek() == '.') { ++pos_; continue; }
            return ks;
        }
    }

    Value* descend(Value& t, const std::string& k, bool for_header) {
        auto& m = t.as_table();
        auto it = m.find(k);
        if (it == m.end()) {
            Value nv{Value::table_type{}};
            nv.set_line(line_);
            it = m.emplace(k, std::move(nv)).first;
            return &it->second;
        }
        Value* v = &it->second;
        if (v->is_array() && for_header) {
            if (v->as_array().empty() || !v->as_array().back().is_table() || v->sealed) fail("key \"" + k + "\" is not a table");
            return &v->as_array().back();
        }
        if (!v->is_table() || v->sealed) fail("key \"" + k + "\" is already defined as a value");
        return v;
    }

    Value* header(Value& root) {
        ++pos_;
        bool aot = false;
        if (peek() == '[') { aot = true; ++pos_; }
        auto ks = dotted_key();
        if (peek() != ']') fail("expected ']'");
        ++pos_;
        if (aot) { if (peek() != ']') fail("expected ']]'"); ++pos_; }
        Value* t = &root;
        for (std::size_t i = 0; i + 1 < ks.size(); i++) t = descend(*t, ks[i], true);
        const std::string& last = ks.back();
        auto& m = t->as_table();
        if (aot) {
            auto it = m.find(last);
            if (it == m.end()) {
                Value arr{Value::array_type{}};
                arr.set_line(line_);
                it = m.emplace(last, std::move(arr)).first;
            } else if (!it->second.is_array() || it->second.sealed) {
                fail("key \"" + last + "\" is not an array of tables");
            }
            Value nt{Value::table_type{}};
            nt.set_line(line_);
            nt.defined = true;
            it->second.as_array().push_back(std::move(nt));
            return &it->second.as_array().back();
        }
        Value* v = descend(*t, last, true);
        if (v->defined) fail("table \"" + last + "\" defined twice");
        v->defined = true;
        return v;
    }

    void key_value(Value& table) {
        auto ks = dotted_key();
        if (peek() != '=') fail("expected '='");
        ++pos_;
        skip_ws();
        Value* t = &table;
        for (std::size_t i = 0; i + 1 < ks.size(); i++) t = descend(*t, ks[i], false);
        auto& m = t->as_table();
        if (m.count(ks.back())) fail("key \"" + ks.back() + "\" defined twice");
        const int l = line_;
        Value v = value();
        v.set_line(l);
        m.emplace(ks.back(), std::move(v));
    }

    Value value() {
        const char c = peek();
        if (c == '"') {
            if (peek(1) == '"' && peek(2) == '"') return Value(ml_basic_string());
            return Value(basic_string(true));
        }
        if (c == '\'') {
            if (peek(1) == '\'' && peek(2) == '\'') return Value(ml_literal_string());
            return Value(literal_string(true));
        }
        if (c == '[') return array();
        if (c == '{') return inline_table();
        if (s_.compare(pos_, 4, "true") == 0 && !bare(peek(4))) { pos_ += 4; return Value(true); }
        if (s_.compare(pos_, 5, "false") == 0 && !bare(peek(5))) { pos_ += 5; return Value(false); }
        return number_or_date();
    }

    Value array() {
        ++pos_;
        Value::array_type a;
        for (;;) {
            skip_ws_nl();
            if (peek() == ']') { ++pos_; break; }
            const int l = line_;
            Value v = value();
            v.set_line(l);
            a.push_back(std::move(v));
            skip_ws_nl();
            if (peek() == ',') { ++pos_; continue; }
            if (peek() == ']') { ++pos_; break; }
            fail("expected ',' or ']' in array");
        }
        Value out{std::move(a)};
        out.sealed = true;
        return out;
    }

    Value inline_table() {
        ++pos_;
        Value t{Value::table_type{}};
        for (;;) {
            skip_ws_nl();
            if (peek() == '}') { ++pos_; break; }
            key_value(t);
            skip_ws_nl();
            if (peek() == ',') { ++pos_; continue; }
            if (peek() == '}') { ++pos_; break; }
            fail("expected ',' or '}' in inline table");
        }
        seal(t);
        return t;
    }
    static void seal(Value& v) {
        v.sealed = true;
        if (v.is_table()) for (auto& kv : v.as_table()) if (kv.second.is_table()) seal(kv.second);
    }

    void append_utf8(std::string& out, unsigned long cp) {
        if (cp < 0x80) out += (char)cp;
        else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
        else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
        else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
    }
    void escape(std::string& out) {
        // at the character after the backslash
        const char e = peek();
        ++pos_;
        switch (e) {
            case 'b': out += '\b'; break;
            case 't': out += '\t'; break;
            case 'n': out += '\n'; break;
            case 'f': out += '\f'; break;
            case 'r': out += '\r'; break;
            case '"': out += '"'; break;
            case '\\': out += '\\'; break;
            case 'u': case 'U': {
                const int n = e == 'u' ? 4 : 8;
                if (pos_ + n > s_.size()) fail("truncated unicode escape");
                append_utf8(out, std::strtoul(s_.substr(pos_, n).c_str(), nullptr, 16));
                pos_ += n;
                break;
            }
            default: fail(std::string("unknown escape \\") + e);
        }
    }
    std::string basic_string(bool) {
        ++pos_;
        std::string out;
        for (;;) {
            if (eof() || peek() == '\n') fail("unterminated string");
            const char c = peek();
            ++pos_;
            if (c == '"') return out;
            if (c == '\\') escape(out);
            else out += c;
        }
    }
    std::string literal_string(bool) {
        ++pos_;
        const std::size_t b = pos_;
        while (!eof() && peek() != '\'' && peek() != '\n') ++pos_;
        if (peek() != '\'') fail("unterminated literal string");
        std::string out = s_.substr(b, pos_ - b);
        ++pos_;
        return out;
    }
    std::string ml_basic_string() {
        pos_ += 3;
        if (peek() == '\r' && peek(1) == '\n') pos_ += 2, ++line_;
        else if (peek() == '\n') ++pos_, ++line_;
        std::string out;
        for (;;) {
            if (eof()) fail("unterminated multi-line string");
            if (peek() == '"' && peek(1) == '"' && peek(2) == '"') {
                pos_ += 3;
                while (peek() == '"') { out += '"'; ++pos_; }  // up to two quotes may precede the delimiter
                return out;
            }
            const char c = peek();
            ++pos_;
            if (c == '\\') {
                std::size_t p = pos_;
                while (p < s_.size() && (s_[p] == ' ' || s_[p] == '\t' || s_[p] == '\r')) ++p;
                if (p < s_.size() && s_[p] == '\n') {  // line-ending backslash: trim following whitespace
                    pos_ = p;
                    while (!eof() && (peek() == ' ' || peek() == '\t' || peek() == '\n' || peek() == '\r')) {
                        if (peek() == '\n') ++line_;
                        ++pos_;
                    }
                } else {
                    escape(out);
                }
            } else {
                if (c == '\n') ++line_;
                out += c;
            }
        }
    }
    std::string ml_literal_string() {
        pos_ += 3;
        if (peek() == '\r' && peek(1) == '\n') pos_ += 2, ++line_;
        else if (peek() == '\n') ++pos_, ++line_;
        std::string out;
        for (;;) {
            if (eof()) fail("unterminated multi-line literal string");
            if (peek() == '\'' && peek(1) == '\'' && peek(2) == '\'') {
                pos_ += 3;
                while (peek() == '\'') { out += '\''; ++pos_; }
                return out;
            }
            if (peek() == '\n') ++line_;
            out += peek();
            ++pos_;
        }
    }

    Value number_or_date() {
        const std::size_t b = pos_;
        while (!eof()) {
            const char c = peek();
            if (std::isalnum((unsigned char)c) || c == '+' || c == '-' || c == '.' || c == '_' || c == ':') ++pos_;
            else break;
        }
        if (pos_ == b) fail("expected a value");
        std::string tok = s_.substr(b, pos_ - b);
        std::string t;
        for (char c : tok) if (c != '_') t += c;
        const std::string u = (t[0] == '+' || t[0] == '-') ? t.substr(1) : t;
        const double sgn = t[0] == '-' ? -1.0 : 1.0;
        if (u == "inf") return Value(sgn * std::numeric_limits<double>::infinity());
        if (u == "nan") return Value(std::numeric_limits<double>::quiet_NaN());
        if (u.size() > 2 && u[0] == '0' && (u[1] == 'x' || u[1] == 'o' || u[1] == 'b')) {
            const int base = u[1] == 'x' ? 16 : (u[1] == 'o' ? 8 : 2);
            char* end = nullptr;
            const long long v = std::strtoll(u.c_str() + 2, &end, base);
            if (*end) fail("malformed integer \"" + tok + "\"");
            return Value((long long)v);
        }
        bool digits_only = !u.empty();
        for (char c : u) if (!std::isdigit((unsigned char)c)) { digits_only = false; break; }
        if (digits_only) {
            char* end = nullptr;
            const long long v = std::strtoll(t.c_str(), &end, 10);
            if (*end) fail("malformed integer \"" + tok + "\"");
            return Value((long long)v);
        }
        bool floaty = std::isdigit((unsigned char)u[0]) != 0;
        for (char c : u) if (!(std::isdigit((unsigned char)c) || c == '.' || c == 'e' || c == 'E' || c == '+' || c == '-')) { floaty = false; break; }
        if (floaty) {
            char* end = nullptr;
            const double v = std::strtod(t.c_str(), &end);
            if (*end) fail("malformed number \"" + tok + "\"");
            return Value(v);
        }
        // date / time: kept verbatim (a following " hh:mm:ss" part is taken too)
        if (std::isdigit((unsigned char)u[0]) && (tok.find('-') != std::string::npos || tok.find(':') != std::string::npos)) {
            if (peek() == ' ' && std::isdigit((unsigned char)peek(1)) && std::isdigit((unsigned char)peek(2)) && peek(3) == ':') {
                ++pos_;
                const std::size_t b2 = pos_;
                while (!eof() && (std::isalnum((unsigned char)peek()) || peek() == ':' || peek() == '.' || peek() == '+' || peek() == '-')) ++pos_;
                tok += " " + s_.substr(b2, pos_ - b2);
            }
            return Value(tok);
        }
        fail("cannot read value \"" + tok + "\"");
    }
};

}  // namespace detail

inline Value parse_string(const std::string& text, const std::string& name = "<string>") {
    return detail::Reader(text, name).parse();
}

inline Value parse_file(const std::string& path) {
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open \"" + path + "\"");
    std::ostringstream ss;
    ss << in.rdbuf();
    const std::string text = ss.str();
    return parse_string(text, path);
}

// JSON dump (used by the parser cross-check against Python's tomllib)
inline void to_json(const Value& v, std::ostream& os) {
    switch (v.type()) {
        case value_t::empty: os << "null"; break;
        case value_t::boolean: os << (v.as_boolean() ? "true" : "false"); break;
        case value_t::integer: os << v.as_integer(); break;
        case value_t::floating: {
            const double d = v.as_floating();
            if (std::isnan(d)) os << "\"nan\"";
            else if (std::isinf(d)) os << (d > 0 ? "\"inf\"" : "\"-inf\"");
            else { std::ostringstream t; t.precision(17); t << d; std::string s = t.str(); if (s.find_first_of(".en") == std::string::npos) s += ".0"; os << s; }
            break;
        }
        case value_t::string: {
            os << '"';
            for (unsigned char c : v.as_string()) {
                if (c == '"' || c == '\\') os << '\\' << c;
                else if (c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); os << b; }
                else os << c;
            }
            os << '"';
            break;
        }
        case value_t::array: {
            os << '[';
            bool first = true;
            for (const auto& e : v.as_array()) { if (!first) os << ','; first = false; to_json(e, os); }
            os << ']';
            break;
        }
        case value_t::table: {
            os << '{';
            bool first = true;
            for (const auto& kv : v.as_table()) { if (!first) os << ','; first = false; to_json(Value(kv.first), os); os << ':'; to_json(kv.second, os); }
            os << '}';
            break;
        }
    }
}

}  // namespace tomllite

#endif  // OCM_TOML_LITE_HPP
