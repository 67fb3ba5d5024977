// toml.hpp — the slice of the toml11 v4 interface the reference's reader is written against
// (`#include <toml.hpp>` in src/input/*.cpp, src/util/Utility.hpp), on top of this repository's TOML reader
// (../../host/toml_lite.hpp).
//
// The reference reads its inputs through toml11 (git submodule, not vendored in its tree, not available
// offline).  With this directory on the include path its reader compiles unmodified — both for the drop-in
// build on the B200 engine (compat/Makefile) and for the checker build (oracle/Makefile, target `ref`).
// Offered: toml::value, array, table, find / find_or / get / get_or, parse, format_error, make_error_info.
// Nothing of toml11 is copied.
#ifndef OCM_REF_SHIM_TOML_HPP
#define OCM_REF_SHIM_TOML_HPP

#include <algorithm>
#include <array>
#include <cassert>
#include <iomanip>
#include <iostream>
#include <cstdint>
#include <map>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../host/toml_lite.hpp"

namespace toml {

using value = tomllite::Value;
using array = value::array_type;
using table = value::table_type;
using value_t = tomllite::value_t;
using type_error = tomllite::type_error;
using syntax_error = tomllite::syntax_error;

inline value parse(const std::string& fname) { return tomllite::parse_file(fname); }

// ---- error formatting ------------------------------------------------------------------------------
struct error_info {
    std::string text;
};
inline void append_locs(std::string&) {}
template <class... Ts>
void append_locs(std::string& out, const value& v, const std::string& msg, const Ts&... rest) {
    out += "\n  --> ";
    if (v.line() > 0) out += "line " + std::to_string(v.line()) + ": ";
    out += msg;
    append_locs(out, rest...);
}
template <class... Ts>
error_info make_error_info(const std::string& title, const Ts&... rest) {
    error_info e{title};
    append_locs(e.text, rest...);
    return e;
}
inline std::string format_error(const error_info& e) { return e.text; }
template <class... Ts>
std::string format_error(const std::string& title, const value& v, const Ts&... rest) {
    return make_error_info(title, v, rest...).text;
}

// ---- get<T> ------------------------------------------------------------------------------------------
template <class T, class = void>
struct getter;

template <class T>
T get(const value& v) { return getter<T>::get(v); }

template <>
struct getter<value> { static value get(const value& v) { return v; } };
template <>
struct getter<bool> { static bool get(const value& v) { return v.as_boolean(); } };
template <>
struct getter<std::string> { static std::string get(const value& v) { return v.as_string(); } };
template <>
struct getter<array> { static array get(const value& v) { return v.as_array(); } };
template <>
struct getter<table> { static table get(const value& v) { return v.as_table(); } };
template <class T>
struct getter<T, std::enable_if_t<std::is_floating_point<T>::value>> {
    static T get(const value& v) { return (T)v.as_number(); }
};
template <class T>
struct getter<T, std::enable_if_t<std::is_integral<T>::value && !std::is_same<T, bool>::value>> {
    static T get(const value& v) { return (T)v.as_integer(); }
};
template <class T>
struct getter<std::vector<T>> {
    static std::vector<T> get(const value& v) {
        std::vector<T> out;
        for (const auto& e : v.as_array()) out.push_back(toml::get<T>(e));
        return out;
    }
};
template <class T, std::size_t N>
struct getter<std::array<T, N>> {
    static std::array<T, N> get(const value& v) {
        const auto& a = v.as_array();
        if (a.size() != N) throw std::out_of_range(v.where() + "array of " + std::to_string(a.size()) + " elements where " + std::to_string(N) + " are expected");
        std::array<T, N> out;
        for (std::size_t i = 0; i < N; i++) out[i] = toml::get<T>(a[i]);
        return out;
    }
};
template <class A, class B>
struct getter<std::pair<A, B>> {
    static std::pair<A, B> get(const value& v) {
        const auto& a = v.as_array();
        if (a.size() != 2) throw std::out_of_range(v.where() + "a pair needs an array of 2 elements");
        return {toml::get<A>(a[0]), toml::get<B>(a[1])};
    }
};
template <class T>
struct getter<std::map<std::string, T>> {
    static std::map<std::string, T> get(const value& v) {
        std::map<std::string, T> out;
        for (const auto& kv : v.as_table()) out.emplace(kv.first, toml::get<T>(kv.second));
        return out;
    }
};
template <class T>
struct getter<std::unordered_map<std::string, T>> {
    static std::unordered_map<std::string, T> get(const value& v) {
        std::unordered_map<std::string, T> out;
        for (const auto& kv : v.as_table()) out.emplace(kv.first, toml::get<T>(kv.second));
        return out;
    }
};

// ---- find ----------------------------------------------------------------------------------------------
inline const value& find(const value& v, const std::string& k) {
    if (!v.is_table()) throw type_error(v.where() + "toml::find(\"" + k + "\"): not a table but " + tomllite::type_name(v.type()));
    if (!v.contains(k)) throw std::out_of_range(v.where() + "key \"" + k + "\" not found");
    return v.at(k);
}
inline value& find(value& v, const std::string& k) {
    if (!v.is_table()) throw type_error(v.where() + "toml::find(\"" + k + "\"): not a table but " + tomllite::type_name(v.type()));
    if (!v.contains(k)) throw std::out_of_range(v.where() + "key \"" + k + "\" not found");
    return v.at(k);
}
inline const value& find(const value& v, std::size_t i) { return v.at(i); }
template <class K2, class... Ks>
const value& find(const value& v, const std::string& k, const K2& k2, const Ks&... ks) {
    return find(find(v, k), k2, ks...);
}
template <class T, class... Ks>
T find(const value& v, const Ks&... ks) { return toml::get<T>(find(v, ks...)); }

// ---- get_or / find_or ------------------------------------------------------------------------------------
template <class T>
T get_or(const value& v, const T& opt) {
    try { return toml::get<T>(v); } catch (...) { return opt; }
}
inline std::string get_or(const value& v, const char* opt) {
    try { return toml::get<std::string>(v); } catch (...) { return opt; }
}
template <class T>
T find_or(const value& v, const std::string& k, const T& opt) {
    if (!v.is_table() || !v.contains(k)) return opt;
    return get_or<T>(v.at(k), opt);
}
inline std::string find_or(const value& v, const std::string& k, const char* opt) {
    if (!v.is_table() || !v.contains(k)) return opt;
    return get_or(v.at(k), opt);
}

}  // namespace toml

#endif  // OCM_REF_SHIM_TOML_HPP
