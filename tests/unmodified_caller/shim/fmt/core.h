// TEST INFRASTRUCTURE.  The few {fmt} features src/motion_planning.cpp uses — "{}" and "{:.Nf}" — so that the text
// it draws on the plot (x, y, v, yaw, acc, steer of every tick) carries real numbers the test can read back.
#pragma once
#include <cstdio>
#include <string>
#include <type_traits>
namespace fmt {
namespace detail {
inline void put(std::string& out, const std::string& spec, double v) {
    char buf[64];
    int prec = 6;
    size_t dot = spec.find('.');
    if (dot != std::string::npos) prec = std::atoi(spec.c_str() + dot + 1);
    std::snprintf(buf, sizeof buf, "%.*f", prec, v);
    out += spec.empty() ? std::to_string(v) : std::string(buf);
}
inline void put(std::string& out, const std::string&, const std::string& v) { out += v; }
inline void put(std::string& out, const std::string&, const char* v) { out += v; }
template <typename I, typename = std::enable_if_t<std::is_integral<I>::value>>
inline void put(std::string& out, const std::string&, I v) { out += std::to_string(v); }
inline void expand(std::string& out, const char* f) { out += f; }
template <typename A, typename... R>
inline void expand(std::string& out, const char* f, const A& a, const R&... rest) {
    while (*f) {
        if (*f == '{') {
            const char* e = f;
            while (*e && *e != '}') ++e;
            std::string spec(f + 1, e);
            if (!spec.empty() && spec[0] == ':') spec.erase(0, 1);
            put(out, spec, a);
            expand(out, *e ? e + 1 : e, rest...);
            return;
        }
        out += *f++;
    }
}
}  // namespace detail
template <typename... A>
inline std::string format(const char* f, const A&... a) {
    std::string out;
    detail::expand(out, f, a...);
    return out;
}
template <typename... A>
inline void print(const char* f, const A&... a) { std::fputs(format(f, a...).c_str(), stderr); }
}  // namespace fmt
