// ORACLE — TEST INFRASTRUCTURE ONLY.  No-op stand-in for {fmt}: the reference only formats log text.
#pragma once
#include <string>
namespace fmt {
template <typename... A>
inline std::string format(const std::string& f, A&&...) { return f; }
template <typename... A>
inline void print(A&&...) {}
}  // namespace fmt
