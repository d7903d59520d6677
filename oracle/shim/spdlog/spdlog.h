// ORACLE — TEST INFRASTRUCTURE ONLY.  No-op stand-in for spdlog: log statements compile to nothing.
#pragma once
#define SPDLOG_ERROR(...) ((void)0)
#define SPDLOG_WARN(...) ((void)0)
#define SPDLOG_INFO(...) ((void)0)
#define SPDLOG_DEBUG(...) ((void)0)
namespace spdlog {
namespace level { enum level_enum { trace, debug, info, warn, err }; }
inline void set_level(int) {}
template <typename... A>
inline void error(A&&...) {}
template <typename... A>
inline void info(A&&...) {}
}  // namespace spdlog
