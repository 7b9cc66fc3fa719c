// artwork/logger/logger.h — stand-in for the author's logging macros used by the reference's solver drivers
// (st17-ceres/src/include/solver.hpp:251-292, st20-g2o/src/include/test_ceres.h:100-151): variadic print to stdout.
#ifndef STBA_COMPAT_ARTWORK_LOGGER_H_
#define STBA_COMPAT_ARTWORK_LOGGER_H_
#include <iostream>
#include <sstream>
namespace ns_log_compat {
inline void put(std::ostream&) {}
template <typename A, typename... R>
inline void put(std::ostream& s, const A& a, const R&... r) { s << a; put(s, r...); }
template <typename... A>
inline void line(const char* tag, const A&... a) { std::ostringstream s; s << tag; put(s, a...); std::cout << s.str() << std::endl; }
}  // namespace ns_log_compat
#define LOG_INFO(...) ns_log_compat::line("[ info ] ", __VA_ARGS__);
#define LOG_PROCESS(...) ns_log_compat::line("[ process ] ", __VA_ARGS__);
#define LOG_PLAINTEXT(...) ns_log_compat::line("", __VA_ARGS__);
#define LOG_WARNING(...) ns_log_compat::line("[ warning ] ", __VA_ARGS__);
#define LOG_ERROR(...) ns_log_compat::line("[ error ] ", __VA_ARGS__);
#define LOG_VAR(...) ns_log_compat::line("[ var ] ", __VA_ARGS__);
#define LOG_ENDL() std::cout << std::endl;
#endif
