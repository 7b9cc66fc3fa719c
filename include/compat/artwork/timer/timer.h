// artwork/timer/timer.h — stand-in for the author's wall-clock timer (st17-ceres/src/include/solver.hpp:253-254, 288):
// ns_timer::Timer with re_start() and last_elapsed(description) -> printable string; plus the logging macros the same
// translation units expect to be visible.
#ifndef STBA_COMPAT_ARTWORK_TIMER_H_
#define STBA_COMPAT_ARTWORK_TIMER_H_
#include <chrono>
#include <mutex>
#include <sstream>
#include <string>

#include "../logger/logger.h"
namespace ns_timer {
class Timer {
 public:
  Timer() { re_start(); }
  void re_start() { start_ = last_ = std::chrono::steady_clock::now(); }
  double last_elapsed_ms() {
    const auto now = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(now - last_).count();
    last_ = now;
    return ms;
  }
  std::string last_elapsed(const std::string& desc) {
    std::ostringstream s;
    s << "{'" << desc << "': " << last_elapsed_ms() << "(ms)}";
    return s.str();
  }
  std::string total_elapsed(const std::string& desc) {
    std::ostringstream s;
    s << "{'" << desc << "': " << std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - start_).count() << "(ms)}";
    return s.str();
  }

 private:
  std::chrono::steady_clock::time_point start_, last_;
};
}  // namespace ns_timer
#endif
