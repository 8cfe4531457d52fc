// super4pcs-b200: GlobalRegistration::Utils::Logger, interface-compatible with the reference's
// src/super4pcs/utils/logger.h:55-119 (three levels; errors to stderr, the rest to stdout).
#ifndef SUPER4PCS_B200_UTILS_LOGGER_H_
#define SUPER4PCS_B200_UTILS_LOGGER_H_

#include <iostream>

namespace GlobalRegistration {
namespace Utils {

enum LogLevel { NoLog = 0, ErrorReport = 1, Verbose = 2 };

class Logger {
 public:
  Logger(LogLevel level = Verbose) : level_(level) {}
  void setLogLevel(LogLevel level) { level_ = level; }
  LogLevel logLevel() const { return level_; }

  /// prints the arguments back to back followed by a newline when the logger's level admits
  /// messages of level `msg`
  template <LogLevel msg, typename... Args>
  void Log(const Args&... args) const {
    if (static_cast<int>(level_) < static_cast<int>(msg)) return;
    // (like the reference, the STREAM depends on the logger's level, not on the message's)
    std::ostream& os = (level_ == ErrorReport) ? std::cerr : std::cout;
    emit(os, args...);
    os << std::endl;
  }

 private:
  static void emit(std::ostream&) {}
  template <typename First, typename... Rest>
  static void emit(std::ostream& os, const First& first, const Rest&... rest) {
    os << first;
    emit(os, rest...);
  }
  LogLevel level_;
};

}  // namespace Utils
}  // namespace GlobalRegistration

#endif  // SUPER4PCS_B200_UTILS_LOGGER_H_
