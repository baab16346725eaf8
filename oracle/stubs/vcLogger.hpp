// TEST INFRASTRUCTURE (oracle/): stand-in for ViennaCore's vcLogger.hpp.
// Errors print and continue (whether upstream aborts is UNVERIFIED).
#pragma once
#include <iostream>
#include <string>
#include <vcTimer.hpp>

namespace viennacore {
enum class LogLevel : unsigned {
  ERROR = 0,
  WARNING = 1,
  INFO = 2,
  INTERMEDIATE = 3,
  TIMING = 4,
  DEBUG = 5
};
class Logger {
public:
  static LogLevel &level() {
    static LogLevel l = LogLevel::WARNING;
    return l;
  }
  static void setLogLevel(LogLevel l) { level() = l; }
  static LogLevel getLogLevel() { return level(); }
  static void emit(LogLevel l, const char *tag, const std::string &msg) {
    if (static_cast<unsigned>(l) <= static_cast<unsigned>(level()))
      std::cerr << "[" << tag << "] " << msg << std::endl;
  }
};
} // namespace viennacore
#define VIENNACORE_LOG_ERROR(msg)                                              \
  ::viennacore::Logger::emit(::viennacore::LogLevel::ERROR, "ERROR", (msg))
#define VIENNACORE_LOG_WARNING(msg)                                            \
  ::viennacore::Logger::emit(::viennacore::LogLevel::WARNING, "WARNING", (msg))
#define VIENNACORE_LOG_DEBUG(msg)                                              \
  ::viennacore::Logger::emit(::viennacore::LogLevel::DEBUG, "DEBUG", (msg))
