// TEST INFRASTRUCTURE (oracle/): stand-in for ViennaCore's vcTimer.hpp
// (nanosecond wall timer; used at rayTraceKernel.hpp:84-85,341,416).
#pragma once
#include <chrono>
namespace viennacore {
struct Timer {
  using clock = std::chrono::steady_clock;
  clock::time_point t0{};
  long long currentDuration = 0; // ns
  long long totalDuration = 0;
  void start() { t0 = clock::now(); }
  void finish() {
    currentDuration = std::chrono::duration_cast<std::chrono::nanoseconds>(
                          clock::now() - t0)
                          .count();
    totalDuration += currentDuration;
  }
  void reset() { currentDuration = totalDuration = 0; }
};
} // namespace viennacore
