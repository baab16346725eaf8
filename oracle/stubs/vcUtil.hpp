// TEST INFRASTRUCTURE (oracle/): stand-in for ViennaCore's vcUtil.hpp.
#pragma once
#include <cstddef>
#include <vcTimer.hpp>
namespace viennacore {
namespace util {
inline void ProgressBar(size_t, size_t) {}
} // namespace util
} // namespace viennacore
