// TEST INFRASTRUCTURE (oracle/): stand-in for ViennaCore's vcRNG.hpp.
// RNG is taken to be std::mt19937_64 and tea<N> the N-round Tiny Encryption
// hash of (val0, val1) popularised by the OptiX SDK samples -- both
// UNVERIFIED against ViennaCore 2.1.2 (absent).  Call site:
// /root/reference/include/viennaray/rayTraceKernel.hpp:120-121.  Only the
// statistical behaviour of the stream matters for parity.
#pragma once
#include <random>

namespace viennacore {
using RNG = std::mt19937_64;

template <unsigned N> unsigned int tea(unsigned int v0, unsigned int v1) {
  unsigned int sum = 0;
  for (unsigned n = 0; n < N; ++n) {
    sum += 0x9e3779b9u;
    v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + sum) ^ ((v1 >> 5) + 0xc8013ea4u);
    v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + sum) ^ ((v0 >> 5) + 0x7e95761eu);
  }
  return v0;
}
} // namespace viennacore
