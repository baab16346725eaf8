// TEST INFRASTRUCTURE (oracle/): stand-in for ViennaCore's vcTestAsserts.hpp
// so the reference's own tests/ can be compiled against oracle/mini_rtc.cpp.
#pragma once
#include <cmath>
#include <cstdlib>
#include <iostream>
#define VC_TEST_ASSERT(cond)                                                   \
  {                                                                            \
    if (!(cond)) {                                                             \
      std::cerr << "VC_TEST_ASSERT failed: " #cond " at " << __FILE__ << ":"   \
                << __LINE__ << std::endl;                                      \
      std::exit(1);                                                            \
    }                                                                          \
  }
#define VC_TEST_ASSERT_ISCLOSE(a, b, eps)                                      \
  {                                                                            \
    if (!(std::fabs(double(a) - double(b)) <= double(eps))) {                  \
      std::cerr << "VC_TEST_ASSERT_ISCLOSE failed: " #a "=" << (a)             \
                << " vs " #b "=" << (b) << " at " << __FILE__ << ":"           \
                << __LINE__ << std::endl;                                      \
      std::exit(1);                                                            \
    }                                                                          \
  }
// runs viennacore::RunTest<T, D>() for the four (float|double) x (2|3) cases
#define VC_RUN_ALL_TESTS                                                       \
  viennacore::RunTest<double, 2>();                                            \
  viennacore::RunTest<double, 3>();                                            \
  viennacore::RunTest<float, 2>();                                             \
  viennacore::RunTest<float, 3>();
