// Pins ndtopt::sinf_libm / cosf_libm (csrc/ndt_opt.cuh), the restatement of glibc's sinf / cosf that the device-resident
// NDT optimiser evaluates on the GPU, against the C library of this machine: bit for bit over a strided sweep of every
// |x| < 120 plus the neighbourhood of the quadrant boundaries.  (The full 2.2e9-value sweep was run once: 0 differences.)
//   g++ -O2 -std=c++17 -I lidar_graph_slam_b200/csrc tests/sincos_check.cpp -o sincos_check -lm && ./sincos_check
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "ndt_opt.cuh"

int main(int argc, char** argv) {
  const uint32_t stride = argc > 1 ? static_cast<uint32_t>(atoi(argv[1])) : 97u;
  long bad = 0, total = 0;
  auto check = [&](float v) {
    const float a = sinf(v), b = lgs::ndtopt::sinf_libm(v), c = cosf(v), d = lgs::ndtopt::cosf_libm(v);
    if (memcmp(&a, &b, 4) || memcmp(&c, &d, 4)) {
      if (bad < 10) printf("x = %a: sinf %a / %a  cosf %a / %a\n", v, a, b, c, d);
      bad++;
    }
    total++;
  };
  for (uint32_t u = 0; u < 0x42f00000u; u += stride) {
    float y;
    memcpy(&y, &u, 4);
    check(y);
    check(-y);
  }
  for (int k = 1; k < 76; k++) {  // around every multiple of pi/4 below 60
    float c = static_cast<float>(k * 0.78539816339744830962);
    uint32_t u;
    memcpy(&u, &c, 4);
    for (int d = -2000; d <= 2000; d++) {
      uint32_t w = u + d;
      float y;
      memcpy(&y, &w, 4);
      check(y);
      check(-y);
    }
  }
  printf("%ld values, %ld failures\n", total, bad);
  return bad ? 1 : 0;
}
