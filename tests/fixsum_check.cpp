// CPU check that the product's order-independent accumulator (lidar_graph_slam_b200/csrc/fixsum.cuh, host path) and the
// oracle's (oracle/exactsum.hpp) are the same function of the multiset of terms: random and adversarial doubles, any
// order, any grouping into partial sums.  Test infrastructure (it includes the oracle header).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <random>
#include <vector>

#include "../lidar_graph_slam_b200/csrc/fixsum.cuh"
#include "../oracle/exactsum.hpp"

static uint64_t bits(double x) {
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return u;
}

int main() {
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> uni(-1.0, 1.0);
  int failures = 0;
  for (int trial = 0; trial < 200; trial++) {
    std::vector<double> terms;
    const int n = 1 + static_cast<int>(rng() % 5000);
    for (int i = 0; i < n; i++) {
      const int kind = static_cast<int>(rng() % 10);
      double v = uni(rng) * std::ldexp(1.0, static_cast<int>(rng() % 80) - 60);   // 2^-60 .. 2^20
      if (kind == 0) v = static_cast<double>(static_cast<float>(v));               // f32-derived, as the cost terms are
      if (kind == 1) v = 0.0;
      if (kind == 2) v = -0.0;
      if (kind == 3) v = std::ldexp(uni(rng), -1060);                              // denormal
      if (kind == 4 && i < 3) v = std::ldexp(uni(rng), 45);                        // a few terms just inside the range (the SUM must stay below 2^47)
      terms.push_back(v);
    }
    if (trial == 7) terms.push_back(std::ldexp(1.0, 46));                          // out of range: both must poison
    if (trial == 8) terms.push_back(std::numeric_limits<double>::quiet_NaN());
    if (trial == 9) terms.push_back(-std::numeric_limits<double>::infinity());
    lgs_oracle::ExactSum ref;
    for (double t : terms) ref.add(t);
    // product accumulator, one pass
    lgs::Fix128 a = lgs::fix_zero();
    for (double t : terms) lgs::fix_add(a, t);
    // product accumulator, shuffled and grouped into partial sums that are then added (what warps / CTAs do)
    std::vector<double> sh(terms);
    for (size_t i = sh.size(); i > 1; i--) std::swap(sh[i - 1], sh[rng() % i]);
    lgs::Fix128 total = lgs::fix_zero();
    size_t i = 0;
    while (i < sh.size()) {
      lgs::Fix128 part = lgs::fix_zero();
      const size_t len = 1 + rng() % 97;
      for (size_t j = 0; j < len && i < sh.size(); j++, i++) lgs::fix_add(part, sh[i]);
      lgs::fix_add(total, part);
    }
    const double r = ref.value(), v1 = lgs::fix_value(a), v2 = lgs::fix_value(total);
    const bool same = (std::isnan(r) && std::isnan(v1) && std::isnan(v2)) || (bits(r) == bits(v1) && bits(r) == bits(v2));
    // and close to the long-double sum (the accumulator is exact above 2^-80 per term)
    long double ld = 0;
    for (double t : terms) ld += t;
    const bool close = std::isnan(r) || std::fabs(static_cast<double>(ld) - r) <= 1e-12 * std::fabs(static_cast<double>(ld)) + n * std::ldexp(1.0, -79);
    if ((!same || !close) && failures < 5) {
      std::printf("trial %d: oracle %.17g product %.17g grouped %.17g long double %.17Lg\n", trial, r, v1, v2, ld);
    }
    if (!same || !close) failures++;
    if ((trial == 7 || trial == 8 || trial == 9) && !std::isnan(r)) {
      std::printf("trial %d: out-of-range term did not poison the sum\n", trial);
      failures++;
    }
  }
  std::printf("fixsum check: %d failures\n", failures);
  return failures ? 1 : 0;
}
