// ORACLE (test infrastructure).  Order-independent accumulation used by the pclomp-GICP functor sums: each term is
// converted to 128-bit fixed point (quantum 2^-80, magnitude truncated) and added as an integer, so the total does
// not depend on the order of the terms.  The reference's own order is unspecified (per-thread partial sums added
// in thread order, gicp_omp_impl.hpp:251,274,291-314: the low bits change with the OpenMP thread count), and its
// BFGS line search compares costs closer than the rounding noise of a plain f64 sum; an exact sum is the one
// statement of "the sum" that every implementation can reproduce bit for bit.  Range |term| < 2^46, else NaN; |sum| < 2^47 (not checked: the
// cost terms are below 2^14 and there are fewer than 2^22 of them).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace lgs_oracle {

struct ExactSum {
  __int128 acc = 0;
  bool poisoned = false;

  void add(double t) {
    uint64_t bits;
    std::memcpy(&bits, &t, 8);
    int ex = static_cast<int>((bits >> 52) & 0x7ff);
    uint64_t man = bits & ((uint64_t(1) << 52) - 1);
    if (ex)
      man |= uint64_t(1) << 52;
    else
      ex = 1;
    const int sh = ex - 995;  // man * 2^(ex - 1075) in units of 2^-80
    if (sh > 73) {
      poisoned = true;
      return;
    }
    unsigned __int128 v = 0;
    if (sh >= 0)
      v = static_cast<unsigned __int128>(man) << sh;
    else if (sh > -64)
      v = man >> (-sh);
    if (bits >> 63)
      acc -= static_cast<__int128>(v);
    else
      acc += static_cast<__int128>(v);
  }

  double value() const {
    if (poisoned) return std::numeric_limits<double>::quiet_NaN();
    const int64_t hi = static_cast<int64_t>(acc >> 64);           // arithmetic shift: floor
    const uint64_t lo = static_cast<uint64_t>(static_cast<unsigned __int128>(acc));
    return static_cast<double>(hi) * 1.52587890625e-05 + static_cast<double>(lo) * 8.271806125530277e-25;  // 2^-16, 2^-80
  }
};

}  // namespace lgs_oracle
