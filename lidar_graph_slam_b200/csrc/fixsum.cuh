// Order-independent summation of doubles: every term is converted (magnitude truncated below 2^-80, exact above it
// for any f32-derived value >= 2^-57 and any f64 >= 2^-28) to a 128-bit two's-complement fixed-point number with
// quantum 2^-80 and added as an integer.  Integer addition is associative, so the sum is bit-identical for every
// decomposition of the work over threads, warps, CTAs and GPUs - unlike the reference's per-thread partial sums
// (gicp_omp_impl.hpp:251,274,291-314), whose low bits depend on the OpenMP thread count.  The BFGS line search of
// the PCL-style GICP compares cost values that differ by less than the rounding noise of a 30 000-term f64 sum;
// with exact sums its trajectory is reproducible.  Range: |term| < 2^46 (larger, inf or NaN poisons the sum -> NaN) and |sum| < 2^47 (not checked: the cost terms are
// below 2^14 and there are fewer than 2^22 of them).
#pragma once
#include <cstdint>

namespace lgs {

struct Fix128 {
  unsigned long long lo;
  long long hi;  // value = (hi * 2^64 + lo) * 2^-80; hi == INT64_MIN marks a poisoned sum
};

#ifdef __CUDACC__
#define LGS_FIX_HD __host__ __device__ __forceinline__
#else
#define LGS_FIX_HD inline
#endif

LGS_FIX_HD Fix128 fix_zero() { return Fix128{0ull, 0ll}; }
LGS_FIX_HD bool fix_bad(const Fix128& a) { return a.hi == INT64_MIN; }

LGS_FIX_HD void fix_add(Fix128& a, const Fix128& b) {
  if (fix_bad(a) || fix_bad(b)) {
    a.hi = INT64_MIN;
    return;
  }
  const unsigned long long lo = a.lo + b.lo;
  a.hi = static_cast<long long>(static_cast<unsigned long long>(a.hi) + static_cast<unsigned long long>(b.hi) + (lo < a.lo ? 1ull : 0ull));
  a.lo = lo;
}

LGS_FIX_HD void fix_add(Fix128& a, double t) {
  unsigned long long bits;
#ifdef __CUDA_ARCH__
  bits = static_cast<unsigned long long>(__double_as_longlong(t));
#else
  __builtin_memcpy(&bits, &t, 8);
#endif
  int ex = static_cast<int>((bits >> 52) & 0x7ffull);
  unsigned long long man = bits & ((1ull << 52) - 1ull);
  if (ex)
    man |= 1ull << 52;
  else
    ex = 1;
  const int sh = ex - 995;  // man * 2^(ex - 1075) in units of 2^-80
  if (sh > 73 || fix_bad(a)) {  // |t| >= 2^46, inf or NaN
    a.hi = INT64_MIN;
    return;
  }
  unsigned long long lo, hi;
  if (sh >= 64) {
    lo = 0ull;
    hi = man << (sh - 64);
  } else if (sh > 0) {
    lo = man << sh;
    hi = man >> (64 - sh);
  } else if (sh > -64) {
    lo = man >> (-sh);
    hi = 0ull;
  } else {
    lo = hi = 0ull;
  }
  if (bits >> 63) {  // subtract the magnitude
    const unsigned long long nlo = a.lo - lo;
    a.hi = static_cast<long long>(static_cast<unsigned long long>(a.hi) - hi - (a.lo < lo ? 1ull : 0ull));
    a.lo = nlo;
  } else {
    const unsigned long long nlo = a.lo + lo;
    a.hi = static_cast<long long>(static_cast<unsigned long long>(a.hi) + hi + (nlo < a.lo ? 1ull : 0ull));
    a.lo = nlo;
  }
}

// nearest-ish double of the sum: two exact-integer conversions and one addition, the same on host and device
LGS_FIX_HD double fix_value(const Fix128& a) {
  if (fix_bad(a)) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(0x7ff8000000000000ll);
#else
    return __builtin_nan("");
#endif
  }
  const double h = static_cast<double>(a.hi) * 1.52587890625e-05;                      // 2^-16
  const double l = static_cast<double>(a.lo) * 8.271806125530277e-25;                  // 2^-80
  return h + l;
}

}  // namespace lgs
