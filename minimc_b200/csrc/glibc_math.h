// Bit-exact restatement of glibc 2.39's double-precision log, sincos, sin and
// cos as executed on x86-64 CPUs with FMA + AVX2 (the `_fma` ifunc variants,
// which every host this repo targets selects), for use inside CUDA kernels.
//
// Why the product needs this: the reference obtains collision distances from
// std::log and directions from sincos / sin / cos of glibc (TransportMethod.cpp:57
// via std::exponential_distribution, Point.cpp:89-96,117-118).  Its
// surface-crossing nudge is an ABSOLUTE 10 eps (Constants.hpp:15), smaller than
// one ulp of most coordinates, so whether a particle that crosses a surface at a
// shallow angle lands in the next cell is decided by the last bit of its
// position.  CUDA's libdevice log / sincos differ from glibc in the last ulp for
// a few percent of arguments, which flipped about one event sequence per 10^5
// histories in geometry-rich decks.  With these functions positions and
// directions are bit-identical to the reference's, so event sequences are too.
//
// What is restated (algorithms as published in glibc, sysdeps/ieee754/dbl-64):
//   * e_log.c   -- Szabolcs Nagy's table-driven log (ARM optimized routines),
//                  N = 128 subintervals, the __FP_FAST_FMA path;
//   * s_sin.c / s_sincos.c -- the IBM Accurate Mathematical Library sin/cos as
//                  simplified in glibc 2.28: do_sin / do_cos around a 1/128-spaced
//                  table of double-double sin/cos, TAYLOR_SIN below 0.126,
//                  reduce_sincos (three-part pi/2) up to |x| < 105414350.
// The C sources leave FMA contraction to the compiler; the contraction used here
// is the one in the shipped binary (read from the disassembly of __log_fma,
// __sincos_fma, __sin_fma, __cos_fma), and tests/test_glibc_math.py verifies bit
// equality against the box's libm for millions of arguments per branch.
// Arguments outside the restated domain (non-finite, negative for log, |x| >=
// 105414350 for sin/cos, subnormal) fall back to the platform's function; the
// transport loop never produces them.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#include "glibc_tables.h"

#if defined(__CUDACC__)
#define MMC_MATH_FN __device__ __forceinline__
#define MMC_MATH_TABLE __device__ const
#else
#define MMC_MATH_FN inline
#define MMC_MATH_TABLE static const
#endif

namespace mmc {
namespace glibc {

MMC_MATH_TABLE double kLogData[MMC_GLIBC_LOG_DATA_N] = {MMC_GLIBC_LOG_DATA};
MMC_MATH_TABLE double kSinCosTab[MMC_GLIBC_SINCOSTAB_N] = {MMC_GLIBC_SINCOSTAB};

MMC_MATH_FN uint64_t as_u64(double x) {
#if defined(__CUDA_ARCH__)
  return static_cast<uint64_t>(__double_as_longlong(x));
#else
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return u;
#endif
}
MMC_MATH_FN double as_f64(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(static_cast<long long>(u));
#else
  double x;
  std::memcpy(&x, &u, 8);
  return x;
#endif
}
// one rounding each, never contracted (the translation unit is built with
// -fmad=false / -ffp-contract=off; the intrinsics make it explicit on device)
MMC_MATH_FN double add(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
MMC_MATH_FN double sub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
MMC_MATH_FN double mul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
MMC_MATH_FN double fma_(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
MMC_MATH_FN double fnma(double a, double b, double c) { return fma_(-a, b, c); }  // -(a*b) + c
MMC_MATH_FN double copysign_(double mag, double sgn) {
  return as_f64((as_u64(mag) & 0x7fffffffffffffffull) | (as_u64(sgn) & 0x8000000000000000ull));
}
MMC_MATH_FN double fabs_(double x) { return as_f64(as_u64(x) & 0x7fffffffffffffffull); }

// ------------------------------------------------------------------------ log
// e_log.c: __log, __FP_FAST_FMA path.  kLogData = {ln2hi, ln2lo, A[0..4], B[0..10], T[128]{invc, logc}}
MMC_MATH_FN double log(double x) {
  const double* A = kLogData + 2;
  const double* B = kLogData + 7;
  const double* T = kLogData + 18;
  const uint64_t ix = as_u64(x);
  const uint32_t top = static_cast<uint32_t>(ix >> 48);
  // LO = asuint64(1.0 - 0x1p-4), HI = asuint64(1.0 + 0x1.09p-4)
  if (ix - 0x3fee000000000000ull < 0x3090000000000ull) {
    if (ix == 0x3ff0000000000000ull) return 0.0;
    const double r = sub(x, 1.0);
    const double r2 = mul(r, r);
    const double r3 = mul(r, r2);
    const double p1 = fma_(r2, B[3], fma_(r, B[2], B[1]));
    const double p4 = fma_(r2, B[6], fma_(r, B[5], B[4]));
    double p7 = fma_(r2, B[9], fma_(r, B[8], B[7]));
    p7 = fma_(r3, B[10], p7);
    const double poly = fma_(fma_(p7, r3, p4), r3, p1);
    // hi + lo = r - r^2/2 computed exactly enough with a 27-bit split of r
    const double rhi = fnma(0x1p27, r, fma_(r, 0x1p27, r));
    const double rlo = sub(r, rhi);
    const double rhi2 = mul(rhi, rhi);
    const double hi = fma_(rhi2, B[0], r);
    double lo = fma_(rhi2, B[0], sub(r, hi));
    lo = fma_(mul(B[0], rlo), add(r, rhi), lo);
    const double y = fma_(poly, r3, lo);
    return add(hi, y);
  }
  if (top - 0x0010u >= 0x7ff0u - 0x0010u) return ::log(x);  // zero, subnormal, negative, inf, nan
  const uint64_t tmp = ix - 0x3fe6000000000000ull;  // OFF
  const int i = static_cast<int>((tmp >> 45) & 127);
  const int k = static_cast<int>(static_cast<int64_t>(tmp) >> 52);
  const uint64_t iz = ix - (tmp & 0xfff0000000000000ull);
  const double invc = T[2 * i], logc = T[2 * i + 1];
  const double z = as_f64(iz);
  const double r = fma_(z, invc, -1.0);
  const double kd = static_cast<double>(k);
  const double w = fma_(kd, kLogData[0], logc);
  const double hi = add(w, r);
  const double lo = fma_(kd, kLogData[1], add(sub(w, hi), r));
  const double r2 = mul(r, r);
  const double p = fma_(fma_(r, A[4], A[3]), r2, fma_(r, A[2], A[1]));
  const double y = fma_(mul(r, r2), p, fma_(r2, A[0], lo));
  return add(y, hi);
}

// -------------------------------------------------------------------- sin/cos
namespace detail {
constexpr double kBig = 0x1.8p45;
constexpr double kSn3 = -0x1.5555555555515p-3, kSn5 = 0x1.11110e829872fp-7;
constexpr double kCs2 = 0.5, kCs4 = -0x1.5555555555535p-5, kCs6 = 0x1.6c16bedd9e239p-10;
constexpr double kS1 = -0x1.5555555555555p-3, kS2 = 0x1.1111111110ecep-7, kS3 = -0x1.a01a019db08b8p-13,
                 kS4 = 0x1.71de27b9a7ed9p-19, kS5 = -0x1.addffc2fcdf59p-26;
constexpr double kHp0 = 0x1.921fb54442d18p+0, kHp1 = 0x1.1a62633145c07p-54;
constexpr double kMp1 = 0x1.921fb58000000p+0, kMp2 = -0x1.dde973c000000p-27;
constexpr double kPp3 = -0x1.cb3b398000000p-55, kPp4 = -0x1.d747f23e32ed7p-83;
constexpr double kHpInv = 0x1.45f306dc9c883p-1, kToInt = 0x1.8p52;
constexpr double kTaylorLimit = 0x1.020c49ba5e354p-3;  // 0.126

struct TableEntry {
  double sn, ssn, cs, ccs;
};

// u = big + |x|: the low word of u indexes the table, |x| - (u - big) is the remainder
MMC_MATH_FN TableEntry lookup(double ax, double& remainder) {
  const double u = add(ax, kBig);
  remainder = sub(ax, sub(u, kBig));
  const int k = static_cast<int>(static_cast<uint32_t>(as_u64(u))) * 4;
  TableEntry e;
  e.sn = kSinCosTab[k];
  e.ssn = kSinCosTab[k + 1];
  e.cs = kSinCosTab[k + 2];
  e.ccs = kSinCosTab[k + 3];
  return e;
}

// TAYLOR_SIN(x*x, x, dx)
MMC_MATH_FN double taylor_sin(double x, double dx) {
  const double xx = mul(x, x);
  const double p = fma_(fma_(fma_(fma_(kS5, xx, kS4), xx, kS3), xx, kS2), xx, kS1);
  const double t = fma_(xx, fma_(p, x, -mul(0.5, dx)), dx);
  return add(x, t);
}

// do_sin(x, dx) for |x| >= 0.126, given the table entry of |x| and r = |x| - (u - big)
MMC_MATH_FN double do_sin_core(double xsign, double r, double dx, const TableEntry& e) {
  const double xx = mul(r, r);
  const double t = fma_(mul(xx, r), fma_(xx, kSn5, kSn3), dx);
  const double s = add(t, r);
  const double q = fma_(fma_(xx, kCs6, kCs4), xx, kCs2);
  const double c = fma_(dx, r, mul(xx, q));
  const double cor = fma_(s, e.cs, fnma(c, e.sn, fma_(s, e.ccs, e.ssn)));
  return copysign_(add(cor, e.sn), xsign);
}

// do_cos(x, dx) given the table entry of |x| and r = |x| - (u - big); dx already sign-adjusted
MMC_MATH_FN double do_cos_core(double r, double dx, const TableEntry& e) {
  const double x = add(r, dx);
  const double xx = mul(x, x);
  const double s = fma_(mul(x, xx), fma_(xx, kSn5, kSn3), x);
  const double q = fma_(fma_(xx, kCs6, kCs4), xx, kCs2);
  const double c = mul(xx, q);
  const double cor = fnma(s, e.sn, fnma(c, e.cs, fnma(s, e.ssn, e.ccs)));
  return add(cor, e.cs);
}

MMC_MATH_FN double do_sin(double x, double dx) {
  const double ax = fabs_(x);
  if (ax < kTaylorLimit) return taylor_sin(x, dx);
  double r;
  const TableEntry e = lookup(ax, r);
  return do_sin_core(x, r, x <= 0 ? -dx : dx, e);
}

MMC_MATH_FN double do_cos(double x, double dx) {
  double r;
  const TableEntry e = lookup(fabs_(x), r);
  return do_cos_core(r, x < 0 ? -dx : dx, e);
}

// reduce_sincos: x = n*pi/2 + (a + da), |x| < 105414350
MMC_MATH_FN int reduce(double x, double& a, double& da) {
  const double t = fma_(x, kHpInv, kToInt);
  const double xn = sub(t, kToInt);
  const int n = static_cast<int>(as_u64(t) & 3);
  const double y = fnma(xn, kMp2, fnma(xn, kMp1, x));
  const double t2 = fnma(xn, kPp3, y);
  double db = fnma(kPp3, xn, sub(y, t2));
  const double b = fnma(xn, kPp4, t2);
  db = add(db, fnma(xn, kPp4, sub(t2, b)));
  a = b;
  da = db;
  return n;
}

MMC_MATH_FN double do_sincos(double a, double da, int n) {
  const double r = (n & 1) ? do_cos(a, da) : do_sin(a, da);
  return (n & 2) ? -r : r;
}
}  // namespace detail

MMC_MATH_FN double sin(double x);
MMC_MATH_FN double cos(double x);

// s_sincos.c: __sincos.  glibc picks one of three range reductions and then evaluates, in every one of them,
// exactly one do_sin and one do_cos on the same reduced argument (a, da) -- only which of the two becomes the sine
// and the signs differ.  Written that way here (reduce; evaluate both; select) so that the lanes of a warp, whose
// arguments fall into different ranges and quadrants, execute ONE pass over the expensive table + polynomial code
// instead of one pass per branch.  The values are glibc's bit for bit (tests/test_glibc_math.py).
MMC_MATH_FN void sincos(double x, double* sinx, double* cosx) {
  using namespace detail;
  const uint32_t k = static_cast<uint32_t>(as_u64(x) >> 32) & 0x7fffffffu;
  if (k >= 0x419921FBu) {  // |x| >= 105414350, inf, nan
    ::sincos(x, sinx, cosx);
    return;
  }
  if (k < 0x3e400000u) {  // |x| < 2^-27
    *sinx = x;
    *cosx = 1.0;
    return;
  }
  double a, da;
  int n = 0, range;
  if (k < 0x3feb6000u) {  // |x| < 0.855469: __sin_local, __cos_local on (x, 0)
    a = x;
    da = 0.0;
    range = 0;
  } else if (k < 0x400368fdu) {  // |x| < 2.426265: pi/2 - |x| in two pieces
    const double y = sub(kHp0, fabs_(x));
    a = add(y, kHp1);
    da = add(sub(y, a), kHp1);
    range = 1;
  } else {  // |x| < 105414350: reduce_sincos
    n = reduce(x, a, da);
    range = 2;
  }
  const double aa = fabs_(a);
  double r;
  const TableEntry e = lookup(aa, r);
  const double sin_like = aa < kTaylorLimit ? taylor_sin(a, da) : do_sin_core(a, r, a <= 0 ? -da : da, e);  // do_sin(a, da)
  const double cos_like = do_cos_core(r, a < 0 ? -da : da, e);                                               // do_cos(a, da)
  if (range == 0) {
    *sinx = sin_like;
    *cosx = cos_like;
  } else if (range == 1) {
    *sinx = copysign_(cos_like, x);
    *cosx = sin_like;
  } else {
    const double sn = (n & 1) ? cos_like : sin_like;        // do_sincos(a, da, n)
    const double cn = ((n + 1) & 1) ? cos_like : sin_like;  // do_sincos(a, da, n + 1)
    *sinx = (n & 2) ? -sn : sn;
    *cosx = ((n + 1) & 2) ? -cn : cn;
  }
}

// __sin(x) and __cos(x) as two separate calls return (the reference's Direction(d, mu, phi) calls std::cos and
// std::sin separately, Point.cpp:117-118), computed in one converged pass like sincos above.  NOT the same values as
// sincos: in the middle range __sin evaluates do_cos on (pi/2 - |x|, hp1) while __cos and __sincos use the
// renormalised pair (a, da).
MMC_MATH_FN void sin_and_cos(double x, double* sinx, double* cosx) {
  using namespace detail;
  const uint32_t k = static_cast<uint32_t>(as_u64(x) >> 32) & 0x7fffffffu;
  if (k >= 0x419921FBu || k < 0x3e500000u) {  // huge / non-finite, or |x| < 2^-26: the scalar functions' own paths
    *sinx = sin(x);
    *cosx = cos(x);
    return;
  }
  double as, das, ac, dac;  // arguments of the one do_sin and the one do_cos
  int n = 0, range;
  if (k < 0x3feb6000u) {  // |x| < 0.855469: sin = do_sin(x, 0), cos = do_cos(x, 0)
    as = ac = x;
    das = dac = 0.0;
    range = 0;
  } else if (k < 0x400368fdu) {  // |x| < 2.426265: sin = copysign(do_cos(y, hp1), x), cos = do_sin(a, da)
    const double y = sub(kHp0, fabs_(x));
    ac = y;
    dac = kHp1;
    as = add(y, kHp1);
    das = add(sub(y, as), kHp1);
    range = 1;
  } else {  // reduce_sincos
    n = reduce(x, as, das);
    ac = as;
    dac = das;
    range = 2;
  }
  const double sin_like = do_sin(as, das);
  const double cos_like = do_cos(ac, dac);
  if (range == 0) {
    *sinx = sin_like;
    *cosx = cos_like;
  } else if (range == 1) {
    *sinx = copysign_(cos_like, x);
    *cosx = sin_like;
  } else {
    const double sn = (n & 1) ? cos_like : sin_like;
    const double cn = ((n + 1) & 1) ? cos_like : sin_like;
    *sinx = (n & 2) ? -sn : sn;
    *cosx = ((n + 1) & 2) ? -cn : cn;
  }
}

// s_sin.c: __sin
MMC_MATH_FN double sin(double x) {
  using namespace detail;
  const uint32_t k = static_cast<uint32_t>(as_u64(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e500000u) return x;                  // |x| < 2^-26
  if (k < 0x3feb6000u) return do_sin(x, 0.0);     // |x| < 0.855469
  if (k < 0x400368fdu) return copysign_(do_cos(sub(kHp0, fabs_(x)), kHp1), x);  // |x| < 2.426265
  if (k < 0x419921FBu) {                          // |x| < 105414350
    double a, da;
    const int n = reduce(x, a, da);
    return do_sincos(a, da, n);
  }
  return ::sin(x);
}

// s_sin.c: __cos
MMC_MATH_FN double cos(double x) {
  using namespace detail;
  const uint32_t k = static_cast<uint32_t>(as_u64(x) >> 32) & 0x7fffffffu;
  if (k < 0x3e400000u) return 1.0;                // |x| < 2^-27
  if (k < 0x3feb6000u) return do_cos(x, 0.0);     // |x| < 0.855469
  if (k < 0x400368fdu) {                          // |x| < 2.426265
    const double y = sub(kHp0, fabs_(x));
    const double a = add(y, kHp1);
    const double da = add(sub(y, a), kHp1);
    return do_sin(a, da);
  }
  if (k < 0x419921FBu) {                          // |x| < 105414350
    double a, da;
    const int n = reduce(x, a, da);
    return do_sincos(a, da, n + 1);
  }
  return ::cos(x);
}

}  // namespace glibc
}  // namespace mmc
