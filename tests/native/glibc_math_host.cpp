// Test-only host build of minimc_b200/csrc/glibc_math.h: compares the restated
// glibc log / sincos / sin / cos with the libm of this box over argument sweeps.
// Built by tests/test_glibc_math.py (g++ -mfma -ffp-contract=off); exports a C
// interface for ctypes and doubles as a command-line fuzzer.
#include "../../minimc_b200/csrc/glibc_math.h"

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

static uint64_t bits(double x) {
  uint64_t u;
  std::memcpy(&u, &x, 8);
  return u;
}

extern "C" {
void mmc_host_log(const double* x, double* y, size_t n) {
  for (size_t i = 0; i < n; i++) y[i] = mmc::glibc::log(x[i]);
}
void mmc_host_sincos(const double* x, double* s, double* c, size_t n) {
  for (size_t i = 0; i < n; i++) mmc::glibc::sincos(x[i], &s[i], &c[i]);
}
#ifdef MMC_HAVE_SIN_COS
void mmc_host_sin(const double* x, double* y, size_t n) {
  for (size_t i = 0; i < n; i++) y[i] = mmc::glibc::sin(x[i]);
}
void mmc_host_cos(const double* x, double* y, size_t n) {
  for (size_t i = 0; i < n; i++) y[i] = mmc::glibc::cos(x[i]);
}
#endif
// Sweeps n arguments from a splitmix64 stream over [lo, hi) (uniform in value
// when mode 0, uniform in bit pattern between lo and hi when mode 1) and
// returns the number of arguments where the restatement and libm differ.
// what: 0 log, 1 sincos, 2 sin, 3 cos, 4 sin_and_cos (vs separate sin, cos)
uint64_t mmc_host_fuzz(int what, double lo, double hi, int mode, uint64_t n, uint64_t seed, double* first_bad) {
  uint64_t bad = 0, s = seed;
  const uint64_t blo = bits(lo), bhi = bits(hi);
  for (uint64_t i = 0; i < n; i++) {
    s += 0x9e3779b97f4a7c15ull;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    z ^= z >> 31;
    double x;
    if (mode == 0) {
      x = lo + (hi - lo) * ((z >> 11) * 0x1p-53);
    } else {
      const uint64_t b = blo + z % (bhi - blo);
      std::memcpy(&x, &b, 8);
    }
    bool ok = true;
    if (what == 0) {
      ok = bits(mmc::glibc::log(x)) == bits(::log(x));
    } else if (what == 1) {
      double s1, c1, s2, c2;
      mmc::glibc::sincos(x, &s1, &c1);
      ::sincos(x, &s2, &c2);
      ok = bits(s1) == bits(s2) && bits(c1) == bits(c2);
    }
#ifdef MMC_HAVE_SIN_COS
    else if (what == 2) {
      ok = bits(mmc::glibc::sin(x)) == bits(::sin(x));
    } else if (what == 3) {
      ok = bits(mmc::glibc::cos(x)) == bits(::cos(x));
    }
#endif
    if (!ok) {
      if (bad == 0 && first_bad) *first_bad = x;
      bad++;
    }
  }
  return bad;
}
}

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 10000000ull;
  struct Case { const char* name; int what; double lo, hi; int mode; };
  const Case cases[] = {
      {"log (0,1) value", 0, 1e-300, 1.0, 0},     {"log near 1", 0, 0.93, 1.07, 0},
      {"log bits [2^-70,2^70)", 0, 0x1p-70, 0x1p70, 1}, {"log [1,1e6)", 0, 1.0, 1e6, 0},
      {"sincos [0,2pi)", 1, 0.0, 6.283185307179586, 0}, {"sincos [-7,7)", 1, -7.0, 7.0, 0},
      {"sincos bits [2^-30,0.9)", 1, 0x1p-30, 0.9, 1},   {"sincos [0,1e5)", 1, 0.0, 1e5, 0},
      {"sincos [0.8,2.5)", 1, 0.8, 2.5, 0},
#ifdef MMC_HAVE_SIN_COS
      {"sin [0,2pi)", 2, 0.0, 6.283185307179586, 0},    {"sin [-7,7)", 2, -7.0, 7.0, 0},
      {"sin bits [2^-30,0.9)", 2, 0x1p-30, 0.9, 1},      {"sin [0,1e5)", 2, 0.0, 1e5, 0},
      {"cos [0,2pi)", 3, 0.0, 6.283185307179586, 0},    {"cos [-7,7)", 3, -7.0, 7.0, 0},
      {"cos bits [2^-30,0.9)", 3, 0x1p-30, 0.9, 1},      {"cos [0,1e5)", 3, 0.0, 1e5, 0},
      {"sin_and_cos [0,2pi)", 4, 0.0, 6.283185307179586, 0}, {"sin_and_cos [-7,7)", 4, -7.0, 7.0, 0},
      {"sin_and_cos bits [2^-30,0.9)", 4, 0x1p-30, 0.9, 1},  {"sin_and_cos [0,1e5)", 4, 0.0, 1e5, 0},
      {"sin_and_cos [0.8,2.5)", 4, 0.8, 2.5, 0},
#endif
  };
  int rc = 0;
  for (const Case& c : cases) {
    double first = 0;
    const uint64_t bad = mmc_host_fuzz(c.what, c.lo, c.hi, c.mode, n, 12345, &first);
    std::printf("%-28s n=%llu mismatches=%llu", c.name, (unsigned long long)n, (unsigned long long)bad);
    if (bad) std::printf("  first at %a", first);
    std::printf("\n");
    rc |= bad != 0;
  }
  return rc;
}
