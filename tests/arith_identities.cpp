// Host-side proof-by-exhaustion harness for the division/rounding identities the CUDA kernels
// rely on (motion_primitive_library_b200/csrc/mplx_device.cuh: div_exact, round_haz, ceil_exact).
// The functions below are the same expressions with std::fma; g++ -O2 -ffp-contract=off.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

static const double MAGIC = 6755399441055744.0;
static inline double div_exact(double a, double b, double binv) {
  double q0 = a * binv;
  double r = std::fma(-b, q0, a);
  return std::fma(r, binv, q0);
}
static inline double round_haz(double x, int &k) {
  double m = x + MAGIC;
  double kd = m - MAGIC;
  double f = x - kd;
  uint64_t bits;
  std::memcpy(&bits, &m, 8);
  k = (int)(uint32_t)bits;
  if (f == 0.5 && x > 0.0) { kd += 1.0; k += 1; }
  if (f == -0.5 && x < 0.0) { kd -= 1.0; k -= 1; }
  return kd;
}
static inline double ceil_exact(double x) {
  double kd = (x + MAGIC) - MAGIC;
  return kd < x ? kd + 1.0 : kd;
}

// sample_index's cell rule (mplx_kernels.cu): for the quotient y = RN((p-origin)/res),
//   reference: pn = (int)std::round(RN(y - 0.5)); inside <=> 0 <= pn < dim   (map_util.h:103-108,51-55)
//   kernel:    inside <=> 2^-55 < y < dim ; pn = floor(y)
static long check_cell_rule() {
  long bad = 0;
  const int dims[] = {1, 2, 5, 199, 256, 512, 799, 100000};
  auto one = [&](double y) {
    for (int dim : dims) {
      const double x = y - 0.5;
      const double rr = std::round(x);
      const bool in_ref = rr >= 0 && rr < dim;
      const bool in_k = (y > 0x1p-55) && (y < (double)dim);
      if (in_ref != in_k) { bad++; continue; }
      // device form: pn = cvt.rmi.s32.f64(y) (saturating, NaN -> 0); inside <=> y > 2^-55 && (unsigned)pn < dim
      int pn_dev;
      if (y != y) pn_dev = 0;
      else if (y >= 2147483648.0) pn_dev = 2147483647;
      else if (y < -2147483648.0) pn_dev = (-2147483647 - 1);
      else pn_dev = (int)std::floor(y);
      const bool in_dev = (y > 0x1p-55) && ((unsigned)pn_dev < (unsigned)dim);
      if (in_dev != in_ref) { bad++; continue; }
      if (in_ref && pn_dev != (int)rr) bad++;
      if (in_ref) {
        const double m = y + MAGIC, kd = m - MAGIC;
        uint64_t bits;
        std::memcpy(&bits, &m, 8);
        const int pn = (int)(uint32_t)bits - (kd > y ? 1 : 0);
        if (pn != (int)rr) bad++;
      }
    }
  };
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> U(0, 1);
  for (int J = -3; J <= 100002; J++) {
    if (J > 900 && J < 99990) continue;
    for (double off : {0.0, 0.25, 0.5, 0.75}) {
      double y = J + off;
      one(y);
      double a = y, b = y;
      for (int i = 0; i < 3; i++) {
        a = std::nextafter(a, 1e300);
        b = std::nextafter(b, -1e300);
        one(a);
        one(b);
      }
    }
    for (int i = 0; i < 20; i++) one(J + U(rng));
  }
  for (int e = -1080; e <= 60; e++)
    for (double mnt : {1.0, 1.25, 1.5, 1.9999999999999998}) {
      one(std::ldexp(mnt, e));
      one(-std::ldexp(mnt, e));
    }
  one(0.0);
  one(-0.0);
  one(0x1p-55);
  one(std::nextafter(0x1p-55, 1.0));
  one(std::nextafter(0x1p-55, 0.0));
  one(std::nan(""));
  one(INFINITY);
  one(-INFINITY);
  return bad;
}

int main(int argc, char **argv) {
  const long nrand = argc > 1 ? atol(argv[1]) : 2000000;
  const long kmax = argc > 2 ? atol(argv[2]) : 60000;
  std::mt19937_64 rng(1);
  const double bs[] = {0.1, 0.05, 0.25, 0.2, 0.01, 0.15, 0.3, 1.0 / 3, 0.07, 0.123456789, 2.5, 0.0625,
                       10.0, 0.45, 0.9999999999999999, 1.0000000000000002, 0.75, 1.9999999999999998, 3.0, 7.0};
  const double steps[] = {0.005, 0.01, 0.025, 0.05, 0.1, 0.125, 1.0 / 30, 1.0 / 900, 1.0 / 1800};
  long bad_div = 0, bad_round = 0, bad_ceil = 0, tot = 0;
  for (double b : bs) {
    const double binv = 1.0 / b;
    for (long k = -kmax; k <= kmax; k++)
      for (double step : steps) {
        const double a0 = k * step;
        for (int d = -2; d <= 2; d++) {
          double a = a0;
          for (int i = 0; i < std::abs(d); i++) a = std::nextafter(a, d > 0 ? 1e300 : -1e300);
          if (std::fabs(a) < 1e-300 && a != 0) continue;  // subnormal numerators are out of scope
          const double q = div_exact(a, b, binv), t = a / b;
          tot++;
          if (std::memcmp(&q, &t, 8) != 0 && !(q == 0 && t == 0)) bad_div++;
          // floatToInt / lattice roundings on top of the quotient
          int ki;
          const double kd = round_haz(t - 0.5, ki);
          if (kd != std::round(t - 0.5) || (std::fabs(kd) < 2e9 && ki != (int)std::round(t - 0.5))) bad_round++;
          const double kd2 = round_haz(t, ki);
          if (kd2 != std::round(t) || (std::fabs(kd2) < 2e9 && ki != (int)std::round(t))) bad_round++;
          if (ceil_exact(t) != std::ceil(t)) bad_ceil++;
        }
      }
    std::uniform_real_distribution<double> U(-1, 1);
    for (long i = 0; i < nrand; i++) {
      const int e = (int)(rng() % 80) - 40;
      const double a = std::ldexp(U(rng), e);
      const double q = div_exact(a, b, binv), t = a / b;
      tot++;
      if (std::memcmp(&q, &t, 8) != 0) bad_div++;
      if (std::fabs(t) < 1e15) {
        int ki;
        if (round_haz(t, ki) != std::round(t)) bad_round++;
        if (ceil_exact(t) != std::ceil(t)) bad_ceil++;
      }
    }
  }
  // exact half-integers and their neighbours
  for (long k = -2000000; k <= 2000000; k++) {
    for (int d = -1; d <= 1; d++) {
      double x = k + 0.5;
      if (d) x = std::nextafter(x, d > 0 ? 1e300 : -1e300);
      int ki;
      const double kd = round_haz(x, ki);
      tot++;
      if (kd != std::round(x) || ki != (int)std::round(x)) bad_round++;
      if (ceil_exact(x) != std::ceil(x)) bad_ceil++;
    }
  }
  const long bad_cell = check_cell_rule();
  printf("cases %ld bad_div %ld bad_round %ld bad_ceil %ld bad_cell %ld\n", tot, bad_div, bad_round, bad_ceil, bad_cell);
  return (bad_div || bad_round || bad_ceil || bad_cell) ? 1 : 0;
}
