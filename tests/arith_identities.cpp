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


// The fixed-point cell of mplx_fx.cu (fx_axis + fx_group) against the literal chain
// (eval_pos + sample_cell of mplx_expand.cuh == Primitive1D::p, floatToInt, isOutside):
//   fraction word >= 128  =>  floor(y_fx + eps) is the reference's cell (and the inside verdict agrees);
//   fraction word <  128  =>  the reference's cell is c' or c'-1.
// Also records the largest |y_fx - y_ref| seen (the bound DESIGN.md derives is 2^-30 below 2^18 cells).
static long check_fx_cell(long nrand, double *max_err_out) {
  const double FX_MAGIC = 1572864.0, FX_EPS = 0x1p-26, FX_RANGE = 262144.0;
  const int HI_BASE = 0x41380000;
  long bad = 0, certain = 0, uncertain = 0;
  double max_err = 0;
  std::mt19937_64 rng(99);
  std::uniform_real_distribution<double> U01(0, 1);
  auto pick = [&](std::initializer_list<double> v) { return *(v.begin() + rng() % v.size()); };
  for (long it = 0; it < nrand; it++) {
    const int ORD = 1 + (int)(rng() % 4);
    const double res = (it & 7) == 7 ? 0.01 + U01(rng) : pick({0.05, 0.1, 0.25, 0.2, 0.5, 1.0, 0.15});
    const double rinv = 1.0 / res;
    const int dim = (int)pick({64, 199, 256, 512, 799, 4096, 100000});
    const bool lattice = (it % 3) != 0;
    const double origin = lattice ? -res * (double)(rng() % (dim + 1)) : -res * dim * U01(rng);
    const double T = pick({1.0, 0.5, 2.0, 0.25, 1.5});
    // state: c5 position, c4 velocity, c3 acceleration, c2 jerk, c1 snap (the last ORD+1 used)
    double c[5];  // c1..c5 -> c[0..4]
    if (lattice) {
      const double cellc = (double)((long)(rng() % (dim + 8)) - 4);
      c[4] = (cellc + pick({0.5, 0.0, 0.25})) * res + origin + pick({0.0, 0.0, 0.5, -0.5, 0.125}) * (double)(rng() % 5);
      c[3] = pick({0.0, 1.0, -1.0, 2.0, -2.0, 3.0, -3.0, 0.5, -0.5, 1.5});
      c[2] = pick({0.0, 1.0, -1.0, 0.5, -0.5, 2.0, -2.0});
      c[1] = pick({0.0, 1.0, -1.0, 2.0, -2.0, 0.5});
      c[0] = pick({0.0, 1.0, -1.0, 4.0, -4.0});
    } else {
      c[4] = origin + res * dim * (U01(rng) * 1.2 - 0.1);
      c[3] = (U01(rng) * 2 - 1) * 6;
      c[2] = (U01(rng) * 2 - 1) * 4;
      c[1] = (U01(rng) * 2 - 1) * 4;
      c[0] = (U01(rng) * 2 - 1) * 8;
    }
    // Axis<ORD>::build: only the last ORD+1 coefficients are non-zero
    if (ORD < 4) c[0] = 0;
    if (ORD < 3) c[1] = 0;
    if (ORD < 2) c[2] = 0;
    // quotients (fill_coef): c1/24 c2/6 c3/2 c4 c5
    const double q4 = c[0] / 24, q3 = c[1] / 6, q2 = c[2] / 2, q1 = c[3], q0 = c[4];
    // fx_axis
    double C[5] = {0, q1 * rinv, q2 * rinv, q3 * rinv, q4 * rinv};
    double bound = (std::fabs(q0) + std::fabs(origin)) * rinv, tp = 1;
    for (int i = 1; i <= ORD; i++) {
      tp *= T;
      bound += std::fabs(C[i]) * tp;
    }
    C[0] = (q0 - origin) * rinv + (FX_MAGIC + FX_EPS);
    if (!(bound < FX_RANGE)) continue;
    const int n = 5 + (int)(rng() % 124);
    const double dt = T / n;
    for (double t = 0; t < T; t += dt) {
      // literal chain: eval_pos (left-to-right, power by repeated multiply) + sample_cell
      const double pw3 = (t * t) * t, pw4 = pw3 * t;
      double p;
      if (ORD == 1) p = q1 * t + q0;
      else if (ORD == 2) p = q2 * t * t + q1 * t + q0;
      else if (ORD == 3) p = q3 * pw3 + q2 * t * t + q1 * t + q0;
      else p = q4 * pw4 + q3 * pw3 + q2 * t * t + q1 * t + q0;
      const double y = (p - origin) / res;
      const bool in_ref = (y > 0x1p-55) && (y < (double)dim);
      const double cell_ref = std::floor(y);
      // fx_group
      double h = C[ORD];
      for (int i = ORD - 1; i >= 1; i--) h = std::fma(h, t, C[i]);
      const double m = std::fma(h, t, C[0]);
      uint64_t bits;
      std::memcpy(&bits, &m, 8);
      const int cell = (int)(uint32_t)(bits >> 32) - HI_BASE;
      const uint32_t fr = (uint32_t)bits;
      const double yfx = (m - FX_MAGIC) - FX_EPS;  // exact: m is a multiple of 2^-32 below 2^21
      if (std::fabs(y) < 1e6) max_err = std::max(max_err, std::fabs(yfx - y));
      if ((double)cell != std::floor(yfx + FX_EPS)) bad++;  // the word extraction itself
      if (fr >= 128u) {
        certain++;
        const bool in_fx = (unsigned)cell < (unsigned)dim;
        if (in_fx != in_ref) bad++;
        if (in_ref && (double)cell != cell_ref) bad++;
      } else {
        uncertain++;
        if (!(cell_ref == (double)cell || cell_ref == (double)cell - 1)) bad++;
        // an uncertain sample that the reference sees inside has both candidates' summary defined:
        // cell == dim is possible only with cell_ref == dim-1 (handled as "not inside" -> ambiguous)
      }
    }
  }
  *max_err_out = max_err;
  printf("fx: certain %ld uncertain %ld max|y_fx-y_ref| %.3g (2^-30 = %.3g)\n", certain, uncertain, max_err, 0x1p-30);
  if (!(max_err < 0x1p-30)) bad++;
  if (uncertain == 0) bad++;
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
  double fx_err = 0;
  const long bad_fx = check_fx_cell(nrand / 4, &fx_err);
  printf("cases %ld bad_div %ld bad_round %ld bad_ceil %ld bad_cell %ld bad_fx %ld\n", tot, bad_div, bad_round, bad_ceil,
         bad_cell, bad_fx);
  return (bad_div || bad_round || bad_ceil || bad_cell || bad_fx) ? 1 : 0;
}
