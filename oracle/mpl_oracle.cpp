/*
 * mpl_oracle.cpp — CPU ORACLE (test infrastructure, never shipped, never on the product path).
 *
 * A scalar restatement of the reference's node-expansion path.  Every arithmetic
 * expression keeps the reference's operand order and association; build with
 *   g++ -O2 -std=c++17 -ffp-contract=off      (no -march=native, no -ffast-math)
 * which mirrors the reference build (CMakeLists.txt:5-8: -std=c++11 -Wall, RelWithDebInfo,
 * baseline x86-64 => no FMA contraction).
 *
 * Citations are path:line relative to /root/reference.
 */
#include "mpl_oracle.h"

#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

namespace {

thread_local int64_t g_samples = 0;

/* Control bits: include/mpl_basis/control.h:10-20 and the bit-field union
 * include/mpl_basis/waypoint.h:47-56 (bit0 pos, bit1 vel, bit2 acc, bit3 jrk, bit4 yaw). */
enum : int {
  C_VEL = 0b00001,
  C_ACC = 0b00011,
  C_JRK = 0b00111,
  C_SNP = 0b01111,
  C_VELxYAW = 0b10001,
  C_ACCxYAW = 0b10011,
  C_JRKxYAW = 0b10111,
  C_SNPxYAW = 0b11111
};

/* power(t, n): include/mpl_basis/math.h:197-205 */
inline double power(double t, int n) {
  double tn = 1;
  while (n > 0) {
    tn *= t;
    n--;
  }
  return tn;
}

/* normalize_angle: include/mpl_basis/math.h:15-19 */
inline double normalize_angle(double angle) {
  while (angle > M_PI) angle -= 2.0 * M_PI;
  while (angle < -M_PI) angle += 2.0 * M_PI;
  return angle;
}

/* quad(): include/mpl_basis/math.h:22-32.  Roots are appended in the reference's order. */
inline int quad(double b, double c, double d, double *out) {
  double p = c * c - 4 * b * d;
  if (p < 0) return 0;
  out[0] = (-c - sqrt(p)) / (2 * b);
  out[1] = (-c + sqrt(p)) / (2 * b);
  return 2;
}

/* cubic(): include/mpl_basis/math.h:35-66.  Not reachable from primitives built by the
 * state+control constructor (leading coefficients are literal zeros) but restated so that
 * solve() is total. */
inline int cubic(double a, double b, double c, double d, double *out) {
  double a2 = b / a;
  double a1 = c / a;
  double a0 = d / a;
  double Q = (3 * a1 - a2 * a2) / 9;
  double R = (9 * a1 * a2 - 27 * a0 - 2 * a2 * a2 * a2) / 54;
  double D = Q * Q * Q + R * R;
  if (D > 0) {
    double S = std::cbrt(R + sqrt(D));
    double T = std::cbrt(R - sqrt(D));
    out[0] = -a2 / 3 + (S + T);
    return 1;
  } else if (D == 0) {
    double S = std::cbrt(R);
    out[0] = -a2 / 3 + S + S;
    out[1] = -a2 / 3 - S;
    return 2;
  } else {
    double theta = acos(R / sqrt(-Q * Q * Q));
    out[0] = 2 * sqrt(-Q) * cos(theta / 3) - a2 / 3;
    out[1] = 2 * sqrt(-Q) * cos((theta + 2 * M_PI) / 3) - a2 / 3;
    out[2] = 2 * sqrt(-Q) * cos((theta + 4 * M_PI) / 3) - a2 / 3;
    return 3;
  }
}

/* quartic(): include/mpl_basis/math.h:69-110 (same remark as cubic). */
inline int quartic(double a, double b, double c, double d, double e, double *out) {
  double a3 = b / a;
  double a2 = c / a;
  double a1 = d / a;
  double a0 = e / a;
  double ys[3];
  cubic(1, -a2, a1 * a3 - 4 * a0, 4 * a2 * a0 - a1 * a1 - a3 * a3 * a0, ys);
  double y1 = ys[0];
  double r = a3 * a3 / 4 - a2 + y1;
  if (r < 0) return 0;
  double R = sqrt(r);
  double D, E;
  if (R != 0) {
    D = sqrt(0.75 * a3 * a3 - R * R - 2 * a2 + 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
    E = sqrt(0.75 * a3 * a3 - R * R - 2 * a2 - 0.25 * (4 * a3 * a2 - 8 * a1 - a3 * a3 * a3) / R);
  } else {
    D = sqrt(0.75 * a3 * a3 - 2 * a2 + 2 * sqrt(y1 * y1 - 4 * a0));
    E = sqrt(0.75 * a3 * a3 - 2 * a2 - 2 * sqrt(y1 * y1 - 4 * a0));
  }
  int n = 0;
  if (!std::isnan(D)) {
    out[n++] = -a3 / 4 + R / 2 + D / 2;
    out[n++] = -a3 / 4 + R / 2 - D / 2;
  }
  if (!std::isnan(E)) {
    out[n++] = -a3 / 4 - R / 2 + E / 2;
    out[n++] = -a3 / 4 - R / 2 - E / 2;
  }
  return n;
}

/* solve(a,b,c,d,e): include/mpl_basis/math.h:117-131 */
inline int solve(double a, double b, double c, double d, double e, double *out) {
  if (a != 0)
    return quartic(a, b, c, d, e, out);
  else if (b != 0)
    return cubic(b, c, d, e, out);
  else if (c != 0)
    return quad(c, d, e, out);
  else if (d != 0) {
    out[0] = -e / d;
    return 1;
  } else
    return 0;
}

/* Primitive1D: include/mpl_basis/primitive.h:21-198.  c[0] is the highest order. */
struct P1 {
  double c[6] = {0, 0, 0, 0, 0, 0};

  /* primitive.h:128-131 */
  double p(double t) const {
    return c[0] / 120 * power(t, 5) + c[1] / 24 * power(t, 4) + c[2] / 6 * power(t, 3) +
           c[3] / 2 * t * t + c[4] * t + c[5];
  }
  /* primitive.h:134-137 */
  double v(double t) const {
    return c[0] / 24 * power(t, 4) + c[1] / 6 * power(t, 3) + c[2] / 2 * t * t + c[3] * t + c[4];
  }
  /* primitive.h:140-142 */
  double a(double t) const { return c[0] / 6 * power(t, 3) + c[1] / 2 * t * t + c[2] * t + c[3]; }
  /* primitive.h:145 */
  double j(double t) const { return c[0] / 2 * t * t + c[1] * t + c[2]; }

  /* primitive.h:152-162: interior stationary points of v, with the unsorted early break */
  int extrema_v(double t, double *ts) const {
    double roots[4];
    int nr = solve(0, c[0] / 6, c[1] / 2, c[2], c[3], roots);
    int n = 0;
    for (int i = 0; i < nr; i++) {
      double it = roots[i];
      if (it > 0 && it < t)
        ts[n++] = it;
      else if (it >= t)
        break;
    }
    return n;
  }
  /* primitive.h:169-179 */
  int extrema_a(double t, double *ts) const {
    double roots[4];
    int nr = solve(0, 0, c[0] / 2, c[1], c[2], roots);
    int n = 0;
    for (int i = 0; i < nr; i++) {
      double it = roots[i];
      if (it > 0 && it < t)
        ts[n++] = it;
      else if (it >= t)
        break;
    }
    return n;
  }
  /* primitive.h:186-193 */
  int extrema_j(double t, double *ts) const {
    int n = 0;
    if (c[0] != 0) {
      double t_sol = -c[1] * 2 / c[0];
      if (t_sol > 0 && t_sol < t) ts[n++] = t_sol;
    }
    return n;
  }

  /* primitive.h:92-122 */
  double J(double t, int control) const {
    if (control == C_VEL || control == C_VELxYAW)
      return c[0] * c[0] / 5184 * power(t, 9) + c[0] * c[1] / 576 * power(t, 8) +
             (c[1] * c[1] / 252 + c[0] * c[2] / 168) * power(t, 7) +
             (c[0] * c[3] / 72 + c[1] * c[2] / 36) * power(t, 6) +
             (c[2] * c[2] / 20 + c[0] * c[4] / 60 + c[1] * c[3] / 15) * power(t, 5) +
             (c[2] * c[3] / 4 + c[1] * c[4] / 12) * power(t, 4) +
             (c[3] * c[3] / 3 + c[2] * c[4] / 3) * power(t, 3) + c[3] * c[4] * t * t +
             c[4] * c[4] * t;
    else if (control == C_ACC || control == C_ACCxYAW)
      return c[0] * c[0] / 252 * power(t, 7) + c[0] * c[1] / 36 * power(t, 6) +
             (c[1] * c[1] / 20 + c[0] * c[2] / 15) * power(t, 5) +
             (c[0] * c[3] / 12 + c[1] * c[2] / 4) * power(t, 4) +
             (c[2] * c[2] / 3 + c[1] * c[3] / 3) * power(t, 3) + c[2] * c[3] * t * t +
             c[3] * c[3] * t;
    else if (control == C_JRK || control == C_JRKxYAW)
      return c[0] * c[0] / 20 * power(t, 5) + c[0] * c[1] / 4 * power(t, 4) +
             (c[1] * c[1] + c[0] * c[2]) / 3 * power(t, 3) + c[1] * c[2] * t * t + c[2] * c[2] * t;
    else if (control == C_SNP || control == C_SNPxYAW)
      return c[0] * c[0] / 3 * power(t, 3) + c[0] * c[1] * t * t + c[1] * c[1] * t;
    else
      return 0;
  }
};

/* Waypoint<Dim> with its flags (include/mpl_basis/waypoint.h:23-58) */
struct WP {
  double pos[3] = {0, 0, 0}, vel[3] = {0, 0, 0}, acc[3] = {0, 0, 0}, jrk[3] = {0, 0, 0};
  double yaw = 0, t = 0;
  int control = 0;
  bool enable_t = false;
  bool use_pos() const { return control & 1; }
  bool use_vel() const { return control & 2; }
  bool use_acc() const { return control & 4; }
  bool use_jrk() const { return control & 8; }
  bool use_yaw() const { return control & 16; }
};

/* boost::hash_combine for a 64-bit size_t, Boost 1.56-1.80 (boost/functional/hash/hash.hpp,
 * hash_combine_impl(uint64&, uint64)); boost::hash<int> is the sign-extending cast.
 * Boost is an un-vendored, un-pinned dependency of the reference (waypoint.h:9,98..121). */
inline void hash_combine(uint64_t &h, int v) {
  uint64_t k = (uint64_t)(int64_t)v;
  const uint64_t m = 0xc6a4a7935bd1e995ULL;
  k *= m;
  k ^= k >> 47;
  k *= m;
  h ^= k;
  h *= m;
  h += 0xe6546b64ULL;
}

/* hash_value(Waypoint): include/mpl_basis/waypoint.h:93-125 */
inline uint64_t hash_value(const WP &key, int Dim, int32_t *lat, int32_t *nlat) {
  uint64_t val = 0;
  int n = 0;
  for (int i = 0; i < Dim; i++) {
    if (key.use_pos()) {
      int id = std::round(key.pos[i] / 0.01);
      hash_combine(val, id);
      if (lat) lat[n] = id;
      n++;
    }
    if (key.use_vel()) {
      int id = std::round(key.vel[i] / 0.1);
      hash_combine(val, id);
      if (lat) lat[n] = id;
      n++;
    }
    if (key.use_acc()) {
      int id = std::round(key.acc[i] / 0.1);
      hash_combine(val, id);
      if (lat) lat[n] = id;
      n++;
    }
    if (key.use_jrk()) {
      int id = std::round(key.jrk[i] / 0.1);
      hash_combine(val, id);
      if (lat) lat[n] = id;
      n++;
    }
  }
  if (key.use_yaw()) {
    int id = std::round(key.yaw / 0.1);
    hash_combine(val, id);
    if (lat) lat[n] = id;
    n++;
  }
  if (key.enable_t) {
    int id = std::round(key.t / 0.1);
    hash_combine(val, id);
    if (lat) lat[n] = id;
    n++;
  }
  if (nlat) *nlat = n;
  return val;
}

/* Primitive<Dim>: include/mpl_basis/primitive.h:205-432 */
struct Prim {
  int Dim;
  double t_;
  int control_;
  P1 prs_[3];
  P1 pr_yaw_;

  /* state + control constructor: primitive.h:220-256 with Primitive1D ctors :34-50 */
  Prim(int dim, const WP &p, const double *u, double t) : Dim(dim), t_(t), control_(p.control) {
    const int base = control_ & 15;
    for (int i = 0; i < Dim; i++) {
      double *c = prs_[i].c;
      if (base == C_SNP) {
        c[0] = 0, c[1] = u[i], c[2] = p.jrk[i], c[3] = p.acc[i], c[4] = p.vel[i], c[5] = p.pos[i];
      } else if (base == C_JRK) {
        c[0] = 0, c[1] = 0, c[2] = u[i], c[3] = p.acc[i], c[4] = p.vel[i], c[5] = p.pos[i];
      } else if (base == C_ACC) {
        c[0] = 0, c[1] = 0, c[2] = 0, c[3] = u[i], c[4] = p.vel[i], c[5] = p.pos[i];
      } else if (base == C_VEL) {
        c[0] = 0, c[1] = 0, c[2] = 0, c[3] = 0, c[4] = u[i], c[5] = p.pos[i];
      }
    }
    if (control_ & 16) { /* pr_yaw_ = Primitive1D(p.yaw, u(Dim)) : primitive.h:34,235,240,244,247 */
      double *c = pr_yaw_.c;
      c[0] = c[1] = c[2] = c[3] = 0;
      c[4] = u[Dim];
      c[5] = p.yaw;
    }
  }

  /* evaluate: primitive.h:321-331 (yaw is recomputed inside the axis loop, as written) */
  WP evaluate(double t) const {
    WP p;
    p.control = control_;
    for (int k = 0; k < Dim; k++) {
      p.pos[k] = prs_[k].p(t);
      p.vel[k] = prs_[k].v(t);
      p.acc[k] = prs_[k].a(t);
      p.jrk[k] = prs_[k].j(t);
      if (p.use_yaw()) p.yaw = normalize_angle(pr_yaw_.p(t));
    }
    return p;
  }

  /* primitive.h:353-363 */
  double max_vel(int k) const {
    double ts[4];
    int n = prs_[k].extrema_v(t_, ts);
    double max_v = std::max(std::abs(prs_[k].v(0)), std::abs(prs_[k].v(t_)));
    for (int i = 0; i < n; i++) {
      double it = ts[i];
      if (it > 0 && it < t_) {
        double v = std::abs(prs_[k].v(it));
        max_v = v > max_v ? v : max_v;
      }
    }
    return max_v;
  }
  /* primitive.h:369-379 */
  double max_acc(int k) const {
    double ts[4];
    int n = prs_[k].extrema_a(t_, ts);
    double max_a = std::max(std::abs(prs_[k].a(0)), std::abs(prs_[k].a(t_)));
    for (int i = 0; i < n; i++) {
      double it = ts[i];
      if (it > 0 && it < t_) {
        double a = std::abs(prs_[k].a(it));
        max_a = a > max_a ? a : max_a;
      }
    }
    return max_a;
  }
  /* primitive.h:384-394 */
  double max_jrk(int k) const {
    double ts[4];
    int n = prs_[k].extrema_j(t_, ts);
    double max_j = std::max(std::abs(prs_[k].j(0)), std::abs(prs_[k].j(t_)));
    for (int i = 0; i < n; i++) {
      double it = ts[i];
      if (it > 0 && it < t_) {
        double j = std::abs(prs_[k].j(it));
        max_j = j > max_j ? j : max_j;
      }
    }
    return max_j;
  }
  /* primitive.h:403-407 */
  double J(int control) const {
    double j = 0;
    for (int k = 0; k < Dim; k++) j += prs_[k].J(t_, control);
    return j;
  }
};

/* validate_xxx: primitive.h:482-496 */
inline bool validate_xxx(const Prim &pr, double max, int xxx) {
  if (max <= 0) return true;
  for (int i = 0; i < pr.Dim; i++) {
    if (xxx == C_VEL && pr.max_vel(i) > max)
      return false;
    else if (xxx == C_ACC && pr.max_acc(i) > max)
      return false;
    else if (xxx == C_JRK && pr.max_jrk(i) > max)
      return false;
  }
  return true;
}

/* 2-vector helpers standing in for the Eigen expressions on the path:
 *   v.norm()        = sqrt(v0*v0 + v1*v1)
 *   v.normalized()  = v / sqrt(squaredNorm) when squaredNorm > 0   (Eigen Dot.h)
 *   a.dot(b)        = a0*b0 + a1*b1
 * Eigen is an un-vendored, un-pinned dependency; only cost terms and the
 * yaw-FOV compare go through these. */
inline double dot2_normalized(double v0, double v1, double c, double s) {
  double z = v0 * v0 + v1 * v1;
  double n0 = v0, n1 = v1;
  if (z > 0) {
    double nn = sqrt(z);
    n0 = v0 / nn;
    n1 = v1 / nn;
  }
  return n0 * c + n1 * s;
}

/* validate_yaw: primitive.h:503-525 */
inline bool validate_yaw(const Prim &pr, double my) {
  if (my <= 0) return true;
  WP ws[2] = {pr.evaluate(0), pr.evaluate(pr.t_)};
  for (const auto &w : ws) {
    double v0 = w.vel[0], v1 = w.vel[1];
    if (v0 != 0 || v1 != 0) {
      double d = dot2_normalized(v0, v1, cos(w.yaw), sin(w.yaw));
      if (d < cos(my)) return false;
    }
  }
  return true;
}

/* validate_primitive: primitive.h:449-475 */
inline bool validate_primitive(const Prim &pr, double mv, double ma, double mj, double myaw) {
  const int c = pr.control_;
  if (c == C_ACC)
    return validate_xxx(pr, mv, C_VEL);
  else if (c == C_JRK)
    return validate_xxx(pr, mv, C_VEL) && validate_xxx(pr, ma, C_ACC);
  else if (c == C_SNP)
    return validate_xxx(pr, mv, C_VEL) && validate_xxx(pr, ma, C_ACC) && validate_xxx(pr, mj, C_JRK);
  else if (c == C_VELxYAW)
    return validate_yaw(pr, myaw);
  else if (c == C_ACCxYAW)
    return validate_yaw(pr, myaw) && validate_xxx(pr, mv, C_VEL);
  else if (c == C_JRKxYAW)
    return validate_yaw(pr, myaw) && validate_xxx(pr, mv, C_VEL) && validate_xxx(pr, ma, C_ACC);
  else if (c == C_SNPxYAW)
    return validate_yaw(pr, myaw) && validate_xxx(pr, mv, C_VEL) && validate_xxx(pr, ma, C_ACC) &&
           validate_xxx(pr, mj, C_JRK);
  else
    return true;
}

/* MapUtil lookups: include/mpl_collision/map_util.h:34-69,103-108 */
inline void floatToInt(const orc_env *e, const double *pt, int *pn) {
  for (int i = 0; i < e->dim; i++) pn[i] = std::round((pt[i] - e->origin[i]) / e->res - 0.5);
}
inline bool isOutside(const orc_env *e, const int *pn) {
  for (int i = 0; i < e->dim; i++)
    if (pn[i] < 0 || pn[i] >= e->mdim[i]) return true;
  return false;
}
inline int64_t getIndex(const orc_env *e, const int *pn) {
  /* reference uses int; 512^3 = 2^27 fits.  64-bit here only so that an out-of-map pn
   * cannot overflow before the isOutside test (the reference computes idx first:
   * env_map.h:102-104, harmless there because idx is not dereferenced when outside). */
  if (e->dim == 2) return (int64_t)pn[0] + (int64_t)e->mdim[0] * pn[1];
  return (int64_t)pn[0] + (int64_t)e->mdim[0] * pn[1] + (int64_t)e->mdim[0] * e->mdim[1] * pn[2];
}

/* traverse_primitive: include/mpl_planner/env/env_map.h:90-132 */
inline double traverse_primitive(const orc_env *e, const Prim &pr) {
  const double inf = std::numeric_limits<double>::infinity();
  double max_v = 0;
  for (int i = 0; i < e->dim; i++) {
    if (pr.max_vel(i) > max_v) max_v = pr.max_vel(i);
  }
  int n = std::max(5, (int)std::ceil(max_v * pr.t_ / e->res));
  double c = 0;
  double dt = pr.t_ / n;
  for (double t = 0; t < pr.t_; t += dt) {
    g_samples++;
    const WP pt = pr.evaluate(t);
    int pn[3] = {0, 0, 0};
    floatToInt(e, pt.pos, pn);
    if (isOutside(e, pn)) return inf;
    const int64_t idx = getIndex(e, pn);
    if (e->region && !e->region[idx]) return inf;
    if (e->potential) {
      if (e->potential[idx] < 100 && e->potential[idx] > 0) {
        /* pt.vel.norm(): sqrt(squaredNorm); Eigen's unrolled fixed-size reduction
         * associates a 3-vector sum as a0 + (a1 + a2) (Redux.h redux_novec_unroller). */
        double nrm2 = e->dim == 2 ? pt.vel[0] * pt.vel[0] + pt.vel[1] * pt.vel[1]
                                  : pt.vel[0] * pt.vel[0] + (pt.vel[1] * pt.vel[1] + pt.vel[2] * pt.vel[2]);
        c += dt * (e->potential_weight * e->potential[idx] + e->gradient_weight * sqrt(nrm2));
      } else if (e->potential[idx] >= 100)
        return inf;
    } else if (e->map[idx] == 100)
      return inf;
    if (e->wyaw > 0 && pt.use_yaw()) {
      double v0 = pt.vel[0], v1 = pt.vel[1];
      if (sqrt(v0 * v0 + v1 * v1) > 1e-5) {
        double v_value = 1 - dot2_normalized(v0, v1, cos(pt.yaw), sin(pt.yaw));
        c += e->wyaw * v_value * dt;
      }
    }
  }
  return c;
}

inline WP to_wp(const orc_waypoint *w, int control) {
  WP r;
  for (int k = 0; k < 3; k++) r.pos[k] = w->pos[k], r.vel[k] = w->vel[k], r.acc[k] = w->acc[k], r.jrk[k] = w->jrk[k];
  r.yaw = w->yaw;
  r.t = w->t;
  r.control = control;
  return r;
}
inline void from_wp(const WP &r, orc_waypoint *w) {
  for (int k = 0; k < 3; k++) w->pos[k] = r.pos[k], w->vel[k] = r.vel[k], w->acc[k] = r.acc[k], w->jrk[k] = r.jrk[k];
  w->yaw = r.yaw;
  w->t = r.t;
}

/* get_succ: include/mpl_planner/env/env_map.h:147-172 */
int get_succ(const orc_env *e, const orc_waypoint *curr_, orc_waypoint *succ, double *cost_out,
             int32_t *action, uint64_t *key, int32_t *lattice) {
  const WP curr = to_wp(curr_, e->control);
  int n_out = 0;
  for (int i = 0; i < e->nU; i++) {
    Prim pr(e->dim, curr, e->U + (size_t)i * e->udim, e->T);
    WP tn = pr.evaluate(e->T);
    /* tn == curr  <=>  hash_value(tn) == hash_value(curr): waypoint.h:133-135 */
    if (hash_value(tn, e->dim, nullptr, nullptr) == hash_value(curr, e->dim, nullptr, nullptr) ||
        !validate_primitive(pr, e->v_max, e->a_max, e->j_max, e->yaw_max))
      continue;
    tn.t = curr.t + e->T;
    bool same_pos = true; /* curr.pos == tn.pos : exact coefficient-wise compare */
    for (int k = 0; k < e->dim; k++) same_pos = same_pos && (curr.pos[k] == tn.pos[k]);
    double cost = same_pos ? 0 : traverse_primitive(e, pr);
    if (!std::isinf(cost)) {
      /* calculate_intrinsic_cost: include/mpl_planner/common/env_base.h:343-345 */
      cost += pr.J(pr.control_) + e->w * e->T;
    }
    from_wp(tn, &succ[n_out]);
    cost_out[n_out] = cost;
    action[n_out] = i;
    int32_t nl = 0;
    int32_t lat[ORC_LATTICE_MAX + 1] = {0};
    uint64_t h = hash_value(tn, e->dim, lat, &nl);
    if (key) key[n_out] = h;
    if (lattice) {
      for (int q = 0; q < ORC_LATTICE_MAX; q++) lattice[(size_t)n_out * ORC_LATTICE_MAX + q] = q < nl ? lat[q] : 0;
    }
    n_out++;
  }
  return n_out;
}

/* env_map<Dim>::is_free(const Primitive&): include/mpl_planner/env/env_map.h:60-76, with
 * Primitive::sample(N) (primitive.h:415-420): N+1 samples at i*(t_/N), no lower bound on N. */
inline bool is_free_primitive(const orc_env *e, const Prim &pr) {
  double max_v = 0;
  for (int i = 0; i < e->dim; i++) {
    if (pr.max_vel(i) > max_v) max_v = pr.max_vel(i);
  }
  int n = std::ceil(max_v * pr.t_ / e->res);
  const double dt = pr.t_ / n;
  for (int i = 0; i <= n; i++) {
    const WP pt = pr.evaluate(i * dt);
    int pn[3] = {0, 0, 0};
    floatToInt(e, pt.pos, pn);
    /* isOccupied(pn) || isOutside(pn): map_util.h:51-55,64-69 */
    if (isOutside(e, pn)) return false;
    const int64_t idx = getIndex(e, pn);
    if (e->map[idx] == 100) return false;
    if (e->region && !e->region[idx]) return false;
  }
  return true;
}

/* inner loop of MapPlanner<Dim>::getLinkedNodes: src/mpl_planner/map_planner.cpp:135-151 */
inline int64_t linked_cells(const orc_env *e, const Prim &pr, int32_t *out, int64_t room) {
  double max_v = 0;
  for (int i = 0; i < e->dim; i++) max_v = std::max(max_v, pr.max_vel(i));
  int n = 1.0 * std::ceil(max_v * pr.t_ / e->res);
  int prev_id = -1;
  int64_t k = 0;
  const double dt = pr.t_ / n;
  for (int i = 0; i <= n; i++) {
    const WP w = pr.evaluate(i * dt);
    int pn[3] = {0, 0, 0};
    floatToInt(e, w.pos, pn);
    /* getIndex in the reference's int arithmetic (map_util.h:34-41), no bounds test here */
    unsigned uid = (unsigned)pn[0] + (unsigned)e->mdim[0] * (unsigned)pn[1];
    if (e->dim == 3) uid += (unsigned)e->mdim[0] * (unsigned)e->mdim[1] * (unsigned)pn[2];
    const int id = (int)uid;
    if (id != prev_id) {
      if (out && k < room)
        for (int d = 0; d < e->dim; d++) out[k * e->dim + d] = pn[d];
      k++;
      prev_id = id;
    }
  }
  return k;
}

}  // namespace

extern "C" {

int orc_get_succ(const orc_env *env, const orc_waypoint *curr, orc_waypoint *succ, double *cost,
                 int32_t *action, uint64_t *key, int32_t *lattice) {
  g_samples = 0;
  return get_succ(env, curr, succ, cost, action, key, lattice);
}

int orc_expand_batch(const orc_env *env, const orc_waypoint *nodes, int n, orc_waypoint *succ,
                     double *cost, int32_t *action, uint64_t *key, int32_t *lattice,
                     int32_t *count, int nthreads) {
  auto work = [&](int lo, int hi) {
    for (int i = lo; i < hi; i++) {
      size_t o = (size_t)i * env->nU;
      count[i] = get_succ(env, &nodes[i], succ + o, cost + o, action + o, key ? key + o : nullptr,
                          lattice ? lattice + o * ORC_LATTICE_MAX : nullptr);
    }
  };
  if (nthreads <= 1) {
    g_samples = 0;
    work(0, n);
    return 0;
  }
  std::vector<std::thread> th;
  for (int r = 0; r < nthreads; r++) {
    int lo = (int)((int64_t)n * r / nthreads), hi = (int)((int64_t)n * (r + 1) / nthreads);
    th.emplace_back(work, lo, hi);
  }
  for (auto &t : th) t.join();
  return 0;
}

int orc_expand_batch_timed(const orc_env *env, const orc_waypoint *nodes, int n, int nthreads,
                           int64_t *total_succ, int64_t *total_samples, double *seconds) {
  if (nthreads < 1) nthreads = 1;
  std::vector<int64_t> ns(nthreads, 0), nsm(nthreads, 0);
  auto work = [&](int r, int lo, int hi) {
    std::vector<orc_waypoint> succ(env->nU);
    std::vector<double> cost(env->nU);
    std::vector<int32_t> act(env->nU);
    std::vector<uint64_t> key(env->nU);
    int64_t s = 0, sm = 0;
    for (int i = lo; i < hi; i++) {
      g_samples = 0;
      s += get_succ(env, &nodes[i], succ.data(), cost.data(), act.data(), key.data(), nullptr);
      sm += g_samples;
    }
    ns[r] = s;
    nsm[r] = sm;
  };
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads == 1) {
    work(0, 0, n);
  } else {
    std::vector<std::thread> th;
    for (int r = 0; r < nthreads; r++) {
      int lo = (int)((int64_t)n * r / nthreads), hi = (int)((int64_t)n * (r + 1) / nthreads);
      th.emplace_back(work, r, lo, hi);
    }
    for (auto &t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  int64_t a = 0, b = 0;
  for (int r = 0; r < nthreads; r++) a += ns[r], b += nsm[r];
  if (total_succ) *total_succ = a;
  if (total_samples) *total_samples = b;
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return 0;
}

uint64_t orc_hash(const orc_env *env, const orc_waypoint *w, int32_t *lattice, int32_t *n_lattice) {
  WP x = to_wp(w, env->control);
  return hash_value(x, env->dim, lattice, n_lattice);
}

int orc_sample_count(double T, int n) {
  int k = 0;
  double dt = T / n;
  for (double t = 0; t < T; t += dt) k++;
  return k;
}

double orc_max_vel(const orc_env *env, const orc_waypoint *curr, int control_idx, int axis) {
  WP c = to_wp(curr, env->control);
  Prim pr(env->dim, c, env->U + (size_t)control_idx * env->udim, env->T);
  return pr.max_vel(axis);
}

int64_t orc_last_samples(void) { return g_samples; }

int orc_edges_is_free(const orc_env *env, const orc_waypoint *parents, const int32_t *actions, int n,
                      uint8_t *out_free, double *out_cost) {
  for (int i = 0; i < n; i++) {
    const WP c = to_wp(&parents[i], env->control);
    Prim pr(env->dim, c, env->U + (size_t)actions[i] * env->udim, env->T); /* forward_action: env_base.h:228-231 */
    out_free[i] = is_free_primitive(env, pr) ? 1 : 0;
    if (out_cost) out_cost[i] = pr.J(pr.control_) + env->w * env->T; /* env_base.h:343-345 */
  }
  return 0;
}

int64_t orc_edges_cells(const orc_env *env, const orc_waypoint *parents, const int32_t *actions, int n,
                        int64_t *out_offset, int32_t *out_cells, int64_t capacity) {
  int64_t total = 0;
  for (int i = 0; i < n; i++) {
    const WP c = to_wp(&parents[i], env->control);
    Prim pr(env->dim, c, env->U + (size_t)actions[i] * env->udim, env->T);
    out_offset[i] = total;
    total += linked_cells(env, pr, out_cells ? out_cells + total * env->dim : nullptr,
                          capacity > total ? capacity - total : 0);
  }
  out_offset[n] = total;
  return total;
}
}
