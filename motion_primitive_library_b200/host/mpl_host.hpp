// mpl_host.hpp — host side of the planner above the libmplx C ABI, C++17, no Eigen/Boost.
//
// Mirrors the reference's C++ surface for the node-expansion path and its caller, with the
// reference's names and argument meaning (citations are path:line in the reference checkout):
//   MPL::MapUtil<Dim>          include/mpl_collision/map_util.h
//   Waypoint<Dim>, hash_value  include/mpl_basis/waypoint.h
//   MPL::env_base<Dim>         include/mpl_planner/common/env_base.h   (params, is_goal, get_heur, get_succ)
//   MPL::env_map_gpu<Dim>      replaces env_map<Dim> (include/mpl_planner/env/env_map.h): get_succ via libmplx
//   MPL::StateSpace / State    include/mpl_planner/common/state_space.h (A* part)
//   MPL::GraphSearch::Astar    include/mpl_planner/common/graph_search.h:39-182, recoverTraj :369-455
//   MPL::MapPlanner::plan      include/mpl_planner/common/planner_base.h:275-325, src/mpl_planner/map_planner.cpp:14-18
// The search bookkeeping (hash map, priority queue) stays on the host, as in the reference; only
// get_succ crosses the boundary.  The only addition is env_base::prefetch(): a hint that lets a
// batching env expand the likely-next open nodes in the same launch.  get_succ is a pure function
// of (node, env), so speculation cannot change which nodes A* expands or in which order.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <cmath>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <type_traits>
#include <deque>
#include <limits>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mplx.h"

typedef double decimal_t;

namespace Control {
enum Control { NONE = 0, VEL = 0x01, ACC = 0x03, JRK = 0x07, SNP = 0x0f, VELxYAW = 0x11, ACCxYAW = 0x13, JRKxYAW = 0x17, SNPxYAW = 0x1f };
}

template <int N>
struct Vecf {
  decimal_t d[N];
  Vecf() { for (int i = 0; i < N; i++) d[i] = 0; }
  decimal_t &operator()(int i) { return d[i]; }
  const decimal_t &operator()(int i) const { return d[i]; }
  Vecf operator-(const Vecf &o) const { Vecf r; for (int i = 0; i < N; i++) r.d[i] = d[i] - o.d[i]; return r; }
  Vecf operator+(const Vecf &o) const { Vecf r; for (int i = 0; i < N; i++) r.d[i] = d[i] + o.d[i]; return r; }
  Vecf operator*(decimal_t k) const { Vecf r; for (int i = 0; i < N; i++) r.d[i] = d[i] * k; return r; }
  Vecf operator/(decimal_t k) const { Vecf r; for (int i = 0; i < N; i++) r.d[i] = d[i] / k; return r; }
  decimal_t lpNormInf() const { decimal_t m = 0; for (int i = 0; i < N; i++) m = std::max(m, std::abs(d[i])); return m; }
  /// Eigen's unrolled fixed-size reduction: a0 + a1 for two elements, a0 + (a1 + a2) for three
  decimal_t dot(const Vecf &o) const {
    if (N == 2) return d[0] * o.d[0] + d[1] * o.d[1];
    return d[0] * o.d[0] + (d[1] * o.d[1] + d[N - 1] * o.d[N - 1]);
  }
  decimal_t norm() const { return std::sqrt(dot(*this)); }
};
template <int N>
struct Veci {
  int d[N];
  Veci() { for (int i = 0; i < N; i++) d[i] = 0; }
  int &operator()(int i) { return d[i]; }
  const int &operator()(int i) const { return d[i]; }
  bool operator!=(const Veci &o) const { for (int i = 0; i < N; i++) if (d[i] != o.d[i]) return true; return false; }
};
template <typename T>
using vec_E = std::vector<T>;
using VecDf = std::vector<decimal_t>;

/// Waypoint<Dim>: include/mpl_basis/waypoint.h:23-58
template <int Dim>
struct Waypoint {
  Vecf<Dim> pos, vel, acc, jrk;
  decimal_t yaw{0}, t{0};
  int control{Control::NONE};
  bool enable_t{false};
  Waypoint() {}
  explicit Waypoint(int c) : control(c) {}
  bool use_pos() const { return control & 1; }
  bool use_vel() const { return control & 2; }
  bool use_acc() const { return control & 4; }
  bool use_jrk() const { return control & 8; }
  bool use_yaw() const { return control & 16; }
};

/// boost::hash_combine, 64-bit, Boost 1.56-1.80 (see DESIGN.md §2 for the version caveat)
inline void hash_combine(std::size_t &h, int v) {
  std::uint64_t k = (std::uint64_t)(std::int64_t)v;
  const std::uint64_t m = 0xc6a4a7935bd1e995ULL;
  k *= m; k ^= k >> 47; k *= m; h ^= k; h *= m; h += 0xe6546b64ULL;
}
/// hash_value(Waypoint): include/mpl_basis/waypoint.h:93-125 (host copy: start/goal nodes only;
/// successor keys come back from the device)
template <int Dim>
std::size_t hash_value(const Waypoint<Dim> &key) {
  std::size_t val = 0;
  for (int i = 0; i < Dim; i++) {
    if (key.use_pos()) { int id = std::round(key.pos(i) / 0.01); hash_combine(val, id); }
    if (key.use_vel()) { int id = std::round(key.vel(i) / 0.1); hash_combine(val, id); }
    if (key.use_acc()) { int id = std::round(key.acc(i) / 0.1); hash_combine(val, id); }
    if (key.use_jrk()) { int id = std::round(key.jrk(i) / 0.1); hash_combine(val, id); }
  }
  if (key.use_yaw()) { int id = std::round(key.yaw / 0.1); hash_combine(val, id); }
  if (key.enable_t) { int id = std::round(key.t / 0.1); hash_combine(val, id); }
  return val;
}

/// power: include/mpl_basis/math.h:197-205
inline decimal_t power(decimal_t t, int n) {
  decimal_t tn = 1;
  while (n > 0) { tn *= t; n--; }
  return tn;
}
/// normalize_angle: include/mpl_basis/math.h:15-19
inline decimal_t normalize_angle(decimal_t angle) {
  while (angle > M_PI) angle -= 2.0 * M_PI;
  while (angle < -M_PI) angle += 2.0 * M_PI;
  return angle;
}

/// Primitive1D: include/mpl_basis/primitive.h:25-197 — the evaluators and the effort integral of one
/// axis, with the reference's operand order (these run on the host, after planning: sampling a
/// recovered trajectory is not on the expansion path).
struct Primitive1D {
  decimal_t c[6] = {0, 0, 0, 0, 0, 0};  // highest order first (primitive.h:34-50)
  /// primitive.h:128-145
  decimal_t p(decimal_t t) const {
    return c[0] / 120 * power(t, 5) + c[1] / 24 * power(t, 4) + c[2] / 6 * power(t, 3) + c[3] / 2 * t * t + c[4] * t + c[5];
  }
  decimal_t v(decimal_t t) const { return c[0] / 24 * power(t, 4) + c[1] / 6 * power(t, 3) + c[2] / 2 * t * t + c[3] * t + c[4]; }
  decimal_t a(decimal_t t) const { return c[0] / 6 * power(t, 3) + c[1] / 2 * t * t + c[2] * t + c[3]; }
  decimal_t j(decimal_t t) const { return c[0] / 2 * t * t + c[1] * t + c[2]; }
  /// primitive.h:92-122
  decimal_t J(decimal_t t, int control) const {
    const int o = control & 15;
    if (o == Control::VEL)
      return c[0] * c[0] / 5184 * power(t, 9) + c[0] * c[1] / 576 * power(t, 8) +
             (c[1] * c[1] / 252 + c[0] * c[2] / 168) * power(t, 7) + (c[0] * c[3] / 72 + c[1] * c[2] / 36) * power(t, 6) +
             (c[2] * c[2] / 20 + c[0] * c[4] / 60 + c[1] * c[3] / 15) * power(t, 5) +
             (c[2] * c[3] / 4 + c[1] * c[4] / 12) * power(t, 4) + (c[3] * c[3] / 3 + c[2] * c[4] / 3) * power(t, 3) +
             c[3] * c[4] * t * t + c[4] * c[4] * t;
    else if (o == Control::ACC)
      return c[0] * c[0] / 252 * power(t, 7) + c[0] * c[1] / 36 * power(t, 6) +
             (c[1] * c[1] / 20 + c[0] * c[2] / 15) * power(t, 5) + (c[0] * c[3] / 12 + c[1] * c[2] / 4) * power(t, 4) +
             (c[2] * c[2] / 3 + c[1] * c[3] / 3) * power(t, 3) + c[2] * c[3] * t * t + c[3] * c[3] * t;
    else if (o == Control::JRK)
      return c[0] * c[0] / 20 * power(t, 5) + c[0] * c[1] / 4 * power(t, 4) + (c[1] * c[1] + c[0] * c[2]) / 3 * power(t, 3) +
             c[1] * c[2] * t * t + c[2] * c[2] * t;
    else if (o == Control::SNP)
      return c[0] * c[0] / 3 * power(t, 3) + c[0] * c[1] * t * t + c[1] * c[1] * t;
    return 0;
  }
};

/// Primitive<Dim> built from a state and a control input: include/mpl_basis/primitive.h:220-256
template <int Dim>
class Primitive {
 public:
  Primitive() {}
  Primitive(const Waypoint<Dim> &p, const VecDf &u, decimal_t t) : t_(t), control_(p.control) {
    const int o = control_ & 15;
    for (int i = 0; i < Dim; i++) {
      decimal_t *c = prs_[i].c;
      if (o == Control::SNP) { c[1] = u[i]; c[2] = p.jrk(i); c[3] = p.acc(i); c[4] = p.vel(i); c[5] = p.pos(i); }
      else if (o == Control::JRK) { c[2] = u[i]; c[3] = p.acc(i); c[4] = p.vel(i); c[5] = p.pos(i); }
      else if (o == Control::ACC) { c[3] = u[i]; c[4] = p.vel(i); c[5] = p.pos(i); }
      else if (o == Control::VEL) { c[4] = u[i]; c[5] = p.pos(i); }
    }
    if (control_ & 16) { pr_yaw_.c[4] = u[Dim]; pr_yaw_.c[5] = p.yaw; }
  }
  /// primitive.h:321-331
  Waypoint<Dim> evaluate(decimal_t t) const {
    Waypoint<Dim> p(control_);
    for (int k = 0; k < Dim; k++) {
      p.pos(k) = prs_[k].p(t);
      p.vel(k) = prs_[k].v(t);
      p.acc(k) = prs_[k].a(t);
      p.jrk(k) = prs_[k].j(t);
      if (p.use_yaw()) p.yaw = normalize_angle(pr_yaw_.p(t));
    }
    return p;
  }
  decimal_t t() const { return t_; }
  int control() const { return control_; }
  const Primitive1D &pr(int k) const { return prs_[k]; }
  const Primitive1D &pr_yaw() const { return pr_yaw_; }
  /// primitive.h:403-410
  decimal_t J(int control) const {
    decimal_t j = 0;
    for (int k = 0; k < Dim; k++) j += prs_[k].J(t_, control);
    return j;
  }
  decimal_t Jyaw() const { return pr_yaw_.J(t_, Control::VEL); }

 private:
  decimal_t t_{0};
  int control_{Control::NONE};
  Primitive1D prs_[Dim];
  Primitive1D pr_yaw_;
};

/// Command<Dim>: include/mpl_basis/trajectory.h:19-28
template <int Dim>
struct Command {
  Vecf<Dim> pos, vel, acc, jrk;
  decimal_t yaw{0}, yaw_dot{0}, t{0};
};

/// Trajectory<Dim>: include/mpl_basis/trajectory.h:42-318 without the Lambda time scaling
/// (scale / scale_down are not provided; lambda = 1, lambda_dot = 0 in the Command evaluation).
template <int Dim>
class Trajectory {
 public:
  Trajectory() {}
  explicit Trajectory(const vec_E<Primitive<Dim>> &prs) : segs(prs) {
    taus.push_back(0);
    for (const auto &pr : prs) taus.push_back(pr.t() + taus.back());
    Ts = taus;
    total_t_ = taus.back();
  }
  /// trajectory.h:67-91
  Waypoint<Dim> evaluate(decimal_t time) const {
    decimal_t tau = time;
    if (tau < 0) tau = 0;
    if (tau > total_t_) tau = total_t_;
    for (std::size_t id = 0; id < segs.size(); id++) {
      if ((tau >= taus[id] && tau < taus[id + 1]) || id == segs.size() - 1) {
        tau -= taus[id];
        Waypoint<Dim> p(segs[id].control());
        for (int j = 0; j < Dim; j++) {
          const Primitive1D &pr = segs[id].pr(j);
          p.pos(j) = pr.p(tau);
          p.vel(j) = pr.v(tau);
          p.acc(j) = pr.a(tau);
          p.jrk(j) = pr.j(tau);
          p.yaw = normalize_angle(segs[id].pr_yaw().p(tau));
        }
        return p;
      }
    }
    return Waypoint<Dim>();
  }
  /// trajectory.h:100-137
  bool evaluate(decimal_t time, Command<Dim> &p) const {
    decimal_t tau = time;
    if (tau < 0) tau = 0;
    if (tau > total_t_) tau = total_t_;
    const decimal_t lambda = 1, lambda_dot = 0;
    for (std::size_t id = 0; id < segs.size(); id++) {
      if (tau >= taus[id] && tau <= taus[id + 1]) {
        tau -= taus[id];
        for (int j = 0; j < Dim; j++) {
          const Primitive1D &pr = segs[id].pr(j);
          p.pos(j) = pr.p(tau);
          p.vel(j) = pr.v(tau) / lambda;
          p.acc(j) = pr.a(tau) / lambda / lambda - p.vel(j) * lambda_dot / lambda / lambda / lambda;
          p.jrk(j) = pr.j(tau) / lambda / lambda - 3 / power(lambda, 3) * p.acc(j) * p.acc(j) * lambda_dot +
                     3 / power(lambda, 4) * p.vel(j) * lambda_dot * lambda_dot;
          p.yaw = normalize_angle(segs[id].pr_yaw().p(tau));
          p.yaw_dot = normalize_angle(segs[id].pr_yaw().v(tau));
          p.t = time;
        }
        return true;
      }
    }
    return false;
  }
  /// trajectory.h:230-237
  vec_E<Command<Dim>> sample(int N) const {
    vec_E<Command<Dim>> ps(N + 1);
    decimal_t dt = total_t_ / N;
    for (int i = 0; i <= N; i++) evaluate(i * dt, ps[i]);
    return ps;
  }
  /// trajectory.h:251-266
  decimal_t J(int control) const {
    decimal_t j = 0;
    for (const auto &seg : segs) j += seg.J(control);
    return j;
  }
  decimal_t Jyaw() const {
    decimal_t j = 0;
    for (const auto &seg : segs) j += seg.Jyaw();
    return j;
  }
  /// trajectory.h:269-289
  std::vector<decimal_t> getSegmentTimes() const {
    std::vector<decimal_t> dts;
    for (int i = 0; i < (int)Ts.size() - 1; i++) dts.push_back(Ts[i + 1] - Ts[i]);
    return dts;
  }
  vec_E<Waypoint<Dim>> getWaypoints() const {
    vec_E<Waypoint<Dim>> ws;
    if (segs.empty()) return ws;
    decimal_t t = 0;
    for (const auto &seg : segs) {
      ws.push_back(seg.evaluate(0));
      ws.back().t = t;
      t += seg.t();
    }
    ws.push_back(segs.back().evaluate(segs.back().t()));
    ws.back().t = t;
    return ws;
  }
  vec_E<Primitive<Dim>> getPrimitives() const { return segs; }
  decimal_t getTotalTime() const { return total_t_; }

  vec_E<Primitive<Dim>> segs;
  std::vector<decimal_t> taus, Ts;
  decimal_t total_t_{0};
};

namespace MPL {
using Tmap = std::vector<signed char>;

/// MapUtil<Dim>: include/mpl_collision/map_util.h (storage + the lookups the planner's host side uses)
template <int Dim>
class MapUtil {
 public:
  Tmap getMap() { return map_; }
  const Tmap &map() const { return map_; }
  decimal_t getRes() { return res_; }
  Veci<Dim> getDim() { return dim_; }
  Vecf<Dim> getOrigin() { return origin_d_; }
  int getIndex(const Veci<Dim> &pn) {
    return Dim == 2 ? pn(0) + dim_(0) * pn(1) : pn(0) + dim_(0) * pn(1) + dim_(0) * dim_(1) * pn(Dim - 1);
  }
  bool isFree(int idx) { return map_[idx] < val_occ && map_[idx] >= val_free; }
  bool isOccupied(int idx) { return map_[idx] == val_occ; }
  bool isOutside(const Veci<Dim> &pn) {
    for (int i = 0; i < Dim; i++) if (pn(i) < 0 || pn(i) >= dim_(i)) return true;
    return false;
  }
  bool isFree(const Veci<Dim> &pn) { return isOutside(pn) ? false : isFree(getIndex(pn)); }
  bool isOccupied(const Veci<Dim> &pn) { return isOutside(pn) ? false : isOccupied(getIndex(pn)); }
  void setMap(const Vecf<Dim> &ori, const Veci<Dim> &dim, const Tmap &map, decimal_t res) {
    map_ = map; dim_ = dim; origin_d_ = ori; res_ = res; version_++;
  }
  Veci<Dim> floatToInt(const Vecf<Dim> &pt) {
    Veci<Dim> pn;
    for (int i = 0; i < Dim; i++) pn(i) = std::round((pt(i) - origin_d_(i)) / res_ - 0.5);
    return pn;
  }
  /// Cells crossed by the segment pt1 -> pt2, the way MapUtil::rayTrace samples it (map_util.h:120-137):
  /// the segment is cut into floor(Linf(diff/res)/0.8) equal steps and the interior points
  /// pt1 + (diff/steps)*i, i = 1..steps-1, are converted with floatToInt; the walk ends at the first
  /// point outside the map, and a cell is reported once per run of consecutive equal cells.  `visit`
  /// returns false to stop early (is_goal stops at the first occupied cell).  The sample points must
  /// be these exact doubles for the goal test / tunnel to agree with the reference, hence the same
  /// three operations per point (scale, multiply by i, add).
  template <typename Visit>
  void walkRay(const Vecf<Dim> &pt1, const Vecf<Dim> &pt2, Visit visit) {
    const Vecf<Dim> span = pt2 - pt1;
    const int steps = (span / res_).lpNormInf() / 0.8;
    const Vecf<Dim> inc = span * (1.0 / steps);
    bool have_last = false;
    Veci<Dim> last;
    for (int i = 1; i < steps; i++) {
      const Veci<Dim> cell = floatToInt(pt1 + inc * i);
      if (isOutside(cell)) return;
      if (!have_last || cell != last) {
        if (!visit(cell)) return;
      }
      last = cell;
      have_last = true;
    }
  }
  vec_E<Veci<Dim>> rayTrace(const Vecf<Dim> &pt1, const Vecf<Dim> &pt2) {
    vec_E<Veci<Dim>> cells;
    walkRay(pt1, pt2, [&](const Veci<Dim> &c) { cells.push_back(c); return true; });
    return cells;
  }
  void freeUnknown() { for (auto &v : map_) if (v == val_unknown) v = val_free; version_++; }
  unsigned long version() const { return version_; }

 protected:
  decimal_t res_{1};
  Vecf<Dim> origin_d_;
  Veci<Dim> dim_;
  Tmap map_;
  unsigned long version_{0};
  int8_t val_occ = 100, val_free = 0, val_unknown = -1;
};
typedef MapUtil<2> OccMapUtil;
typedef MapUtil<3> VoxelMapUtil;

/// env_base<Dim>: include/mpl_planner/common/env_base.h (the members the A* path touches)
template <int Dim>
class env_base {
 public:
  virtual ~env_base() {}
  /// env_base.h:23-41
  virtual bool is_goal(const Waypoint<Dim> &state) const {
    if (state.t >= t_max_) return true;
    bool goaled = (state.pos - goal_node_.pos).lpNormInf() <= tol_pos_;
    if (goaled && tol_vel_ >= 0) goaled = (state.vel - goal_node_.vel).lpNormInf() <= tol_vel_;
    if (goaled && tol_acc_ >= 0) goaled = (state.acc - goal_node_.acc).lpNormInf() <= tol_acc_;
    if (goaled && tol_yaw_ >= 0) goaled = std::abs(state.yaw - goal_node_.yaw) <= tol_yaw_;
    return goaled;
  }
  /// env_base.h:48-64 (heur_ignore_dynamics_ == true, the default: :368)
  virtual decimal_t get_heur(const Waypoint<Dim> &state) const { return get_heur(state, hash_value(state)); }
  /// the same with the state's lattice key already known (the device returns it with the successor);
  /// `goal_node_ == state` is hash equality (waypoint.h:133-135), the goal's hash is cached by set_goal
  virtual decimal_t get_heur(const Waypoint<Dim> &state, std::size_t state_key) const {
    if (goal_key_ == state_key) return 0;
    return cal_heur(state, goal_node_);
  }
  /// cal_heur: env_base.h:55-64, the heur_ignore_dynamics_ branch (the reference's default, :368): the Linf
  /// distance over v_max.  The minimum-time heuristics with dynamics (env_base.h:66-211: closed-form
  /// quartic for ACC states, Eigen's PolynomialSolver for JRK) are outside the expansion path this
  /// library rebuilds and are not provided: set_heur_ignore_dynamics(false) is ignored.
  virtual decimal_t cal_heur(const Waypoint<Dim> &state, const Waypoint<Dim> &goal) const {
    if (v_max_ > 0) return w_ * (state.pos - goal.pos).lpNormInf() / v_max_;
    return w_ * (state.pos - goal.pos).lpNormInf();
  }
  /// env_base.h:305-306
  /// Only the default (true) is supported; false is reported and ignored (the reference's style: a
  /// message and no exception, planner_base.h:283-287), the search keeps the admissible Linf heuristic.
  bool set_heur_ignore_dynamics(bool ignore) {
    if (!ignore) {
      std::fprintf(stderr, "[mpl_host] setHeurIgnoreDynamics(false): the minimum-time heuristic with dynamics "
                           "(env_base.h:66-211) is not provided; keeping the default Linf heuristic\n");
      return false;
    }
    heur_ignore_dynamics_ = true;
    return true;
  }
  /// env_base.h:228-231
  void forward_action(const Waypoint<Dim> &curr, int action_id, Primitive<Dim> &pr) const {
    pr = Primitive<Dim>(curr, U_[action_id], dt_);
  }
  void set_u(const vec_E<VecDf> &U) { U_ = U; touch(); }
  void set_v_max(decimal_t v) { v_max_ = v; touch(); }
  void set_a_max(decimal_t a) { a_max_ = a; touch(); }
  void set_j_max(decimal_t j) { j_max_ = j; touch(); }
  void set_yaw_max(decimal_t yaw) { yaw_max_ = yaw; touch(); }
  void set_dt(decimal_t dt) { dt_ = dt; touch(); }
  void set_tol_pos(decimal_t pos) { tol_pos_ = pos; }
  void set_tol_vel(decimal_t vel) { tol_vel_ = vel; }
  void set_tol_acc(decimal_t acc) { tol_acc_ = acc; }
  void set_tol_yaw(decimal_t yaw) { tol_yaw_ = yaw; }
  void set_w(decimal_t w) { w_ = w; touch(); }
  void set_wyaw(decimal_t wyaw) { wyaw_ = wyaw; touch(); }
  void set_t_max(int t) { t_max_ = t; }
  bool set_goal(const Waypoint<Dim> &state) { goal_node_ = state; goal_key_ = hash_value(state); return true; }
  virtual void set_potential_weight(decimal_t) {}
  virtual void set_gradient_weight(decimal_t) {}
  virtual void set_potential_map(const std::vector<int8_t> &) {}
  virtual void set_search_region(const std::vector<bool> &r) { search_region_ = r; touch(); }
  decimal_t get_dt() const { return dt_; }
  virtual bool is_free(const Vecf<Dim> &) const { return true; }
  /// env_base.h:358-362
  virtual void get_succ(const Waypoint<Dim> &curr, vec_E<Waypoint<Dim>> &succ, std::vector<decimal_t> &succ_cost,
                        std::vector<int> &action_idx) const = 0;
  /// Extensions for batching envs (results of get_succ are pure, so none of this can change what
  /// the search expands):
  ///  - wants_candidates(key): true when the env would have to launch for this node and could take
  ///    more nodes in the same launch;
  ///  - prefetch(cands, keys): the open nodes the search is likely to pop next;
  ///  - last_succ_keys(): lattice keys of the successors returned by the last get_succ call, if the
  ///    env already has them (the device computes them); nullptr = hash on the host.
  /// Stored-edge queries of the incremental planner, batched.  Edge k is the primitive
  /// forward_action(parents[k], actions[k]) (env_base.h:228-231):
  ///  - is_free_edges: is_free(pr) (env_base.h:338-341, env_map.h:60-76) and
  ///    calculate_intrinsic_cost(pr) (env_base.h:343-345) per edge;
  ///  - edge_cells: the cells getLinkedNodes visits along each edge (map_planner.cpp:135-151),
  ///    edge k owning cells[offset[k]*Dim .. offset[k+1]*Dim); an env that can also returns the
  ///    inverted table (table_voxel sorted, table_edge the edge of each entry, edges of one voxel in
  ///    emission order), else leaves both empty and the planner sorts on the host.
  virtual void is_free_edges(const vec_E<Waypoint<Dim>> &, const std::vector<int> &, std::vector<uint8_t> &,
                             std::vector<decimal_t> &) const {
    throw std::runtime_error("this env does not serve edge re-validation");
  }
  virtual void edge_cells(const vec_E<Waypoint<Dim>> &, const std::vector<int> &, std::vector<long long> &,
                          std::vector<int> &, std::vector<int> &, std::vector<int> &) const {
    throw std::runtime_error("this env does not serve edge re-validation");
  }
  /// MapPlanner::setSearchRegion(path, dense) (src/mpl_planner/map_planner.cpp:46-95) served by the
  /// env that owns the grids: build the tunnel of half-width ceil(radius/res) cells around the path
  /// and install it as the search region.
  virtual void search_region_from_path(const vec_E<Vecf<Dim>> &, const Vecf<Dim> &, bool) {
    throw std::runtime_error("this env does not build search regions");
  }
  virtual bool wants_candidates(std::size_t) const { return false; }
  virtual void prefetch(const vec_E<Waypoint<Dim>> &, const std::vector<std::size_t> &) const {}
  virtual const std::size_t *last_succ_keys() const { return nullptr; }
  virtual void begin_plan() const {}

  bool heur_ignore_dynamics_{true};
  decimal_t w_{10.0}, wyaw_{1.0};
  decimal_t tol_pos_{0.5}, tol_vel_{-1.0}, tol_acc_{-1.0}, tol_yaw_{-1.0};
  decimal_t v_max_{-1.0}, a_max_{-1.0}, j_max_{-1.0}, yaw_max_{-1.0};
  decimal_t t_max_{std::numeric_limits<decimal_t>::infinity()};
  decimal_t dt_{1.0};
  vec_E<VecDf> U_;
  Waypoint<Dim> goal_node_;
  std::size_t goal_key_{hash_value(Waypoint<Dim>())};
  std::vector<bool> search_region_;
  mutable vec_E<Vecf<Dim>> expanded_nodes_;

 protected:
  void touch() { params_version_++; }
  unsigned long params_version_{1};
};

/// The host-only members of env_map<Dim> (include/mpl_planner/env/env_map.h:25-51): goal test with
/// ray-trace and the start-point free test.  They run on the host against the MapUtil copy.
template <int Dim>
class env_map_host : public env_base<Dim> {
 public:
  explicit env_map_host(std::shared_ptr<MapUtil<Dim>> map_util) : map_util_(map_util) {}
  /// env_map.h:25-45
  bool is_goal(const Waypoint<Dim> &state) const override {
    bool goaled = (state.pos - this->goal_node_.pos).lpNormInf() <= this->tol_pos_;
    if (goaled && this->tol_vel_ >= 0) goaled = (state.vel - this->goal_node_.vel).lpNormInf() <= this->tol_vel_;
    if (goaled && this->tol_acc_ >= 0) goaled = (state.acc - this->goal_node_.acc).lpNormInf() <= this->tol_acc_;
    if (goaled && this->tol_yaw_ >= 0) goaled = std::abs(state.yaw - this->goal_node_.yaw) <= this->tol_yaw_;
    if (goaled) {
      // env_map.h:31-35: the straight line to the goal must not cross an occupied cell
      map_util_->walkRay(state.pos, this->goal_node_.pos, [&](const Veci<Dim> &c) {
        if (map_util_->isOccupied(c)) goaled = false;
        return goaled;
      });
    }
    return goaled;
  }
  /// env_map.h:48-51
  bool is_free(const Vecf<Dim> &pt) const override { return map_util_->isFree(map_util_->floatToInt(pt)); }

 protected:
  std::shared_ptr<MapUtil<Dim>> map_util_;
};

/// env_map_gpu<Dim>: env_map<Dim> (include/mpl_planner/env/env_map.h) with get_succ served by
/// libmplx (CUDA).  Throws std::runtime_error when the engine cannot be created: no CPU fallback.
template <int Dim>
class env_map_gpu : public env_map_host<Dim> {
  using env_map_host<Dim>::map_util_;

 public:
  explicit env_map_gpu(std::shared_ptr<MapUtil<Dim>> map_util, int device = 0) : env_map_host<Dim>(map_util) {
    if (mplx_create(Dim, device, &ctx_) != MPLX_OK) throw std::runtime_error(mplx_last_error());
  }
  ~env_map_gpu() override { mplx_destroy(ctx_); }
  env_map_gpu(const env_map_gpu &) = delete;

  void set_potential_map(const std::vector<int8_t> &map) override { potential_map_ = map; potential_on_device_ = false; this->touch(); }
  void set_search_region(const std::vector<bool> &r) override { region_on_device_ = false; env_base<Dim>::set_search_region(r); }
  void set_potential_weight(decimal_t w) override { potential_weight_ = w; this->touch(); }
  void set_gradient_weight(decimal_t w) override { gradient_weight_ = w; this->touch(); }
  void set_control(int control) { control_ = control; this->touch(); }
  /// nodes speculatively expanded per launch (1 = plain one-node get_succ)
  void set_speculation(int k) { speculate_ = std::max(1, k); }
  /// record the node of every get_succ call, in call order = the A* pop order (graph_search.h:66-75);
  /// the replay frontier of the benchmark (SURVEY.md §8d i)
  void set_trace(std::vector<mplx_waypoint> *trace) { trace_ = trace; }

  void begin_plan() const override { cache_.clear(); stats_nodes_ = stats_calls_ = stats_hits_ = 0; }

  /// env_map.h:147-172.  Served from the speculation cache when possible.
  void get_succ(const Waypoint<Dim> &curr, vec_E<Waypoint<Dim>> &succ, std::vector<decimal_t> &succ_cost,
                std::vector<int> &action_idx) const override {
    succ.clear(); succ_cost.clear(); action_idx.clear();
    this->expanded_nodes_.push_back(curr.pos);  // env_map.h:154, at real pop time
    if (trace_) trace_->push_back(to_pod(curr));
    const std::size_t key = hash_value(curr);
    auto it = cache_.find(key);
    if (it == cache_.end()) {
      batch_.clear(); batch_keys_.clear();
      batch_.push_back(curr); batch_keys_.push_back(key);
      for (std::size_t c = 0; c < pending_.size() && (int)batch_.size() < speculate_; c++) {
        const std::size_t k = pending_keys_[c];
        if (k != key && !cache_.count(k)) { batch_.push_back(pending_[c]); batch_keys_.push_back(k); }
      }
      pending_.clear(); pending_keys_.clear();
      expand_batch();
      it = cache_.find(key);
    } else {
      stats_hits_++;
    }
    Entry &e = it->second;
    for (std::size_t j = 0; j < e.cost.size(); j++) succ.push_back(from_pod(e.succ[j], curr.control));
    succ_cost.assign(e.cost.begin(), e.cost.end());
    action_idx.assign(e.action.begin(), e.action.end());
    last_keys_.swap(e.key);
    cache_.erase(it);  // A* expands a node once
  }
  bool wants_candidates(std::size_t key) const override { return speculate_ > 1 && !cache_.count(key); }
  void prefetch(const vec_E<Waypoint<Dim>> &cands, const std::vector<std::size_t> &keys) const override {
    pending_ = cands;
    pending_keys_ = keys;
  }
  const std::size_t *last_succ_keys() const override { return last_keys_.data(); }

  /// MapPlanner::updatePotentialMap on the device (mplx_update_potential_map): the MapUtil grid is
  /// replaced by the potential field, which also becomes the potential map (map_planner.cpp:387-388).
  void update_potential_map(const Vecf<Dim> &radius, decimal_t pow_, const Vecf<Dim> &range, const Vecf<Dim> &pos) {
    sync();
    Tmap out(map_util_->map().size());
    check(mplx_update_potential_map(ctx_, radius.d, pow_, range.d, pos.d, potential_weight_, gradient_weight_,
                                    (int8_t *)out.data()));
    map_util_->setMap(map_util_->getOrigin(), map_util_->getDim(), out, map_util_->getRes());
    map_version_ = map_util_->version();  // the device already holds this grid
    potential_map_.assign(out.begin(), out.end());
    potential_on_device_ = true;
  }
  /// MapPlanner::setSearchRegion on the device (mplx_set_search_region_path)
  void search_region_from_path(const vec_E<Vecf<Dim>> &path, const Vecf<Dim> &radius, bool dense) override {
    set_search_region_path(path, radius, dense);
  }
  void set_search_region_path(const vec_E<Vecf<Dim>> &path, const Vecf<Dim> &radius, bool dense) {
    sync();
    std::vector<double> flat;
    for (const auto &p : path) for (int k = 0; k < Dim; k++) flat.push_back(p(k));
    std::vector<uint8_t> out(map_util_->map().size());
    check(mplx_set_search_region_path(ctx_, flat.data(), (int)path.size(), radius.d, dense ? 1 : 0, out.data()));
    this->search_region_.assign(out.begin(), out.end());
    region_on_device_ = true;
  }

  /// env_map::is_free(pr) for stored edges on the device (mplx_edges_is_free)
  void is_free_edges(const vec_E<Waypoint<Dim>> &parents, const std::vector<int> &actions, std::vector<uint8_t> &free,
                     std::vector<decimal_t> &cost) const override {
    sync();
    std::vector<mplx_waypoint> in(parents.size());
    for (std::size_t i = 0; i < parents.size(); i++) in[i] = to_pod(parents[i]);
    free.assign(parents.size(), 0);
    cost.assign(parents.size(), 0);
    check(mplx_edges_is_free(ctx_, in.data(), actions.data(), (int)parents.size(), free.data(), cost.data()));
  }
  /// the getLinkedNodes voxel walk for stored edges on the device (mplx_edges_cells)
  void edge_cells(const vec_E<Waypoint<Dim>> &parents, const std::vector<int> &actions, std::vector<long long> &offset,
                  std::vector<int> &cells, std::vector<int> &table_voxel, std::vector<int> &table_edge) const override {
    sync();
    std::vector<mplx_waypoint> in(parents.size());
    for (std::size_t i = 0; i < parents.size(); i++) in[i] = to_pod(parents[i]);
    offset.assign(parents.size() + 1, 0);
    int64_t total = 0;
    std::size_t cap = std::max<std::size_t>(1, 32 * parents.size());
    for (int attempt = 0;; attempt++) {
      cells.resize(cap * Dim); table_voxel.resize(cap); table_edge.resize(cap);
      const int rc = mplx_edges_cells(ctx_, in.data(), actions.data(), (int)parents.size(), (int64_t *)offset.data(),
                                      cells.data(), (int64_t)cap, &total, table_voxel.data(), table_edge.data());
      if (rc == MPLX_OK) break;
      if (attempt > 0 || total <= (int64_t)cap) check(rc);
      cap = (std::size_t)total;  // sized from the reported total
    }
    cells.resize((std::size_t)total * Dim); table_voxel.resize((std::size_t)total); table_edge.resize((std::size_t)total);
  }

  /// Packed batched expansion for lock-step drivers (mplx_expand_packed, +inf successors dropped
  /// on the device: A* skips them, graph_search.h:81).  Results stay in the env's buffers until the
  /// next call: record r of node i is r in [p_offset[i], p_offset[i] + p_count[i]).
  ///
  /// keys_only: per record only {key, action} cross PCIe (10 B instead of 66 B for 3-D ACC).  Possible
  /// for occupancy planning (keys_only_possible()): the finite edge cost is calculate_intrinsic_cost,
  /// a function of the action alone (action_cost()), and the coordinates of a successor are needed only
  /// when its key is new to the search (graph_search.h:84-88), where forward_from_pod() evaluates them
  /// on the host with the reference's own operand order (the same bits the device produces).
  void expand_packed(const std::vector<mplx_waypoint> &nodes, bool keys_only = false) const {
    sync();
    const int n = (int)nodes.size(), nU = (int)this->U_.size();
    const std::size_t cap = (std::size_t)n * nU;
    const int nstate = Dim * __builtin_popcount(control_ & 15) + ((control_ & 16) ? 1 : 0);
    p_count.reserve(n); p_offset.reserve(n); p_action.reserve(cap); p_key.reserve(cap);
    if (!keys_only) { p_state.reserve(cap * nstate); p_cost.reserve(cap); }
    mplx_packed_out out{p_count.data(), (int64_t *)p_offset.data(), keys_only ? nullptr : p_state.data(),
                        keys_only ? nullptr : p_cost.data(), p_action.data(), p_key.data(), (int64_t)cap, 0, 0};
    check(mplx_expand_packed(ctx_, nodes.data(), n, MPLX_PACK_DROP_INF, &out));
    p_nstate = out.nstate;
    stats_nodes_ += n;
    stats_calls_++;
  }
  /// No potential field and no yaw control: traverse_primitive contributes 0 to every finite cost.
  bool keys_only_possible() const { return potential_map_.empty() && !(control_ & 16); }
  /// calculate_intrinsic_cost (env_base.h:343-345) of Primitive(., U[action], dt): for a primitive built
  /// from a state and a control only the control's own term of Primitive1D::J is non-zero
  /// (primitive.h:92-122), so the cost does not depend on the state.
  decimal_t action_cost(int action) const {
    if (action_cost_version_ != this->params_version_ || action_cost_.size() != this->U_.size()) {
      action_cost_.resize(this->U_.size());
      const Waypoint<Dim> zero(control_);
      for (std::size_t a = 0; a < this->U_.size(); a++) {
        const Primitive<Dim> pr(zero, this->U_[a], this->dt_);
        action_cost_[a] = pr.J(control_ & 15) + this->w_ * this->dt_;
      }
      action_cost_version_ = this->params_version_;
    }
    return action_cost_[action];
  }
  /// env_map.h:156-161 on the host for one successor: tn = Primitive(curr, U[action], dt).evaluate(dt),
  /// tn.t = curr.t + dt.
  Waypoint<Dim> forward_from_pod(const mplx_waypoint &curr, int action) const {
    const Waypoint<Dim> c = from_pod(curr, control_);
    const Primitive<Dim> pr(c, this->U_[action], this->dt_);
    Waypoint<Dim> tn = pr.evaluate(this->dt_);
    tn.t = c.t + this->dt_;
    return tn;
  }
  /// Rebuild the successor Waypoint of packed record r (parent `curr`): the state fields come from
  /// the record; the rest are copies, not results (include/mplx.h, mplx_packed_out): the first
  /// derivative above the state order is 0 + U[action] (Primitive1D::v/a/j at T with literal-zero
  /// higher coefficients, primitive.h:134-145), higher ones 0, yaw 0 without a yaw control,
  /// t = curr.t + dt (env_map.h:161).
  Waypoint<Dim> packed_waypoint(std::size_t r, const mplx_waypoint &curr) const {
    Waypoint<Dim> w(control_);
    const double *st = p_state.data() + r * p_nstate;
    const int nf = __builtin_popcount(control_ & 15);
    const VecDf &u = this->U_[p_action[r]];
    int c = 0;
    for (int d = 0; d < Dim; d++) w.pos(d) = st[c++];
    for (int d = 0; d < Dim; d++) w.vel(d) = nf >= 2 ? st[c++] : 0.0 + u[d];
    for (int d = 0; d < Dim; d++) w.acc(d) = nf >= 3 ? st[c++] : (nf == 2 ? 0.0 + u[d] : 0.0);
    for (int d = 0; d < Dim; d++) w.jrk(d) = nf >= 4 ? st[c++] : (nf == 3 ? 0.0 + u[d] : 0.0);
    w.yaw = (control_ & 16) ? st[c++] : 0.0;
    w.t = curr.t + this->dt_;
    return w;
  }
  /// page-locked host array (mplx_host_alloc): the packed records cross PCIe by DMA straight into
  /// it; grows, never shrinks
  template <typename T>
  struct Pinned {
    T *p = nullptr;
    std::size_t cap = 0;
    Pinned() {}
    Pinned(const Pinned &) = delete;
    ~Pinned() { if (p) mplx_host_free(p); }
    void reserve(std::size_t n) {
      if (n <= cap) return;
      if (p) mplx_host_free(p);
      cap = n + n / 4;
      p = (T *)mplx_host_alloc(cap * sizeof(T));
      if (!p) { cap = 0; throw std::runtime_error(mplx_last_error()); }
    }
    T *data() const { return p; }
    T &operator[](std::size_t i) const { return p[i]; }
  };
  mutable Pinned<int32_t> p_count;
  mutable Pinned<long long> p_offset;
  mutable Pinned<double> p_state, p_cost;
  mutable Pinned<uint16_t> p_action;
  mutable Pinned<uint64_t> p_key;
  mutable int p_nstate = 0;
  mutable std::vector<decimal_t> action_cost_;
  mutable unsigned long action_cost_version_ = ~0ul;
  static mplx_waypoint pod(const Waypoint<Dim> &w) { return to_pod(w); }
  int control() const { return control_; }

  long stats_nodes() const { return stats_nodes_; }
  long stats_calls() const { return stats_calls_; }
  long stats_hits() const { return stats_hits_; }
  long launches() const { return (long)mplx_launch_count(ctx_); }

 private:
  struct Entry { std::vector<mplx_waypoint> succ; std::vector<double> cost; std::vector<int> action; std::vector<std::size_t> key; };
  static void check(int rc) { if (rc != MPLX_OK) throw std::runtime_error(mplx_last_error()); }
  static mplx_waypoint to_pod(const Waypoint<Dim> &w) {
    mplx_waypoint p{};
    for (int d = 0; d < Dim; d++) { p.pos[d] = w.pos(d); p.vel[d] = w.vel(d); p.acc[d] = w.acc(d); p.jrk[d] = w.jrk(d); }
    p.yaw = w.yaw; p.t = w.t;
    return p;
  }
  static Waypoint<Dim> from_pod(const mplx_waypoint &p, int control) {
    Waypoint<Dim> w(control);
    for (int d = 0; d < Dim; d++) { w.pos(d) = p.pos[d]; w.vel(d) = p.vel[d]; w.acc(d) = p.acc[d]; w.jrk(d) = p.jrk[d]; }
    w.yaw = p.yaw; w.t = p.t;
    return w;
  }
  void sync() const {
    if (map_version_ != map_util_->version()) {
      const Veci<Dim> dim = map_util_->getDim();
      const Vecf<Dim> ori = map_util_->getOrigin();
      check(mplx_set_map(ctx_, map_util_->map().data(), dim.d, ori.d, map_util_->getRes()));
      map_version_ = map_util_->version();
      sent_version_ = 0;
      // mplx_set_map drops the device's potential map and tunnel (their sizes are tied to the grid); the
      // reference keeps both across MapUtil::setMap, so they are re-sent from the host copies below
      potential_on_device_ = region_on_device_ = false;
    }
    if (sent_version_ != this->params_version_) {
      if (this->U_.empty()) throw std::runtime_error("env_map_gpu: set_u() was not called");
      const int udim = (int)this->U_.front().size();
      std::vector<double> U;
      for (const auto &u : this->U_) for (int k = 0; k < udim; k++) U.push_back(u[k]);
      check(mplx_set_params(ctx_, control_, this->dt_, this->w_, this->wyaw_, this->v_max_, this->a_max_,
                            this->j_max_, this->yaw_max_, U.data(), (int)this->U_.size(), udim));
      if (!potential_on_device_)
        check(mplx_set_potential(ctx_, potential_map_.empty() ? nullptr : potential_map_.data(), potential_weight_,
                                 gradient_weight_));
      else  // the field was built on the device: only the weights may have changed
        check(mplx_set_potential_weights(ctx_, potential_weight_, gradient_weight_));
      if (!region_on_device_) {
        if (this->search_region_.empty()) check(mplx_set_search_region(ctx_, nullptr));
        else {
          std::vector<uint8_t> r(this->search_region_.begin(), this->search_region_.end());
          check(mplx_set_search_region(ctx_, r.data()));
        }
      }
      sent_version_ = this->params_version_;
    }
  }
  void expand_batch() const {
    sync();
    const int n = (int)batch_.size(), nU = (int)this->U_.size();
    in_.resize(n);
    for (int i = 0; i < n; i++) in_[i] = to_pod(batch_[i]);
    const std::size_t slots = (std::size_t)n * nU;
    count_.resize(n); succ_.resize(slots); cost_.resize(slots); action_.resize(slots); key_.resize(slots);
    mplx_succ_out out{count_.data(), succ_.data(), cost_.data(), action_.data(), key_.data(), nullptr};
    check(mplx_expand(ctx_, in_.data(), n, &out));
    for (int i = 0; i < n; i++) {
      Entry &e = cache_[batch_keys_[i]];
      const std::size_t o = (std::size_t)i * nU;
      e.succ.assign(succ_.begin() + o, succ_.begin() + o + count_[i]);
      e.cost.assign(cost_.begin() + o, cost_.begin() + o + count_[i]);
      e.action.assign(action_.begin() + o, action_.begin() + o + count_[i]);
      e.key.assign(key_.begin() + o, key_.begin() + o + count_[i]);
    }
    stats_nodes_ += n;
    stats_calls_++;
  }

  mplx_ctx *ctx_ = nullptr;
  std::vector<mplx_waypoint> *trace_ = nullptr;
  int control_ = Control::NONE, speculate_ = 1;
  mutable bool potential_on_device_ = false, region_on_device_ = false;
  std::vector<int8_t> potential_map_;
  decimal_t potential_weight_{0.1}, gradient_weight_{0.0};
  mutable unsigned long map_version_ = ~0ul, sent_version_ = 0;
  mutable std::unordered_map<std::size_t, Entry> cache_;
  mutable vec_E<Waypoint<Dim>> pending_, batch_;
  mutable std::vector<std::size_t> pending_keys_, batch_keys_, last_keys_;
  mutable std::vector<uint64_t> key_;
  mutable std::vector<mplx_waypoint> in_, succ_;
  mutable std::vector<int32_t> count_, action_;
  mutable std::vector<double> cost_;
  mutable long stats_nodes_ = 0, stats_calls_ = 0, stats_hits_ = 0;
};

/// Bump allocator for the overflow buffers of a search's predecessor lists: one rewind() returns
/// everything at once, so recycling the states of a finished search costs nothing per state.
class BumpPool {
 public:
  BumpPool() {}
  BumpPool(const BumpPool &) = delete;
  BumpPool &operator=(const BumpPool &) = delete;
  ~BumpPool() {
    for (char *b : blocks_) std::free(b);
    for (char *b : big_) std::free(b);
  }
  void *alloc(std::size_t n) {
    n = (n + 15) & ~(std::size_t)15;
    if (n > kBytes) {  // never for predecessor lists; kept correct
      char *p = (char *)std::malloc(n);
      if (!p) throw std::bad_alloc();
      big_.push_back(p);
      return p;
    }
    if (blocks_.empty() || used_ + n > kBytes) {
      if (!blocks_.empty() && cur_ + 1 < blocks_.size()) {
        cur_++;
      } else {
        char *b = (char *)std::malloc(kBytes);
        if (!b) throw std::bad_alloc();
        blocks_.push_back(b);
        cur_ = blocks_.size() - 1;
      }
      used_ = 0;
    }
    void *p = blocks_[cur_] + used_;
    used_ += n;
    return p;
  }
  void rewind() {
    cur_ = 0;
    used_ = 0;
    for (char *b : big_) std::free(b);
    big_.clear();
  }

 private:
  static constexpr std::size_t kBytes = 1 << 16;
  std::vector<char *> blocks_, big_;
  std::size_t cur_ = 0, used_ = 0;
};
/// The pool the SmallVecs of the calling thread grow into (set for the duration of one relax step by
/// PoolScope); null = malloc.
inline BumpPool *&tl_pred_pool() {
  static thread_local BumpPool *p = nullptr;
  return p;
}
struct PoolScope {
  BumpPool *prev;
  explicit PoolScope(BumpPool *p) : prev(tl_pred_pool()) { tl_pred_pool() = p; }
  ~PoolScope() { tl_pred_pool() = prev; }
};

/// A vector with room for N elements inside the object: most states have one or two predecessors,
/// so the common case needs no heap allocation.  Only what the planner uses.  A buffer grown while a
/// PoolScope is active lives in that pool (never freed individually); otherwise it is malloc'ed.
template <typename T, int N>
class SmallVec {
 public:
  SmallVec() {}
  SmallVec(const SmallVec &) = delete;
  SmallVec &operator=(const SmallVec &) = delete;
  ~SmallVec() { if (heap_ && owned_) std::free(heap_); }
  std::size_t size() const { return n_; }
  bool empty() const { return n_ == 0; }
  void clear() { n_ = 0; }
  T *begin() { return data(); }
  T *end() { return data() + n_; }
  const T *begin() const { return data(); }
  const T *end() const { return data() + n_; }
  T &operator[](std::size_t i) { return data()[i]; }
  const T &operator[](std::size_t i) const { return data()[i]; }
  void push_back(const T &v) {
    if (n_ == cap_) grow();
    data()[n_++] = v;
  }

 private:
  static_assert(std::is_trivially_copyable<T>::value, "SmallVec holds plain records");
  T *data() { return heap_ ? heap_ : inl_; }
  const T *data() const { return heap_ ? heap_ : inl_; }
  void grow() {
    const unsigned nc = cap_ * 2;
    BumpPool *pool = tl_pred_pool();
    T *p = (T *)(pool ? pool->alloc(sizeof(T) * nc) : std::malloc(sizeof(T) * nc));
    if (!p) throw std::bad_alloc();
    std::memcpy(p, data(), sizeof(T) * n_);
    if (heap_ && owned_) std::free(heap_);
    heap_ = p;
    owned_ = pool == nullptr;
    cap_ = nc;
  }
  T inl_[N];
  T *heap_ = nullptr;
  unsigned n_ = 0, cap_ = N;
  bool owned_ = false;
};

/// State: include/mpl_planner/common/state_space.h:36-74 (A* members)
template <int Dim>
struct State {
  Waypoint<Dim> coord;
  std::size_t key;
  /// pred_coord / pred_action_id / pred_action_cost of the reference (state_space.h:47-52), kept as
  /// one array of records (a predecessor is identified by its lattice key)
  struct Pred {
    State *node;
    decimal_t action_cost;
    int action_id;
  };
  SmallVec<Pred, 1> pred;
  /// succ_coord / succ_action_id / succ_action_cost (state_space.h:40-45), LPA* only.  The
  /// successor is looked up again by its lattice key when the node is re-expanded (hm_[succ_coord],
  /// graph_search.h:282), so a state pruned from the space in between is re-created, as there.
  struct Succ {
    State *node;
    decimal_t action_cost;
    int action_id;
  };
  std::vector<Succ> succ;
  int heap_idx = -1;
  decimal_t g = std::numeric_limits<decimal_t>::infinity();
  decimal_t rhs = std::numeric_limits<decimal_t>::infinity();
  decimal_t h = std::numeric_limits<decimal_t>::infinity();
  bool iterationopened = false, iterationclosed = false;
  State(const Waypoint<Dim> &c, std::size_t k) : coord(c), key(k) {}
};

/// priorityQueue: boost::heap::d_ary_heap<pair<f, State*>, arity<2>, mutable_<true>,
/// compare<compare_pair>> (state_space.h:16-34) restated: binary heap of (key, state) pairs with
/// position handles; push = append + sift up, pop = swap root with last + sift down, increase =
/// sift up, erase = move the element to the root, then pop; a child replaces its parent unless it
/// compares strictly lower.
template <int Dim>
class PriorityQueue {
 public:
  using S = State<Dim>;
  using Item = std::pair<decimal_t, S *>;
  /// compare_pair (state_space.h:16-27): true when p1 has LOWER priority than p2; ties on the key
  /// are broken on the states' current min(g, rhs)
  static bool lower(const Item &p1, const Item &p2) {
    if (p1.first == p2.first) return std::min(p1.second->g, p1.second->rhs) > std::min(p2.second->g, p2.second->rhs);
    return p1.first > p2.first;
  }
  bool empty() const { return q_.empty(); }
  std::size_t size() const { return q_.size(); }
  const Item &top() const { return q_.front(); }
  const std::vector<Item> &raw() const { return q_; }
  void clear() { q_.clear(); }
  void push(decimal_t key, S *s) { s->heap_idx = (int)q_.size(); q_.emplace_back(key, s); siftup(s->heap_idx); }
  void pop() {
    swap_at(0, (int)q_.size() - 1);
    q_.back().second->heap_idx = -1;
    q_.pop_back();
    if (!q_.empty()) siftdown(0);
  }
  /// (*heapkey).first = key; pq_.increase(heapkey)  (graph_search.h:118-119)
  void increase(S *s, decimal_t key) { q_[s->heap_idx].first = key; siftup(s->heap_idx); }
  void erase(S *s) {
    int i = s->heap_idx;
    while (i != 0) {
      const int p = (i - 1) / 2;
      swap_at(p, i);
      i = p;
    }
    pop();
  }

 private:
  void swap_at(int a, int b) { std::swap(q_[a], q_[b]); q_[a].second->heap_idx = a; q_[b].second->heap_idx = b; }
  void siftup(int i) {
    while (i != 0) {
      int p = (i - 1) / 2;
      if (lower(q_[p], q_[i])) { swap_at(p, i); i = p; } else return;
    }
  }
  void siftdown(int i) {
    const int n = (int)q_.size();
    while (2 * i + 1 < n) {
      int c = 2 * i + 1;
      if (c + 1 < n && lower(q_[c], q_[c + 1])) c = c + 1;  // first maximum among the children
      if (!lower(q_[c], q_[i])) { swap_at(c, i); i = c; } else return;
    }
  }
  std::vector<Item> q_;
};

/// lattice key -> state, open addressing with linear probing (the keys are already well-mixed 64-bit
/// hashes).  Only what the state space needs of the reference's hashMap (state_space.h:77-79):
/// look-up, insert-if-absent, swap; iteration goes through StateSpace::order_.  A slot is empty
/// while its value is null, so a key is only present once a state has been stored for it.
template <typename V>
class KeyMap {
 public:
  KeyMap() { rehash(64); }
  V *find(std::size_t k) const {
    for (std::size_t i = slot_of(k);; i = (i + 1) & mask_) {
      if (!tab_[i].second) return nullptr;
      if (tab_[i].first == k) return tab_[i].second;
    }
  }
  /// the value slot of k, entered (null) if absent; the caller must store a non-null value before the
  /// next call when `created`
  V *&obtain(std::size_t k, bool &created) {
    if ((size_ + 1) * 2 > tab_.size()) rehash(tab_.size() * 2);
    for (std::size_t i = slot_of(k);; i = (i + 1) & mask_) {
      if (!tab_[i].second) {
        tab_[i].first = k;
        size_++;
        created = true;
        return tab_[i].second;
      }
      if (tab_[i].first == k) {
        created = false;
        return tab_[i].second;
      }
    }
  }
  std::size_t size() const { return size_; }
  /// forget every key, keep the table
  void clear() {
    std::fill(tab_.begin(), tab_.end(), std::pair<std::size_t, V *>(0, nullptr));
    size_ = 0;
  }
  void swap(KeyMap &o) { tab_.swap(o.tab_); std::swap(mask_, o.mask_); std::swap(size_, o.size_); }

 private:
  std::size_t slot_of(std::size_t k) const { return ((k * 0x9e3779b97f4a7c15ULL) >> 20) & mask_; }
  void rehash(std::size_t n) {
    std::vector<std::pair<std::size_t, V *>> old(n);
    old.swap(tab_);
    mask_ = n - 1;
    for (const auto &e : old)
      if (e.second) {
        std::size_t i = slot_of(e.first);
        while (tab_[i].second) i = (i + 1) & mask_;
        tab_[i] = e;
      }
  }
  std::vector<std::pair<std::size_t, V *>> tab_;
  std::size_t mask_ = 0, size_ = 0;
};

/// StateSpace: include/mpl_planner/common/state_space.h:81-287
template <int Dim>
struct StateSpace {
  using S = State<Dim>;
  PriorityQueue<Dim> pq_;
  /// hashMap (state_space.h:77-79): lattice key -> state.  The states live in an arena owned by the
  /// space; order_ lists the states of hm_ in insertion order, which is the iteration order this
  /// planner defines for `for (it : hm_)` (boost::unordered_map leaves it unspecified; the loops
  /// of getSubStateSpace :184-192 and getLinkedNodes depend on it).
  KeyMap<S> hm_;
  std::vector<S *> order_;
  /// states are carved out of 256-state blocks: one allocation per block, addresses never move; blocks
  /// survive reset() and are handed out again
  struct Arena {
    static constexpr std::size_t kBlock = 256;
    std::vector<S *> blocks;
    std::size_t in_use = 0;  // blocks holding states
    std::size_t used = kBlock;
    S *emplace(const Waypoint<Dim> &c, std::size_t k) {
      if (used == kBlock) {
        if (in_use == blocks.size()) blocks.push_back((S *)::operator new(sizeof(S) * kBlock));
        in_use++;
        used = 0;
      }
      return new (blocks[in_use - 1] + used++) S(c, k);
    }
    /// end the life of every state; `trivial`: no state owns memory (A*-only search whose predecessor
    /// lists grew into the space's pool), so nothing has to be visited
    void rewind(bool trivial) {
      if (!trivial)
        for (std::size_t b = 0; b < in_use; b++) {
          const std::size_t n = b + 1 == in_use ? used : kBlock;
          for (std::size_t i = 0; i < n; i++) blocks[b][i].~S();
        }
      in_use = 0;
      used = kBlock;
    }
    ~Arena() {
      for (S *b : blocks) ::operator delete(b);
    }
  } arena_;
  /// overflow buffers of the predecessor lists of an A* search (AstarStepper::consume)
  BumpPool pred_pool_;
  /// true while only AstarStepper touched the states: none of them owns memory
  bool trivial_states_ = true;
  ~StateSpace() { arena_.rewind(trivial_states_); }
  /// Forget the search but keep every allocation (state blocks, predecessor pool, hash table, heap and
  /// order arrays) for the next one on this space.
  void reset(decimal_t eps) {
    arena_.rewind(trivial_states_);
    pred_pool_.rewind();
    trivial_states_ = true;
    hm_.clear();
    order_.clear();
    pq_.clear();
    best_child_.clear();
    expand_iteration_ = 0;
    eps_ = eps;
    start_t_ = start_g_ = start_rhs_ = 0;
  }
  /// hm_[coord] of a key that is not in the map yet: create the state and enter it
  S *make_state(const Waypoint<Dim> &c, std::size_t k) {
    S *n = arena_.emplace(c, k);
    bool created;
    hm_.obtain(k, created) = n;
    order_.push_back(n);
    return n;
  }
  S *find(std::size_t k) const { return hm_.find(k); }
  /// `StatePtr &p = hm_[coord]; if (!p) p = make_shared<State>(coord)` (graph_search.h:84-87) with
  /// one hash-map operation; coord() is only evaluated for a new state
  template <typename MakeCoord>
  S *get_or_make(std::size_t k, MakeCoord coord, bool &created) {
    S *&slot = hm_.obtain(k, created);
    if (created) {
      slot = arena_.emplace(coord(), k);
      order_.push_back(slot);
    }
    return slot;
  }
  decimal_t eps_;
  decimal_t dt_{1};
  std::vector<S *> best_child_;
  int expand_iteration_ = 0;
  /// state_space.h:96-100
  decimal_t start_t_{0}, start_g_{0}, start_rhs_{0};
  explicit StateSpace(decimal_t eps = 1) : eps_(eps) {}

  /// state_space.h:106-111
  decimal_t getInitTime() const { return best_child_.empty() ? 0 : best_child_.front()->coord.t; }

  /// calculateKey: state_space.h:283-285
  decimal_t calculateKey(const S *node) const { return std::min(node->g, node->rhs) + eps_ * node->h; }

  /// updateNode: state_space.h:254-280
  void updateNode(S *currNode_ptr) {
    if (currNode_ptr->rhs != start_rhs_) {  // compares VALUES, as the reference does
      currNode_ptr->rhs = std::numeric_limits<decimal_t>::infinity();
      for (const auto &pr : currNode_ptr->pred)
        if (currNode_ptr->rhs > pr.node->g + pr.action_cost) currNode_ptr->rhs = pr.node->g + pr.action_cost;
    }
    if (currNode_ptr->iterationopened && !currNode_ptr->iterationclosed) {
      pq_.erase(currNode_ptr);
      currNode_ptr->iterationclosed = true;
    }
    if (currNode_ptr->g != currNode_ptr->rhs) {
      pq_.push(calculateKey(currNode_ptr), currNode_ptr);
      currNode_ptr->iterationopened = true;
      currNode_ptr->iterationclosed = false;
    }
  }

  /// getSubStateSpace: state_space.h:116-197 — re-root the graph at best_child_[time_step]
  void getSubStateSpace(int time_step) {
    if (best_child_.empty()) return;
    const decimal_t inf = std::numeric_limits<decimal_t>::infinity();
    S *currNode_ptr = best_child_[time_step];
    start_g_ = currNode_ptr->g;
    start_rhs_ = currNode_ptr->rhs;
    start_t_ = currNode_ptr->coord.t;
    currNode_ptr->pred.clear();
    for (S *it : order_) {
      it->g = inf;
      it->rhs = inf;
      it->pred.clear();
    }
    currNode_ptr->g = start_g_;
    currNode_ptr->rhs = start_rhs_;

    KeyMap<S> new_hm;
    std::vector<S *> new_order;
    PriorityQueue<Dim> epq;
    epq.push(currNode_ptr->rhs, currNode_ptr);
    bool fresh;
    new_hm.obtain(currNode_ptr->key, fresh) = currNode_ptr;
    new_order.push_back(currNode_ptr);
    while (!epq.empty()) {
      currNode_ptr = epq.top().second;
      epq.pop();
      for (std::size_t i = 0; i < currNode_ptr->succ.size(); i++) {
        const std::size_t skey = currNode_ptr->succ[i].node->key;
        S *&slot = new_hm.obtain(skey, fresh);
        if (fresh) {
          slot = find(skey);  // hm_[succ_coord]; the reference reports a "critical bug" when absent
          if (!slot) throw std::logic_error("getSubStateSpace: successor is not in the state space");
          new_order.push_back(slot);
        }
        S *succNode_ptr = slot;
        int id = -1;
        for (std::size_t k = 0; k < succNode_ptr->pred.size(); k++)
          if (succNode_ptr->pred[k].node->key == currNode_ptr->key) { id = (int)k; break; }
        if (id == -1)
          succNode_ptr->pred.push_back(typename S::Pred{currNode_ptr, currNode_ptr->succ[i].action_cost,
                                                        currNode_ptr->succ[i].action_id});
        const decimal_t tentative_rhs = currNode_ptr->rhs + currNode_ptr->succ[i].action_cost;
        if (tentative_rhs < succNode_ptr->rhs) {
          succNode_ptr->rhs = tentative_rhs;
          if (succNode_ptr->iterationclosed) {
            succNode_ptr->g = succNode_ptr->rhs;
            epq.push(succNode_ptr->rhs, succNode_ptr);
          }
        }
      }
    }
    hm_.swap(new_hm);
    order_.swap(new_order);
    pq_.clear();
    for (S *it : order_)
      if (it->iterationopened && !it->iterationclosed) pq_.push(calculateKey(it), it);
  }

  /// One (state, i-th predecessor) reference: std::pair<Coord, int> of the reference (state_space.h:200,225)
  using EdgeRef = std::pair<S *, int>;

  /// increaseCost: state_space.h:200-221
  void increaseCost(const std::vector<EdgeRef> &states) {
    const decimal_t inf = std::numeric_limits<decimal_t>::infinity();
    for (const auto &affected_node : states) {
      S *succNode_ptr = affected_node.first;
      const int i = affected_node.second;
      if (!std::isinf(succNode_ptr->pred[i].action_cost)) {
        succNode_ptr->pred[i].action_cost = inf;
        updateNode(succNode_ptr);
        S *parent = succNode_ptr->pred[i].node;
        const int succ_act_id = succNode_ptr->pred[i].action_id;
        for (auto &sc : parent->succ)
          if (succ_act_id == sc.action_id) { sc.action_cost = inf; break; }
      }
    }
  }

  /// decreaseCost: state_space.h:223-251.  is_free(pr) of every still-blocked edge is asked in ONE
  /// batched query up front (the map does not change during the call, so the answers are the ones
  /// the reference's per-edge calls would get); the updates are then applied in list order.
  template <typename Env>
  void decreaseCost(const std::vector<EdgeRef> &states, const Env &ENV) {
    vec_E<Waypoint<Dim>> parents;
    std::vector<int> actions;
    std::vector<int> slot(states.size(), -1);
    for (std::size_t k = 0; k < states.size(); k++) {
      const auto &pr = states[k].first->pred[states[k].second];
      if (std::isinf(pr.action_cost)) {
        slot[k] = (int)parents.size();
        parents.push_back(pr.node->coord);  // forward_action(parent_key, action_id): env_base.h:228-231
        actions.push_back(pr.action_id);
      }
    }
    std::vector<uint8_t> free;
    std::vector<decimal_t> cost;
    if (!parents.empty()) ENV.is_free_edges(parents, actions, free, cost);
    for (std::size_t k = 0; k < states.size(); k++) {
      S *succNode_ptr = states[k].first;
      const int i = states[k].second;
      if (std::isinf(succNode_ptr->pred[i].action_cost) && free[slot[k]]) {
        succNode_ptr->pred[i].action_cost = cost[slot[k]];  // calculate_intrinsic_cost(pr)
        updateNode(succNode_ptr);
        S *parent = succNode_ptr->pred[i].node;
        const int succ_act_id = succNode_ptr->pred[i].action_id;
        for (auto &sc : parent->succ)
          if (succ_act_id == sc.action_id) { sc.action_cost = succNode_ptr->pred[i].action_cost; break; }
      }
    }
  }
};

/// One edge of the recovered trajectory (the reference stores Primitive<Dim>; a primitive built
/// by the state+control constructor is fully described by its start node and action id,
/// env_base.h:228-231).
template <int Dim>
struct Edge { Waypoint<Dim> from; int action_id; };

/// recoverTraj: graph_search.h:369-455 (shared by Astar and LPAstar)
template <int Dim>
bool recoverTraj(State<Dim> *currNode_ptr, StateSpace<Dim> &ss, std::size_t start_key, std::vector<Edge<Dim>> &traj) {
  using S = State<Dim>;
  const decimal_t inf = std::numeric_limits<decimal_t>::infinity();
  ss.best_child_.clear();
  bool find_traj = false;
  std::vector<Edge<Dim>> prs;
  while (!currNode_ptr->pred.empty()) {
    ss.best_child_.push_back(currNode_ptr);
    int min_id = -1;
    decimal_t min_rhs = inf, min_g = inf;
    for (unsigned int i = 0; i < currNode_ptr->pred.size(); i++) {
      const S *pred = currNode_ptr->pred[i].node;
      const decimal_t ac = currNode_ptr->pred[i].action_cost;
      if (min_rhs > pred->g + ac) {
        min_rhs = pred->g + ac;
        min_g = pred->g;
        min_id = i;
      } else if (!std::isinf(ac) && min_rhs == pred->g + ac) {
        if (min_g < pred->g) {
          min_g = pred->g;
          min_id = i;
        }
      }
    }
    if (min_id >= 0) {
      int action_idx = currNode_ptr->pred[min_id].action_id;
      currNode_ptr = currNode_ptr->pred[min_id].node;
      prs.push_back(Edge<Dim>{currNode_ptr->coord, action_idx});  // forward_action(coord, action): env_base.h:228-231
    } else
      break;
    if (currNode_ptr->key == start_key) {
      ss.best_child_.push_back(currNode_ptr);
      find_traj = true;
      break;
    }
  }
  std::reverse(prs.begin(), prs.end());
  std::reverse(ss.best_child_.begin(), ss.best_child_.end());
  traj = find_traj ? prs : std::vector<Edge<Dim>>();
  return find_traj;
}

/// The A* loop of include/mpl_planner/common/graph_search.h:39-182 cut at the get_succ call, so
/// that a driver can either run it to completion for one query (GraphSearch::Astar below) or
/// advance many queries in lock-step and expand their current nodes in ONE device launch
/// (MultiQueryPlanner).  The order of operations inside an iteration is the reference's:
/// pop + close (:66-68) -> get_succ (:75) -> relax successors (:79-143) -> goal test (:146) ->
/// max_expand (:149-155) -> empty queue (:157-161).
template <int Dim>
class AstarStepper {
 public:
  using S = State<Dim>;
  AstarStepper(const env_base<Dim> *env, std::shared_ptr<StateSpace<Dim>> ss, int max_expand)
      : ENV(env), ss_ptr(ss), max_expand_(max_expand) {}

  /// graph_search.h:43-61
  void start(const Waypoint<Dim> &start_coord) {
    start_key_ = hash_value(start_coord);
    if (ENV->is_goal(start_coord)) {
      status_ = DONE_TRIVIAL;
      return;
    }
    if (ss_ptr->pq_.empty()) {
      S *n = ss_ptr->make_state(start_coord, start_key_);
      n->g = 0;
      n->h = ss_ptr->eps_ == 0 ? 0 : ENV->get_heur(start_coord);
      ss_ptr->pq_.push(n->g + ss_ptr->eps_ * n->h, n);
      n->iterationopened = true;
      n->iterationclosed = false;
    }
    status_ = RUNNING;
  }
  bool active() const { return status_ == RUNNING; }
  const std::vector<std::pair<decimal_t, S *>> &open_heap() const { return ss_ptr->pq_.raw(); }

  /// graph_search.h:64-68: the node this iteration expands
  const Waypoint<Dim> &pop() {
    expand_iteration_++;
    curr_ = ss_ptr->pq_.top().second;
    ss_ptr->pq_.pop();
    curr_->iterationclosed = true;
    return curr_->coord;
  }

  /// graph_search.h:79-161 with the successors of the node returned by pop()
  template <typename SuccAt, typename KeyAt>
  void consume(int n_succ, SuccAt succ_at, const decimal_t *succ_cost, const int *succ_act_id, KeyAt key_at) {
    const PoolScope pool(&ss_ptr->pred_pool_);  // predecessor lists grow into the space's pool
    for (int s = 0; s < n_succ; ++s) {
      if (std::isinf(succ_cost[s])) continue;  // graph_search.h:81
      const std::size_t skey = key_at(s);
      bool created;
      S *succNode_ptr = ss_ptr->get_or_make(skey, [&] { return succ_at(s); }, created);
      if (created) succNode_ptr->h = ss_ptr->eps_ == 0 ? 0 : ENV->get_heur(succNode_ptr->coord, skey);
      succNode_ptr->pred.push_back(typename S::Pred{curr_, succ_cost[s], succ_act_id[s]});
      const decimal_t tentative_gval = curr_->g + succ_cost[s];
      if (tentative_gval < succNode_ptr->g) {
        succNode_ptr->g = tentative_gval;
        const decimal_t fval = succNode_ptr->g + (ss_ptr->eps_) * succNode_ptr->h;
        if (succNode_ptr->iterationopened && !succNode_ptr->iterationclosed) {
          ss_ptr->pq_.increase(succNode_ptr, fval);
        } else {
          ss_ptr->pq_.push(fval, succNode_ptr);
          succNode_ptr->iterationopened = true;
        }
      }
    }
    if (ENV->is_goal(curr_->coord)) {
      status_ = DONE_GOAL;
    } else if (max_expand_ > 0 && expand_iteration_ >= max_expand_) {
      status_ = FAILED;
    } else if (ss_ptr->pq_.empty()) {
      status_ = FAILED;
    }
    if (status_ != RUNNING) ss_ptr->expand_iteration_ = expand_iteration_;
  }

  /// graph_search.h:163-181: cost + recovered trajectory
  decimal_t finish(std::vector<Edge<Dim>> &traj) {
    const decimal_t inf = std::numeric_limits<decimal_t>::infinity();
    traj.clear();
    if (status_ == DONE_TRIVIAL) return 0;
    if (status_ != DONE_GOAL) return inf;
    if (recoverTraj<Dim>(curr_, *ss_ptr, start_key_, traj)) return curr_->g;
    return inf;
  }
  int expanded() const { return expand_iteration_; }

 private:
  enum Status { IDLE, RUNNING, DONE_TRIVIAL, DONE_GOAL, FAILED };
  const env_base<Dim> *ENV;
  std::shared_ptr<StateSpace<Dim>> ss_ptr;
  int max_expand_;
  Status status_ = IDLE;
  int expand_iteration_ = 0;
  std::size_t start_key_ = 0;
  S *curr_ = nullptr;
};

/// GraphSearch::Astar: include/mpl_planner/common/graph_search.h:39-182
template <int Dim>
class GraphSearch {
 public:
  explicit GraphSearch(bool verbose = false, int lookahead = 0) : verbose_(verbose), lookahead_(lookahead) {}

  decimal_t Astar(const Waypoint<Dim> &start_coord, const std::shared_ptr<env_base<Dim>> &ENV,
                  std::shared_ptr<StateSpace<Dim>> &ss_ptr, std::vector<Edge<Dim>> &traj, int max_expand = -1) {
    AstarStepper<Dim> st(ENV.get(), ss_ptr, max_expand);
    st.start(start_coord);
    vec_E<Waypoint<Dim>> succ_coord;
    std::vector<decimal_t> succ_cost;
    std::vector<int> succ_act_id;
    vec_E<Waypoint<Dim>> cands;
    std::vector<std::size_t> cand_keys;
    while (st.active()) {
      const auto &raw = st.open_heap();
      if (lookahead_ > 0 && !raw.empty() && ENV->wants_candidates(raw[0].second->key)) {
        // hint: the open nodes nearest the root of the heap are the likeliest next pops
        cands.clear();
        cand_keys.clear();
        for (std::size_t i = 1; i < raw.size() && (int)cands.size() < lookahead_; i++) {
          cands.push_back(raw[i].second->coord);
          cand_keys.push_back(raw[i].second->key);
        }
        ENV->prefetch(cands, cand_keys);
      }
      const Waypoint<Dim> &curr = st.pop();
      ENV->get_succ(curr, succ_coord, succ_cost, succ_act_id);
      const std::size_t *keys = ENV->last_succ_keys();
      st.consume((int)succ_coord.size(), [&](int s) -> const Waypoint<Dim> & { return succ_coord[s]; },
                 succ_cost.data(), succ_act_id.data(),
                 [&](int s) { return keys ? keys[s] : hash_value(succ_coord[s]); });
    }
    if (verbose_ && std::isinf(st.finish(traj))) printf("[GraphSearch] no trajectory (max expansions or empty queue)\n");
    return st.finish(traj);
  }

  /// LPAstar: include/mpl_planner/common/graph_search.h:194-365.  +inf successors are kept (they
  /// may be re-opened by decreaseCost).  An empty queue at entry reads as a +inf top key (the
  /// reference dereferences pq_.top() there).
  decimal_t LPAstar(const Waypoint<Dim> &start_coord, const std::shared_ptr<env_base<Dim>> &ENV,
                    std::shared_ptr<StateSpace<Dim>> &ss_ptr, std::vector<Edge<Dim>> &traj, int max_expand = -1) {
    ss_ptr->trivial_states_ = false;  // LPA* keeps successor lists and malloc'ed predecessor lists in the states
    using S = State<Dim>;
    const decimal_t inf = std::numeric_limits<decimal_t>::infinity();
    traj.clear();
    if (ENV->is_goal(start_coord)) return 0;
    const std::size_t start_key = hash_value(start_coord);
    S *currNode_ptr = ss_ptr->find(start_key);
    if (!currNode_ptr) {
      currNode_ptr = ss_ptr->make_state(start_coord, start_key);
      currNode_ptr->g = inf;
      currNode_ptr->rhs = 0;
      currNode_ptr->h = ss_ptr->eps_ == 0 ? 0 : ENV->get_heur(start_coord);
      ss_ptr->pq_.push(ss_ptr->calculateKey(currNode_ptr), currNode_ptr);
      currNode_ptr->iterationopened = true;
      currNode_ptr->iterationclosed = false;
    }
    // goal node: the previous goal if it still is one, else a detached placeholder (:222-240)
    S goal_placeholder{Waypoint<Dim>(), 0};
    S *goalNode_ptr = &goal_placeholder;
    if (!ss_ptr->best_child_.empty() && ENV->is_goal(ss_ptr->best_child_.back()->coord)) {
      goalNode_ptr = ss_ptr->best_child_.back();
    } else {
      goalNode_ptr->g = inf;
      goalNode_ptr->rhs = inf;
      goalNode_ptr->h = 0;
    }

    int expand_iteration = 0;
    vec_E<Waypoint<Dim>> succ_coord, cands;
    std::vector<decimal_t> succ_cost;
    std::vector<int> succ_act_id;
    std::vector<std::size_t> succ_key, cand_keys;
    auto top_key = [&] { return ss_ptr->pq_.empty() ? inf : ss_ptr->pq_.top().first; };
    while (top_key() < ss_ptr->calculateKey(goalNode_ptr) || goalNode_ptr->rhs != goalNode_ptr->g) {
      if (ss_ptr->pq_.empty()) return inf;
      expand_iteration++;
      const auto &raw = ss_ptr->pq_.raw();
      if (lookahead_ > 0 && raw[0].second->succ.empty() && ENV->wants_candidates(raw[0].second->key)) {
        cands.clear();
        cand_keys.clear();
        for (std::size_t i = 1; i < raw.size() && (int)cands.size() < lookahead_; i++)
          if (raw[i].second->succ.empty()) {
            cands.push_back(raw[i].second->coord);
            cand_keys.push_back(raw[i].second->key);
          }
        ENV->prefetch(cands, cand_keys);
      }
      currNode_ptr = ss_ptr->pq_.top().second;
      ss_ptr->pq_.pop();
      currNode_ptr->iterationclosed = true;
      if (currNode_ptr->g > currNode_ptr->rhs)
        currNode_ptr->g = currNode_ptr->rhs;
      else {
        currNode_ptr->g = inf;
        ss_ptr->updateNode(currNode_ptr);
      }

      // successors: stored ones if the node was explored before, else get_succ (:259-271)
      const bool explored = !currNode_ptr->succ.empty();
      succ_key.clear();
      if (explored) {
        succ_coord.clear(); succ_cost.clear(); succ_act_id.clear();
        for (const auto &sc : currNode_ptr->succ) {
          succ_coord.push_back(sc.node->coord);
          succ_key.push_back(sc.node->key);
          succ_cost.push_back(sc.action_cost);
          succ_act_id.push_back(sc.action_id);
        }
      } else {
        ENV->get_succ(currNode_ptr->coord, succ_coord, succ_cost, succ_act_id);
        const std::size_t *keys = ENV->last_succ_keys();
        for (std::size_t s = 0; s < succ_coord.size(); s++) succ_key.push_back(keys ? keys[s] : hash_value(succ_coord[s]));
        currNode_ptr->succ.resize(succ_coord.size());
      }

      for (std::size_t s = 0; s < succ_coord.size(); ++s) {
        bool created;
        S *succNode_ptr = ss_ptr->get_or_make(succ_key[s], [&] { return succ_coord[s]; }, created);
        if (created) succNode_ptr->h = ss_ptr->eps_ == 0 ? 0 : ENV->get_heur(succNode_ptr->coord, succ_key[s]);
        currNode_ptr->succ[s] = typename S::Succ{succNode_ptr, succ_cost[s], succ_act_id[s]};
        int id = -1;
        for (std::size_t i = 0; i < succNode_ptr->pred.size(); i++)
          if (succNode_ptr->pred[i].node->key == currNode_ptr->key) { id = (int)i; break; }
        if (id == -1) succNode_ptr->pred.push_back(typename S::Pred{currNode_ptr, succ_cost[s], succ_act_id[s]});
        ss_ptr->updateNode(succNode_ptr);
      }

      if (ENV->is_goal(currNode_ptr->coord)) goalNode_ptr = currNode_ptr;
      if (max_expand > 0 && expand_iteration >= max_expand) return inf;
      if (ss_ptr->pq_.empty()) return inf;
    }
    ss_ptr->expand_iteration_ = expand_iteration;
    if (recoverTraj<Dim>(goalNode_ptr, *ss_ptr, start_key, traj)) return goalNode_ptr->g - ss_ptr->start_g_;
    return inf;
  }

 private:
  bool verbose_;
  int lookahead_;
};

/// PlannerBase + MapPlanner: include/mpl_planner/common/planner_base.h, planner/map_planner.h
template <int Dim>
class PlannerBase {
 public:
  explicit PlannerBase(bool verbose = false) : planner_verbose_(verbose) {}
  virtual ~PlannerBase() {}
  bool initialized() { return !(ss_ptr_ == nullptr); }
  std::vector<Edge<Dim>> getTraj() const { return traj_; }
  /// planner_base.h getTraj(): the recovered trajectory as piece-wise polynomials
  /// (recoverTraj's forward_action per edge, graph_search.h:417-419)
  Trajectory<Dim> getTrajectory() const {
    vec_E<Primitive<Dim>> prs;
    for (const auto &e : traj_) {
      Primitive<Dim> pr;
      ENV_->forward_action(e.from, e.action_id, pr);
      prs.push_back(pr);
    }
    return Trajectory<Dim>(prs);
  }
  decimal_t getTrajCost() const { return traj_cost_; }
  int getExpandedNum() const { return ss_ptr_ ? ss_ptr_->expand_iteration_ : 0; }
  vec_E<Vecf<Dim>> getExpandedNodes() const { return ENV_->expanded_nodes_; }
  /// getCloseSet (planner_base.h): states with iterationclosed
  std::vector<const State<Dim> *> getCloseSetStates() const {
    std::vector<const State<Dim> *> v;
    for (const auto *st : ss_ptr_->order_) if (st->iterationclosed) v.push_back(st);
    return v;
  }
  std::size_t getOpenSetSize() const {
    std::size_t n = 0;
    for (const auto *st : ss_ptr_->order_) if (st->iterationopened && !st->iterationclosed) n++;
    return n;
  }
  void setVmax(decimal_t v) { ENV_->set_v_max(v); }
  void setAmax(decimal_t a) { ENV_->set_a_max(a); }
  void setJmax(decimal_t j) { ENV_->set_j_max(j); }
  void setYawmax(decimal_t yaw) { ENV_->set_yaw_max(yaw); }
  void setTmax(decimal_t t) { ENV_->set_t_max(t); }
  void setDt(decimal_t dt) { ENV_->set_dt(dt); }
  void setW(decimal_t w) { ENV_->set_w(w); }
  void setWyaw(decimal_t w) { ENV_->set_wyaw(w); }
  void setEpsilon(decimal_t eps) { epsilon_ = eps; }
  void setMaxNum(int num) { max_num_ = num; }
  void setU(const vec_E<VecDf> &U) { ENV_->set_u(U); }
  void setTol(decimal_t tol_pos, decimal_t tol_vel = -1, decimal_t tol_acc = -1) {
    ENV_->set_tol_pos(tol_pos); ENV_->set_tol_vel(tol_vel); ENV_->set_tol_acc(tol_acc);
  }
  void setLookahead(int k) { lookahead_ = k; }
  /// planner_base.h:233-237
  void setHeurIgnoreDynamics(bool ignore) { ENV_->set_heur_ignore_dynamics(ignore); }
  /// planner_base.h:170-176
  void setLPAstar(bool use_lpastar) { use_lpastar_ = use_lpastar; }
  /// planner_base.h:155: prune the state space to the subtree under best_child_[time_step]
  void getSubStateSpace(int time_step) { ss_ptr_->getSubStateSpace(time_step); }
  /// planner_base.h:164-167
  void reset() { ss_ptr_ = nullptr; traj_.clear(); }
  const std::shared_ptr<StateSpace<Dim>> &stateSpace() const { return ss_ptr_; }

  /// planner_base.h:275-325
  bool plan(const Waypoint<Dim> &start, const Waypoint<Dim> &goal) {
    if (!ENV_->is_free(start.pos)) {
      printf("[PlannerBase] start is not free!\n");
      return false;
    }
    GraphSearch<Dim> planner(planner_verbose_, lookahead_);
    // A*: a fresh state space per plan; LPA*: only at the first plan (planner_base.h:293-303)
    if (!use_lpastar_ || !initialized()) ss_ptr_.reset(new StateSpace<Dim>(epsilon_));
    ENV_->set_goal(goal);
    ENV_->expanded_nodes_.clear();
    ENV_->begin_plan();
    ss_ptr_->dt_ = ENV_->get_dt();
    if (use_lpastar_)
      traj_cost_ = planner.LPAstar(start, ENV_, ss_ptr_, traj_, max_num_);
    else
      traj_cost_ = planner.Astar(start, ENV_, ss_ptr_, traj_, max_num_);
    if (std::isinf(traj_cost_)) return false;
    return true;
  }

 protected:
  std::shared_ptr<env_base<Dim>> ENV_;
  std::shared_ptr<StateSpace<Dim>> ss_ptr_;
  std::vector<Edge<Dim>> traj_;
  decimal_t traj_cost_ = 0;
  decimal_t epsilon_ = 1.0;
  int max_num_ = -1;
  int lookahead_ = 0;
  bool planner_verbose_;
  bool use_lpastar_ = false;
};

template <int Dim>
class MapPlanner : public PlannerBase<Dim> {
 public:
  explicit MapPlanner(bool verbose = false) : PlannerBase<Dim>(verbose) {}
  /// src/mpl_planner/map_planner.cpp:14-18 — installs the GPU env instead of env_map<Dim>
  virtual void setMapUtil(const std::shared_ptr<MapUtil<Dim>> &map_util, int device = 0) {
    gpu_env_.reset(new env_map_gpu<Dim>(map_util, device));
    this->ENV_ = gpu_env_;
    map_util_ = map_util;
    // One launch per popped node costs ~40 us of launch + PCIe latency against ~25 us of CPU get_succ, so the
    // default expands the node together with the best open nodes it is likely to pop next (results are
    // identical for any value: get_succ is a pure function of the node); setSpeculation(1) = one node per call.
    setSpeculation(kDefaultSpeculation);
  }
  static constexpr int kDefaultSpeculation = 32;
  /// Any env_base implementation (the closed-set equality tests install a CPU checker env here).
  void setEnv(const std::shared_ptr<env_base<Dim>> &env, const std::shared_ptr<MapUtil<Dim>> &map_util = nullptr) {
    this->ENV_ = env;
    gpu_env_.reset();
    if (map_util) map_util_ = map_util;
  }
  void setControl(int control) { if (gpu_env_) gpu_env_->set_control(control); }
  void setSpeculation(int k) { if (gpu_env_) gpu_env_->set_speculation(k); this->setLookahead(k > 1 ? 4 * k : 0); }
  /// map_planner.cpp:20-43
  void setPotentialRadius(const Vecf<Dim> &radius) { potential_radius_ = radius; }
  void setPotentialMapRange(const Vecf<Dim> &range) { potential_map_range_ = range; }
  void setSearchRadius(const Vecf<Dim> &radius) { search_radius_ = radius; }
  /// map_planner.cpp:323-391, on the device
  void updatePotentialMap(const Vecf<Dim> &pos) {
    if (!gpu_env_) throw std::runtime_error("updatePotentialMap needs the GPU env (setMapUtil)");
    gpu_env_->update_potential_map(potential_radius_, pow_, potential_map_range_, pos);
  }
  /// map_planner.cpp:46-95, on the device
  void setSearchRegion(const vec_E<Vecf<Dim>> &path, bool dense = false) {
    this->ENV_->search_region_from_path(path, search_radius_, dense);
  }
  /// Trajectory::getWaypoints() positions of the last plan (trajectory.h:277-289): the start state of
  /// every primitive and the end state of the last one = the states recoverTraj walked, start to goal
  vec_E<Vecf<Dim>> getWaypointPositions() const {
    vec_E<Vecf<Dim>> path;
    if (this->ss_ptr_)
      for (const auto *st : this->ss_ptr_->best_child_) path.push_back(st->coord.pos);
    return path;
  }
  /// iterativePlan: src/mpl_planner/map_planner.cpp:393-433 — replan inside a tunnel around the previous
  /// trajectory until the cost stops changing (or max_num iterations).  raw_path = the waypoint
  /// positions of the trajectory to start from (getWaypointPositions() of an earlier plan).
  bool iterativePlan(const Waypoint<Dim> &start, const Waypoint<Dim> &goal, const vec_E<Vecf<Dim>> &raw_path,
                     int max_num = 3) {
    vec_E<Vecf<Dim>> path = raw_path;
    double prev_traj_cost = 0;
    iterations_ = 0;
    int cnt = 0;
    while (cnt < max_num) {
      cnt++;
      iterations_ = cnt;
      setSearchRegion(path, false);
      if (!this->plan(start, goal)) return false;
      if (prev_traj_cost == this->traj_cost_) break;
      prev_traj_cost = this->traj_cost_;
      path = getWaypointPositions();
    }
    return true;
  }
  /// plan() calls made by the last iterativePlan
  int iterations() const { return iterations_; }
  void setPotentialWeight(decimal_t w) { this->ENV_->set_potential_weight(w); }
  void setGradientWeight(decimal_t w) { this->ENV_->set_gradient_weight(w); }
  env_map_gpu<Dim> *gpu_env() { return gpu_env_.get(); }

  /// The reference's lhm_ (map_planner.h:15-16,101: voxel index -> the (state, i-th predecessor)
  /// edges through it) kept as a table sorted by voxel index; the edges of one voxel are in the
  /// order getLinkedNodes' push_backs would have left them.
  struct LinkedTable {
    std::vector<int> voxel, edge;                          // sorted by voxel (stable)
    std::vector<std::pair<State<Dim> *, int>> owner;       // edge -> (state, pred index)
    void clear() { voxel.clear(); edge.clear(); owner.clear(); }
    /// append the edges through voxel `id`, in lhm_[id] order
    void collect(int id, std::vector<std::pair<State<Dim> *, int>> &out) const {
      auto r = std::equal_range(voxel.begin(), voxel.end(), id);
      for (auto it = r.first; it != r.second; ++it) out.push_back(owner[edge[it - voxel.begin()]]);
    }
  };

  /// getLinkedNodes: src/mpl_planner/map_planner.cpp:124-157.  Every stored edge (state, i-th
  /// predecessor) is walked through the grid — on the device for the GPU env, all edges in one
  /// batch, which also sorts the (voxel, edge) pairs into the table — and the voxel centres are
  /// returned in the reference's order.
  vec_E<Vecf<Dim>> getLinkedNodes() const {
    using S = State<Dim>;
    const auto t_begin = std::chrono::steady_clock::now();
    lhm_.clear();
    vec_E<Vecf<Dim>> linked_pts;
    vec_E<Waypoint<Dim>> parents;
    std::vector<int> actions;
    std::size_t n_edges = 0;
    for (const S *st : this->ss_ptr_->order_) n_edges += st->pred.size();
    parents.reserve(n_edges); actions.reserve(n_edges); lhm_.owner.reserve(n_edges);
    for (S *st : this->ss_ptr_->order_)
      for (std::size_t i = 0; i < st->pred.size(); i++) {
        parents.push_back(st->pred[i].node->coord);
        actions.push_back(st->pred[i].action_id);
        lhm_.owner.emplace_back(st, (int)i);
      }
    std::vector<long long> offset;
    std::vector<int> cells;
    if (parents.empty()) return linked_pts;
    const auto t_env0 = std::chrono::steady_clock::now();
    this->ENV_->edge_cells(parents, actions, offset, cells, lhm_.voxel, lhm_.edge);
    const auto t_env1 = std::chrono::steady_clock::now();
    const std::size_t total = cells.size() / Dim;
    const decimal_t res = map_util_->getRes();
    const Vecf<Dim> ori = map_util_->getOrigin();
    const bool host_table = lhm_.voxel.size() != total;
    std::vector<int> ids;
    if (host_table) ids.resize(total);
    linked_pts.resize(total);
    for (std::size_t c = 0; c < total; c++) {
      Veci<Dim> pn;
      for (int k = 0; k < Dim; k++) pn(k) = cells[c * Dim + k];
      for (int k = 0; k < Dim; k++) linked_pts[c](k) = (pn(k) + 0.5) * res + ori(k);  // intToFloat: map_util.h:110-113
      if (host_table) ids[c] = map_util_->getIndex(pn);
    }
    if (host_table) {  // an env without the device sort: stable sort of the (voxel, edge) pairs here
      std::vector<int> owner_of(total), perm(total);
      for (std::size_t e = 0; e < parents.size(); e++)
        for (long long c = offset[e]; c < offset[e + 1]; c++) owner_of[c] = (int)e;
      for (std::size_t c = 0; c < total; c++) perm[c] = (int)c;
      std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return ids[x] < ids[y]; });
      lhm_.voxel.resize(total); lhm_.edge.resize(total);
      for (std::size_t c = 0; c < total; c++) { lhm_.voxel[c] = ids[perm[c]]; lhm_.edge[c] = owner_of[perm[c]]; }
    }
    if (std::getenv("MPLH_TRACE")) {
      auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
      };
      std::fprintf(stderr, "[getLinkedNodes] %zu edges, %zu voxels: gather %.1f ms, env walk %.1f ms, points/table %.1f ms\n",
                   parents.size(), total, ms(t_begin, t_env0), ms(t_env0, t_env1), ms(t_env1, std::chrono::steady_clock::now()));
    }
    return linked_pts;
  }
  /// updateBlockedNodes: map_planner.cpp:159-171
  void updateBlockedNodes(const vec_E<Veci<Dim>> &blocked_pns) {
    std::vector<std::pair<State<Dim> *, int>> blocked_nodes;
    for (const auto &it : blocked_pns) lhm_.collect(map_util_->getIndex(it), blocked_nodes);
    this->ss_ptr_->increaseCost(blocked_nodes);
  }
  /// updateClearedNodes: map_planner.cpp:173-185; the is_free(pr) re-validation runs batched in the env
  void updateClearedNodes(const vec_E<Veci<Dim>> &cleared_pns) {
    std::vector<std::pair<State<Dim> *, int>> cleared_nodes;
    for (const auto &it : cleared_pns) lhm_.collect(map_util_->getIndex(it), cleared_nodes);
    this->ss_ptr_->decreaseCost(cleared_nodes, *this->ENV_);
  }
  const LinkedTable &linkedTable() const { return lhm_; }

 protected:
  mutable LinkedTable lhm_;
  int iterations_ = 0;
  std::shared_ptr<MapUtil<Dim>> map_util_;
  std::shared_ptr<env_map_gpu<Dim>> gpu_env_;
  Vecf<Dim> potential_radius_, potential_map_range_, search_radius_;
  decimal_t pow_{1.0};  // map_planner.h:113
};
typedef MapPlanner<2> OccMapPlanner;
typedef MapPlanner<3> VoxelMapPlanner;

/// CPUs this process may actually use: the hardware thread count capped by the cgroup v2 CPU quota
/// (/sys/fs/cgroup/cpu.max = "<quota> <period>"); oversubscribing a quota only adds throttling.
inline int effective_cpus() {
  int n = (int)std::thread::hardware_concurrency();
  if (n < 1) n = 1;
  if (FILE *f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {
    char q[64];
    long period = 0;
    if (std::fscanf(f, "%63s %ld", q, &period) == 2 && period > 0 && q[0] != 'm') {
      const long quota = std::atol(q);
      const int cap = (int)((quota + period - 1) / period);
      if (cap >= 1 && cap < n) n = cap;
    }
    std::fclose(f);
  }
  return n;
}

/// Minimal persistent worker pool for the host side of the lock-step driver: the per-query
/// bookkeeping (hash map, heap) of different queries is independent, so it is spread over the host
/// cores while the device expands the next batch's nodes.
class WorkerPool {
 public:
  explicit WorkerPool(int n_threads) {
    n_ = n_threads < 1 ? 1 : n_threads;
    for (int t = 1; t < n_; t++) th_.emplace_back([this] { loop(); });
  }
  ~WorkerPool() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
      gen_++;
    }
    cv_.notify_all();
    for (auto &t : th_) t.join();
  }
  int size() const { return n_; }
  /// fn(i) for i in [0, count), dynamically chunked; returns when all are done
  void run(std::size_t count, const std::function<void(std::size_t)> &fn) {
    if (count == 0) return;
    if (n_ == 1 || count < 64) {
      for (std::size_t i = 0; i < count; i++) fn(i);
      return;
    }
    {
      std::lock_guard<std::mutex> lk(m_);
      fn_ = &fn;
      count_ = count;
      next_.store(0);
      pending_ = n_ - 1;
      gen_++;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(m_);
    done_.wait(lk, [this] { return pending_ == 0; });
  }

 private:
  void work() {
    const std::size_t chunk = 16;
    for (;;) {
      const std::size_t b = next_.fetch_add(chunk);
      if (b >= count_) break;
      const std::size_t e = b + chunk < count_ ? b + chunk : count_;
      for (std::size_t i = b; i < e; i++) (*fn_)(i);
    }
  }
  void loop() {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return gen_ != seen; });
        seen = gen_;
        if (stop_) return;
      }
      work();
      {
        std::lock_guard<std::mutex> lk(m_);
        if (--pending_ == 0) done_.notify_one();
      }
    }
  }
  int n_;
  std::vector<std::thread> th_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(std::size_t)> *fn_ = nullptr;
  std::size_t count_ = 0;
  std::atomic<std::size_t> next_{0};
  int pending_ = 0;
  unsigned long gen_ = 0;
  bool stop_ = false;
};

/// Lock-step batched A* over many independent (start, goal) queries on one map — BASELINE.json
/// config 5.  Every iteration pops the best open node of each live query (AstarStepper::pop) and
/// expands all of them in ONE device launch (env_map_gpu::expand_packed); each query then relaxes
/// its own successors.  Per query the sequence of expanded nodes is exactly the one
/// MapPlanner::plan() produces for that query alone.
template <int Dim>
class MultiQueryPlanner {
 public:
  struct Result {
    bool valid = false;
    decimal_t cost = std::numeric_limits<decimal_t>::infinity();
    int expanded = 0;
    std::size_t n_closed = 0;
    std::vector<int> actions;
  };
  explicit MultiQueryPlanner(const std::shared_ptr<MapUtil<Dim>> &map_util, int device = 0)
      : map_util_(map_util), gpu_(new env_map_gpu<Dim>(map_util, device)) {}
  /// the shared env: set U, limits, weights, control, tolerances on it
  env_map_gpu<Dim> &env() { return *gpu_; }
  long iterations() const { return iterations_; }
  /// seconds spent in the three phases of the last plan(): pop, device expansion (incl. PCIe), relax
  double t_pop() const { return t_pop_; }
  double t_device() const { return t_dev_; }
  double t_relax() const { return t_relax_; }
  long nodes_expanded() const { return nodes_; }

  /// host threads used for the per-query bookkeeping (default: all cores)
  void setHostThreads(int n) { host_threads_ = n; }
  /// keys-only result stream for occupancy planning (default on; see env_map_gpu::expand_packed)
  void setKeysOnly(bool on) { keys_only_ = on; }

  std::vector<Result> plan(const vec_E<Waypoint<Dim>> &starts, const vec_E<Waypoint<Dim>> &goals, decimal_t eps,
                           int max_expand) {
    // the search states of the previous plan() are recycled, not freed: a planner that answers batch
    // after batch allocates (and page-faults) its state memory once
    if (ss_.size() != starts.size()) release();
    WorkerPool pool(host_threads_ > 0 ? host_threads_ : effective_cpus());
    const std::size_t Q = starts.size();
    auto &envs = envs_;
    auto &ss = ss_;
    auto &st = st_;
    envs.resize(Q); ss.resize(Q); st.resize(Q);
    std::vector<Result> res(Q);
    pool.run(Q, [&](std::size_t q) {
      if (!envs[q]) envs[q].reset(new QueryEnv(map_util_));
      QueryEnv &e = *envs[q];
      // every env_base parameter the goal test and the heuristic read (the expansion itself runs on gpu_)
      e.w_ = gpu_->w_; e.wyaw_ = gpu_->wyaw_; e.v_max_ = gpu_->v_max_; e.a_max_ = gpu_->a_max_; e.j_max_ = gpu_->j_max_;
      e.yaw_max_ = gpu_->yaw_max_; e.dt_ = gpu_->dt_; e.t_max_ = gpu_->t_max_;
      e.tol_pos_ = gpu_->tol_pos_; e.tol_vel_ = gpu_->tol_vel_; e.tol_acc_ = gpu_->tol_acc_; e.tol_yaw_ = gpu_->tol_yaw_;
      e.set_goal(goals[q]);
      if (ss[q]) ss[q]->reset(eps);
      else ss[q].reset(new StateSpace<Dim>(eps));
      st[q].reset(new AstarStepper<Dim>(&e, ss[q], max_expand));
      if (e.is_free(starts[q].pos)) st[q]->start(starts[q]);  // planner_base.h:283-287
    });
    std::vector<mplx_waypoint> batch;
    std::vector<std::size_t> who;
    iterations_ = nodes_ = 0;
    t_pop_ = t_dev_ = t_relax_ = 0;
    for (;;) {
      // pop phase: one node per live query (independent heaps -> parallel), then compact
      who.clear();
      for (std::size_t q = 0; q < Q; q++)
        if (st[q]->active()) who.push_back(q);
      if (who.empty()) break;
      batch.resize(who.size());
      auto t0 = std::chrono::steady_clock::now();
      pool.run(who.size(), [&](std::size_t b) { batch[b] = env_map_gpu<Dim>::pod(st[who[b]]->pop()); });
      auto t1 = std::chrono::steady_clock::now();
      const bool keys_only = keys_only_ && gpu_->keys_only_possible();
      if (keys_only) gpu_->action_cost(0);  // build the table before the parallel phase
      gpu_->expand_packed(batch, keys_only);
      auto t2 = std::chrono::steady_clock::now();
      iterations_++;
      nodes_ += (long)batch.size();
      // relax phase: every query consumes its own successors (independent state spaces -> parallel)
      pool.run(who.size(), [&](std::size_t b) {
        const std::size_t r0 = (std::size_t)gpu_->p_offset[b];
        const int cnt = gpu_->p_count[b];
        int act[kMaxSucc];
        for (int j = 0; j < cnt; j++) act[j] = gpu_->p_action[r0 + j];
        if (keys_only) {
          decimal_t cst[kMaxSucc];
          for (int j = 0; j < cnt; j++) cst[j] = gpu_->action_cost(act[j]);
          st[who[b]]->consume(cnt, [&](int s) { return gpu_->forward_from_pod(batch[b], act[s]); }, cst, act,
                              [&](int s) { return (std::size_t)gpu_->p_key[r0 + s]; });
        } else {
          st[who[b]]->consume(cnt, [&](int s) { return gpu_->packed_waypoint(r0 + s, batch[b]); },
                              gpu_->p_cost.data() + r0, act, [&](int s) { return (std::size_t)gpu_->p_key[r0 + s]; });
        }
      });
      auto t3 = std::chrono::steady_clock::now();
      t_pop_ += std::chrono::duration<double>(t1 - t0).count();
      t_dev_ += std::chrono::duration<double>(t2 - t1).count();
      t_relax_ += std::chrono::duration<double>(t3 - t2).count();
    }
    // trace back and count on the pool; the state spaces stay until release() / the next plan()
    pool.run(Q, [&](std::size_t q) {
      std::vector<Edge<Dim>> traj;
      res[q].cost = st[q]->finish(traj);
      res[q].valid = !std::isinf(res[q].cost);
      res[q].expanded = st[q]->expanded();
      for (const auto &e : traj) res[q].actions.push_back(e.action_id);
      for (const auto *stt : ss[q]->order_) if (stt->iterationclosed) res[q].n_closed++;
    });
    return res;
  }
  /// Free the search states of the last plan() (tens of millions of states for a large batch), on
  /// the host cores.  Called by the next plan() and the destructor.
  void release() {
    if (ss_.empty()) return;
    WorkerPool pool(host_threads_ > 0 ? host_threads_ : effective_cpus());
    pool.run(ss_.size(), [&](std::size_t q) {
      st_[q].reset();
      ss_[q].reset();
      envs_[q].reset();
    });
    st_.clear(); ss_.clear(); envs_.clear();
  }
  ~MultiQueryPlanner() { release(); }

 private:
  // per-query host env: goal test + heuristic only (its get_succ is never called)
  struct QueryEnv : env_map_host<Dim> {
    using env_map_host<Dim>::env_map_host;
    void get_succ(const Waypoint<Dim> &, vec_E<Waypoint<Dim>> &, std::vector<decimal_t> &, std::vector<int> &) const override {}
  };
  std::vector<std::unique_ptr<QueryEnv>> envs_;
  std::vector<std::shared_ptr<StateSpace<Dim>>> ss_;
  std::vector<std::unique_ptr<AstarStepper<Dim>>> st_;
  std::shared_ptr<MapUtil<Dim>> map_util_;
  std::unique_ptr<env_map_gpu<Dim>> gpu_;
  long iterations_ = 0, nodes_ = 0;
  double t_pop_ = 0, t_dev_ = 0, t_relax_ = 0;
  int host_threads_ = 0;
  bool keys_only_ = true;
  static constexpr int kMaxSucc = 1024;  // |U| upper bound of libmplx
};
}  // namespace MPL
