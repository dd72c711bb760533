// ref_driver.cpp — links the UNMODIFIED reference headers (from /root/reference/include,
// compiled where they lie) behind the oracle's C ABI, so that the restatement in
// mpl_oracle.cpp can be checked against the reference's own code.  TEST INFRASTRUCTURE.
// Build: oracle/Makefile target `ref` -> oracle/_ref/libmplref.so (git-ignored).
#include <mpl_planner/env/env_map.h>

#include <chrono>
#include <thread>

#include "mpl_oracle.h"

namespace {
template <int Dim>
struct Ref {
  std::shared_ptr<MPL::MapUtil<Dim>> mu;
  std::shared_ptr<MPL::env_map<Dim>> env;
  int control;

  static std::shared_ptr<MPL::MapUtil<Dim>> make_map(const orc_env *e) {
    std::shared_ptr<MPL::MapUtil<Dim>> m(new MPL::MapUtil<Dim>);
    Vecf<Dim> ori;
    Veci<Dim> dim;
    size_t n = 1;
    for (int k = 0; k < Dim; k++) {
      ori(k) = e->origin[k];
      dim(k) = e->mdim[k];
      n *= (size_t)e->mdim[k];
    }
    MPL::Tmap data(e->map, e->map + n);
    m->setMap(ori, dim, data, e->res);
    return m;
  }

  // The env holds the MapUtil by shared_ptr (env_map.h:288): worker threads share one read-only
  // grid, exactly as several planners sharing a map_util would in the reference.
  explicit Ref(const orc_env *e, std::shared_ptr<MPL::MapUtil<Dim>> shared = nullptr) : control(e->control) {
    mu = shared ? shared : make_map(e);
    size_t n = 1;
    for (int k = 0; k < Dim; k++) n *= (size_t)e->mdim[k];
    env.reset(new MPL::env_map<Dim>(mu));
    vec_E<VecDf> U;
    for (int i = 0; i < e->nU; i++) {
      VecDf u(e->udim);
      for (int k = 0; k < e->udim; k++) u(k) = e->U[(size_t)i * e->udim + k];
      U.push_back(u);
    }
    env->set_u(U);
    env->set_dt(e->T);
    env->set_w(e->w);
    env->set_wyaw(e->wyaw);
    env->set_v_max(e->v_max);
    env->set_a_max(e->a_max);
    env->set_j_max(e->j_max);
    env->set_yaw_max(e->yaw_max);
    if (e->potential) {
      env->set_potential_map(std::vector<int8_t>(e->potential, e->potential + n));
      env->set_potential_weight(e->potential_weight);
      env->set_gradient_weight(e->gradient_weight);
    }
    if (e->region) {
      std::vector<bool> r(n);
      for (size_t i = 0; i < n; i++) r[i] = e->region[i] != 0;
      env->set_search_region(r);
    }
  }

  int get_succ(const orc_waypoint *c, orc_waypoint *succ, double *cost, int32_t *action, uint64_t *key) {
    Waypoint<Dim> curr((Control::Control)control);
    for (int k = 0; k < Dim; k++) {
      curr.pos(k) = c->pos[k];
      curr.vel(k) = c->vel[k];
      curr.acc(k) = c->acc[k];
      curr.jrk(k) = c->jrk[k];
    }
    curr.yaw = c->yaw;
    curr.t = c->t;
    vec_E<Waypoint<Dim>> s;
    std::vector<decimal_t> sc;
    std::vector<int> sa;
    env->get_succ(curr, s, sc, sa);
    env->expanded_nodes_.clear();  // debug collections grow without bound (env_map.h:154,166)
    env->expanded_edges_.clear();
    for (size_t i = 0; i < s.size(); i++) {
      orc_waypoint &o = succ[i];
      for (int k = 0; k < 3; k++) {
        o.pos[k] = k < Dim ? s[i].pos(k) : 0;
        o.vel[k] = k < Dim ? s[i].vel(k) : 0;
        o.acc[k] = k < Dim ? s[i].acc(k) : 0;
        o.jrk[k] = k < Dim ? s[i].jrk(k) : 0;
      }
      o.yaw = s[i].yaw;
      o.t = s[i].t;
      cost[i] = sc[i];
      action[i] = sa[i];
      if (key) key[i] = hash_value(s[i]);
    }
    return (int)s.size();
  }
};

template <int Dim>
int batch(const orc_env *e, const orc_waypoint *nodes, int n, orc_waypoint *succ, double *cost, int32_t *action,
          uint64_t *key, int32_t *count, int nthreads) {
  auto shared = Ref<Dim>::make_map(e);
  auto work = [&](int lo, int hi) {
    Ref<Dim> r(e, shared);  // one env per thread: get_succ is not re-entrant (env_base.h:402-404)
    for (int i = lo; i < hi; i++) {
      size_t o = (size_t)i * e->nU;
      count[i] = r.get_succ(&nodes[i], succ + o, cost + o, action + o, key ? key + o : nullptr);
    }
  };
  if (nthreads <= 1) {
    work(0, n);
    return 0;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++)
    th.emplace_back(work, (int)((int64_t)n * t / nthreads), (int)((int64_t)n * (t + 1) / nthreads));
  for (auto &t : th) t.join();
  return 0;
}

template <int Dim>
int timed(const orc_env *e, const orc_waypoint *nodes, int n, int nthreads, int64_t *total_succ, double *seconds) {
  if (nthreads < 1) nthreads = 1;
  std::vector<std::unique_ptr<Ref<Dim>>> envs;
  auto shared = Ref<Dim>::make_map(e);
  for (int t = 0; t < nthreads; t++) envs.emplace_back(new Ref<Dim>(e, shared));  // set-up outside the clock
  std::vector<int64_t> ns(nthreads, 0);
  auto work = [&](int t, int lo, int hi) {
    std::vector<orc_waypoint> succ(e->nU);
    std::vector<double> cost(e->nU);
    std::vector<int32_t> act(e->nU);
    int64_t s = 0;
    for (int i = lo; i < hi; i++) s += envs[t]->get_succ(&nodes[i], succ.data(), cost.data(), act.data(), nullptr);
    ns[t] = s;
  };
  auto t0 = std::chrono::steady_clock::now();
  if (nthreads == 1)
    work(0, 0, n);
  else {
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++)
      th.emplace_back(work, t, (int)((int64_t)n * t / nthreads), (int)((int64_t)n * (t + 1) / nthreads));
    for (auto &t : th) t.join();
  }
  auto t1 = std::chrono::steady_clock::now();
  int64_t a = 0;
  for (auto v : ns) a += v;
  if (total_succ) *total_succ = a;
  if (seconds) *seconds = std::chrono::duration<double>(t1 - t0).count();
  return 0;
}

template <int Dim>
int edges_free(const orc_env *e, const orc_waypoint *parents, const int32_t *actions, int n, uint8_t *out_free,
               double *out_cost) {
  Ref<Dim> r(e);
  for (int i = 0; i < n; i++) {
    Waypoint<Dim> curr((Control::Control)e->control);
    for (int k = 0; k < Dim; k++) {
      curr.pos(k) = parents[i].pos[k];
      curr.vel(k) = parents[i].vel[k];
      curr.acc(k) = parents[i].acc[k];
      curr.jrk(k) = parents[i].jrk[k];
    }
    curr.yaw = parents[i].yaw;
    curr.t = parents[i].t;
    Primitive<Dim> pr;
    r.env->forward_action(curr, actions[i], pr);
    out_free[i] = r.env->is_free(pr) ? 1 : 0;
    if (out_cost) out_cost[i] = r.env->calculate_intrinsic_cost(pr);
  }
  return 0;
}
}  // namespace

extern "C" {
int ref_expand_batch(const orc_env *e, const orc_waypoint *nodes, int n, orc_waypoint *succ, double *cost,
                     int32_t *action, uint64_t *key, int32_t *count, int nthreads) {
  return e->dim == 2 ? batch<2>(e, nodes, n, succ, cost, action, key, count, nthreads)
                     : batch<3>(e, nodes, n, succ, cost, action, key, count, nthreads);
}
int ref_expand_batch_timed(const orc_env *e, const orc_waypoint *nodes, int n, int nthreads, int64_t *total_succ,
                           double *seconds) {
  return e->dim == 2 ? timed<2>(e, nodes, n, nthreads, total_succ, seconds)
                     : timed<3>(e, nodes, n, nthreads, total_succ, seconds);
}
// env_map::is_free(pr) and calculate_intrinsic_cost(pr) of the reference for stored edges
// (parent, action): pr built by env_base::forward_action (env_base.h:228-231).
int ref_edges_is_free(const orc_env *e, const orc_waypoint *parents, const int32_t *actions, int n,
                      uint8_t *out_free, double *out_cost) {
  return e->dim == 2 ? edges_free<2>(e, parents, actions, n, out_free, out_cost)
                     : edges_free<3>(e, parents, actions, n, out_free, out_cost);
}
const char *ref_info(void) {
  return "unmodified /root/reference/include headers (env_map.h, env_base.h, primitive.h, waypoint.h, math.h, "
         "map_util.h) + oracle/shim Eigen/Boost stand-ins; g++ -O2 -std=c++11 (reference CMakeLists.txt:5-8)";
}
}
