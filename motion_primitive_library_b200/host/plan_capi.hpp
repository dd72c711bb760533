// plan_capi.hpp — a flat C struct interface to one MPL::MapPlanner<Dim>::plan() call, shared by
// libmpl_host.so (GPU env) and the test harness (same planner, CPU checker env).
#pragma once
#include "mpl_host.hpp"

#include "plan_capi_types.h"

namespace mplh {
template <int Dim>
Waypoint<Dim> wp_from(const mplx_waypoint &p, int control) {
  Waypoint<Dim> w(control);
  for (int d = 0; d < Dim; d++) { w.pos(d) = p.pos[d]; w.vel(d) = p.vel[d]; w.acc(d) = p.acc[d]; w.jrk(d) = p.jrk[d]; }
  w.yaw = p.yaw; w.t = p.t;
  return w;
}
template <int Dim>
std::shared_ptr<MPL::MapUtil<Dim>> make_map(const mplh_plan_args *a) {
  auto mu = std::make_shared<MPL::MapUtil<Dim>>();
  Vecf<Dim> ori; Veci<Dim> dim; size_t n = 1;
  for (int k = 0; k < Dim; k++) { ori(k) = a->origin[k]; dim(k) = a->mdim[k]; n *= (size_t)a->mdim[k]; }
  mu->setMap(ori, dim, MPL::Tmap(a->map, a->map + n), a->res);
  return mu;
}
// Configure `planner` (whose env is already installed) from the flat args, run plan(), and
// export the closed set (sorted lattice keys) and the trajectory's action ids.
template <int Dim>
void run(MPL::MapPlanner<Dim> &planner, const mplh_plan_args *a, mplh_plan_result *r, uint64_t *closed_keys,
         int cap_closed, int32_t *actions, int cap_actions) {
  vec_E<VecDf> U;
  for (int i = 0; i < a->nU; i++) U.push_back(VecDf(a->U + (size_t)i * a->udim, a->U + (size_t)(i + 1) * a->udim));
  planner.setU(U);
  planner.setVmax(a->v_max); planner.setAmax(a->a_max); planner.setJmax(a->j_max); planner.setYawmax(a->yaw_max);
  planner.setDt(a->T); planner.setW(a->w); planner.setWyaw(a->wyaw); planner.setEpsilon(a->eps);
  planner.setTol(a->tol_pos, a->tol_vel, a->tol_acc);
  planner.setMaxNum(a->max_num);
  const Waypoint<Dim> start = wp_from<Dim>(a->start, a->control), goal = wp_from<Dim>(a->goal, a->control);
  auto t0 = std::chrono::steady_clock::now();
  r->valid = planner.plan(start, goal) ? 1 : 0;
  r->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  r->cost = planner.getTrajCost();
  r->expanded = planner.getExpandedNum();
  std::vector<uint64_t> keys;
  if (planner.initialized())
    for (const auto *s : planner.getCloseSetStates()) keys.push_back((uint64_t)s->key);
  std::sort(keys.begin(), keys.end());
  r->n_closed = (int)keys.size();
  r->n_open = planner.initialized() ? (int)planner.getOpenSetSize() : 0;
  for (int i = 0; i < (int)keys.size() && i < cap_closed; i++) closed_keys[i] = keys[i];
  const auto traj = planner.getTraj();
  r->n_actions = (int)traj.size();
  for (int i = 0; i < (int)traj.size() && i < cap_actions; i++) actions[i] = traj[i].action_id;
}
}  // namespace mplh
