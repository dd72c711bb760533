// ref_b200_driver.cpp — the reference's planner (unmodified headers + src/mpl_planner/map_planner.cpp,
// compiled where they lie) with integration/env_map_b200.h installed through the virtual
// MapPlanner::setMapUtil.  Same flat C interface as oracle/ref_planner_driver.cpp (refp_plan), so the
// test compares the two planners field by field.  TEST INFRASTRUCTURE for the drop-in boundary.
#include <mpl_planner/planner/map_planner.h>

#include <algorithm>
#include <chrono>

#include "../motion_primitive_library_b200/host/plan_capi_types.h"
#include "env_map_b200.h"

namespace {
template <int Dim>
struct Planner : MPL::MapPlannerB200<Dim> {
  explicit Planner(bool v, int device) : MPL::MapPlannerB200<Dim>(v, device) {}
  using MPL::MapPlanner<Dim>::ss_ptr_;
};

template <int Dim>
int plan(const mplh_plan_args *a, mplh_plan_result *r, uint64_t *closed_keys, int cap_closed, int32_t *actions,
         int cap_actions) {
  Planner<Dim> planner(false, a->device);
  std::shared_ptr<MPL::MapUtil<Dim>> mu(new MPL::MapUtil<Dim>);
  Vecf<Dim> ori;
  Veci<Dim> dim;
  size_t n = 1;
  for (int k = 0; k < Dim; k++) {
    ori(k) = a->origin[k];
    dim(k) = a->mdim[k];
    n *= (size_t)a->mdim[k];
  }
  mu->setMap(ori, dim, MPL::Tmap(a->map, a->map + n), a->res);
  planner.setMapUtil(mu);  // -> env_map_b200
  vec_E<VecDf> U;
  for (int i = 0; i < a->nU; i++) {
    VecDf u(a->udim);
    for (int k = 0; k < a->udim; k++) u(k) = a->U[(size_t)i * a->udim + k];
    U.push_back(u);
  }
  planner.setU(U);
  planner.setVmax(a->v_max);
  planner.setAmax(a->a_max);
  planner.setJmax(a->j_max);
  planner.setYawmax(a->yaw_max);
  planner.setDt(a->T);
  planner.setW(a->w);
  planner.setWyaw(a->wyaw);
  planner.setEpsilon(a->eps);
  planner.setTol(a->tol_pos, a->tol_vel, a->tol_acc);
  planner.setMaxNum(a->max_num);
  Waypoint<Dim> start((Control::Control)a->control), goal((Control::Control)a->control);
  for (int d = 0; d < Dim; d++) {
    start.pos(d) = a->start.pos[d]; start.vel(d) = a->start.vel[d]; start.acc(d) = a->start.acc[d]; start.jrk(d) = a->start.jrk[d];
    goal.pos(d) = a->goal.pos[d]; goal.vel(d) = a->goal.vel[d]; goal.acc(d) = a->goal.acc[d]; goal.jrk(d) = a->goal.jrk[d];
  }
  start.yaw = a->start.yaw;
  goal.yaw = a->goal.yaw;
  auto t0 = std::chrono::steady_clock::now();
  r->valid = planner.plan(start, goal) ? 1 : 0;
  r->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  r->cost = planner.getTrajCost();
  r->expanded = planner.initialized() ? planner.getExpandedNum() : 0;
  std::vector<uint64_t> keys;
  int n_open = 0;
  if (planner.initialized()) {
    for (const auto &it : planner.ss_ptr_->hm_) {
      if (!it.second) continue;
      if (it.second->iterationclosed) keys.push_back((uint64_t)hash_value(it.second->coord));
      else if (it.second->iterationopened) n_open++;
    }
  }
  std::sort(keys.begin(), keys.end());
  r->n_closed = (int)keys.size();
  r->n_open = n_open;
  for (int i = 0; i < (int)keys.size() && i < cap_closed; i++) closed_keys[i] = keys[i];
  const auto prs = planner.getTraj().getPrimitives();
  r->n_actions = (int)prs.size();
  const int order = __builtin_popcount(a->control & 15);
  for (int i = 0; i < (int)prs.size() && i < cap_actions; i++) {
    int found = -1;
    for (int u = 0; u < a->nU && found < 0; u++) {
      bool same = true;
      for (int d = 0; d < Dim; d++) same = same && prs[i].pr(d).coeff()(5 - order) == a->U[(size_t)u * a->udim + d];
      if (same && (a->control & 16)) same = prs[i].pr_yaw().coeff()(4) == a->U[(size_t)u * a->udim + Dim];
      if (same) found = u;
    }
    actions[i] = found;
  }
  return 0;
}
thread_local char g_err[512] = "";
}  // namespace

extern "C" {
const char *refb_last_error(void) { return g_err; }
int refb_plan(const mplh_plan_args *a, mplh_plan_result *r, uint64_t *closed_keys, int cap_closed, int32_t *actions,
              int cap_actions) {
  *r = mplh_plan_result{};
  try {
    return a->dim == 2 ? plan<2>(a, r, closed_keys, cap_closed, actions, cap_actions)
                       : plan<3>(a, r, closed_keys, cap_closed, actions, cap_actions);
  } catch (const std::exception &e) {
    snprintf(g_err, sizeof g_err, "%s", e.what());
    return 1;
  }
}
}
