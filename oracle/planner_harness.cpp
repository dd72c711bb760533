// planner_harness.cpp — TEST INFRASTRUCTURE.  The product's host planner (mpl_host.hpp:
// StateSpace, GraphSearch::Astar, MapPlanner::plan) driven by a CPU env whose get_succ is the
// ORACLE (orc_get_succ).  tests/test_planner_e2e_gpu.py runs the same plan through
// libmpl_host.so (GPU env) and requires identical closed sets, costs and action sequences.
#include "../motion_primitive_library_b200/host/plan_capi.hpp"
#include "mpl_oracle.h"

namespace {
template <int Dim>
class env_map_oracle : public MPL::env_map_host<Dim> {
 public:
  env_map_oracle(std::shared_ptr<MPL::MapUtil<Dim>> mu, const mplh_plan_args *a) : MPL::env_map_host<Dim>(mu), a_(a) {}
  void get_succ(const Waypoint<Dim> &curr, vec_E<Waypoint<Dim>> &succ, std::vector<decimal_t> &succ_cost,
                std::vector<int> &action_idx) const override {
    succ.clear(); succ_cost.clear(); action_idx.clear();
    this->expanded_nodes_.push_back(curr.pos);
    orc_env e{};
    e.dim = Dim; e.control = a_->control; e.T = this->dt_; e.w = this->w_; e.wyaw = this->wyaw_;
    e.v_max = this->v_max_; e.a_max = this->a_max_; e.j_max = this->j_max_; e.yaw_max = this->yaw_max_;
    e.nU = a_->nU; e.udim = a_->udim; e.U = a_->U;
    for (int k = 0; k < 3; k++) { e.mdim[k] = k < Dim ? a_->mdim[k] : 1; e.origin[k] = k < Dim ? a_->origin[k] : 0; }
    e.res = a_->res; e.map = a_->map; e.potential = a_->potential;
    e.potential_weight = a_->potential_weight; e.gradient_weight = a_->gradient_weight; e.region = nullptr;
    orc_waypoint c{};
    for (int d = 0; d < Dim; d++) { c.pos[d] = curr.pos(d); c.vel[d] = curr.vel(d); c.acc[d] = curr.acc(d); c.jrk[d] = curr.jrk(d); }
    c.yaw = curr.yaw; c.t = curr.t;
    std::vector<orc_waypoint> s(e.nU); std::vector<double> cost(e.nU); std::vector<int32_t> act(e.nU);
    const int n = orc_get_succ(&e, &c, s.data(), cost.data(), act.data(), nullptr, nullptr);
    for (int j = 0; j < n; j++) {
      Waypoint<Dim> w(curr.control);
      for (int d = 0; d < Dim; d++) { w.pos(d) = s[j].pos[d]; w.vel(d) = s[j].vel[d]; w.acc(d) = s[j].acc[d]; w.jrk(d) = s[j].jrk[d]; }
      w.yaw = s[j].yaw; w.t = s[j].t;
      succ.push_back(w); succ_cost.push_back(cost[j]); action_idx.push_back(act[j]);
    }
  }
 private:
  const mplh_plan_args *a_;
};
}  // namespace

extern "C" int orcp_plan(const mplh_plan_args *a, mplh_plan_result *r, uint64_t *closed_keys, int cap_closed,
                         int32_t *actions, int cap_actions) {
  *r = mplh_plan_result{};
  auto go = [&](auto dimtag) {
    constexpr int Dim = decltype(dimtag)::value;
    MPL::MapPlanner<Dim> planner(false);
    planner.setEnv(std::make_shared<env_map_oracle<Dim>>(mplh::make_map<Dim>(a), a));
    mplh::run<Dim>(planner, a, r, closed_keys, cap_closed, actions, cap_actions);
  };
  if (a->dim == 2) go(std::integral_constant<int, 2>());
  else go(std::integral_constant<int, 3>());
  return 0;
}
