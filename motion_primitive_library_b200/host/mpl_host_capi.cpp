// libmpl_host.so — C entry point that runs MPL::MapPlanner<Dim>::plan() with the GPU env
// (env_map_gpu -> libmplx).  Used by the Python tests and tools; C++ users include mpl_host.hpp.
#include <cstdlib>

#include "plan_capi.hpp"

static thread_local std::string g_err;

extern "C" {
const char *mplh_last_error(void) { return g_err.c_str(); }

int mplh_plan(const mplh_plan_args *a, mplh_plan_result *r, uint64_t *closed_keys, int cap_closed, int32_t *actions,
              int cap_actions) {
  try {
    *r = mplh_plan_result{};
    auto go = [&](auto dimtag) {
      constexpr int Dim = decltype(dimtag)::value;
      MPL::MapPlanner<Dim> planner(false);
      planner.setMapUtil(mplh::make_map<Dim>(a), a->device);  // installs env_map_gpu (map_planner.cpp:14-18)
      planner.setControl(a->control);
      planner.setSpeculation(a->speculate);
      if (a->potential) {
        size_t n = 1;
        for (int k = 0; k < Dim; k++) n *= (size_t)a->mdim[k];
        planner.gpu_env()->set_potential_map(std::vector<int8_t>(a->potential, a->potential + n));
        planner.setPotentialWeight(a->potential_weight);
        planner.setGradientWeight(a->gradient_weight);
      }
      mplh::run<Dim>(planner, a, r, closed_keys, cap_closed, actions, cap_actions);
      r->gpu_nodes = planner.gpu_env()->stats_nodes();
      r->gpu_calls = planner.gpu_env()->stats_calls();
      r->gpu_launches = planner.gpu_env()->launches();
    };
    if (a->dim == 2) go(std::integral_constant<int, 2>());
    else if (a->dim == 3) go(std::integral_constant<int, 3>());
    else { g_err = "dim must be 2 or 3"; return 1; }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 2;
  }
}

/* MapPlanner::plan() with the GPU env, then the recovered Trajectory (include/mpl_basis/trajectory.h):
 * sample(N), getTotalTime / J / Jyaw, getWaypoints and evaluate(t); layouts in plan_capi.hpp. */
int mplh_plan_trajectory(const mplh_plan_args *a, int N, mplh_plan_result *r, double *samples, double *totals,
                         double *waypoints, int cap_wp, int32_t *n_wp, double *mids) {
  try {
    *r = mplh_plan_result{};
    auto go = [&](auto dimtag) {
      constexpr int Dim = decltype(dimtag)::value;
      MPL::MapPlanner<Dim> planner(false);
      planner.setMapUtil(mplh::make_map<Dim>(a), a->device);
      planner.setControl(a->control);
      planner.setSpeculation(a->speculate);
      mplh::run_trajectory<Dim>(planner, a, N, r, samples, totals, waypoints, cap_wp, n_wp, mids);
    };
    if (a->dim == 2) go(std::integral_constant<int, 2>());
    else if (a->dim == 3) go(std::integral_constant<int, 3>());
    else { g_err = "dim must be 2 or 3"; return 1; }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 2;
  }
}

/* MapPlanner::plan() followed by MapPlanner::iterativePlan() (map_planner.cpp:393-433) with the GPU
 * env: every iteration builds the tunnel around the previous trajectory on the device
 * (mplx_set_search_region_path) and replans inside it.  info[0] = plan() calls made by
 * iterativePlan, info[1] = its return value. */
int mplh_iterative_plan(const mplh_plan_args *a, const double *search_radius, int max_iter, mplh_plan_result *first,
                        mplh_plan_result *last, int32_t *info, uint64_t *closed_keys, int cap_closed, int32_t *actions,
                        int cap_actions) {
  try {
    *first = mplh_plan_result{};
    *last = mplh_plan_result{};
    auto go = [&](auto dimtag) {
      constexpr int Dim = decltype(dimtag)::value;
      MPL::MapPlanner<Dim> planner(false);
      planner.setMapUtil(mplh::make_map<Dim>(a), a->device);
      planner.setControl(a->control);
      planner.setSpeculation(a->speculate);
      if (a->potential) {
        size_t n = 1;
        for (int k = 0; k < Dim; k++) n *= (size_t)a->mdim[k];
        planner.gpu_env()->set_potential_map(std::vector<int8_t>(a->potential, a->potential + n));
        planner.setPotentialWeight(a->potential_weight);
        planner.setGradientWeight(a->gradient_weight);
      }
      mplh::run_iterative<Dim>(planner, a, search_radius, max_iter, first, last, info, closed_keys, cap_closed, actions,
                               cap_actions);
    };
    if (a->dim == 2) go(std::integral_constant<int, 2>());
    else if (a->dim == 3) go(std::integral_constant<int, 3>());
    else { g_err = "dim must be 2 or 3"; return 1; }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 2;
  }
}

/* A scripted LPA* session (plan_capi_types.h: PLAN / LINK / BLOCK / CLEAR / SUBTREE steps) on one
 * MapPlanner with the GPU env: get_succ, the getLinkedNodes voxel walk and the is_free(pr)
 * re-validation of decreaseCost all run on the device.  outs has n_steps entries; the action ids of
 * the trajectory found by PLAN step k go to actions[k*cap_actions ...]. */
int mplh_lpa_run(const mplh_plan_args *a, const mplh_lpa_step *steps, int n_steps, mplh_lpa_out *outs,
                 int32_t *actions, int cap_actions) {
  try {
    auto go = [&](auto dimtag) {
      constexpr int Dim = decltype(dimtag)::value;
      MPL::MapPlanner<Dim> planner(false);
      auto mu = mplh::make_map<Dim>(a);
      planner.setMapUtil(mu, a->device);
      planner.setControl(a->control);
      planner.setSpeculation(a->speculate);
      mplh::run_lpa<Dim>(planner, mu, a, steps, n_steps, outs, actions, cap_actions);
    };
    if (a->dim == 2) go(std::integral_constant<int, 2>());
    else if (a->dim == 3) go(std::integral_constant<int, 3>());
    else { g_err = "dim must be 2 or 3"; return 1; }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 2;
  }
}

/* Lock-step batched A* over n_q (start, goal) pairs on the map/params of `a` (a->start/goal are
 * ignored).  totals[0] = lock-step iterations (= device launches of the expansion kernel),
 * totals[1] = nodes expanded over all queries, totals[2] = wall seconds of the search,
 * totals[3..5] = seconds in the pop / device expansion (incl. PCIe) / relax phases,
 * totals[6] = seconds spent freeing the search states afterwards (7 doubles). */
int mplh_plan_batch(const mplh_plan_args *a, const mplx_waypoint *starts, const mplx_waypoint *goals, int n_q,
                    mplh_query_result *out, double *totals) {
  try {
    auto go = [&](auto dimtag) {
      constexpr int Dim = decltype(dimtag)::value;
      MPL::MultiQueryPlanner<Dim> mq(mplh::make_map<Dim>(a), a->device);
      auto &e = mq.env();
      if (const char *t = std::getenv("MPLH_THREADS")) mq.setHostThreads(std::atoi(t));
      vec_E<VecDf> U;
      for (int i = 0; i < a->nU; i++) U.push_back(VecDf(a->U + (size_t)i * a->udim, a->U + (size_t)(i + 1) * a->udim));
      e.set_u(U); e.set_control(a->control);
      e.set_v_max(a->v_max); e.set_a_max(a->a_max); e.set_j_max(a->j_max); e.set_yaw_max(a->yaw_max);
      e.set_dt(a->T); e.set_w(a->w); e.set_wyaw(a->wyaw);
      e.set_tol_pos(a->tol_pos); e.set_tol_vel(a->tol_vel); e.set_tol_acc(a->tol_acc);
      if (a->potential) {
        size_t n = 1;
        for (int k = 0; k < Dim; k++) n *= (size_t)a->mdim[k];
        e.set_potential_map(std::vector<int8_t>(a->potential, a->potential + n));
        e.set_potential_weight(a->potential_weight);
        e.set_gradient_weight(a->gradient_weight);
      }
      vec_E<Waypoint<Dim>> S, G;
      for (int q = 0; q < n_q; q++) {
        S.push_back(mplh::wp_from<Dim>(starts[q], a->control));
        G.push_back(mplh::wp_from<Dim>(goals[q], a->control));
      }
      auto t0 = std::chrono::steady_clock::now();
      auto res = mq.plan(S, G, a->eps, a->max_num);
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      for (int q = 0; q < n_q; q++) {
        out[q].valid = res[q].valid ? 1 : 0;
        out[q].cost = res[q].cost;
        out[q].expanded = res[q].expanded;
        out[q].n_closed = (int)res[q].n_closed;
        out[q].n_actions = (int)res[q].actions.size();
      }
      auto t1 = std::chrono::steady_clock::now();
      mq.release();
      const double secs_release = std::chrono::duration<double>(std::chrono::steady_clock::now() - t1).count();
      if (totals) {
        totals[0] = (double)mq.iterations(); totals[1] = (double)mq.nodes_expanded(); totals[2] = secs;
        totals[3] = mq.t_pop(); totals[4] = mq.t_device(); totals[5] = mq.t_relax(); totals[6] = secs_release;
      }
    };
    if (a->dim == 2) go(std::integral_constant<int, 2>());
    else if (a->dim == 3) go(std::integral_constant<int, 3>());
    else { g_err = "dim must be 2 or 3"; return 1; }
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 2;
  }
}
}
