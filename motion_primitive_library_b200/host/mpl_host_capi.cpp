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
      if (a->speculate > 0) planner.setSpeculation(a->speculate);  // 0 = the library default
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

/* mplh_plan that also returns the nodes A* expanded, in pop order (graph_search.h:66-75): the replay
 * frontier of the benchmark.  *n_trace = nodes recorded (<= cap_trace). */
int mplh_plan_trace(const mplh_plan_args *a, mplh_plan_result *r, mplx_waypoint *trace, int cap_trace, int32_t *n_trace) {
  try {
    *r = mplh_plan_result{};
    std::vector<mplx_waypoint> tr;
    auto go = [&](auto dimtag) {
      constexpr int Dim = decltype(dimtag)::value;
      MPL::MapPlanner<Dim> planner(false);
      planner.setMapUtil(mplh::make_map<Dim>(a), a->device);
      planner.setControl(a->control);
      if (a->speculate > 0) planner.setSpeculation(a->speculate);  // 0 = the library default
      planner.gpu_env()->set_trace(&tr);
      std::vector<uint64_t> closed(1);
      std::vector<int32_t> actions(1);
      mplh::run<Dim>(planner, a, r, closed.data(), 0, actions.data(), 0);
      planner.gpu_env()->set_trace(nullptr);
    };
    if (a->dim == 2) go(std::integral_constant<int, 2>());
    else if (a->dim == 3) go(std::integral_constant<int, 3>());
    else { g_err = "dim must be 2 or 3"; return 1; }
    const int n = (int)std::min<std::size_t>(tr.size(), (std::size_t)cap_trace);
    for (int i = 0; i < n; i++) trace[i] = tr[i];
    if (n_trace) *n_trace = n;
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
      if (a->speculate > 0) planner.setSpeculation(a->speculate);  // 0 = the library default
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
      if (a->speculate > 0) planner.setSpeculation(a->speculate);  // 0 = the library default
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
      if (a->speculate > 0) planner.setSpeculation(a->speculate);  // 0 = the library default
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

/* Lock-step batched A* over many (start, goal) pairs on one map (MPL::MultiQueryPlanner) as a
 * session: open once (map upload, parameters), plan any number of query sets — the search states of a
 * set are recycled for the next one, so a steady stream of batches allocates its state memory once —
 * and close.  totals (7 doubles): [0] lock-step iterations (= device launches of the expansion
 * kernel), [1] nodes expanded over all queries, [2] wall seconds of the search, [3..5] seconds in the
 * pop / device expansion (incl. PCIe) / relax phases, [6] 0 (states are kept for the next set). */
}  // extern "C"
namespace {
struct BatchSession {
  int dim = 0;
  int control = 0;
  void *mq = nullptr;  // MPL::MultiQueryPlanner<dim>*
};
template <int Dim>
MPL::MultiQueryPlanner<Dim> *open_mq(const mplh_plan_args *a) {
  auto *mq = new MPL::MultiQueryPlanner<Dim>(mplh::make_map<Dim>(a), a->device);
  auto &e = mq->env();
  if (const char *t = std::getenv("MPLH_THREADS")) mq->setHostThreads(std::atoi(t));
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
  return mq;
}
template <int Dim>
void plan_mq(MPL::MultiQueryPlanner<Dim> *mq, int control, const mplx_waypoint *starts, const mplx_waypoint *goals, int n_q,
             double eps, int max_num, mplh_query_result *out, double *totals) {
  vec_E<Waypoint<Dim>> S, G;
  for (int q = 0; q < n_q; q++) {
    S.push_back(mplh::wp_from<Dim>(starts[q], control));
    G.push_back(mplh::wp_from<Dim>(goals[q], control));
  }
  auto t0 = std::chrono::steady_clock::now();
  auto res = mq->plan(S, G, eps, max_num);
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  for (int q = 0; q < n_q; q++) {
    out[q].valid = res[q].valid ? 1 : 0;
    out[q].cost = res[q].cost;
    out[q].expanded = res[q].expanded;
    out[q].n_closed = (int)res[q].n_closed;
    out[q].n_actions = (int)res[q].actions.size();
  }
  if (totals) {
    totals[0] = (double)mq->iterations(); totals[1] = (double)mq->nodes_expanded(); totals[2] = secs;
    totals[3] = mq->t_pop(); totals[4] = mq->t_device(); totals[5] = mq->t_relax(); totals[6] = 0.0;
  }
}
}  // namespace
extern "C" {

void *mplh_batch_open(const mplh_plan_args *a) {
  try {
    if (a->dim != 2 && a->dim != 3) throw std::runtime_error("dim must be 2 or 3");
    BatchSession *s = new BatchSession();
    s->dim = a->dim;
    s->control = a->control;
    s->mq = a->dim == 2 ? (void *)open_mq<2>(a) : (void *)open_mq<3>(a);
    return s;
  } catch (const std::exception &e) {
    g_err = e.what();
    return nullptr;
  }
}

int mplh_batch_plan(void *session, const mplx_waypoint *starts, const mplx_waypoint *goals, int n_q, double eps, int max_num,
                    mplh_query_result *out, double *totals) {
  try {
    BatchSession *s = (BatchSession *)session;
    if (!s || !s->mq) throw std::runtime_error("null session");
    if (s->dim == 2) plan_mq<2>((MPL::MultiQueryPlanner<2> *)s->mq, s->control, starts, goals, n_q, eps, max_num, out, totals);
    else plan_mq<3>((MPL::MultiQueryPlanner<3> *)s->mq, s->control, starts, goals, n_q, eps, max_num, out, totals);
    return 0;
  } catch (const std::exception &e) {
    g_err = e.what();
    return 1;
  }
}

/* Frees the session incl. the search states it kept; seconds spent are returned in *release_seconds. */
int mplh_batch_close(void *session, double *release_seconds) {
  BatchSession *s = (BatchSession *)session;
  if (!s) return 0;
  auto t0 = std::chrono::steady_clock::now();
  if (s->dim == 2) delete (MPL::MultiQueryPlanner<2> *)s->mq;
  else delete (MPL::MultiQueryPlanner<3> *)s->mq;
  delete s;
  if (release_seconds) *release_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return 0;
}

/* One-shot form: open, plan one set, close.  totals[6] = seconds spent freeing the search states. */
int mplh_plan_batch(const mplh_plan_args *a, const mplx_waypoint *starts, const mplx_waypoint *goals, int n_q,
                    mplh_query_result *out, double *totals) {
  void *s = mplh_batch_open(a);
  if (!s) return 1;
  const int rc = mplh_batch_plan(s, starts, goals, n_q, a->eps, a->max_num, out, totals);
  double rel = 0;
  mplh_batch_close(s, &rel);
  if (totals) totals[6] = rel;
  return rc;
}
}
