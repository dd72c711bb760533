// libmpl_host.so — C entry point that runs MPL::MapPlanner<Dim>::plan() with the GPU env
// (env_map_gpu -> libmplx).  Used by the Python tests and tools; C++ users include mpl_host.hpp.
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
}
