// plan_capi_types.h — the flat C structs of the one-shot planner entry points (shared by
// libmpl_host.so and the test harnesses; no C++ types).
#pragma once
#include <stdint.h>

#include "../../include/mplx.h"

extern "C" {
typedef struct {
  int32_t dim, control;
  const int8_t *map;
  int32_t mdim[3];
  double origin[3];
  double res;
  const double *U;
  int32_t nU, udim;
  double T, w, wyaw, eps;
  double v_max, a_max, j_max, yaw_max;
  double tol_pos, tol_vel, tol_acc;
  mplx_waypoint start, goal;
  int32_t max_num;     /* PlannerBase::setMaxNum */
  int32_t speculate;   /* nodes expanded per launch (GPU env); 1 = no speculation */
  int32_t device;
  const int8_t *potential; /* optional */
  double potential_weight, gradient_weight;
  int32_t heur_ignore_dynamics; /* PlannerBase::setHeurIgnoreDynamics; 1 = the reference's default */
} mplh_plan_args;

typedef struct {
  int32_t valid;          /* plan() return value */
  double cost;            /* getTrajCost() */
  int32_t expanded;       /* expand iterations (get_succ calls made by A*) */
  int32_t n_closed;       /* closed states */
  int32_t n_open;
  int32_t n_actions;      /* edges of the recovered trajectory */
  int64_t gpu_nodes;      /* nodes sent to the device (>= expanded when speculating) */
  int64_t gpu_calls;      /* mplx_expand calls */
  int64_t gpu_launches;
  double seconds;         /* wall time of plan() */
} mplh_plan_result;
}

extern "C" {
/* A scripted incremental-planning (LPA*) session on one planner: the steps run in order.
 *   PLAN     plan(start, goal) with setLPAstar(true); start is args->start until a SUBTREE step
 *   LINK     getLinkedNodes(): rebuild the voxel -> edges table from the current graph
 *   BLOCK    mark n cells occupied in the map, then updateBlockedNodes(cells)
 *   CLEAR    mark n cells free in the map, then updateClearedNodes(cells)
 *   SUBTREE  start := best_child_[n]; getSubStateSpace(n)   (re-root at the n-th trajectory node) */
enum { MPLH_OP_PLAN = 0, MPLH_OP_LINK = 1, MPLH_OP_BLOCK = 2, MPLH_OP_CLEAR = 3, MPLH_OP_SUBTREE = 4 };
typedef struct {
  int32_t op;
  int32_t n;            /* BLOCK/CLEAR: number of cells; SUBTREE: time step */
  const int32_t *cells; /* BLOCK/CLEAR: n * dim cell coordinates */
} mplh_lpa_step;
typedef struct {
  int32_t valid;        /* PLAN: plan() return value */
  double cost;          /* PLAN: getTrajCost() */
  int32_t expanded;     /* PLAN: expand iterations of this call */
  int32_t n_actions;    /* PLAN: edges of the recovered trajectory */
  int32_t n_states, n_closed, n_open; /* after the step: states in the space / closed / in the queue */
  uint64_t state_hash;  /* after the step: FNV-1a over (key, g, rhs, opened, closed) of all states by key */
  int64_t n_linked;     /* LINK: voxel centres returned */
  uint64_t linked_hash; /* LINK: FNV-1a over the sorted (voxel index, state key, pred index) table */
  double seconds;
} mplh_lpa_out;
}

extern "C" {
/* Batched queries (config 5): shares every field of mplh_plan_args except start/goal. */
typedef struct {
  int32_t valid;
  double cost;
  int32_t expanded;
  int32_t n_closed;
  int32_t n_actions;
} mplh_query_result;
}

