/*
 * mpl_oracle.h — C ABI of the CPU ORACLE for the node-expansion hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it, and only as the checker / the timed CPU baseline.
 *
 * The oracle restates, operation by operation, the reference's
 *   env_map<Dim>::get_succ            include/mpl_planner/env/env_map.h:147-172
 *   env_map<Dim>::traverse_primitive  include/mpl_planner/env/env_map.h:90-132
 * and everything beneath them (see mpl_oracle.cpp for per-function citations).
 *
 * Parity status: pinned against the UNMODIFIED reference headers compiled with a
 * minimal Eigen/Boost shim (oracle/_ref, built by oracle/Makefile from /root/reference
 * where it lies) and against the analytic known-answer vectors of SURVEY.md §10
 * (tests/test_oracle_kat.py).  The 64-bit hash mix is Boost-version dependent and the
 * reference pins no hash values: the canonical identity is the int32 lattice tuple.
 */
#ifndef MPL_ORACLE_H
#define MPL_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Waypoint<Dim> payload (include/mpl_basis/waypoint.h:33-38). 2D uses [0],[1]. 112 B. */
typedef struct {
  double pos[3], vel[3], acc[3], jrk[3];
  double yaw;
  double t;
} orc_waypoint;

/* Everything env_map::get_succ reads besides `curr` (env_base.h:368-400, env_map.h:288-296,
 * map_util.h:300-313). */
typedef struct {
  int32_t dim;     /* 2 or 3 */
  int32_t control; /* Control::Control bits (include/mpl_basis/control.h:10-20) */
  double T;        /* dt_ : primitive duration */
  double w, wyaw;
  double v_max, a_max, j_max, yaw_max;
  int32_t nU, udim; /* |U| and length of each control vector (Dim or Dim+1) */
  const double *U;  /* nU*udim, row-major */
  int32_t mdim[3];
  double origin[3];
  double res;
  const int8_t *map;       /* x-fastest, occ=100 free=0 unknown=-1 */
  const int8_t *potential; /* NULL = potential_map_.empty() */
  double potential_weight, gradient_weight;
  const uint8_t *region; /* NULL = search_region_.empty(); else 1 byte per voxel */
} orc_env;

#define ORC_LATTICE_MAX 13 /* 3 axes * 4 fields + yaw */

/* One get_succ call. Outputs have capacity nU. lattice may be NULL (else nU*13 int32).
 * Returns the number of successors emitted (control-index order). */
int orc_get_succ(const orc_env *env, const orc_waypoint *curr, orc_waypoint *succ,
                 double *cost, int32_t *action, uint64_t *key, int32_t *lattice);

/* Batch helper: expands n nodes, output segment for node i starts at i*nU; count[i]
 * entries valid. nthreads<=1 runs serial (the reference's execution model). */
int orc_expand_batch(const orc_env *env, const orc_waypoint *nodes, int n, orc_waypoint *succ,
                     double *cost, int32_t *action, uint64_t *key, int32_t *lattice,
                     int32_t *count, int nthreads);

/* Same work as orc_expand_batch but discards the outputs (thread-local scratch), and
 * returns the total number of successors + samples visited through the two out params;
 * used to time the CPU baseline without a huge output allocation. */
int orc_expand_batch_timed(const orc_env *env, const orc_waypoint *nodes, int n, int nthreads,
                           int64_t *total_succ, int64_t *total_samples, double *seconds);

/* hash_value(Waypoint) (include/mpl_basis/waypoint.h:93-125); lattice may be NULL. */
uint64_t orc_hash(const orc_env *env, const orc_waypoint *w, int32_t *lattice, int32_t *n_lattice);

/* Diagnostics for the KATs. */
int orc_sample_count(double T, int n); /* iterations of for(t=0;t<T;t+=T/n) */
double orc_max_vel(const orc_env *env, const orc_waypoint *curr, int control_idx, int axis);
int64_t orc_last_samples(void); /* samples visited by the last orc_get_succ on this thread */

/* Stored-edge queries of the incremental planner: edge = Primitive(parents[i], U[actions[i]], T).
 * orc_edges_is_free: env_map::is_free(pr) (env_map.h:60-76) and, optionally, the intrinsic cost
 * (env_base.h:343-345).  orc_edges_cells: the voxel walk of MapPlanner::getLinkedNodes
 * (map_planner.cpp:135-151); entries of edge i are [out_offset[i], out_offset[i+1]) of out_cells
 * (dim int32 each); returns the total (entries beyond capacity are counted, not written). */
int orc_edges_is_free(const orc_env *env, const orc_waypoint *parents, const int32_t *actions, int n,
                      uint8_t *out_free, double *out_cost);
int64_t orc_edges_cells(const orc_env *env, const orc_waypoint *parents, const int32_t *actions, int n,
                        int64_t *out_offset, int32_t *out_cells, int64_t capacity);

#ifdef __cplusplus
}
#endif
#endif
