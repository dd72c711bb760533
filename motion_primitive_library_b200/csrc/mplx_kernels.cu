// mplx_kernels.cu — expand(frontier x U): the batched body of env_map<Dim>::get_succ
// (include/mpl_planner/env/env_map.h:147-172) for sm_100a.
//
// Work decomposition (v1): one thread per (frontier node, control) primitive; a CTA covers
// NPB = 256/|U| whole nodes so that the stable, control-ordered compaction of a node's
// successors (env_map.h:155-170 push_back order) is a CTA-local ballot/popcount.
// The voxel grid / potential grid / tunnel bitmask live in HBM (staged once by mplx_set_*),
// and are read with read-only (ld.global.nc) loads; every sample loop exits at the first
// blocking sample exactly like traverse_primitive's `return inf` (env_map.h:104-121).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false  (see mplx_device.cuh).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/mplx.h"
#include "mplx_device.cuh"
#include "mplx_kernels.h"

namespace mplx {

constexpr int kThreads = 256;

template <int DIM, int ORD, bool YAW>
struct PrimState {
  Axis<ORD> ax[DIM];
  double yaw_u, yaw0;  // pr_yaw_ = [0,0,0,0,u(Dim),yaw]  (primitive.h:34,235-248)
};

// traverse_primitive: include/mpl_planner/env/env_map.h:90-132
template <int DIM, int ORD, bool YAW, bool STATS>
__device__ __forceinline__ double traverse(const EnvParams &P, const PrimState<DIM, ORD, YAW> &pr,
                                           unsigned &n_samples) {
  const double T = P.T;
  double max_v = 0;
#pragma unroll
  for (int i = 0; i < DIM; i++) {
    double mv = pr.ax[i].max_vel(T);
    if (mv > max_v) max_v = mv;
  }
  int n = max(5, (int)ceil(max_v * T / P.res));
  double c = 0;
  const double dt = T / n;
  const bool need_vel = YAW || (P.pot != nullptr && P.grad_w != 0.0);
  for (double t = 0; t < T; t += dt) {
    if (STATS) n_samples++;
    const double t2 = t * t;
    const double pw3 = t2 * t;
    const double pw4 = pw3 * t;
    int pn[DIM];
    bool outside = false;
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      double pk = pr.ax[k].template p<false>(t, pw3, pw4);
      // floatToInt: map_util.h:103-108
      pn[k] = (int)round((pk - P.origin[k]) / P.res - 0.5);
      outside = outside || pn[k] < 0 || pn[k] >= P.mdim[k];
    }
    if (outside) return INFINITY;
    // getIndex: map_util.h:34-41 (inside the map it cannot overflow: the grid is <2^31 cells)
    int idx = pn[0] + P.mdim[0] * pn[1];
    if (DIM == 3) idx += P.mdim[0] * P.mdim[1] * pn[2];
    if (P.region_bits != nullptr) {
      if (!((__ldg(P.region_bits + (idx >> 5)) >> (idx & 31)) & 1u)) return INFINITY;
    }
    double vel[DIM];
    if (need_vel) {
#pragma unroll
      for (int k = 0; k < DIM; k++) vel[k] = pr.ax[k].v(t, pw3);
    }
    if (P.pot != nullptr) {
      const int pv = (int)__ldg(P.pot + idx);
      if (pv < 100 && pv > 0) {
        double g = 0.0;
        if (P.grad_w != 0.0) {
          // pt.vel.norm(): Eigen's unrolled reduction a0 + (a1 + a2)
          double n2 = DIM == 2 ? vel[0] * vel[0] + vel[1] * vel[1]
                               : vel[0] * vel[0] + (vel[1] * vel[1] + vel[DIM - 1] * vel[DIM - 1]);
          g = P.grad_w * sqrt(n2);
        } else {
          g = 0.0;  // gradient_weight_(0) * norm == +0 for finite norm
        }
        c += dt * (P.pot_w * pv + g);
      } else if (pv >= 100)
        return INFINITY;
    } else if (__ldg(P.map + idx) == 100)
      return INFINITY;
    if (YAW) {
      if (P.wyaw > 0) {
        const double v0 = vel[0], v1 = vel[1];
        if (sqrt(v0 * v0 + v1 * v1) > 1e-5) {
          const double yaw = normalize_angle(pr.yaw_u * t + pr.yaw0);
          double sn, cs;
          sincos(yaw, &sn, &cs);
          const double v_value = 1 - dot2_normalized(v0, v1, cs, sn);
          c += P.wyaw * v_value * dt;
        }
      }
    }
  }
  return c;
}

// validate_yaw: include/mpl_basis/primitive.h:503-525
template <int DIM, int ORD, bool YAW>
__device__ __forceinline__ bool validate_yaw(const EnvParams &P, const PrimState<DIM, ORD, YAW> &pr) {
  if (P.yaw_max <= 0) return true;
  const double T = P.T;
  const double pw3T = (T * T) * T;
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const double t = e == 0 ? 0.0 : T;
    const double v0 = pr.ax[0].v(t, e == 0 ? 0.0 : pw3T);
    const double v1 = pr.ax[1].v(t, e == 0 ? 0.0 : pw3T);
    if (v0 != 0 || v1 != 0) {
      const double yaw = normalize_angle(0.0 + pr.yaw_u * t + pr.yaw0);
      double sn, cs;
      sincos(yaw, &sn, &cs);
      const double d = dot2_normalized(v0, v1, cs, sn);
      if (d < P.cos_yaw_max) return false;
    }
  }
  return true;
}

template <int DIM, int ORD, bool YAW, bool STATS>
__global__ void __launch_bounds__(kThreads)
expand_kernel(const EnvParams P, const mplx_waypoint *__restrict__ nodes, int n_nodes, int npb,
              int32_t *__restrict__ out_count, mplx_waypoint *__restrict__ out_succ,
              double *__restrict__ out_cost, int32_t *__restrict__ out_action,
              uint64_t *__restrict__ out_key, int32_t *__restrict__ out_lattice) {
  __shared__ uint32_t vbits[kMaxU / 32 + 9];
  __shared__ unsigned long long s_stats[2];
  const int nU = P.nU;
  const int items = npb * nU;
  const int node0 = blockIdx.x * npb;
  const int words = (items + 31) >> 5;
  if (STATS && threadIdx.x < 2) s_stats[threadIdx.x] = 0;

  for (int base = 0; base < items; base += kThreads) {
    const int item = base + threadIdx.x;
    const int nl = item / nU;
    const int ci = item - nl * nU;
    const int ni = node0 + nl;
    const bool active = item < items && ni < n_nodes;

    bool emit = false;
    double cost = 0;
    mplx_waypoint tn;
    int lat[MPLX_LATTICE_MAX];
    uint64_t key = 0;
    unsigned n_samples = 0;

    if (active) {
      const mplx_waypoint *cp = nodes + ni;
      const double *u = P.U + (size_t)ci * P.udim;
      PrimState<DIM, ORD, YAW> pr;
      double cpos[DIM];
      uint64_t hcurr = 0;
      // Primitive(curr, U[i], dt): primitive.h:220-256 ; hash_value(curr): waypoint.h:93-125
#pragma unroll
      for (int k = 0; k < DIM; k++) {
        const double p = cp->pos[k], v = cp->vel[k], a = cp->acc[k], j = cp->jrk[k];
        cpos[k] = p;
        pr.ax[k].build(__ldg(u + k), p, v, a, j);
        hash_combine(hcurr, lattice_id(p, 0.01));
        if (ORD >= 2) hash_combine(hcurr, lattice_id(v, 0.1));
        if (ORD >= 3) hash_combine(hcurr, lattice_id(a, 0.1));
        if (ORD >= 4) hash_combine(hcurr, lattice_id(j, 0.1));
      }
      if (YAW) {
        pr.yaw_u = __ldg(u + DIM);
        pr.yaw0 = cp->yaw;
        hash_combine(hcurr, lattice_id(cp->yaw, 0.1));
      }
      // tn = pr.evaluate(dt): primitive.h:321-331 (all four derivative vectors are filled)
      const double T = P.T;
      const double pw3T = (T * T) * T, pw4T = pw3T * T;
      int nl_ = 0;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        if (k < DIM) {
          tn.pos[k] = pr.ax[k].template p<true>(T, pw3T, pw4T);
          tn.vel[k] = pr.ax[k].v(T, pw3T);
          tn.acc[k] = pr.ax[k].a(T);
          tn.jrk[k] = pr.ax[k].j(T);
          int id = lattice_id(tn.pos[k], 0.01);
          hash_combine(key, id);
          lat[nl_++] = id;
          if (ORD >= 2) { id = lattice_id(tn.vel[k], 0.1); hash_combine(key, id); lat[nl_++] = id; }
          if (ORD >= 3) { id = lattice_id(tn.acc[k], 0.1); hash_combine(key, id); lat[nl_++] = id; }
          if (ORD >= 4) { id = lattice_id(tn.jrk[k], 0.1); hash_combine(key, id); lat[nl_++] = id; }
        } else {
          tn.pos[k] = tn.vel[k] = tn.acc[k] = tn.jrk[k] = 0.0;
        }
      }
      tn.yaw = 0.0;
      if (YAW) {
        // pr_yaw_.p(t) = 0/120*.. + c4*t + c5 with the leading +0 sum (primitive.h:128-131,328)
        tn.yaw = normalize_angle(0.0 + pr.yaw_u * T + pr.yaw0);
        int id = lattice_id(tn.yaw, 0.1);
        hash_combine(key, id);
        lat[nl_++] = id;
      }
#pragma unroll
      for (int q = 0; q < MPLX_LATTICE_MAX; q++)
        if (q >= nl_) lat[q] = 0;
      tn.t = cp->t + T;  // env_map.h:161

      // tn == curr (hash equality, waypoint.h:133-135) || !validate_primitive (primitive.h:449-475)
      bool ok = key != hcurr;
      if (ok && YAW) ok = validate_yaw<DIM, ORD, YAW>(P, pr);
      if (ok && ORD >= 2 && P.v_max > 0) {
#pragma unroll
        for (int k = 0; k < DIM; k++) ok = ok && !(pr.ax[k].max_vel(T) > P.v_max);
      }
      if (ok && ORD >= 3 && P.a_max > 0) {
#pragma unroll
        for (int k = 0; k < DIM; k++) ok = ok && !(pr.ax[k].max_acc(T) > P.a_max);
      }
      if (ok && ORD >= 4 && P.j_max > 0) {
#pragma unroll
        for (int k = 0; k < DIM; k++) ok = ok && !(pr.ax[k].max_jrk(T) > P.j_max);
      }
      emit = ok;
      if (ok) {
        bool same = true;  // curr.pos == tn.pos (env_map.h:163)
#pragma unroll
        for (int k = 0; k < DIM; k++) same = same && (cpos[k] == tn.pos[k]);
        cost = same ? 0.0 : traverse<DIM, ORD, YAW, STATS>(P, pr, n_samples);
        if (!isinf(cost)) {
          // calculate_intrinsic_cost: env_base.h:343-345 ; Primitive::J: primitive.h:403-407
          double J = pr.ax[0].J(T);
#pragma unroll
          for (int k = 1; k < DIM; k++) J += pr.ax[k].J(T);
          cost += J + P.w * T;
        }
      }
    }

    // ---- stable per-node compaction (control order) -------------------------------------
    const unsigned bal = __ballot_sync(0xffffffffu, emit);
    if ((threadIdx.x & 31) == 0 && (item >> 5) < words) vbits[item >> 5] = bal;
    __syncthreads();
    if (active) {
      const int s = nl * nU;  // first item of my node
      int rank = 0;
      for (int wd = s >> 5; wd <= (item >> 5); wd++) {
        uint32_t m = vbits[wd];
        const int lo = wd << 5;
        if (s > lo) m &= ~0u << (s - lo);
        if (item < lo + 32) m &= (item - lo) ? (~0u >> (32 - (item - lo))) : 0u;
        rank += __popc(m);
      }
      if (ci == nU - 1) out_count[ni] = rank + (emit ? 1 : 0);
      if (emit) {
        const size_t slot = (size_t)ni * nU + rank;
        if (out_succ) out_succ[slot] = tn;
        if (out_cost) out_cost[slot] = cost;
        if (out_action) out_action[slot] = ci;
        if (out_key) out_key[slot] = key;
        if (out_lattice) {
#pragma unroll
          for (int q = 0; q < MPLX_LATTICE_MAX; q++) out_lattice[slot * MPLX_LATTICE_MAX + q] = lat[q];
        }
      }
    }
    if (STATS) {
      atomicAdd(&s_stats[0], (unsigned long long)n_samples);
      if (emit) atomicAdd(&s_stats[1], 1ull);
    }
    __syncthreads();
  }
  if (STATS && threadIdx.x < 2 && P.stats) atomicAdd(&P.stats[threadIdx.x], s_stats[threadIdx.x]);
}

template <int DIM, int ORD, bool YAW>
static cudaError_t launch_t(const EnvParams &P, const mplx_waypoint *d_nodes, int n_nodes,
                            const mplx_succ_out &o, cudaStream_t st) {
  const int npb = P.nU >= kThreads ? 1 : kThreads / P.nU;
  const int grid = (n_nodes + npb - 1) / npb;
  if (P.stats)
    expand_kernel<DIM, ORD, YAW, true><<<grid, kThreads, 0, st>>>(
        P, d_nodes, n_nodes, npb, o.count, o.succ, o.cost, o.action, o.key, o.lattice);
  else
    expand_kernel<DIM, ORD, YAW, false><<<grid, kThreads, 0, st>>>(
        P, d_nodes, n_nodes, npb, o.count, o.succ, o.cost, o.action, o.key, o.lattice);
  return cudaGetLastError();
}

template <int DIM>
static cudaError_t launch_d(const EnvParams &P, const mplx_waypoint *d_nodes, int n_nodes,
                            const mplx_succ_out &o, cudaStream_t st) {
  const bool yaw = (P.control & 16) != 0;
  switch (P.control & 15) {
    case MPLX_VEL: return yaw ? launch_t<DIM, 1, true>(P, d_nodes, n_nodes, o, st) : launch_t<DIM, 1, false>(P, d_nodes, n_nodes, o, st);
    case MPLX_ACC: return yaw ? launch_t<DIM, 2, true>(P, d_nodes, n_nodes, o, st) : launch_t<DIM, 2, false>(P, d_nodes, n_nodes, o, st);
    case MPLX_JRK: return yaw ? launch_t<DIM, 3, true>(P, d_nodes, n_nodes, o, st) : launch_t<DIM, 3, false>(P, d_nodes, n_nodes, o, st);
    case MPLX_SNP: return yaw ? launch_t<DIM, 4, true>(P, d_nodes, n_nodes, o, st) : launch_t<DIM, 4, false>(P, d_nodes, n_nodes, o, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_expand(const EnvParams &P, const mplx_waypoint *d_nodes, int n_nodes,
                          const mplx_succ_out &o, cudaStream_t st) {
  if (n_nodes <= 0) return cudaSuccess;
  return P.dim == 2 ? launch_d<2>(P, d_nodes, n_nodes, o, st) : launch_d<3>(P, d_nodes, n_nodes, o, st);
}

// ---- set-up kernels -------------------------------------------------------------------

// std::vector<bool> search_region_ (env_base.h:400) arrives as one byte per voxel; pack it
// to 1 bit per voxel so the per-sample test is a 4-byte read-only load.
__global__ void pack_region_kernel(const uint8_t *__restrict__ bytes, size_t nvox,
                                   uint32_t *__restrict__ bits) {
  const size_t nwords = (nvox + 31) >> 5;
  for (size_t wd = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wd < nwords;
       wd += (size_t)gridDim.x * blockDim.x) {
    uint32_t m = 0;
    const size_t b0 = wd << 5;
#pragma unroll 8
    for (int b = 0; b < 32; b++) {
      const size_t i = b0 + b;
      if (i < nvox && bytes[i]) m |= 1u << b;
    }
    bits[wd] = m;
  }
}

cudaError_t launch_pack_region(const uint8_t *d_bytes, size_t nvox, uint32_t *d_bits, cudaStream_t st) {
  const size_t nwords = (nvox + 31) >> 5;
  int grid = (int)((nwords + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  pack_region_kernel<<<grid, 256, 0, st>>>(d_bytes, nvox, d_bits);
  return cudaGetLastError();
}

}  // namespace mplx
