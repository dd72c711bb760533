// mplx_kernels.cu — expand(frontier x U): the batched body of env_map<Dim>::get_succ
// (include/mpl_planner/env/env_map.h:147-172) for sm_100a.
//
// Work decomposition.  A CTA of 256 threads covers NPB = 256/|U| whole frontier nodes, one
// thread per (node, control) primitive:
//   phase A  build the primitive (primitive.h:220-256), evaluate the end state tn
//            (:321-331), the dynamic validity (primitive.h:449-525) and, for the survivors, the
//            lattice key (waypoint.h:93-125) in registers;
//   phase B  stable, control-ordered compaction of each node's successors with warp ballots
//            (the push_back order of env_map.h:155-170) and write-out of tn/key/action;
//   phase C  traverse_primitive (env_map.h:90-132), three interchangeable implementations:
//     expand_reg_kernel  (default) thread = primitive, quotients in registers, the reference's
//                        t += dt loop in groups of 4 samples with group-level control flow;
//     expand_flat_kernel the samples of a warp's 32 primitives laid end to end and dealt to the
//                        lanes (quotients staged in shared memory, sample times from the table
//                        that reproduces the running sum, atomicMin of the first blocking sample);
//     expand_seq_kernel  the literal per-primitive loop; serves |U| > 256.
// The occupancy grid is read as 1 bit/voxel (16 MiB at 512^3: L2-resident); potential-field
// planning reads the int8 grid.  No tensor cores: there is no dense contraction on this path.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false  (see mplx_device.cuh).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/mplx.h"
#include "mplx_device.cuh"
#include "mplx_kernels.h"
#include "mplx_prim.cuh"

namespace mplx {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kNoBlock = 0x7fffffff;

// ---- voxel classification shared by all sample loops ---------------------------------------
// Two steps so that a caller can issue the loads of several samples before consuming any.
// voxel_fetch returns one packed word for a valid in-map index:
//   bits 0..7  the potential value (int8) when a potential map is set, else the occupancy bit
//   bit  8     1 = inside the tunnel (search_region_ empty counts as inside)
// voxel_classify applies env_map.h:104-121 to it: true = the sample blocks the primitive,
// otherwise the potential term is added to `term`.
typedef unsigned int VoxelRaw;
constexpr VoxelRaw kVoxelNone = 0x100u;  // in-region, free
__device__ __forceinline__ VoxelRaw voxel_fetch(const EnvParams &P, int idx) {
  unsigned r = 0x100u;
  if (P.region_bits != nullptr) r = ((__ldg(P.region_bits + (idx >> 5)) >> (idx & 31)) & 1u) << 8;
  if (P.pot != nullptr)
    r |= (unsigned)(unsigned char)__ldg(P.pot + idx);
  else
    r |= (__ldg(P.occ_bits + (idx >> 5)) >> (idx & 31)) & 1u;
  return r;
}
__device__ __forceinline__ bool voxel_classify(const EnvParams &P, VoxelRaw r, double dt, double vnorm_w,
                                               double &term) {
  if (!(r & 0x100u)) return true;  // outside the tunnel (env_map.h:104-106)
  if (P.pot != nullptr) {
    const int pv = (int)(signed char)(r & 0xffu);
    if (pv < 100 && pv > 0)
      term += dt * (P.pot_w * pv + vnorm_w);
    else if (pv >= 100)
      return true;
    return false;
  }
  return r & 1u;
}
__device__ __forceinline__ bool voxel_blocks(const EnvParams &P, int idx, double dt, double vnorm_w,
                                             double &term) {
  return voxel_classify(P, voxel_fetch(P, idx), dt, vnorm_w, term);
}

// floatToInt + isOutside + getIndex (map_util.h:103-108, 51-55, 34-41) for one sample.
// pk[] are the sample's position coordinates.  Returns the cell index or -1 when outside.
//
// With the exact quotient y = RN((p - origin)/res) the reference's
//     pn = (int)std::round(RN(y - 0.5));   inside <=> 0 <= pn < dim
// is equivalent to
//     inside <=> 2^-55 < y < dim;          pn = floor(y)
// because y - 0.5 is exact for y >= 0.5, rounds into (-0.5, 0) for 2^-55 < y < 0.5 and to
// exactly -0.5 (-> -1, half away from zero) for 0 <= y <= 2^-55; a tie x = J - 0.5 with J >= 1
// rounds up to J = floor(y).  (tests/arith_identities.cpp: check_cell_rule.)
template <int DIM>
__device__ __forceinline__ int sample_index(const EnvParams &P, const double (&pk)[DIM]) {
  int pn[DIM];
  bool inside = true;
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    const double y = div_exact(pk[k] - P.origin[k], P.res, P.rinv);
    // floor(y) on the (otherwise idle) conversion pipe: saturates for |y| >= 2^31 and gives 0
    // for NaN, both of which the two tests below classify as outside
    pn[k] = __double2int_rd(y);
    inside = inside && (y > 0x1p-55) && ((unsigned)pn[k] < (unsigned)P.mdim[k]);
  }
  if (!inside) return -1;
  int idx = pn[0] + P.mdim[0] * pn[1];
  if (DIM == 3) idx += P.mdim[0] * P.mdim[1] * pn[DIM - 1];
  return idx;
}

// The yaw-alignment term of env_map.h:122-129.
__device__ __forceinline__ double yaw_term(const EnvParams &P, double v0, double v1, double yaw, double dt) {
  if (sqrt(v0 * v0 + v1 * v1) > 1e-5) {
    double sn, cs;
    sincos(yaw, &sn, &cs);
    const double v_value = 1 - dot2_normalized(v0, v1, cs, sn);
    return P.wyaw * v_value * dt;
  }
  return 0.0;
}

// pt.vel.norm() scaled by gradient_weight_ (env_map.h:115-116); Eigen's unrolled reduction
// associates a 3-vector sum as a0 + (a1 + a2).
template <int DIM>
__device__ __forceinline__ double grad_term(const EnvParams &P, const double (&vel)[DIM]) {
  if (P.grad_w == 0.0) return 0.0;  // gradient_weight_(0) * norm == +0 for a finite norm
  const double n2 = DIM == 2 ? vel[0] * vel[0] + vel[1] * vel[1]
                             : vel[0] * vel[0] + (vel[1] * vel[1] + vel[DIM - 1] * vel[DIM - 1]);
  return P.grad_w * sqrt(n2);
}

// traverse_primitive, literal per-primitive loop: include/mpl_planner/env/env_map.h:90-132.
// max_v is the caller's max_i pr.max_vel(i) (the reference recomputes it at :91-94).
template <int DIM, int ORD, bool YAW>
__device__ __forceinline__ double traverse_loop(const EnvParams &P, const double *cf, bool need_vel,
                                                double max_v, unsigned &n_samples) {
  using CL = CoefLayout<DIM, ORD, YAW>;
  const double T = P.T;
  const int n = max(5, (int)ceil(max_v * T / P.res));
  double c = 0;
  const double dt = T / n;
  const int NC = CL::ncoef(need_vel);
  for (double t = 0; t < T; t += dt) {
    n_samples++;
    double pk[DIM], vel[DIM];
    eval_pos<DIM, ORD>(cf, t, pk);
    const int idx = sample_index<DIM>(P, pk);
    if (idx < 0) return INFINITY;
    double gterm = 0.0;
    if (need_vel) {
      eval_vel<DIM, ORD>(cf + CL::NCP, t, vel);
      gterm = grad_term<DIM>(P, vel);
    }
    double term = 0.0;
    if (voxel_blocks(P, idx, dt, gterm, term)) return INFINITY;
    c += term;
    if (YAW) {
      if (P.wyaw > 0) c += yaw_term(P, vel[0], vel[1], normalize_angle(cf[NC - 2] * t + cf[NC - 1]), dt);
    }
  }
  return c;
}

// Out-of-line copy for the flat kernel's rare fallback (cf lives in shared memory there).
template <int DIM, int ORD, bool YAW>
__device__ __noinline__ double traverse_loop_cold(const EnvParams *P, const double *cf, bool need_vel,
                                                  double max_v, unsigned *n_samples) {
  unsigned ns = 0;
  const double c = traverse_loop<DIM, ORD, YAW>(*P, cf, need_vel, max_v, ns);
  *n_samples = ns;
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

struct OutPtrs {
  int32_t *count;
  mplx_waypoint *succ;
  double *cost;
  int32_t *action;
  uint64_t *key;
  int32_t *lattice;
};

// Phases A and B for one item (all threads of the CTA must call it: it contains a barrier).
// LAT: the caller asked for the lattice ints (mplx_succ_out.lattice); without it the 13-entry
// array never exists (it would cost 13 registers through phase A).
template <int DIM, int ORD, bool YAW, bool LAT>
__device__ __forceinline__ void phase_ab(const EnvParams &P, const mplx_waypoint *__restrict__ nodes,
                                         int n_nodes, int item, int items, int nU, int node0,
                                         uint32_t *vbits, int words, const OutPtrs &o,
                                         PrimState<DIM, ORD, YAW> &pr, bool &emit, bool &same,
                                         double &max_v, size_t &slot) {
  const int nl = item / nU;
  const int ci = item - nl * nU;
  const int ni = node0 + nl;
  const bool active = item < items && ni < n_nodes;
  emit = false;
  same = true;
  max_v = 0;
  slot = 0;

  mplx_waypoint tn;
  int lat[LAT ? MPLX_LATTICE_MAX : 1];
  uint64_t key = 0;
  if (active) {
    const mplx_waypoint *cp = nodes + ni;
    const double *u = P.U + (size_t)ci * P.udim;
    // Primitive(curr, U[i], dt): primitive.h:220-256
#pragma unroll
    for (int k = 0; k < DIM; k++) pr.ax[k].build(__ldg(u + k), cp->pos[k], cp->vel[k], cp->acc[k], cp->jrk[k]);
    if (YAW) {
      pr.yaw_u = __ldg(u + DIM);
      pr.yaw0 = cp->yaw;
    }
    // tn = pr.evaluate(dt): primitive.h:321-331 (all four derivative vectors are filled)
    const double T = P.T;
    const double pw3T = (T * T) * T, pw4T = pw3T * T;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (k < DIM) {
        tn.pos[k] = pr.ax[k].template p<true>(T, pw3T, pw4T);
        tn.vel[k] = pr.ax[k].v(T, pw3T);
        tn.acc[k] = pr.ax[k].a(T);
        tn.jrk[k] = pr.ax[k].j(T);
        same = same && (pr.ax[k].c5 == tn.pos[k]);  // curr.pos == tn.pos (env_map.h:163)
      } else {
        tn.pos[k] = tn.vel[k] = tn.acc[k] = tn.jrk[k] = 0.0;
      }
    }
    tn.yaw = 0.0;
    // pr_yaw_.p(t) = 0/120*.. + c4*t + c5 with the leading +0 sum (primitive.h:128-131,328)
    if (YAW) tn.yaw = normalize_angle(0.0 + pr.yaw_u * T + pr.yaw0);
    tn.t = cp->t + T;  // env_map.h:161

    // !validate_primitive (primitive.h:449-475) first: the test is pure, and a primitive that fails
    // it is dropped whatever its key, so the lattice key is only computed for the survivors
    bool ok = true;
    if (YAW) ok = validate_yaw<DIM, ORD, YAW>(P, pr);
    // max_vel per axis serves validate_xxx(VEL) (primitive.h:482-496) and traverse (env_map.h:91-94)
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      const double mv = pr.ax[k].max_vel(T);
      if (ORD >= 2 && P.v_max > 0) ok = ok && !(mv > P.v_max);
      if (mv > max_v) max_v = mv;
    }
    if (ok && ORD >= 3 && P.a_max > 0) {
#pragma unroll
      for (int k = 0; k < DIM; k++) ok = ok && !(pr.ax[k].max_acc(T) > P.a_max);
    }
    if (ok && ORD >= 4 && P.j_max > 0) {
#pragma unroll
      for (int k = 0; k < DIM; k++) ok = ok && !(pr.ax[k].max_jrk(T) > P.j_max);
    }
    if (ok) {
      // tn == curr  <=>  hash_value(tn) == hash_value(curr)  (waypoint.h:133-135, 93-125)
      uint64_t hcurr = 0;
      int nl_ = 0;
#pragma unroll
      for (int k = 0; k < DIM; k++) {
        hash_combine(hcurr, lattice_id(cp->pos[k], 0.01, 100.0));
        if (ORD >= 2) hash_combine(hcurr, lattice_id(cp->vel[k], 0.1, 10.0));
        if (ORD >= 3) hash_combine(hcurr, lattice_id(cp->acc[k], 0.1, 10.0));
        if (ORD >= 4) hash_combine(hcurr, lattice_id(cp->jrk[k], 0.1, 10.0));
        int id = lattice_id(tn.pos[k], 0.01, 100.0);
        hash_combine(key, id);
        if (LAT) lat[nl_++] = id;
        if (ORD >= 2) { id = lattice_id(tn.vel[k], 0.1, 10.0); hash_combine(key, id); if (LAT) lat[nl_++] = id; }
        if (ORD >= 3) { id = lattice_id(tn.acc[k], 0.1, 10.0); hash_combine(key, id); if (LAT) lat[nl_++] = id; }
        if (ORD >= 4) { id = lattice_id(tn.jrk[k], 0.1, 10.0); hash_combine(key, id); if (LAT) lat[nl_++] = id; }
      }
      if (YAW) {
        hash_combine(hcurr, lattice_id(cp->yaw, 0.1, 10.0));
        const int id = lattice_id(tn.yaw, 0.1, 10.0);
        hash_combine(key, id);
        if (LAT) lat[nl_++] = id;
      }
      if (LAT) {
#pragma unroll
        for (int q = 0; q < MPLX_LATTICE_MAX; q++)
          if (q >= nl_) lat[q] = 0;
      }
      ok = key != hcurr;
    }
    emit = ok;
  }

  // ---- phase B: stable per-node compaction (control order) ----
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
    if (ci == nU - 1) o.count[ni] = rank + (emit ? 1 : 0);
    if (emit) {
      slot = (size_t)ni * nU + rank;
      if (o.succ) o.succ[slot] = tn;
      if (o.action) o.action[slot] = ci;
      if (o.key) o.key[slot] = key;
      if (LAT && o.lattice) {
#pragma unroll
        for (int q = 0; q < MPLX_LATTICE_MAX; q++) o.lattice[slot * MPLX_LATTICE_MAX + q] = lat[q];
      }
    }
  }
}

// calculate_intrinsic_cost: env_base.h:343-345 ; Primitive::J: primitive.h:403-407
template <int DIM, int ORD, bool YAW>
__device__ __forceinline__ double intrinsic_cost(const EnvParams &P, const PrimState<DIM, ORD, YAW> &pr) {
  double J = pr.ax[0].J(P.T);
#pragma unroll
  for (int k = 1; k < DIM; k++) J += pr.ax[k].J(P.T);
  return J + P.w * P.T;
}

// ---- sequential kernel (any |U| <= kMaxU) ---------------------------------------------------
template <int DIM, int ORD, bool YAW>
__global__ void __launch_bounds__(kThreads)
expand_seq_kernel(const __grid_constant__ EnvParams P, const mplx_waypoint *__restrict__ nodes, int n_nodes,
                  int npb, const __grid_constant__ OutPtrs o) {
  const bool need_vel = YAW || (P.pot != nullptr && P.grad_w != 0.0);
  __shared__ uint32_t vbits[kMaxU / 32 + 9];
  __shared__ unsigned long long s_stats[2];
  const int nU = P.nU;
  const int items = npb * nU;
  const int node0 = blockIdx.x * npb;
  const int words = (items + 31) >> 5;
  if (threadIdx.x < 2) s_stats[threadIdx.x] = 0;
  for (int base = 0; base < items; base += kThreads) {
    PrimState<DIM, ORD, YAW> pr;
    bool emit, same;
    double max_v;
    size_t slot;
    phase_ab<DIM, ORD, YAW, true>(P, nodes, n_nodes, base + threadIdx.x, items, nU, node0, vbits, words, o, pr, emit,
                            same, max_v, slot);
    unsigned n_samples = 0;
    if (emit) {
      double cf[CoefLayout<DIM, ORD, YAW>::NCMAX];
      fill_coef<DIM, ORD, YAW>(pr, need_vel, cf);
      double cost = same ? 0.0 : traverse_loop<DIM, ORD, YAW>(P, cf, need_vel, max_v, n_samples);
      if (!isinf(cost)) cost += intrinsic_cost<DIM, ORD, YAW>(P, pr);
      if (o.cost) o.cost[slot] = cost;
    }
    if (P.stats) {
      atomicAdd(&s_stats[0], (unsigned long long)n_samples);
      if (emit) atomicAdd(&s_stats[1], 1ull);
    }
    __syncthreads();
  }
  if (threadIdx.x < 2 && P.stats) atomicAdd(&P.stats[threadIdx.x], s_stats[threadIdx.x]);
}

// ---- register kernel (|U| <= 256): thread = primitive through all three phases ----------------
// Phase C keeps the primitive's quotients in registers and walks the reference's own loop
// `for (t = 0; t < T; t += dt)` (env_map.h:99) four samples at a time: the four cell indices are
// computed and their voxel loads issued back to back, then the samples are classified IN ORDER,
// so the first blocking sample ends the primitive exactly where the reference returns inf, the
// potential / yaw sums accumulate in the reference's order, and up to three samples past a block
// are computed for nothing (they are bounds-checked like any other).  No shared-memory staging,
// no atomics; lanes whose primitive is invalid or short idle while the longest one finishes.
// Phase C of the register kernel: the reference's loop `for (t = 0; t < T; t += dt)`
// (env_map.h:99) in groups of UNR samples with group-level control flow only.  All UNR samples of
// a group are evaluated unconditionally (a sample past T or outside the map just gets no load),
// their voxel loads are issued back to back, and two decisions close the group:
//   some valid sample blocks  -> the primitive's cost is inf (the reference returns at the first
//                                such sample; later ones cannot change an inf);
//   the group reached t >= T  -> the loop has ended, return the accumulated cost;
// otherwise the terms of the group are added in sample order and the next group starts.
// n_samples counts what the reference's loop visits (up to and including the first blocking
// sample) and is only maintained when the stats counters are on.
template <int DIM, int ORD, bool YAW, int UNR>
__device__ __forceinline__ double traverse_groups(const EnvParams &P, const double (&cf)[CoefLayout<DIM, ORD, YAW>::NCMAX],
                                                 bool need_vel, double dt, unsigned &n_samples) {
  using CL = CoefLayout<DIM, ORD, YAW>;
  const double T = P.T;
  const int NC = CL::ncoef(need_vel);
  const bool plain = P.pot == nullptr && P.region_bits == nullptr && !YAW;
  double c = 0;
  double t = 0;
  for (;;) {
    double ts[UNR];
    int idx[UNR];
    bool valid[UNR];
#pragma unroll
    for (int j = 0; j < UNR; j++) {
      ts[j] = t;
      valid[j] = t < T;
      double pk[DIM];
      eval_pos<DIM, ORD>(cf, t, pk);
      idx[j] = sample_index<DIM>(P, pk);  // -1 when outside the map
      t += dt;                            // the reference's running sum
    }
    bool blocked[UNR];
    double term[UNR];
    if (plain) {
      // occupancy planning: the only question per sample is the voxel bit
      unsigned word[UNR];
#pragma unroll
      for (int j = 0; j < UNR; j++) {
        word[j] = 0;
        if (valid[j] && idx[j] >= 0) word[j] = __ldg(P.occ_bits + (idx[j] >> 5));
      }
#pragma unroll
      for (int j = 0; j < UNR; j++) {
        blocked[j] = idx[j] < 0 || ((word[j] >> (idx[j] & 31)) & 1u);
        term[j] = 0.0;
      }
    } else {
      VoxelRaw raw[UNR];
#pragma unroll
      for (int j = 0; j < UNR; j++) {
        raw[j] = kVoxelNone;
        if (valid[j] && idx[j] >= 0) raw[j] = voxel_fetch(P, idx[j]);
      }
#pragma unroll
      for (int j = 0; j < UNR; j++) {
        term[j] = 0.0;
        blocked[j] = idx[j] < 0;
        if (valid[j] && !blocked[j]) {
          double vel[DIM];
          double gterm = 0.0;
          if (need_vel) {
            eval_vel<DIM, ORD>(cf + CL::NCP, ts[j], vel);
            gterm = grad_term<DIM>(P, vel);
          }
          blocked[j] = voxel_classify(P, raw[j], dt, gterm, term[j]);
          if (YAW) {
            if (!blocked[j] && P.wyaw > 0)
              term[j] += yaw_term(P, vel[0], vel[1], normalize_angle(cf[NC - 2] * ts[j] + cf[NC - 1]), dt);
          }
        }
      }
    }
    bool any_blocked = false;
#pragma unroll
    for (int j = 0; j < UNR; j++) any_blocked = any_blocked || (valid[j] && blocked[j]);
    if (P.stats) {
      bool open = true;  // still before the first blocking sample
#pragma unroll
      for (int j = 0; j < UNR; j++) {
        if (open && valid[j]) n_samples++;
        open = open && !(valid[j] && blocked[j]);
      }
    }
    if (any_blocked) return INFINITY;
    if (!plain) {
#pragma unroll
      for (int j = 0; j < UNR; j++)
        if (valid[j]) c += term[j];
    }
    if (!valid[UNR - 1]) return c;  // this group contained the end of the loop
  }
}

// n = max(5, (int)ceil(max_v*T/res)), dt = T/n  (env_map.h:95,98): exact quotient + ceiling;
// T/n from the table for n <= kNMax, a true division beyond it.
__device__ __forceinline__ int sample_count_n(const EnvParams &P, double max_v, double &dt) {
  const double nd = ceil_exact(div_exact(max_v * P.T, P.res, P.rinv));
  const int n = nd < 5.0 ? 5 : (nd < 2.0e9 ? (int)nd : 2000000000);
  dt = n <= kNMax ? __ldg(P.tdt + n) : P.T / n;
  return n;
}

template <int DIM, int ORD, bool YAW, bool VEL, int UNR, int MINB, bool LAT>
__global__ void __launch_bounds__(kThreads, MINB)
expand_reg_kernel(const __grid_constant__ EnvParams P, const mplx_waypoint *__restrict__ nodes, int n_nodes,
                  int npb, const __grid_constant__ OutPtrs o) {
  __shared__ uint32_t vbits[9];
  __shared__ unsigned long long s_stats[2];
  const int nU = P.nU;
  const int items = npb * nU;  // <= 256
  const int node0 = blockIdx.x * npb;
  const int words = (items + 31) >> 5;
  if (threadIdx.x < 2) s_stats[threadIdx.x] = 0;
  PrimState<DIM, ORD, YAW> pr;
  bool emit, same;
  double max_v;
  size_t slot;
  phase_ab<DIM, ORD, YAW, LAT>(P, nodes, n_nodes, threadIdx.x, items, nU, node0, vbits, words, o, pr, emit, same,
                               max_v, slot);
  unsigned n_samples = 0;
  if (emit) {
    double cost = 0.0;
    const double intrinsic = intrinsic_cost<DIM, ORD, YAW>(P, pr);  // before the loop: pr dies here
    if (!same) {
      double cf[CoefLayout<DIM, ORD, YAW>::NCMAX];
      fill_coef<DIM, ORD, YAW>(pr, VEL, cf);
      double dt;
      sample_count_n(P, max_v, dt);
      cost = traverse_groups<DIM, ORD, YAW, UNR>(P, cf, VEL, dt, n_samples);
    }
    if (!isinf(cost)) cost += intrinsic;
    if (o.cost) o.cost[slot] = cost;
  }
  if (P.stats) {
    atomicAdd(&s_stats[0], (unsigned long long)n_samples);
    if (emit) atomicAdd(&s_stats[1], 1ull);
    __syncthreads();
    if (threadIdx.x < 2) atomicAdd(&P.stats[threadIdx.x], s_stats[threadIdx.x]);
  }
}

// ---- flat kernel (|U| <= 256) -----------------------------------------------------------------
// Per-warp shared-memory slab (doubles first so everything stays 8-byte aligned):
//   coef [32][NC]  loop-invariant polynomial quotients of each lane's primitive
//   cost [32]      accumulated potential / yaw cost        dt [32]  T/n
//   n[32] first[32]  (int)        owner[32*maxns] (uint16: lane<<8 | sample k; k <= kNMax < 256)
template <int DIM, int ORD, bool YAW>
struct FlatLayout : CoefLayout<DIM, ORD, YAW> {
  using CoefLayout<DIM, ORD, YAW>::ncoef;
  __host__ __device__ static size_t warp_bytes(bool need_vel, int maxns) {
    size_t b = (size_t)32 * ncoef(need_vel) * 8 + 32 * 8 * 2 + 32 * 4 * 2 + (size_t)32 * maxns * 2;
    return (b + 15) & ~(size_t)15;
  }
};

template <int DIM, int ORD, bool YAW>
__global__ void __launch_bounds__(kThreads, 4)
expand_flat_kernel(const __grid_constant__ EnvParams P, const mplx_waypoint *__restrict__ nodes, int n_nodes,
                   int npb, const __grid_constant__ OutPtrs o, int maxns, int need_vel_i) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t vbits[9];
  __shared__ unsigned long long s_stats[2];
  using L = FlatLayout<DIM, ORD, YAW>;
  const bool need_vel = need_vel_i != 0;
  const int NC = L::ncoef(need_vel);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wb = smem + (size_t)warp * L::warp_bytes(need_vel, maxns);
  double *w_coef = reinterpret_cast<double *>(wb);
  double *w_cost = w_coef + 32 * NC;
  double *w_dt = w_cost + 32;
  int *w_n = reinterpret_cast<int *>(w_dt + 32);
  int *w_first = w_n + 32;
  unsigned short *w_owner = reinterpret_cast<unsigned short *>(w_first + 32);

  const int nU = P.nU;
  const int items = npb * nU;  // <= 256
  const int node0 = blockIdx.x * npb;
  const int words = (items + 31) >> 5;
  if (threadIdx.x < 2) s_stats[threadIdx.x] = 0;

  PrimState<DIM, ORD, YAW> pr;
  bool emit, same;
  double max_v;
  size_t slot;
  phase_ab<DIM, ORD, YAW, true>(P, nodes, n_nodes, threadIdx.x, items, nU, node0, vbits, words, o, pr, emit, same,
                                max_v, slot);

  // ---- phase C set-up: coefficient slot, n, sample count ----
  const double T = P.T;
  fill_coef<DIM, ORD, YAW>(pr, need_vel, w_coef + lane * NC);
  int n = 0, ns = 0;
  bool seq = false;
  double cost_seq = 0.0;
  unsigned seq_samples = 0;
  if (emit && !same) {
    // n = max(5, (int)ceil(max_v*T/res))  (env_map.h:95), exact quotient and ceiling
    const double nd = ceil_exact(div_exact(max_v * T, P.res, P.rinv));
    if (nd <= (double)P.maxn) {
      n = max(5, (int)nd);
      ns = __ldg(P.tcount + n);
    } else {
      seq = true;  // beyond the table: literal loop in this lane, coefficients from its smem slot
      cost_seq = traverse_loop_cold<DIM, ORD, YAW>(&P, w_coef + lane * NC, need_vel, max_v, &seq_samples);
    }
  }
  int incl = ns;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  const int start = incl - ns;
  const int S = __shfl_sync(0xffffffffu, incl, 31);
  w_cost[lane] = 0.0;
  w_dt[lane] = __ldg(P.tdt + n);  // T/n, env_map.h:98
  w_n[lane] = n;
  w_first[lane] = kNoBlock;
  for (int k = 0; k < ns; k++) w_owner[start + k] = (unsigned short)((lane << 8) | k);
  __syncwarp();

  // ---- phase C: the warp's samples, dealt round-robin to its lanes, two per lane per trip ----
  // prep: owner/time/coefficients -> cell index (or -1 outside, -2 nothing to do);
  // then the voxel loads of both samples are issued before either is consumed.
  struct Prep {
    int i, k, idx;
    double t;
  };
  auto prep = [&](int s, bool valid) {
    Prep r;
    const unsigned ow = valid ? (unsigned)w_owner[s] : 0u;
    r.i = (int)(ow >> 8);
    r.k = (int)(ow & 255u);
    r.idx = -2;
    r.t = 0.0;
    // an earlier sample of this primitive already blocks: the result is inf whatever this one says
    if (!valid || *(volatile int *)(w_first + r.i) < r.k) return r;
    r.t = __ldg(P.ttab + w_n[r.i] * kTStride + r.k);
    double pk[DIM];
    eval_pos<DIM, ORD>(w_coef + r.i * NC, r.t, pk);
    r.idx = sample_index<DIM>(P, pk);
    return r;
  };
  auto finish = [&](const Prep &r, VoxelRaw raw) {
    if (r.idx == -2) return;
    if (r.idx < 0) {
      atomicMin(w_first + r.i, r.k);
      return;
    }
    const double *cf = w_coef + r.i * NC;
    double vel[DIM];
    double gterm = 0.0;
    if (need_vel) {
      eval_vel<DIM, ORD>(cf + L::NCP, r.t, vel);
      gterm = grad_term<DIM>(P, vel);
    }
    const double dt = w_dt[r.i];
    double term = 0.0;
    if (voxel_classify(P, raw, dt, gterm, term)) {
      atomicMin(w_first + r.i, r.k);
      return;
    }
    if (YAW) {
      if (P.wyaw > 0) term += yaw_term(P, vel[0], vel[1], normalize_angle(cf[NC - 2] * r.t + cf[NC - 1]), dt);
    }
    if (term != 0.0) atomicAdd(w_cost + r.i, term);
  };
  for (int s = lane; s < S; s += 64) {
    const Prep a = prep(s, true);
    const Prep b = prep(s + 32, s + 32 < S);
    VoxelRaw ra = kVoxelNone, rb = kVoxelNone;
    if (a.idx >= 0) ra = voxel_fetch(P, a.idx);
    if (b.idx >= 0) rb = voxel_fetch(P, b.idx);
    finish(a, ra);
    finish(b, rb);
  }
  __syncwarp();

  if (emit) {
    const int fb = w_first[lane];
    double cost = same ? 0.0 : seq ? cost_seq : (fb != kNoBlock ? (double)INFINITY : w_cost[lane]);
    if (!isinf(cost)) cost += intrinsic_cost<DIM, ORD, YAW>(P, pr);
    if (o.cost) o.cost[slot] = cost;
    if (P.stats) {
      // samples the reference loop visits: up to and including the first blocking one
      const unsigned visited = seq ? seq_samples : (fb != kNoBlock ? (unsigned)fb + 1u : (unsigned)ns);
      atomicAdd(&s_stats[0], (unsigned long long)visited);
      atomicAdd(&s_stats[1], 1ull);
    }
  }
  if (P.stats) {
    __syncthreads();
    if (threadIdx.x < 2) atomicAdd(&P.stats[threadIdx.x], s_stats[threadIdx.x]);
  }
}

template <int DIM, int ORD, bool YAW>
static cudaError_t launch_t(const EnvParams &P, const mplx_waypoint *d_nodes, int n_nodes,
                            const mplx_succ_out &so, cudaStream_t st, int force_seq) {
  const OutPtrs o{so.count, so.succ, so.cost, so.action, so.key, so.lattice};
  const int npb = P.nU >= kThreads ? 1 : kThreads / P.nU;
  const int grid = (n_nodes + npb - 1) / npb;
  if (P.nU > kThreads || force_seq == 1) {
    expand_seq_kernel<DIM, ORD, YAW><<<grid, kThreads, 0, st>>>(P, d_nodes, n_nodes, npb, o);
    return cudaGetLastError();
  }
  const bool nv = YAW || (P.pot != nullptr && P.grad_w != 0.0);
  if (force_seq != 3) {  // register kernel (default)
    // samples in flight per lane: 4, or 2 when every primitive of the plan has a short loop
    // (n <= 15: a group of 4 would mostly run past the end of the loop)
    const bool short_loops = P.maxn <= 15;
    const bool lat = o.lattice != nullptr;
#define MPLX_LAUNCH_REG(VEL, UNR, LAT) \
  expand_reg_kernel<DIM, ORD, YAW, VEL, UNR, 4, LAT><<<grid, kThreads, 0, st>>>(P, d_nodes, n_nodes, npb, o)
    if (nv) {
      if (short_loops) { if (lat) MPLX_LAUNCH_REG(true, 2, true); else MPLX_LAUNCH_REG(true, 2, false); }
      else { if (lat) MPLX_LAUNCH_REG(true, 4, true); else MPLX_LAUNCH_REG(true, 4, false); }
    } else {
      if (short_loops) { if (lat) MPLX_LAUNCH_REG(YAW, 2, true); else MPLX_LAUNCH_REG(YAW, 2, false); }
      else { if (lat) MPLX_LAUNCH_REG(YAW, 4, true); else MPLX_LAUNCH_REG(YAW, 4, false); }
    }
#undef MPLX_LAUNCH_REG
    return cudaGetLastError();
  }
  using L = FlatLayout<DIM, ORD, YAW>;
  const bool need_vel = YAW || (P.pot != nullptr && P.grad_w != 0.0);
  const int maxns = P.maxn + 1;
  const size_t smem = kWarps * L::warp_bytes(need_vel, maxns);
  static thread_local size_t configured = 0;  // per template instantiation and host thread
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(expand_flat_kernel<DIM, ORD, YAW>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  expand_flat_kernel<DIM, ORD, YAW><<<grid, kThreads, smem, st>>>(P, d_nodes, n_nodes, npb, o, maxns,
                                                                   need_vel ? 1 : 0);
  return cudaGetLastError();
}

template <int DIM>
static cudaError_t launch_d(const EnvParams &P, const mplx_waypoint *d_nodes, int n_nodes,
                            const mplx_succ_out &o, cudaStream_t st, int fs) {
  const bool yaw = (P.control & 16) != 0;
  switch (P.control & 15) {
    case MPLX_VEL: return yaw ? launch_t<DIM, 1, true>(P, d_nodes, n_nodes, o, st, fs) : launch_t<DIM, 1, false>(P, d_nodes, n_nodes, o, st, fs);
    case MPLX_ACC: return yaw ? launch_t<DIM, 2, true>(P, d_nodes, n_nodes, o, st, fs) : launch_t<DIM, 2, false>(P, d_nodes, n_nodes, o, st, fs);
    case MPLX_JRK: return yaw ? launch_t<DIM, 3, true>(P, d_nodes, n_nodes, o, st, fs) : launch_t<DIM, 3, false>(P, d_nodes, n_nodes, o, st, fs);
    case MPLX_SNP: return yaw ? launch_t<DIM, 4, true>(P, d_nodes, n_nodes, o, st, fs) : launch_t<DIM, 4, false>(P, d_nodes, n_nodes, o, st, fs);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_expand(const EnvParams &P, const mplx_waypoint *d_nodes, int n_nodes,
                          const mplx_succ_out &o, cudaStream_t st, int force_seq) {
  if (n_nodes <= 0) return cudaSuccess;
  return P.dim == 2 ? launch_d<2>(P, d_nodes, n_nodes, o, st, force_seq)
                    : launch_d<3>(P, d_nodes, n_nodes, o, st, force_seq);
}

// ---- set-up kernels -------------------------------------------------------------------------

// std::vector<bool> search_region_ (env_base.h:400) arrives as one byte per voxel; the grid
// arrives as int8.  Both are packed to 1 bit per voxel (OCC: bit = (map == 100), isOccupied
// map_util.h:48) so the per-sample test is a 4-byte read-only load from an L2-resident array.
template <bool OCC>
__global__ void pack_bits_kernel(const int8_t *__restrict__ bytes, size_t nvox, uint32_t *__restrict__ bits) {
  const size_t nwords = (nvox + 31) >> 5;
  for (size_t wd = (size_t)blockIdx.x * blockDim.x + threadIdx.x; wd < nwords;
       wd += (size_t)gridDim.x * blockDim.x) {
    uint32_t m = 0;
    const size_t b0 = wd << 5;
#pragma unroll 8
    for (int b = 0; b < 32; b++) {
      const size_t i = b0 + b;
      if (i < nvox && (OCC ? bytes[i] == 100 : bytes[i] != 0)) m |= 1u << b;
    }
    bits[wd] = m;
  }
}

cudaError_t launch_pack_bits(const int8_t *d_bytes, size_t nvox, uint32_t *d_bits, bool occ, cudaStream_t st) {
  const size_t nwords = (nvox + 31) >> 5;
  int grid = (int)((nwords + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  if (occ)
    pack_bits_kernel<true><<<grid, 256, 0, st>>>(d_bytes, nvox, d_bits);
  else
    pack_bits_kernel<false><<<grid, 256, 0, st>>>(d_bytes, nvox, d_bits);
  return cudaGetLastError();
}

// Sample-time table: thread n runs the reference loop `for (t = 0; t < T; t += T/n)`
// (env_map.h:98-99) once and records every t_k and the iteration count.
__global__ void build_ttab_kernel(double T, double *__restrict__ ttab, int *__restrict__ tcount,
                                  double *__restrict__ tdt) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n > kNMax) return;
  if (n < 1) {
    tcount[n] = 0;
    tdt[n] = 0.0;
    return;
  }
  const double dt = T / n;
  tdt[n] = dt;
  int k = 0;
  for (double t = 0; t < T; t += dt) {
    if (k < kTStride) ttab[n * kTStride + k] = t;
    k++;
  }
  tcount[n] = k;
}

cudaError_t launch_build_ttab(double T, double *d_ttab, int *d_tcount, double *d_tdt, cudaStream_t st) {
  build_ttab_kernel<<<(kNMax + 1 + 127) / 128, 128, 0, st>>>(T, d_ttab, d_tcount, d_tdt);
  return cudaGetLastError();
}

}  // namespace mplx
