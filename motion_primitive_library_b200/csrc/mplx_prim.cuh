// mplx_prim.cuh — the primitive in registers and its sample evaluators, shared by the expansion
// kernels (mplx_kernels.cu) and the edge re-validation kernels (mplx_edges.cu).
#pragma once
#include "mplx_device.cuh"

namespace mplx {

template <int DIM, int ORD, bool YAW>
struct PrimState {
  Axis<ORD> ax[DIM];
  double yaw_u, yaw0;  // pr_yaw_ = [0,0,0,0,u(Dim),yaw]  (primitive.h:34,235-248)
};

// Loop-invariant coefficient layout shared by both sample loops ("cf"):
//   [axis 0..DIM-1: position quotients  (c1/24) (c2/6) (c3/2) c4 c5   — the last ORD+1 of them]
//   [axis 0..DIM-1: velocity quotients  (c1/6) (c2/2) c3 c4           — the last ORD, only if need_vel]
//   [yaw_u, yaw0                                                      — only if YAW]
template <int DIM, int ORD, bool YAW>
struct CoefLayout {
  static constexpr int NCP = DIM * (ORD + 1);
  static constexpr int NCV = DIM * ORD;
  static constexpr int NCMAX = NCP + NCV + (YAW ? 2 : 0);
  __host__ __device__ static int ncoef(bool need_vel) { return NCP + (need_vel ? NCV : 0) + (YAW ? 2 : 0); }
};

template <int DIM, int ORD, bool YAW>
__device__ __forceinline__ void fill_coef(const PrimState<DIM, ORD, YAW> &pr, bool need_vel, double *cf) {
  int c = 0;
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    if (ORD >= 4) cf[c++] = pr.ax[k].c1 / 24;
    if (ORD >= 3) cf[c++] = pr.ax[k].c2 / 6;
    if (ORD >= 2) cf[c++] = pr.ax[k].c3 / 2;
    cf[c++] = pr.ax[k].c4;
    cf[c++] = pr.ax[k].c5;
  }
  if (need_vel) {
#pragma unroll
    for (int k = 0; k < DIM; k++) {
      if (ORD >= 4) cf[c++] = pr.ax[k].c1 / 6;
      if (ORD >= 3) cf[c++] = pr.ax[k].c2 / 2;
      if (ORD >= 2) cf[c++] = pr.ax[k].c3;
      cf[c++] = pr.ax[k].c4;
    }
  }
  if (YAW) {
    cf[c++] = pr.yaw_u;
    cf[c++] = pr.yaw0;
  }
}

// Primitive1D::p (primitive.h:128-131) for all axes at time t, from the quotients.
template <int DIM, int ORD>
__device__ __forceinline__ void eval_pos(const double *cf, double t, double (&pk)[DIM]) {
  const double pw3 = (t * t) * t;
  const double pw4 = pw3 * t;
  int c = 0;
#pragma unroll
  for (int a = 0; a < DIM; a++) {
    double acc;
    if (ORD == 1) {
      acc = cf[c] * t + cf[c + 1];
    } else if (ORD == 2) {
      acc = cf[c] * t * t + cf[c + 1] * t + cf[c + 2];
    } else if (ORD == 3) {
      acc = cf[c] * pw3 + cf[c + 1] * t * t + cf[c + 2] * t + cf[c + 3];
    } else {
      acc = cf[c] * pw4 + cf[c + 1] * pw3 + cf[c + 2] * t * t + cf[c + 3] * t + cf[c + 4];
    }
    pk[a] = acc;
    c += ORD + 1;
  }
}

// Primitive1D::v (primitive.h:134-137) for all axes at time t (cf points at the velocity block).
template <int DIM, int ORD>
__device__ __forceinline__ void eval_vel(const double *cf, double t, double (&vel)[DIM]) {
  const double pw3 = (t * t) * t;
  int c = 0;
#pragma unroll
  for (int a = 0; a < DIM; a++) {
    double acc = 0.0;
    if (ORD == 1) {
      acc = acc + cf[c];
    } else if (ORD == 2) {
      acc = acc + cf[c] * t + cf[c + 1];
    } else if (ORD == 3) {
      acc = acc + cf[c] * t * t + cf[c + 1] * t + cf[c + 2];
    } else {
      acc = acc + cf[c] * pw3 + cf[c + 1] * t * t + cf[c + 2] * t + cf[c + 3];
    }
    vel[a] = acc;
    c += ORD;
  }
}

}  // namespace mplx
