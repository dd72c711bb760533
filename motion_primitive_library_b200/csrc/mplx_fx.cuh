// mplx_fx.cuh — the fixed-point cell evaluation shared by the occupancy-planning kernels
// (mplx_fx.cu: per-thread phase A with a helper warp; mplx_fxn.cu: node-cooperative rows + flat
// sample items).  See mplx_fx.cu for the derivation and the certainty rule.
#pragma once
#include "mplx_expand.cuh"

namespace mplx {

constexpr double kFxMagic = 1572864.0;   // 1.5 * 2^20: ulp(2^20..2^21) = 2^-32
constexpr int kFxHiBase = 0x41380000;    // high word of kFxMagic; + floor(y) for |y| < 2^19
constexpr double kFxEps = 0x1p-26;       // certainty margin in cells (error bound 2^-30)
constexpr unsigned kFxUnc = 128u;        // 2*eps in units of 2^-32
constexpr double kFxRange = 131072.0;    // 2^17: start coordinates (cells) the error bound covers (see the kernel)

// Cell-unit coefficients of one axis: y(t) = C[ORD] t^ORD + .. + C[1] t + C[0], C[0] carrying
// -origin/res + eps + magic.
template <int ORD>
__device__ __forceinline__ void fx_axis(const Axis<ORD> &ax, double origin, double rinv, double (&C)[ORD + 1]) {
  // position quotients as Primitive1D::p uses them (primitive.h:128-131): c1/24 c2/6 c3/2 c4 c5
  double q[5];
  q[0] = ax.c5;
  q[1] = ax.c4;
  q[2] = ax.c3 / 2;
  q[3] = ax.c2 / 6;
  q[4] = ax.c1 / 24;
#pragma unroll
  for (int i = 1; i <= ORD; i++) C[i] = q[i] * rinv;
  C[0] = (q[0] - origin) * rinv + (kFxMagic + kFxEps);
}

// One group of UNR samples.  `left` as in sample_group.  Returns 2 when a CERTAIN sample blocks,
// else 1 when the group held the loop's end, else 0; bit j of `amb` = sample j is ambiguous.
// Per sample one 32-bit word is loaded: the occupancy word of the cell when the sample is certain,
// the candidate-summary word when it is uncertain (the two are interleaved, mplx_device.cuh occ2);
// a sample outside the map reads as all-ones (blocked / never "all candidates free").  The word is
// rotated so that the cell's bit lands on bit j, and the group is decided on the OR of those bits.
__device__ __forceinline__ unsigned rotr_wrap(unsigned x, unsigned s) {
  unsigned r;
  asm("shf.r.wrap.b32 %0, %1, %1, %2;" : "=r"(r) : "r"(x), "r"(s));  // only the low 5 bits of s count
  return r;
}

template <int DIM, int ORD, int UNR, bool REGION>
__device__ __forceinline__ int fx_group(const EnvParams &P, const unsigned *__restrict__ base,
                                        const double (&C)[DIM][ORD + 1], double dt, int left, double &t, unsigned &amb) {
  unsigned w[UNR], rot[UNR];
  unsigned uncm = 0;
#pragma unroll
  for (int j = 0; j < UNR; j++) {
    int cell[DIM];
    unsigned lo[DIM];
    bool inside = true;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      double h = C[a][ORD];
#pragma unroll
      for (int i = ORD - 1; i >= 1; i--) h = __fma_rn(h, t, C[a][i]);
      const double m = __fma_rn(h, t, C[a][0]);
      cell[a] = __double2hiint(m) - kFxHiBase;
      lo[a] = (unsigned)__double2loint(m);
      inside = inside && ((unsigned)cell[a] < (unsigned)P.mdim[a]);
    }
    const unsigned fr = DIM == 3 ? min(min(lo[0], lo[1]), lo[DIM - 1]) : min(lo[0], lo[1]);
    int idx = cell[0] + P.mdim[0] * cell[1];
    if (DIM == 3) idx += P.mdim[0] * P.mdim[1] * cell[DIM - 1];
    const unsigned ub = fr < kFxUnc ? 1u : 0u;
    uncm += ub << j;
    rot[j] = (unsigned)(idx - j);  // rotate right by it: the cell's bit lands on bit j
    w[j] = 0xffffffffu;
    // a sample past the loop's end (j >= left) may be loaded too: its bit is masked below
    if (REGION) {
      // no candidate summary for the tunnel: an uncertain sample is ambiguous (word stays all-ones);
      // a certain one is blocked when occupied or outside the tunnel (env_map.h:104-106)
      if (inside && !ub) {
        const unsigned wi = (unsigned)idx >> 5;
        w[j] = __ldg(base + 2 * wi) | ~__ldg(P.region_bits + wi);
      }
    } else {
      if (inside) w[j] = __ldg(base + ((((unsigned)idx >> 4) & ~1u) | ub));
    }
    t += dt;  // the reference's running sum (env_map.h:99)
  }
  unsigned r = 0;
#pragma unroll
  for (int j = 0; j < UNR; j++) r |= rotr_wrap(w[j], rot[j]) & (1u << j);
  if (left < UNR) r &= (1u << left) - 1u;
  amb = r & uncm;
  if (r & ~uncm) return 2;
  return left <= UNR ? 1 : 0;
}

// The two halves of fx_group for a software-pipelined loop: fx_issue evaluates the UNR cells starting
// at time t and issues their word loads; fx_decide consumes the words.  A caller that issues group g+1
// before deciding group g hides the L2 latency of the voxel words behind the next group's arithmetic.
template <int UNR>
struct FxWords {
  unsigned w[UNR], rot[UNR];
  unsigned uncm;
};

template <int DIM, int ORD, int UNR, bool REGION>
__device__ __forceinline__ void fx_issue(const EnvParams &P, const unsigned *__restrict__ base,
                                         const double (&C)[DIM][ORD + 1], double dt, double &t, FxWords<UNR> &G) {
  G.uncm = 0;
#pragma unroll
  for (int j = 0; j < UNR; j++) {
    int cell[DIM];
    unsigned lo[DIM];
    bool inside = true;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
      double h = C[a][ORD];
#pragma unroll
      for (int i = ORD - 1; i >= 1; i--) h = __fma_rn(h, t, C[a][i]);
      const double m = __fma_rn(h, t, C[a][0]);
      cell[a] = __double2hiint(m) - kFxHiBase;
      lo[a] = (unsigned)__double2loint(m);
      inside = inside && ((unsigned)cell[a] < (unsigned)P.mdim[a]);
    }
    const unsigned fr = DIM == 3 ? min(min(lo[0], lo[1]), lo[DIM - 1]) : min(lo[0], lo[1]);
    int idx = cell[0] + P.mdim[0] * cell[1];
    if (DIM == 3) idx += P.mdim[0] * P.mdim[1] * cell[DIM - 1];
    const unsigned ub = fr < kFxUnc ? 1u : 0u;
    G.uncm += ub << j;
    G.rot[j] = (unsigned)(idx - j);
    G.w[j] = 0xffffffffu;
    if (REGION) {
      if (inside && !ub) {
        const unsigned wi = (unsigned)idx >> 5;
        G.w[j] = __ldg(base + 2 * wi) | ~__ldg(P.region_bits + wi);
      }
    } else {
      if (inside) G.w[j] = __ldg(base + ((((unsigned)idx >> 4) & ~1u) | ub));
    }
    t += dt;  // the reference's running sum (env_map.h:99)
  }
}

template <int UNR>
__device__ __forceinline__ int fx_decide(const FxWords<UNR> &G, int left, unsigned &amb) {
  unsigned r = 0;
#pragma unroll
  for (int j = 0; j < UNR; j++) r |= rotr_wrap(G.w[j], G.rot[j]) & (1u << j);
  if (left < UNR) r &= (1u << left) - 1u;
  amb = r & G.uncm;
  if (r & ~G.uncm) return 2;
  return left <= UNR ? 1 : 0;
}

// The whole sample loop of one primitive, software-pipelined two groups deep.  Returns 1 when a
// certain sample blocks, else 0 with the ambiguous samples in amask (k < 64) / full (some k >= 64).
template <int DIM, int ORD, int UNR, bool REGION>
__device__ __forceinline__ int fx_traverse(const EnvParams &P, const double (&C)[DIM][ORD + 1], double dt, int count,
                                           unsigned long long &amask, bool &full) {
  const unsigned *__restrict__ base = reinterpret_cast<const unsigned *>(P.occ2);
  amask = 0;
  full = false;
  double t = 0;
  int left = count, k0 = 0;
  FxWords<UNR> A, B;
  fx_issue<DIM, ORD, UNR, REGION>(P, base, C, dt, t, A);
  for (;;) {
    unsigned amb;
    int st;
    // ---- group in A; group after it goes to B ----
    if (left > UNR) fx_issue<DIM, ORD, UNR, REGION>(P, base, C, dt, t, B);
    st = fx_decide<UNR>(A, left, amb);
    if (st == 2) return 1;
    if (amb) {
      if (k0 + UNR <= 64) amask |= (unsigned long long)amb << k0; else full = true;
    }
    if (st == 1) return 0;
    left -= UNR;
    k0 += UNR;
    // ---- group in B; group after it goes to A ----
    if (left > UNR) fx_issue<DIM, ORD, UNR, REGION>(P, base, C, dt, t, A);
    st = fx_decide<UNR>(B, left, amb);
    if (st == 2) return 1;
    if (amb) {
      if (k0 + UNR <= 64) amask |= (unsigned long long)amb << k0; else full = true;
    }
    if (st == 1) return 0;
    left -= UNR;
    k0 += UNR;
  }
}

// hash_value(curr) once per node instead of once per (node, control): waypoint.h:93-125.
template <int DIM, int ORD>
__device__ __forceinline__ uint64_t curr_hash(const mplx_waypoint *cp) {
  uint64_t h = 0;
#pragma unroll
  for (int k = 0; k < DIM; k++) {
    hash_combine(h, lattice_id(cp->pos[k], 0.01, 100.0));
    if (ORD >= 2) hash_combine(h, lattice_id(cp->vel[k], 0.1, 10.0));
    if (ORD >= 3) hash_combine(h, lattice_id(cp->acc[k], 0.1, 10.0));
    if (ORD >= 4) hash_combine(h, lattice_id(cp->jrk[k], 0.1, 10.0));
  }
  return h;
}

constexpr unsigned kFxSegments = 64;  // the global queue is cut in segments (one counter each) to spread the atomics

// One ambiguous primitive handed to the exact re-evaluation kernel (mplx_fxn.cu).
struct FxAmbRec {
  unsigned slot;             // output slot of the successor
  int node;                  // frontier index
  unsigned short action;     // control index
  unsigned char n;           // max(5, ceil(max_v*T/res)) <= kNMax
  unsigned char full;        // 1: every sample of the loop, 0: the samples of `amask`
  unsigned long long amask;  // ambiguous samples k < 64
};

}  // namespace mplx
