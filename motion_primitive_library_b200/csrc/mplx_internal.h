// mplx_internal.h — ctx layout and small helpers shared by the libmplx translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mplx.h"
#include "mplx_device.cuh"
#include "mplx_kernels.h"

namespace mplx {
int fail(int code, const char *fmt, ...);
}
using mplx::fail;

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t e_ = (call);                                                                  \
    if (e_ != cudaSuccess) {                                                                  \
      cudaGetLastError();                                                                     \
      return fail(e_ == cudaErrorMemoryAllocation ? MPLX_ERR_ALLOC : MPLX_ERR_CUDA,           \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    }                                                                                         \
  } while (0)

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void **)&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

// a function-local DevBuf that frees itself on every exit path
template <typename T>
struct ScopedDevBuf : DevBuf<T> {
  ScopedDevBuf() = default;
  ScopedDevBuf(const ScopedDevBuf &) = delete;
  ScopedDevBuf &operator=(const ScopedDevBuf &) = delete;
  ~ScopedDevBuf() { this->release(); }
};

template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaHostAlloc((void **)&p, n * sizeof(T), cudaHostAllocDefault);
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

inline bool is_pinned(const void *p) {
  if (!p) return false;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Global queue of ambiguous primitives of the fixed-point kernels (mplx_fxn.cu), one per stream in use.
struct FxQueue {
  DevBuf<unsigned char> q;
  DevBuf<unsigned> n;
  mplx::FxScratch view;
  // room for a quarter of the batch's primitives (~5 % are ambiguous on the bench maps) + slack per segment
  cudaError_t reserve(size_t n_prims) {
    const size_t cap = ((n_prims / 4 + 64 * 4096 + 63) / 64) * 64;
    cudaError_t e = q.reserve(cap * mplx::fx_amb_record_bytes());
    if (e == cudaSuccess) e = n.reserve(64);
    if (e != cudaSuccess) return e;
    view.q = q.p;
    view.n = n.p;
    view.cap = (unsigned)(q.cap / mplx::fx_amb_record_bytes() / 64 * 64);
    return cudaSuccess;
  }
  void release() { q.release(); n.release(); view = mplx::FxScratch(); }
};

// One set of per-chunk device buffers + its stream: two sets let chunk k+1 compute while
// chunk k's results cross PCIe (mplx_expand_packed).
constexpr int kPackBufs = 4;  // chunk buffer sets of the packed pipeline
struct ChunkBufs {
  cudaStream_t st = nullptr;
  cudaEvent_t ready = nullptr;    // the chunk's records are packed and its record count is on the host
  cudaEvent_t drained = nullptr;  // the chunk's device-to-host copies have left the buffers
  DevBuf<mplx_waypoint> nodes, succ;
  DevBuf<int32_t> count, action;
  DevBuf<double> cost;
  DevBuf<uint64_t> key;
  // packed stream
  DevBuf<int32_t> kcount;
  DevBuf<long long> offset, total;
  DevBuf<double> pstate, pcost;
  DevBuf<uint16_t> paction;
  DevBuf<uint64_t> pkey;
  PinBuf<long long> h_total;
  FxQueue fxq;
  void release() {
    fxq.release();
    nodes.release(); succ.release(); count.release(); action.release(); cost.release(); key.release();
    kcount.release(); offset.release(); total.release(); pstate.release(); pcost.release(); paction.release();
    pkey.release(); h_total.release();
    if (ready) cudaEventDestroy(ready);
    if (drained) cudaEventDestroy(drained);
    if (st) cudaStreamDestroy(st);
    ready = drained = nullptr;
    st = nullptr;
  }
};

// staging of the edge re-validation queries (mplx_edges.cu)
struct EdgeBufs {
  DevBuf<mplx_waypoint> parents;
  DevBuf<int32_t> actions, cells, ids, owner, ids_sorted, owner_sorted;
  DevBuf<uint8_t> free_, scan_tmp;
  DevBuf<double> cost;
  DevBuf<long long> count, offset;
  void release() {
    parents.release(); actions.release(); cells.release(); free_.release(); scan_tmp.release(); cost.release();
    count.release(); offset.release(); ids.release(); owner.release(); ids_sorted.release(); owner_sorted.release();
  }
};

struct mplx_ctx {
  int dim = 0, device = 0;
  cudaStream_t stream = nullptr;
  // static data in HBM
  DevBuf<int8_t> map, pot;
  DevBuf<uint32_t> region, occ;
  DevBuf<uint2> occ2;
  DevBuf<unsigned char> prow, row_axis;  // per-axis value tables of U (EnvParams::prow ...)
  DevBuf<double> row_u;
  int n_rows = 0;
  size_t occ2_window = 0;  // bytes of occ2 covered by the L2 access-policy window (0 = none)  // {occupancy, candidate summary} words of the fixed-point kernel
  DevBuf<double> U, ttab, tdt;
  DevBuf<int> tcount;
  int force_seq = 0;
  DevBuf<unsigned long long> stats;
  bool has_map = false, has_pot = false, has_region = false, has_params = false, stats_on = false;
  size_t nvox = 0;
  mplx::EnvParams P;
  // per-call staging (host-buffer entry point)
  DevBuf<mplx_waypoint> d_nodes, d_succ;
  DevBuf<int32_t> d_count, d_action, d_lattice;
  DevBuf<double> d_cost;
  DevBuf<uint64_t> d_key;
  PinBuf<mplx_waypoint> h_nodes, h_succ;
  PinBuf<int32_t> h_count, h_action, h_lattice;
  PinBuf<double> h_cost;
  PinBuf<uint64_t> h_key;
  ChunkBufs cb[kPackBufs];
  cudaStream_t d2h_stream = nullptr;  // result copies of the packed pipeline
  FxQueue fxq;
  EdgeBufs eb;
  int64_t launches = 0;
  unsigned long long last_stats[2] = {0, 0};
};

extern "C" {
int mplx_bind(mplx_ctx *ctx);
int mplx_check_ready(mplx_ctx *c, int n_nodes);
}
