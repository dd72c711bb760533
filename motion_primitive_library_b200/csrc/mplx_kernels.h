// mplx_kernels.h — host-callable launchers of the sm_100a kernels (internal).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/mplx.h"

namespace mplx {
struct EnvParams;
constexpr int kMaxU = 1024;  // |U| upper bound (125 is the largest set the reference's users build)

cudaError_t launch_expand(const EnvParams &P, const mplx_waypoint *d_nodes, int n_nodes,
                          const mplx_succ_out &o, cudaStream_t st);
cudaError_t launch_pack_region(const uint8_t *d_bytes, size_t nvox, uint32_t *d_bits, cudaStream_t st);
}  // namespace mplx
