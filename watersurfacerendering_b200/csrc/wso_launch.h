// wso_launch.h — host-side launch interface of the kernels in wso_kernels.cu.
#pragma once
#include <cuda_runtime.h>

namespace wso {

template <int CAP>
struct LaunchArgsT;
using LaunchArgs = LaunchArgsT<64>;  // kMaxChunk (wso_kernels.cuh)

static constexpr int kMinLogN = 4;   // 16
static constexpr int kMaxLogN = 13;  // 8192 (single-CTA line transforms; larger grids: slab path)

// Enqueue K1 (evolve + first transform), K2h (height extrema) and K2 (second transform + pack) for
// n_items tile-frames described by args.items[0..n_items).
// ev: NULL, or 4 events recorded before K1, between the kernels and after K2 (opt-in profiling).
cudaError_t launch_compute_waves(int logn, const LaunchArgs& args, int n_items, cudaStream_t stream,
                                 bool first_use, cudaEvent_t* ev);
int kernels_per_launch();

// Slab-decomposed path (one grid over several devices).  phase 0: K1 on this device's column pairs,
// 1: K2h on its row items, 2: K2 (pair = force the two-CTA cluster variant; always used when a line pair exceeds
// one SM's shared memory).
cudaError_t launch_slab_phase(int logn, int phase, const LaunchArgsT<1>& args, bool pair, cudaStream_t stream);
bool slab_size_supported(int logn);

}  // namespace wso
