// wso_slab_kernels.cu — __global__ entry points of the slab-decomposed path (one large grid over P devices,
// BASELINE config 5) and of the cluster-pair K2 that a 16384-point line pair needs.
#include <cooperative_groups.h>

#include "wso_kernels.cuh"
#include "wso_launch.h"

namespace cg = cooperative_groups;

namespace wso {

using SlabArgs = LaunchArgsT<1>;  // a slab launch is one tile-frame

// CTA tiling of the slab kernels per size.  PAIR: K2 as two-CTA clusters (one line per CTA).
template <int LOGN> struct SlabCfg;
template <> struct SlabCfg<6>  { static constexpr int CP = 2, NF = 2, RI = 2, RH = 2; };   // small sizes: tests only
template <> struct SlabCfg<8>  { static constexpr int CP = 4, NF = 2, RI = 2, RH = 4; };
template <> struct SlabCfg<11> { static constexpr int CP = 4, NF = 1, RI = 1, RH = 2; };
template <> struct SlabCfg<12> { static constexpr int CP = 4, NF = 1, RI = 1, RH = 2; };
template <> struct SlabCfg<13> { static constexpr int CP = 2, NF = 1, RI = 1, RH = 1; };
template <> struct SlabCfg<14> { static constexpr int CP = 1, NF = 1, RI = 1, RH = 1; };

template <int LOGN>
__global__ void __launch_bounds__(Pass1<LOGN, SlabCfg<LOGN>::CP, SlabCfg<LOGN>::NF, true>::T)
wso_slab_pass1_kernel(const __grid_constant__ SlabArgs args) {
    extern __shared__ __align__(16) float2 smem[];
    DeviceExec ex;
    // The field groups of one column-pair block are neighbours in the launch order (blockIdx.x = field group): they read
    // the same spectrum records, so all but the first find them in L2 (with the column-pair blocks fastest, a rank's
    // whole share of the records - 268 MB per rank for 16384^2 on 8 devices - lay between two readers of the same record).
#ifdef WSO_EXP_SLAB_K1_XMAJOR
    Pass1<LOGN, SlabCfg<LOGN>::CP, SlabCfg<LOGN>::NF, true>::run(ex, smem, blockIdx.x, blockIdx.y + args.slab_field0, 0, args);
#else
    Pass1<LOGN, SlabCfg<LOGN>::CP, SlabCfg<LOGN>::NF, true>::run(ex, smem, blockIdx.y, blockIdx.x + args.slab_field0, 0, args);
#endif
}

template <int LOGN>
__global__ void __launch_bounds__(Pass2<LOGN, SlabCfg<LOGN>::RH, true, true>::T)
wso_slab_heights_kernel(const __grid_constant__ SlabArgs args) {
    extern __shared__ __align__(16) float2 smem[];
    DeviceExec ex;
    Pass2<LOGN, SlabCfg<LOGN>::RH, true, true>::run(ex, smem, blockIdx.x, 0, 0, args);
}

// a whole line pair (2 * N complex + padding) in one CTA: possible up to N = 8192
template <int LOGN>
struct PairOnly {
    static constexpr bool value =
        (size_t)2 * SlabCfg<LOGN>::RI * LineStride<(1 << LOGN)>::value * sizeof(float2) > (size_t)227 * 1024 ||
        2 * SlabCfg<LOGN>::RI * (1 << LOGN) / kValsPerThread > 1024;
};

template <int LOGN>
__global__ void __launch_bounds__(PairOnly<LOGN>::value ? 32 : 2 * SlabCfg<LOGN>::RI * (1 << LOGN) / kValsPerThread)
wso_slab_pass2_kernel(const __grid_constant__ SlabArgs args) {
    if constexpr (!PairOnly<LOGN>::value) {
        extern __shared__ __align__(16) float2 smem[];
        DeviceExec ex;
        Pass2<LOGN, SlabCfg<LOGN>::RI, false, true>::run(ex, smem, blockIdx.x, blockIdx.y, 0, args);
    }
}

// K2 as thread-block clusters of two CTAs: blockIdx.x = 2 * (local row item) + rank in the pair.  Each CTA transforms
// one of the two lines of the row item in its own shared memory; after a cluster barrier CTA 0 packs output row m'
// and CTA 1 output row N-m', each reading the partner's line through distributed shared memory.
template <int LOGN>
__global__ void __launch_bounds__(Pass2<LOGN, 1, false, true, true>::T)
wso_slab_pass2_pair_kernel(const __grid_constant__ SlabArgs args) {
    extern __shared__ __align__(16) float2 smem[];
    using P2 = Pass2<LOGN, 1, false, true, true>;
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank();
    const int bx = (int)blockIdx.x >> 1;
    DeviceExec ex;
    P2::transform(ex, smem, bx, blockIdx.y, 0, crank, args);
    cluster.sync();  // both lines complete and visible cluster-wide
    const float2* peer = cluster.map_shared_rank(smem, crank ^ 1);
    P2::pack(ex, smem, peer, bx, blockIdx.y, 0, crank, args);
    cluster.sync();  // the partner may still be reading this CTA's line
}

// The exchange step of the slab path as ONE kernel over peer memory: tile transposes of the staged K1 output
//   stage[f][half][jl][d*Hl + ml]  (this device, coalesced by K1)   ->   W_d[src][ml][f][half][jl]  (owner d of row item ml)
// where W_d is the receive buffer of device d - mapped peer memory over NVLink (CUDA IPC) or, when the exchange is left to
// a collective library, this device's block d of a send buffer.  Reads and writes are 256-byte rows of a 32 x 32 tile
// (8-byte elements) through a padded shared-memory tile.
__global__ void __launch_bounds__(256) wso_slab_exchange_kernel(const float2* __restrict__ stage, XposeDst dst, int hl_log,
                                                                int h_log, int fh0, int world_log, int rank) {
    __shared__ float2 tile[32][33];
    const int Hl = 1 << hl_log;
    // blockIdx.z = ((f - field0) * 2 + half) * world + i: the destinations rotate fastest and start behind this device, so
    // that at any moment the P devices store to P different owners (all NVLink ingress ports busy, no incast on one)
    const int world = 1 << world_log;
    const int i = blockIdx.z & (world - 1);
    const int d = (rank + 1 + i) & (world - 1);
    const int fh = (blockIdx.z >> world_log) + fh0;  // (f, half)
    const int jl0 = blockIdx.y * 32, ml0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float2* src = stage + (((size_t)fh << hl_log) << h_log) + ((size_t)d << hl_log);
#pragma unroll
    for (int r = 0; r < 32; r += 8) tile[ty + r][tx] = src[((size_t)(jl0 + ty + r) << h_log) + ml0 + tx];
    __syncthreads();
    float2* out = dst.p[d] + ((size_t)fh << hl_log);
#pragma unroll
    for (int r = 0; r < 32; r += 8) out[(((size_t)(ml0 + ty + r) * 8) << hl_log) + jl0 + tx] = tile[tx][ty + r];
    (void)Hl;
}

cudaError_t launch_slab_exchange(const float2* stage, const XposeDst& dst, int world, int rank, int hl_log, int h_log,
                                 int field0, int nfields, cudaStream_t stream) {
    const int Hl = 1 << hl_log;
    if (Hl < 32 || field0 < 0 || nfields < 1 || field0 + nfields > 4) return cudaErrorInvalidValue;
    int world_log = 0;
    while ((1 << world_log) < world) ++world_log;
    wso_slab_exchange_kernel<<<dim3(Hl / 32, Hl / 32, nfields * 2 * world), 256, 0, stream>>>(stage, dst, hl_log, h_log,
                                                                                             field0 * 2, world_log, rank);
    return cudaGetLastError();
}

template <class Kern>
static cudaError_t opt_in_smem(Kern kern, int bytes) {
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

template <int LOGN>
static cudaError_t slab_launch(int phase, const SlabArgs& args, bool pair, int nfields, cudaStream_t stream) {
    using C = SlabCfg<LOGN>;
    using P1 = Pass1<LOGN, C::CP, C::NF, true>;
    using PH = Pass2<LOGN, C::RH, true, true>;
    using PP = Pass2<LOGN, 1, false, true, true>;
    constexpr bool kPairOnly = PairOnly<LOGN>::value;
    const int Hl = (1 << (LOGN - 1)) >> args.slab_shift;
    if (Hl < C::CP || Hl < C::RI || Hl < C::RH || Hl < 1) return cudaErrorInvalidValue;
    cudaError_t e;
    if (phase == 0) {
        if ((e = opt_in_smem(wso_slab_pass1_kernel<LOGN>, P1::SMEM_BYTES)) != cudaSuccess) return e;
        // nfields packed fields from field group args.slab_field0 on (all of them: slab_nfields == 0)
        const int groups = nfields > 0 ? nfields / C::NF : 4 / C::NF;
        if (groups < 1 || (nfields > 0 && nfields % C::NF != 0)) return cudaErrorInvalidValue;
#ifdef WSO_EXP_SLAB_K1_XMAJOR
        wso_slab_pass1_kernel<LOGN><<<dim3(Hl / C::CP, groups, 1), P1::T, P1::SMEM_BYTES, stream>>>(args);
#else
        wso_slab_pass1_kernel<LOGN><<<dim3(groups, Hl / C::CP, 1), P1::T, P1::SMEM_BYTES, stream>>>(args);
#endif
    } else if (phase == 1) {
        if ((e = opt_in_smem(wso_slab_heights_kernel<LOGN>, PH::SMEM_BYTES)) != cudaSuccess) return e;
        wso_slab_heights_kernel<LOGN><<<dim3(Hl / C::RH, 1, 1), PH::T, PH::SMEM_BYTES, stream>>>(args);
    } else if (pair || kPairOnly) {
        if ((e = opt_in_smem(wso_slab_pass2_pair_kernel<LOGN>, PP::SMEM_BYTES)) != cudaSuccess) return e;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * Hl, 2, 1);
        cfg.blockDim = dim3(PP::T, 1, 1);
        cfg.dynamicSmemBytes = PP::SMEM_BYTES;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if ((e = cudaLaunchKernelEx(&cfg, wso_slab_pass2_pair_kernel<LOGN>, args)) != cudaSuccess) return e;
    } else {
        if constexpr (!kPairOnly) {
            using P2 = Pass2<LOGN, C::RI, false, true>;
            if ((e = opt_in_smem(wso_slab_pass2_kernel<LOGN>, P2::SMEM_BYTES)) != cudaSuccess) return e;
            wso_slab_pass2_kernel<LOGN><<<dim3(Hl / C::RI, 2, 1), P2::T, P2::SMEM_BYTES, stream>>>(args);
        }
    }
    return cudaGetLastError();
}

int slab_fields_per_group(int logn) {
    switch (logn) {
        case 6: return SlabCfg<6>::NF;
        case 8: return SlabCfg<8>::NF;
        case 11: return SlabCfg<11>::NF;
        case 12: return SlabCfg<12>::NF;
        case 13: return SlabCfg<13>::NF;
        case 14: return SlabCfg<14>::NF;
        default: return 4;
    }
}

cudaError_t launch_slab_phase(int logn, int phase, const LaunchArgsT<1>& args, bool pair, int nfields, cudaStream_t stream) {
    switch (logn) {
#ifndef WSO_ONLY_LOGN
        case 6: return slab_launch<6>(phase, args, pair, nfields, stream);
        case 8: return slab_launch<8>(phase, args, pair, nfields, stream);
        case 11: return slab_launch<11>(phase, args, pair, nfields, stream);
        case 12: return slab_launch<12>(phase, args, pair, nfields, stream);
        case 13: return slab_launch<13>(phase, args, pair, nfields, stream);
        case 14: return slab_launch<14>(phase, args, pair, nfields, stream);
#endif
        default: return cudaErrorInvalidValue;
    }
}

bool slab_size_supported(int logn) { return logn == 6 || logn == 8 || (logn >= 11 && logn <= 14); }

}  // namespace wso
