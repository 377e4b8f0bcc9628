// wso_kernels.cu — __global__ entry points and launch dispatch for sm_100a.
#include "wso_launch.h"

#include "wso_kernels.cuh"

namespace wso {

// CTA tiling per tile size.  CP: column pairs per K1 CTA, NF: packed fields per K1 CTA,
// RI: row items (each = output rows m' and N-m') per K2 CTA.  CTA threads = lines * N / 16.
template <int LOGN> struct Cfg;
// RH: row items per K2h CTA (one line each).
template <> struct Cfg<4>  { static constexpr int CP = 8, NF = 4, RI = 8, RH = 8; };
template <> struct Cfg<5>  { static constexpr int CP = 8, NF = 4, RI = 8, RH = 16; };
template <> struct Cfg<6>  { static constexpr int CP = 8, NF = 4, RI = 8, RH = 16; };
template <> struct Cfg<7>  { static constexpr int CP = 4, NF = 4, RI = 8, RH = 16; };
template <> struct Cfg<8>  { static constexpr int CP = 4, NF = 4, RI = 4, RH = 8; };
// (the WSO_TUNE_* macros exist for tuning sweeps: tools/tune_build.sh builds variant libraries)
#ifndef WSO_TUNE_CP9
#define WSO_TUNE_CP9 4
#define WSO_TUNE_NF9 4
#define WSO_TUNE_RI9 4
#define WSO_TUNE_RH9 8
#endif
#ifndef WSO_TUNE_CP10
#define WSO_TUNE_CP10 4
#define WSO_TUNE_NF10 2
#define WSO_TUNE_RI10 4
#define WSO_TUNE_RH10 8
#endif
#ifndef WSO_TUNE_CP11
#define WSO_TUNE_CP11 4
#define WSO_TUNE_NF11 2
#define WSO_TUNE_RI11 2
#define WSO_TUNE_RH11 4
#endif
template <> struct Cfg<9>  { static constexpr int CP = WSO_TUNE_CP9, NF = WSO_TUNE_NF9, RI = WSO_TUNE_RI9, RH = WSO_TUNE_RH9; };
template <> struct Cfg<10> { static constexpr int CP = WSO_TUNE_CP10, NF = WSO_TUNE_NF10, RI = WSO_TUNE_RI10, RH = WSO_TUNE_RH10; };
template <> struct Cfg<11> { static constexpr int CP = WSO_TUNE_CP11, NF = WSO_TUNE_NF11, RI = WSO_TUNE_RI11, RH = WSO_TUNE_RH11; };
template <> struct Cfg<12> { static constexpr int CP = 4, NF = 1, RI = 2, RH = 4; };
template <> struct Cfg<13> { static constexpr int CP = 2, NF = 1, RI = 1, RH = 2; };

// 1024 resident threads per SM at <= 64 registers: every thread carries 16 complex values between barriers
constexpr int min_blocks(int threads) { return threads >= 1024 ? 1 : (1024 / threads > 8 ? 8 : 1024 / threads); }

template <int LOGN>
__global__ void __launch_bounds__(Pass1<LOGN, Cfg<LOGN>::CP, Cfg<LOGN>::NF>::T,
                                  min_blocks(Pass1<LOGN, Cfg<LOGN>::CP, Cfg<LOGN>::NF>::T))
wso_pass1_kernel(const __grid_constant__ LaunchArgs args) {
    extern __shared__ __align__(16) float2 smem[];
    DeviceExec ex;
    Pass1<LOGN, Cfg<LOGN>::CP, Cfg<LOGN>::NF>::run(ex, smem, blockIdx.x, blockIdx.y, blockIdx.z, args);
}

template <int LOGN>
__global__ void __launch_bounds__(Pass2<LOGN, Cfg<LOGN>::RI, false>::T, min_blocks(Pass2<LOGN, Cfg<LOGN>::RI, false>::T))
wso_pass2_kernel(const __grid_constant__ LaunchArgs args) {
    extern __shared__ __align__(16) float2 smem[];
    DeviceExec ex;
    Pass2<LOGN, Cfg<LOGN>::RI, false>::run(ex, smem, blockIdx.x, blockIdx.y, blockIdx.z, args);
}

// K2h: height extrema (min/max -> amplitude A) ahead of K2, so that K2 writes disp.y already normalised and no
// second pass over the displacement map is needed (reference: the serial NormalizeHeights sweep, WSTessendorf.cpp:443-455)
template <int LOGN>
__global__ void __launch_bounds__(Pass2<LOGN, Cfg<LOGN>::RH, true>::T, min_blocks(Pass2<LOGN, Cfg<LOGN>::RH, true>::T))
wso_heights_kernel(const __grid_constant__ LaunchArgs args) {
    extern __shared__ __align__(16) float2 smem[];
    DeviceExec ex;
    Pass2<LOGN, Cfg<LOGN>::RH, true>::run(ex, smem, blockIdx.x, 0, blockIdx.z, args);
}

template <int LOGN>
static cudaError_t launch_all(const LaunchArgs& args, int n_items, cudaStream_t stream, bool first_use,
                              cudaEvent_t* ev) {
    using P1 = Pass1<LOGN, Cfg<LOGN>::CP, Cfg<LOGN>::NF>;
    using P2 = Pass2<LOGN, Cfg<LOGN>::RI, false>;
    using PH = Pass2<LOGN, Cfg<LOGN>::RH, true>;
    if (first_use) {
        cudaError_t e;
        e = cudaFuncSetAttribute(wso_pass1_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, P1::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(wso_pass2_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(wso_heights_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, PH::SMEM_BYTES);
        if (e != cudaSuccess) return e;
    }
    if (ev) cudaEventRecord(ev[0], stream);
    const dim3 g1(P1::H / Cfg<LOGN>::CP, 4 / Cfg<LOGN>::NF, n_items);
    wso_pass1_kernel<LOGN><<<g1, P1::T, P1::SMEM_BYTES, stream>>>(args);
    if (ev) cudaEventRecord(ev[1], stream);
    const dim3 gh(PH::H / Cfg<LOGN>::RH, 1, n_items);
    wso_heights_kernel<LOGN><<<gh, PH::T, PH::SMEM_BYTES, stream>>>(args);
    if (ev) cudaEventRecord(ev[2], stream);
    const dim3 g2(P2::H / Cfg<LOGN>::RI, 2, n_items);
    wso_pass2_kernel<LOGN><<<g2, P2::T, P2::SMEM_BYTES, stream>>>(args);
    if (ev) cudaEventRecord(ev[3], stream);
    return cudaGetLastError();
}

cudaError_t launch_compute_waves(int logn, const LaunchArgs& args, int n_items, cudaStream_t stream,
                                 bool first_use, cudaEvent_t* ev) {
    switch (logn) {
        case 4: return launch_all<4>(args, n_items, stream, first_use, ev);
        case 5: return launch_all<5>(args, n_items, stream, first_use, ev);
        case 6: return launch_all<6>(args, n_items, stream, first_use, ev);
        case 7: return launch_all<7>(args, n_items, stream, first_use, ev);
        case 8: return launch_all<8>(args, n_items, stream, first_use, ev);
        case 9: return launch_all<9>(args, n_items, stream, first_use, ev);
        case 10: return launch_all<10>(args, n_items, stream, first_use, ev);
        case 11: return launch_all<11>(args, n_items, stream, first_use, ev);
        case 12: return launch_all<12>(args, n_items, stream, first_use, ev);
        case 13: return launch_all<13>(args, n_items, stream, first_use, ev);
        default: return cudaErrorInvalidValue;
    }
}

int kernels_per_launch() { return 3; }

}  // namespace wso
