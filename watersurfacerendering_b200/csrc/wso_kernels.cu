// wso_kernels.cu — __global__ entry points and launch dispatch for sm_100a.
#include "wso_launch.h"

#include "wso_kernels.cuh"

namespace wso {

// CTA tiling per tile size.  CP: column pairs per K1 CTA, NF: packed fields per K1 CTA,
// RI: row items (each = output rows m' and N-m') per K2 CTA.  CTA threads = lines * N / 16.
template <int LOGN> struct Cfg;
template <> struct Cfg<4>  { static constexpr int CP = 8, NF = 4, RI = 8; };
template <> struct Cfg<5>  { static constexpr int CP = 8, NF = 4, RI = 8; };
template <> struct Cfg<6>  { static constexpr int CP = 8, NF = 4, RI = 8; };
template <> struct Cfg<7>  { static constexpr int CP = 4, NF = 4, RI = 8; };
template <> struct Cfg<8>  { static constexpr int CP = 4, NF = 4, RI = 4; };
template <> struct Cfg<9>  { static constexpr int CP = 4, NF = 2, RI = 2; };
template <> struct Cfg<10> { static constexpr int CP = 4, NF = 2, RI = 4; };
template <> struct Cfg<11> { static constexpr int CP = 4, NF = 1, RI = 2; };
template <> struct Cfg<12> { static constexpr int CP = 4, NF = 1, RI = 2; };
template <> struct Cfg<13> { static constexpr int CP = 2, NF = 1, RI = 1; };

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
__global__ void __launch_bounds__(Pass2<LOGN, Cfg<LOGN>::RI>::T, min_blocks(Pass2<LOGN, Cfg<LOGN>::RI>::T))
wso_pass2_kernel(const __grid_constant__ LaunchArgs args) {
    extern __shared__ __align__(16) float2 smem[];
    DeviceExec ex;
    Pass2<LOGN, Cfg<LOGN>::RI>::run(ex, smem, blockIdx.x, blockIdx.y, blockIdx.z, args);
}

// K3: disp.y *= 1/A per slot; also publishes A.  reference: WSTessendorf.cpp:443-455
__global__ void __launch_bounds__(256)
wso_normalize_kernel(const __grid_constant__ LaunchArgs args, unsigned n2) {
    const BatchItem item = args.items[blockIdx.y];
    const float mn = args.minmax[2 * item.slot + 0];
    const float mx = args.minmax[2 * item.slot + 1];
    const float a = amplitude_of(mn, mx);
    const float inv = __fdiv_rn(1.0f, a);
    if (blockIdx.x == 0 && threadIdx.x == 0) args.amp_out[item.slot] = a;
    float4* d = args.disp + (size_t)item.slot * n2;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        float* y = reinterpret_cast<float*>(d + i) + 1;
        *y = __fmul_rn(*y, inv);
    }
}

template <int LOGN>
static cudaError_t launch_all(const LaunchArgs& args, int n_items, cudaStream_t stream, bool first_use,
                              cudaEvent_t* ev) {
    using P1 = Pass1<LOGN, Cfg<LOGN>::CP, Cfg<LOGN>::NF>;
    using P2 = Pass2<LOGN, Cfg<LOGN>::RI>;
    constexpr int N = 1 << LOGN;
    if (first_use) {
        cudaError_t e;
        e = cudaFuncSetAttribute(wso_pass1_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, P1::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(wso_pass2_kernel<LOGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2::SMEM_BYTES);
        if (e != cudaSuccess) return e;
    }
    if (ev) cudaEventRecord(ev[0], stream);
    const dim3 g1(P1::H / Cfg<LOGN>::CP, 4 / Cfg<LOGN>::NF, n_items);
    wso_pass1_kernel<LOGN><<<g1, P1::T, P1::SMEM_BYTES, stream>>>(args);
    if (ev) cudaEventRecord(ev[1], stream);
    const dim3 g2(P2::H / Cfg<LOGN>::RI, 2, n_items);
    wso_pass2_kernel<LOGN><<<g2, P2::T, P2::SMEM_BYTES, stream>>>(args);
    if (ev) cudaEventRecord(ev[2], stream);
    const unsigned n2 = (unsigned)N * N;
    unsigned nb = (n2 + 256 * 4 - 1) / (256 * 4);
    if (nb > 1184) nb = 1184;  // 8 x 148 SMs
    wso_normalize_kernel<<<dim3(nb, n_items), 256, 0, stream>>>(args, n2);
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
