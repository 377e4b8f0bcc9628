// wso_kernels2.cu — __global__ entry points and launch code of the warp-per-line kernels (wso_kernels2.cuh) for sm_100a.
#include <cstdlib>
#include <mutex>

#include "wso_kernels2.cuh"
#include "wso_launch.h"

namespace wso {

// Tiling per tile size.  K1: CP column pairs x NF packed fields per CTA, one line group of L = N/32 threads per
// (column pair, field) -> CP*NF*L threads; K2 / K2h: GPC line groups per CTA.  The register budget (and with it the
// number of resident warps) follows from the CTA size: K1 runs 512-thread CTAs at <= 128 registers (16 warps per SM).
#ifndef WSO_V2_CP9
#define WSO_V2_CP9 8
#define WSO_V2_GPC9 8
#endif
#ifndef WSO_V2_CP10
#define WSO_V2_CP10 4
#define WSO_V2_GPC10 4
#endif
#ifndef WSO_V2_CP11
#define WSO_V2_CP11 2
#define WSO_V2_GPC11 2
#endif
#ifndef WSO_V2_NF
#define WSO_V2_NF 4
#endif
#ifndef WSO_V2_REGS1
#define WSO_V2_REGS1 128  // register budget per K1 thread
#endif
#ifndef WSO_V2_MINB2
#define WSO_V2_MINB2 3
#endif
#ifndef WSO_V2_REGSH
#define WSO_V2_REGSH 128  // register budget per K2h thread (no pack phase: one transformed line and the twiddles)
#endif
#ifndef WSO_V2_NBUFH
#define WSO_V2_NBUFH 1
#endif
#ifndef WSO_V2_NBUF2
#define WSO_V2_NBUF2 2  // line buffers per group of the map kernel
#endif
constexpr int min_blocks_for(int threads, int regs) {
    return (65536 / regs) / threads < 1 ? 1 : (65536 / regs) / threads;
}
template <int LOGN> struct Cfg2;
template <> struct Cfg2<9>  { static constexpr int CP = WSO_V2_CP9,  NF = WSO_V2_NF, GPC = WSO_V2_GPC9; };
template <> struct Cfg2<10> { static constexpr int CP = WSO_V2_CP10, NF = WSO_V2_NF, GPC = WSO_V2_GPC10; };
template <> struct Cfg2<11> { static constexpr int CP = WSO_V2_CP11, NF = WSO_V2_NF, GPC = WSO_V2_GPC11; };

template <int LOGN, class Args>
__global__ void __launch_bounds__(v2::Pass1W<LOGN, Cfg2<LOGN>::CP, Cfg2<LOGN>::NF>::T,
                                   min_blocks_for(v2::Pass1W<LOGN, Cfg2<LOGN>::CP, Cfg2<LOGN>::NF>::T, WSO_V2_REGS1))
wso_pass1w_kernel(const __grid_constant__ Args args) {
    extern __shared__ __align__(128) float2 smem[];
    DevCtx cx;
    v2::Pass1W<LOGN, Cfg2<LOGN>::CP, Cfg2<LOGN>::NF>::run(cx, smem, blockIdx.x, blockIdx.y, blockIdx.z, args);
}

// blockIdx.y: 0 = displacement map, 1 = normal map
template <int LOGN>
using PassM = v2::Pass2W<LOGN, Cfg2<LOGN>::GPC, WSO_V2_NBUF2>;

template <int LOGN, class Args>
__global__ void __launch_bounds__(PassM<LOGN>::T, WSO_V2_MINB2)
wso_pass2w_kernel(const __grid_constant__ Args args, int n_items) {
    extern __shared__ __align__(128) float2 smem[];
    DevCtx cx;
    using P2 = PassM<LOGN>;
    if (blockIdx.y == 0) P2::template run<0>(cx, smem, blockIdx.x, gridDim.x, n_items, args);
    else P2::template run<1>(cx, smem, blockIdx.x, gridDim.x, n_items, args);
}

template <int LOGN>
using PassH = v2::Pass2W<LOGN, Cfg2<LOGN>::GPC, WSO_V2_NBUFH>;

template <int LOGN, class Args>
__global__ void __launch_bounds__(PassH<LOGN>::T, min_blocks_for(PassH<LOGN>::T, WSO_V2_REGSH))
wso_heightsw_kernel(const __grid_constant__ Args args, int n_items) {
    extern __shared__ __align__(128) float2 smem[];
    DevCtx cx;
    PassH<LOGN>::template run<2>(cx, smem, blockIdx.x, gridDim.x, n_items, args);
}

namespace {

template <class... KArgs, class... Actual>
cudaError_t launch_pdl2(void (*kern)(KArgs...), dim3 grid, int threads, int smem, cudaStream_t stream, Actual&&... a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3((unsigned)threads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Actual>(a)...);
}

// per (size, device): shared-memory opt-in and the persistent grid sizes, resolved once
struct DevPlan {
    bool ready = false;
    cudaError_t err = cudaSuccess;
    int ctas_k2 = 0;   // resident CTAs of wso_pass2w_kernel on the whole device
    int ctas_k2h = 0;  // ... of wso_heightsw_kernel
};
constexpr int kMaxDev = 32;

template <int LOGN>
const DevPlan& plan_for_device() {
    static DevPlan plans[kMaxDev];
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= kMaxDev) dev = 0;
    std::lock_guard<std::mutex> lk(mu);
    DevPlan& p = plans[dev];
    if (p.ready) return p;
    using P1 = v2::Pass1W<LOGN, Cfg2<LOGN>::CP, Cfg2<LOGN>::NF>;
    using P2 = PassM<LOGN>;
    auto k1 = wso_pass1w_kernel<LOGN, LaunchArgs>;
    auto k2 = wso_pass2w_kernel<LOGN, LaunchArgs>;
    auto kh = wso_heightsw_kernel<LOGN, LaunchArgs>;
    cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, P1::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, P2::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(kh, cudaFuncAttributeMaxDynamicSharedMemorySize, PassH<LOGN>::SMEM_BYTES);
    int sms = 0, occ2 = 0, occh = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, k2, P2::T, P2::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occh, kh, PassH<LOGN>::T, PassH<LOGN>::SMEM_BYTES);
    if (e == cudaSuccess && (occ2 < 1 || occh < 1)) e = cudaErrorLaunchOutOfResources;
    p.err = e;
    if (const char* env = std::getenv("WSO_V2_OCC")) {  // tuning: cap the resident CTAs per SM of the persistent kernels
        const int cap = std::atoi(env);
        if (cap >= 1) {
            occ2 = occ2 < cap ? occ2 : cap;
            occh = occh < cap ? occh : cap;
        }
    }
    p.ctas_k2 = sms * occ2;
    p.ctas_k2h = sms * occh;
    p.ready = true;
    return p;
}

template <int LOGN>
cudaError_t launch_k1(const LaunchArgs& args, int n_items, cudaStream_t stream) {
    const DevPlan& p = plan_for_device<LOGN>();
    if (p.err != cudaSuccess) return p.err;
    using P1 = v2::Pass1W<LOGN, Cfg2<LOGN>::CP, Cfg2<LOGN>::NF>;
    const dim3 grid(P1::H / Cfg2<LOGN>::CP, 4 / Cfg2<LOGN>::NF, n_items);
    return launch_pdl2(wso_pass1w_kernel<LOGN, LaunchArgs>, grid, P1::T, P1::SMEM_BYTES, stream, args);
}
template <int LOGN>
cudaError_t launch_k2h(const LaunchArgs& args, int n_items, cudaStream_t stream) {
    const DevPlan& p = plan_for_device<LOGN>();
    if (p.err != cudaSuccess) return p.err;
    using PH = PassH<LOGN>;
    const int units = n_items * PH::H;
    int ctas = (units + Cfg2<LOGN>::GPC - 1) / Cfg2<LOGN>::GPC;
    if (ctas > p.ctas_k2h) ctas = p.ctas_k2h;
    return launch_pdl2(wso_heightsw_kernel<LOGN, LaunchArgs>, dim3(ctas, 1, 1), PH::T, PH::SMEM_BYTES, stream, args, n_items);
}
template <int LOGN>
cudaError_t launch_k2(const LaunchArgs& args, int n_items, cudaStream_t stream) {
    const DevPlan& p = plan_for_device<LOGN>();
    if (p.err != cudaSuccess) return p.err;
    using P2 = PassM<LOGN>;
    const int units = n_items * P2::H;
    constexpr int upc = Cfg2<LOGN>::GPC / 2;  // a row item of a map occupies two line groups
    int ctas = (units + upc - 1) / upc;
    if (ctas > p.ctas_k2 / 2) ctas = p.ctas_k2 / 2;  // the two maps share the device
    if (ctas < 1) ctas = 1;
    return launch_pdl2(wso_pass2w_kernel<LOGN, LaunchArgs>, dim3(ctas, 2, 1), P2::T, P2::SMEM_BYTES, stream, args, n_items);
}

}  // namespace

bool warp_core_supported(int logn) { return logn >= 9 && logn <= 11; }

cudaError_t launch_warp_core(int logn, int which, const LaunchArgs& args, int n_items, cudaStream_t stream) {
#define WSO_V2_CASE(LG)                                              \
    case LG:                                                         \
        if (which == 0) return launch_k1<LG>(args, n_items, stream); \
        if (which == 1) return launch_k2h<LG>(args, n_items, stream); \
        return launch_k2<LG>(args, n_items, stream);
    switch (logn) {
#ifdef WSO_ONLY_LOGN
        WSO_V2_CASE(WSO_ONLY_LOGN)
#else
        WSO_V2_CASE(9)
        WSO_V2_CASE(10)
        WSO_V2_CASE(11)
#endif
        default: return cudaErrorNotSupported;
    }
#undef WSO_V2_CASE
}

}  // namespace wso
