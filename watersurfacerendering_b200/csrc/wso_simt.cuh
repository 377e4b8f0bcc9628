// wso_simt.cuh — the small set of SIMT primitives the warp-per-line kernels (wso_kernels2.cuh) are written against.
//
// The kernel bodies are templates over a context type Ctx.  On the device Ctx = DevCtx: every member is one PTX
// instruction (warp shuffle, named barrier, mbarrier, 1-D bulk copy cp.async.bulk, programmatic dependent launch).
// tests/emu/fiber_simt.h provides the same interface over cooperatively scheduled fibers (one per CUDA thread) so that
// the identical bodies - shuffles, barriers and all - run thread for thread on the CPU in the GPU-less build
// container.  That host context is test infrastructure; the product library contains DevCtx only.
#pragma once

#include <stdint.h>

#include "wso_device.cuh"

namespace wso {

#if defined(__CUDACC__)
struct DevCtx {
    int tid;  // threadIdx.x

    __device__ __forceinline__ explicit DevCtx() : tid((int)threadIdx.x) {}
    __device__ __forceinline__ int lane() const { return tid & 31; }

    // ---- warp level ------------------------------------------------------------------------------------------
    __device__ __forceinline__ float shfl(float v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    __device__ __forceinline__ float2 shfl(float2 v, int src) const {
        return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
    }
    __device__ __forceinline__ float2 shfl_xor(float2 v, int m) const {
        return make_float2(__shfl_xor_sync(0xffffffffu, v.x, m), __shfl_xor_sync(0xffffffffu, v.y, m));
    }
    __device__ __forceinline__ float shfl_xor(float v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    __device__ __forceinline__ void syncwarp() const { __syncwarp(); }

    // ---- CTA level -------------------------------------------------------------------------------------------
    __device__ __forceinline__ void cta_sync() const { __syncthreads(); }
    // named barrier over `count` threads (whole warps), id in 1..15
    __device__ __forceinline__ void bar(int id, int count) const {
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
    }

    // ---- asynchronous 1-D bulk copy global -> shared (UBLKCP), completion counted on an mbarrier ------------------
    static __device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
    __device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) const {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(saddr(bar)), "r"(count) : "memory");
    }
    // make freshly initialised mbarriers visible to the async proxy
    __device__ __forceinline__ void mbar_init_fence() const {
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // order this thread's earlier generic-proxy accesses to shared memory before later async-proxy (bulk copy) writes
    __device__ __forceinline__ void fence_async_smem() const {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    // arm `bar` for `bytes` and start copying [src, src+bytes) to dst (16-byte aligned, bytes % 16 == 0); one thread
    __device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) const {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(saddr(bar)), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(saddr(dst)),
            "l"(src), "r"(bytes), "r"(saddr(bar))
            : "memory");
    }
    __device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) const {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "WSO_WAIT:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra WSO_DONE;\n"
            "bra WSO_WAIT;\n"
            "WSO_DONE:\n"
            "}" ::"r"(saddr(bar)),
            "r"(parity)
            : "memory");
    }

    // ---- programmatic dependent launch (see DeviceExec in wso_device.cuh) ------------------------------------------
    __device__ __forceinline__ void pdl_wait() const { asm volatile("griddepcontrol.wait;" ::: "memory"); }
    __device__ __forceinline__ void pdl_release() const { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

    // ---- height extrema: sign-aware float min/max through integer atomics ---------------------------------------
    __device__ __forceinline__ void atomic_minmax(float* out, float mn, float mx) const {
        if (mn >= 0.0f) atomicMin(reinterpret_cast<int*>(out), __float_as_int(mn));
        else atomicMax(reinterpret_cast<unsigned int*>(out), __float_as_uint(mn));
        if (mx >= 0.0f) atomicMax(reinterpret_cast<int*>(out + 1), __float_as_int(mx));
        else atomicMin(reinterpret_cast<unsigned int*>(out + 1), __float_as_uint(mx));
    }
};
#endif

}  // namespace wso
