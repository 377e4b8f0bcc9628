// TEST INFRASTRUCTURE — a cooperative-fiber SIMT emulator for the warp-per-line kernels.
//
// wso_kernels2.cuh is written against a context type (wso_simt.cuh: DevCtx on the GPU).  HostCtx below offers the same
// members on the CPU: every CUDA thread of a CTA is a fiber with its own stack; a warp shuffle, __syncwarp, named
// barrier or __syncthreads suspends the fiber until all participants have arrived, exactly the convergence the device
// code relies on.  A barrier nobody can complete is reported as a deadlock instead of hanging.  Bulk copies complete at
// issue (the source must be final by then on the device as well).  Never linked into the product library.
#pragma once

#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

#include "wso_device.cuh"

#if !defined(__x86_64__)
#error "fiber_simt.h: the context switch below is written for x86-64 (System V)"
#endif

extern "C" void wso_fiber_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.globl wso_fiber_switch
.type wso_fiber_switch,@function
wso_fiber_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size wso_fiber_switch,.-wso_fiber_switch
)");

namespace wso {

class FiberCta;
static thread_local FiberCta* g_fiber_cta = nullptr;

class FiberCta {
  public:
    static constexpr size_t kStack = 256 * 1024;
    // barrier keys: three per warp (shuffle publish / shuffle consume / __syncwarp), __syncthreads, named barriers 0..15
    static constexpr int kWarpKeys = 3 * 32, kCtaKey = kWarpKeys, kNamedKey = kCtaKey + 1, kBarriers = kNamedKey + 16;

    explicit FiberCta(int nthreads) : T_(nthreads), sp_(nthreads, nullptr), state_(nthreads, 0), slots_(nthreads) {
        stacks_ = static_cast<char*>(malloc(kStack * (size_t)T_));  // untouched pages stay uncommitted
        if (!stacks_) abort();
    }
    ~FiberCta() { free(stacks_); }
    FiberCta(const FiberCta&) = delete;

    int threads() const { return T_; }
    int current() const { return cur_; }

    // run body(tid) for every thread of the CTA to completion
    void run(const std::function<void(int)>& body) {
        body_ = &body;
        for (Bar& b : bars_) {
            b.arrived = 0;
            b.waiters.clear();
        }
        for (int t = 0; t < T_; ++t) {
            state_[t] = 0;
            char* top = stacks_ + kStack * (size_t)(t + 1);
            uintptr_t p = reinterpret_cast<uintptr_t>(top) & ~uintptr_t(15);
            void** sp = reinterpret_cast<void**>(p);
            *--sp = nullptr;                                    // fake return address: entry sees rsp % 16 == 8
            *--sp = reinterpret_cast<void*>(&FiberCta::entry);  // `ret` target of the first switch
            for (int i = 0; i < 6; ++i) *--sp = nullptr;        // rbp rbx r12 r13 r14 r15
            sp_[t] = sp;
        }
        FiberCta* prev = g_fiber_cta;
        g_fiber_cta = this;
        int done = 0, idle = 0;
        int t = 0;
        while (done < T_) {
            if (state_[t] == 0) {
                idle = 0;
                cur_ = t;
                wso_fiber_switch(&sched_sp_, sp_[t]);
                if (state_[t] == 2) ++done;
            } else if (++idle > T_) {
                fprintf(stderr, "fiber_simt: deadlock - %d of %d threads finished, the rest wait at barriers:", done, T_);
                for (size_t k = 0; k < bars_.size(); ++k)
                    if (bars_[k].arrived) fprintf(stderr, " [barrier %zu: %d arrived]", k, bars_[k].arrived);
                fprintf(stderr, "\n");
                abort();
            }
            t = (t + 1 == T_) ? 0 : t + 1;
        }
        g_fiber_cta = prev;
    }

    // all `count` participants of `key` must arrive before any continues
    void barrier(int key, int count) {
        Bar& b = bars_[key];
        // a named barrier in use with one participant count must not be joined with another (the device traps)
        if (b.arrived > 0 && b.count != count) {
            fprintf(stderr, "fiber_simt: barrier %d joined with count %d while %d threads wait with count %d\n", key,
                    count, b.arrived, b.count);
            abort();
        }
        b.count = count;
        if (++b.arrived == count) {
            for (int w : b.waiters) state_[w] = 0;
            b.waiters.clear();
            b.arrived = 0;
            return;
        }
        b.waiters.push_back(cur_);
        state_[cur_] = 1;
        const int me = cur_;
        wso_fiber_switch(&sp_[me], sched_sp_);
    }

    // let the other fibers run (a fiber polling for something another fiber produces)
    void yield() {
        const int me = cur_;
        wso_fiber_switch(&sp_[me], sched_sp_);
    }

    float2& slot(int tid) { return slots_[tid]; }

  private:
    struct Bar {
        int arrived = 0;
        int count = 0;
        std::vector<int> waiters;
    };
    static void entry() {
        FiberCta* c = g_fiber_cta;
        const int me = c->cur_;
        (*c->body_)(me);
        c->state_[me] = 2;
        wso_fiber_switch(&c->sp_[me], c->sched_sp_);
        abort();  // a finished fiber is never resumed
    }

    int T_;
    int cur_ = 0;
    char* stacks_ = nullptr;
    void* sched_sp_ = nullptr;
    std::vector<void*> sp_;
    std::vector<uint8_t> state_;  // 0 runnable, 1 blocked, 2 done
    std::vector<float2> slots_;
    std::vector<Bar> bars_ = std::vector<Bar>(kBarriers);
    const std::function<void(int)>* body_ = nullptr;
};

// Same members as DevCtx (wso_simt.cuh).
struct HostCtx {
    int tid;
    FiberCta* cta;

    HostCtx(FiberCta* c, int t) : tid(t), cta(c) {}
    int lane() const { return tid & 31; }

    float2 shfl(float2 v, int src) const {
        const int w = tid >> 5;
        cta->slot(tid) = v;
        cta->barrier(w, 32);
        const float2 r = cta->slot((w << 5) + (src & 31));
        cta->barrier(32 + w, 32);
        return r;
    }
    float shfl(float v, int src) const { return shfl(make_float2(v, 0.0f), src).x; }
    float2 shfl_xor(float2 v, int m) const { return shfl(v, lane() ^ m); }
    float shfl_xor(float v, int m) const { return shfl(v, lane() ^ m); }
    void syncwarp() const { cta->barrier(64 + (tid >> 5), 32); }
    void cta_sync() const { cta->barrier(FiberCta::kCtaKey, cta->threads()); }
    void bar(int id, int count) const { cta->barrier(FiberCta::kNamedKey + id, count); }

    void mbar_init(uint64_t* bar, unsigned) const { *bar = 0; }
    void mbar_init_fence() const {}
    void fence_async_smem() const {}
    void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) const {
        memcpy(dst, src, bytes);
        *bar += 1;  // completed copies (the device flips a phase bit instead)
    }
    // Device semantics: returns once the phase with this parity has completed.  The emulated copies complete at issue;
    // a waiter polls (yielding to the other fibers) until the issuing thread has got there, and a phase nobody ever
    // issues a copy for is reported instead of hanging.
    void mbar_wait(uint64_t* bar, unsigned parity) const {
        for (long spins = 0; ((unsigned)*(volatile uint64_t*)bar & 1u) == (parity & 1u); ++spins) {
            if (spins > 1000000L) {
                fprintf(stderr, "fiber_simt: wait on an mbarrier phase (parity %u) no copy was issued for\n", parity);
                abort();
            }
            cta->yield();
        }
    }
    void pdl_wait() const {}
    void pdl_release() const {}
    void atomic_minmax(float* out, float mn, float mx) const {
        if (mn < out[0]) out[0] = mn;
        if (mx > out[1]) out[1] = mx;
    }
};

}  // namespace wso
