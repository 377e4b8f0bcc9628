// TEST INFRASTRUCTURE — CPU emulation of the warp-per-line kernel bodies (wso_kernels2.cuh).
//
// The bodies are compiled by g++ against HostCtx (fiber_simt.h): every CUDA thread is a fiber, warp shuffles and
// barriers block until all participants arrive.  Checks the lane/register index logic (mirror exchanges, second-stage
// thread assignment, two-for-one separation, bulk-copy double buffering, special rows) against the oracle without a
// GPU.  Never linked into the product library; tests only.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>
#include <chrono>

#include "fiber_simt.h"
#include "wso_kernels2.cuh"

using namespace wso;

namespace {

struct Setup {
    std::vector<float2> tw;
    std::vector<float4> rec, recs;
    TileDev td;
};

// same record / table construction as upload_h0() in wso_api.cu
static bool make_setup(int N, const float* amp_t, const float* omega_t, const float* kv, float omega0, float lambda,
                       Setup& s) {
    s.tw.resize(N);
    for (int k = 0; k < N; ++k) {
        const double a = 2.0 * 3.14159265358979323846 * k / N;
        s.tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    s.rec.resize((size_t)N * N);
    int jmax = 0;
    for (size_t i = 0; i < (size_t)N * N; ++i) {
        const float jf = std::nearbyint(omega_t[i] / omega0);
        if (!(jf >= 0.0f && jf < (float)kMaxTable) || jf * omega0 != omega_t[i]) return false;
        if ((int)jf > jmax) jmax = (int)jf;
    }
    for (int n = 0; n < N; ++n)
        for (int m = 0; m < N; ++m) {
            const size_t i = (size_t)n * N + m;
            const float d = kv[n] * kv[n] + kv[m] * kv[m];
            const float inv = std::sqrt(d) > 0.00001f ? 1.0f / std::sqrt(d) : 0.0f;
            const int j = (int)std::nearbyint(omega_t[i] / omega0);
            float wf;
            std::memcpy(&wf, &j, 4);
            s.rec[h0_index(n, m, N, 0)] = make_float4(amp_t[2 * i], amp_t[2 * i + 1], inv, wf);
        }
    const int hN = N / 2;
    s.recs.assign((size_t)hN * hN * 2, make_float4(0.f, 0.f, 0.f, 0.f));
    for (int j = 1; j < hN; ++j)
        for (int i = 1; i < hN; ++i) {
            const float4 a0 = s.rec[h0_index(j, i, N, 0)], a3 = s.rec[h0_index(N - j, N - i, N, 0)];
            const float4 a1 = s.rec[h0_index(N - j, i, N, 0)], a2 = s.rec[h0_index(j, N - i, N, 0)];
            if (std::memcmp(&a0.w, &a3.w, 4) != 0 || std::memcmp(&a1.w, &a2.w, 4) != 0) return false;
            s.recs[hs_index((int)j, (int)i, 0, (int)hN)] = make_float4(a0.x + a3.x, a0.y + a3.y, a0.z, a0.w);
            s.recs[hs_index((int)j, (int)i, 1, (int)hN)] = make_float4(a1.x + a2.x, a1.y + a2.y, a1.z, a1.w);
        }
    s.td.h0 = s.rec.data();
    s.td.hs = s.recs.data();
    s.td.kv = kv;
    s.td.lambda = lambda;
    s.td.omega0 = omega0;
    s.td.table_len = jmax + 1;
    s.td.use_pairs = 1;
    s.td.j0 = 0;
    return true;
}

// K1 (CP column pairs x NF fields per CTA), K2h and K2 (GPC line groups per CTA, NBX CTAs walking the row items)
template <int LOGN, int CP, int NF, int GPC, int NBX, int NBUFH = 1>
static int run_cfg(int n_items, const float* amp_t, const float* omega_t, const float* kv, float omega0, float lambda,
                   const float* times, float* disp, float* norm, float* minmax, float* amp_out, float* w_out) {
    constexpr int N = 1 << LOGN, H = N / 2;
    using P1 = v2::Pass1W<LOGN, CP, NF>;
    using P2 = v2::Pass2W<LOGN, GPC>;
    using PH = v2::Pass2W<LOGN, GPC, NBUFH>;  // K2h: one or two line buffers per group
    if (n_items < 1 || n_items > kSmallChunk) return -3;
    Setup s;
    if (!make_setup(N, amp_t, omega_t, kv, omega0, lambda, s)) return -2;
    std::vector<float2> W((size_t)n_items * H * 4 * N, make_float2(NAN, NAN));
    LaunchArgsSmall args;
    std::memset(&args, 0, sizeof(args));
    args.tw = s.tw.data();
    args.W = W.data();
    args.disp = reinterpret_cast<float4*>(disp);
    args.norm = reinterpret_cast<float4*>(norm);
    args.minmax = minmax;
    args.amp_out = amp_out;
    for (int i = 0; i < n_items; ++i) {
        args.td[i] = s.td;
        args.items[i].tile = 0;
        args.items[i].slot = (uint32_t)i;
        args.items[i].t = times[i];
    }
    auto T0 = std::chrono::steady_clock::now();
    {
        std::vector<float2> smem((P1::SMEM_BYTES + 7) / 8);
        FiberCta cta(P1::T);
        for (int bz = 0; bz < n_items; ++bz)
            for (int by = 0; by < 4 / NF; ++by)
                for (int bx = 0; bx < H / CP; ++bx) {
                    for (auto& x : smem) x = make_float2(NAN, NAN);
                    cta.run([&](int tid) {
                        HostCtx cx(&cta, tid);
                        P1::run(cx, smem.data(), bx, by, bz, args);
                    });
                }
    }
    if (getenv("WSO_EMU_TIME")) fprintf(stderr, "K1 %.2fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count());
    if (w_out) std::memcpy(w_out, W.data(), W.size() * sizeof(float2));
    {
        std::vector<float2> smem((P2::SMEM_BYTES + 7) / 8);
        FiberCta cta(P2::T);
        for (int bx = 0; bx < NBX; ++bx) {  // K2h
            for (auto& x : smem) x = make_float2(NAN, NAN);
            cta.run([&](int tid) {
                HostCtx cx(&cta, tid);
                PH::template run<2>(cx, smem.data(), bx, NBX, n_items, args);
            });
        }
        for (int map = 0; map < 2; ++map)  // K2
            for (int bx = 0; bx < NBX; ++bx) {
                for (auto& x : smem) x = make_float2(NAN, NAN);
                cta.run([&](int tid) {
                    HostCtx cx(&cta, tid);
                    if (map == 0) P2::template run<0>(cx, smem.data(), bx, NBX, n_items, args);
                    else P2::template run<1>(cx, smem.data(), bx, NBX, n_items, args);
                });
            }
    }
    if (getenv("WSO_EMU_TIME")) fprintf(stderr, "all %.2fs\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - T0).count());
    return 0;
}

}  // namespace

// disp / norm: [n_items][N][N][4]; minmax: [n_items][2]; amp_out: [n_items]; w_out: NULL or [n_items][N/2][4][N] complex
extern "C" int wso_emu2_compute(int logn, int variant, int n_items, const float* amp_t, const float* omega_t,
                                const float* kv, float omega0, float lambda, const float* times, float* disp,
                                float* norm, float* minmax, float* amp_out, float* w_out) {
#define CFG(LG, V, CP, NF, GPC, NBX) \
    if (logn == LG && variant == V)  \
        return run_cfg<LG, CP, NF, GPC, NBX, 1 + (V & 1)>(n_items, amp_t, omega_t, kv, omega0, lambda, times, disp, norm, minmax, amp_out, w_out);
    CFG(9, 0, 8, 4, 8, 3)
    CFG(9, 1, 4, 2, 4, 5)
    CFG(9, 2, 2, 1, 2, 7)
    CFG(10, 0, 4, 4, 4, 3)
    CFG(10, 1, 8, 2, 2, 5)
    CFG(10, 2, 2, 4, 6, 7)
    CFG(10, 3, 1, 1, 2, 2)
    CFG(11, 0, 2, 4, 2, 3)
    CFG(11, 1, 4, 2, 6, 5)
    CFG(11, 2, 8, 1, 4, 4)
#undef CFG
    return -1;
}
