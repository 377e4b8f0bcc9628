// TEST INFRASTRUCTURE — CPU emulation of the CUDA kernel bodies.
//
// Compiles watersurfacerendering_b200/csrc/wso_kernels.cuh with g++ and steps every CTA thread by thread
// (HostExec) so that the index / Hermitian-layout logic of the sm_100a kernels can be checked against the
// oracle in the GPU-less build container.  Never linked into the product library; tests only.
#include <cstdio>
#include <cstring>
#include <vector>

#include "wso_kernels.cuh"

using namespace wso;

template <int LOGN, int CP, int NF, int RI>
static int run_cfg(const float* amp_t, const float* omega_t, const float* kv, float omega0, float lambda,
                   float t, float* disp, float* norm, float* minmax, float* amp_out, float* w_out) {
    constexpr int N = 1 << LOGN, H = N / 2;
    using P1 = Pass1<LOGN, CP, NF>;
    using P2 = Pass2<LOGN, RI, false>;
    using PH = Pass2<LOGN, RI, true>;
    std::vector<float2> tw(N);
    for (int k = 0; k < N; ++k) {
        const double a = 2.0 * 3.14159265358979323846 * k / N;
        tw[k] = make_float2((float)cos(a), (float)sin(a));
    }
    // same record / table decision as upload_h0() in wso_api.cu (omega0 <= 0: force the direct sincos path)
    std::vector<float4> rec((size_t)N * N);
    int jmax = 0;
    bool table_ok = omega0 > 0.0f;
    for (size_t i = 0; i < (size_t)N * N && table_ok; ++i) {
        const float jf = std::nearbyint(omega_t[i] / omega0);
        if (!(jf >= 0.0f && jf < (float)kMaxTable) || jf * omega0 != omega_t[i]) table_ok = false;
        else if ((int)jf > jmax) jmax = (int)jf;
    }
    for (int n = 0; n < N; ++n)
        for (int m = 0; m < N; ++m) {
            const size_t i = (size_t)n * N + m;
            const float d = kv[n] * kv[n] + kv[m] * kv[m];
            const float inv = std::sqrt(d) > 0.00001f ? 1.0f / std::sqrt(d) : 0.0f;
            float wf = omega_t[i];
            if (table_ok) {
                const int j = (int)std::nearbyint(omega_t[i] / omega0);
                std::memcpy(&wf, &j, 4);
            }
            rec[i] = make_float4(amp_t[2 * i], amp_t[2 * i + 1], inv, wf);
        }
    const int hN = N / 2;
    std::vector<float4> recs((size_t)hN * hN * 2, make_float4(0.f, 0.f, 0.f, 0.f));
    bool pairs_ok = true;
    for (int j = 1; j < hN; ++j)
        for (int i = 1; i < hN; ++i) {
            const float4 a0 = rec[(size_t)j * N + i], a3 = rec[(size_t)(N - j) * N + (N - i)];
            const float4 a1 = rec[(size_t)(N - j) * N + i], a2 = rec[(size_t)j * N + (N - i)];
            if (std::memcmp(&a0.w, &a3.w, 4) != 0 || std::memcmp(&a1.w, &a2.w, 4) != 0) pairs_ok = false;
            recs[((size_t)j * hN + i) * 2 + 0] = make_float4(a0.x + a3.x, a0.y + a3.y, a0.z, a0.w);
            recs[((size_t)j * hN + i) * 2 + 1] = make_float4(a1.x + a2.x, a1.y + a2.y, a1.z, a1.w);
        }
    TileDev td;
    td.h0 = rec.data();
    td.hs = recs.data();
    td.kv = kv;
    td.lambda = lambda;
    td.omega0 = omega0;
    td.table_len = table_ok ? jmax + 1 : 0;
    td.use_pairs = pairs_ok ? 1 : 0;
    std::vector<float2> W((size_t)H * 4 * N);
    LaunchArgs args;
    std::memset(&args, 0, sizeof(args));
    args.td[0] = td;
    args.tw = tw.data();
    args.W = W.data();
    args.disp = reinterpret_cast<float4*>(disp);
    args.norm = reinterpret_cast<float4*>(norm);
    args.minmax = minmax;
    args.amp_out = amp_out;
    args.items[0].tile = 0;
    args.items[0].slot = 0;
    args.items[0].t = t;
    {
        std::vector<float2> smem(P1::SMEM_BYTES / sizeof(float2));
        std::vector<ThreadState> st(P1::T);
        for (int by = 0; by < 4 / NF; ++by)
            for (int bx = 0; bx < H / CP; ++bx) {
                for (auto& v : smem) v = make_float2(NAN, NAN);
                HostExec ex{P1::T, st.data()};
                P1::run(ex, smem.data(), bx, by, 0, args);
            }
    }
    if (w_out) std::memcpy(w_out, W.data(), W.size() * sizeof(float2));
    {   // K2h: height extrema
        std::vector<float2> smem(PH::SMEM_BYTES / sizeof(float2));
        std::vector<ThreadState> st(PH::T);
        for (int bx = 0; bx < H / RI; ++bx) {
            for (auto& v : smem) v = make_float2(NAN, NAN);
            HostExec ex{PH::T, st.data()};
            PH::run(ex, smem.data(), bx, 0, 0, args);
        }
    }
    {
        std::vector<float2> smem(P2::SMEM_BYTES / sizeof(float2));
        std::vector<ThreadState> st(P2::T);
        for (int by = 0; by < 2; ++by)
            for (int bx = 0; bx < H / RI; ++bx) {
                for (auto& v : smem) v = make_float2(NAN, NAN);
                HostExec ex{P2::T, st.data()};
                P2::run(ex, smem.data(), bx, by, 0, args);
            }
    }
    return 0;
}

extern "C" int wso_emu_compute(int logn, int variant, const float* amp_t, const float* omega_t,
                               const float* kv, float omega0, float lambda, float t, float* disp, float* norm,
                               float* minmax, float* amp_out, float* w_out) {
#define CFG(L, V, CP, NF, RI) \
    if (logn == L && variant == V) return run_cfg<L, CP, NF, RI>(amp_t, omega_t, kv, omega0, lambda, t, disp, norm, minmax, amp_out, w_out);
    CFG(4, 0, 8, 4, 8)
    CFG(4, 1, 2, 1, 1)
    CFG(5, 0, 8, 4, 8)
    CFG(6, 0, 8, 4, 8)
    CFG(6, 1, 4, 2, 2)
    CFG(6, 2, 1, 1, 1)
    CFG(7, 0, 4, 4, 8)
    CFG(8, 0, 4, 4, 4)
    CFG(8, 1, 4, 2, 2)
    CFG(9, 0, 4, 4, 4)
    CFG(9, 1, 4, 2, 2)
    CFG(10, 0, 4, 2, 4)
    CFG(11, 0, 4, 1, 2)
#undef CFG
    return -1;
}
