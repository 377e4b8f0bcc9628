// Dev tool: single-tile latency of ComputeWaves through the C ABI, without Python in the loop.
//   lat_bench [N=512] [frames=2000]
// Prints (a) back-to-back throughput of wso_compute_batch(1 frame) calls and (b) the device time of one isolated
// tile-frame (events around one call, median of 200), both in microseconds.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "wsocean.h"

#define CK(x) do { int rc_ = (x); if (rc_ != 0) { std::fprintf(stderr, "%s -> %d: %s\n", #x, rc_, wso_last_error(ctx)); return 1; } } while (0)

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 512u;
    const int frames = argc > 2 ? std::atoi(argv[2]) : 2000;
    wso_ctx* ctx = nullptr;
    wso_params p;
    wso_default_params(&p);
    p.tile_size = n;
    p.tile_length = 1000.0f * n / 512.0f;
    CK(wso_create(&p, 0, 1, 4, &ctx));
    CK(wso_prepare(ctx, 0, 1, 1234));
    cudaStream_t s;
    cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    CK(wso_set_stream(ctx, s));
    float t = 0.0f;
    for (int i = 0; i < 50; ++i) { t += 0.05f; CK(wso_compute_batch(ctx, 1, nullptr, &t, 0)); }
    CK(wso_sync(ctx));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    // (a) back to back
    auto c0 = std::chrono::steady_clock::now();
    cudaEventRecord(e0, s);
    for (int i = 0; i < frames; ++i) { t += 0.05f; CK(wso_compute_batch(ctx, 1, nullptr, &t, (uint32_t)(i & 3))); }
    auto c1 = std::chrono::steady_clock::now();
    cudaEventRecord(e1, s);
    CK(wso_sync(ctx));
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double cpu_us = std::chrono::duration<double, std::micro>(c1 - c0).count() / frames;
    // (b) isolated
    std::vector<float> v;
    for (int i = 0; i < 200; ++i) {
        t += 0.05f;
        cudaEventRecord(e0, s);
        CK(wso_compute_batch(ctx, 1, nullptr, &t, 0));
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float m = 0; cudaEventElapsedTime(&m, e0, e1);
        v.push_back(m * 1e3f);
    }
    std::sort(v.begin(), v.end());
    std::printf("{\"tile_size\": %u, \"frames\": %d, \"back_to_back_us_per_frame\": %.3f, \"cpu_enqueue_us_per_frame\": %.3f, "
                "\"isolated_us_median\": %.3f, \"isolated_us_min\": %.3f}\n",
                n, frames, ms * 1e3 / frames, cpu_us, v[v.size() / 2], v[0]);
    wso_destroy(ctx);
    return 0;
}
