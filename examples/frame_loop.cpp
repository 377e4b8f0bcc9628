// frame_loop.cpp — headless stand-in for the reference's only caller of the hot path.
//
// Replays what WaterSurfaceMesh does with the surface model every frame
// (reference: src/scene/WaterSurfaceMesh.cpp:123-154 PrepareModelTess/Update and :701-755
// CopyModelTessDataToStagingBuffer) against the drop-in `class WSTessendorf` of
// include/wso_tessendorf_adaptor.hpp:
//     m_TimeCtr += dt * m_AnimSpeed;  A = model.ComputeWaves(m_TimeCtr);
//     memcpy(staging + offset, GetDisplacements().data(), 16*N*N);  memcpy(..., GetNormals().data(), 16*N*N);
// and reports end-to-end tile-frames/s including the device->host copies.
//
//   frame_loop [N=512] [frames=200] [seed=1234]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "wso_tessendorf_adaptor.hpp"

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 512u;
    const int frames = argc > 2 ? std::atoi(argv[2]) : 200;
    const unsigned seed = argc > 3 ? (unsigned)std::atoi(argv[3]) : 1234u;
    try {
        WSTessendorf model(n, 1000.0f * (float)n / 512.0f);
        std::srand(seed);  // the reference app seeds once at start-up (core/Application.cpp:21)
        model.Prepare();

        const size_t map_bytes = sizeof(WSTessendorf::Displacement) * model.GetDisplacementCount();
        std::vector<unsigned char> staging(2 * map_bytes);  // [Displacements][Normals], as in the reference layout
        const float dt = 1.0f / 60.0f, anim_speed = 3.0f;   // reference: WaterSurfaceMesh.h:208
        float time_ctr = 0.0f, amp = 0.0f;

        amp = model.ComputeWaves(time_ctr);  // PrepareModelTess: one pass to initialise the maps
        const auto t0 = std::chrono::steady_clock::now();
        for (int f = 0; f < frames; ++f) {
            time_ctr += dt * anim_speed;
            amp = model.ComputeWaves(time_ctr);
            std::memcpy(staging.data(), model.GetDisplacements().data(), map_bytes);
            std::memcpy(staging.data() + map_bytes, model.GetNormals().data(), map_bytes);
        }
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

        double sum = 0.0;
        const float* d = reinterpret_cast<const float*>(staging.data());
        for (size_t i = 0; i < 2 * map_bytes / sizeof(float); ++i) sum += d[i];
        std::printf("{\"tile_size\": %u, \"frames\": %d, \"ms_per_frame\": %.4f, \"tile_frames_per_s\": %.1f, "
                    "\"t_last\": %.6f, \"amplitude_last\": %.8g, \"min_height\": %.8g, \"max_height\": %.8g, "
                    "\"checksum\": %.10g}\n",
                    n, frames, s / frames * 1e3, frames / s, time_ctr, amp, model.GetMinHeight(),
                    model.GetMaxHeight(), sum);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "frame_loop: %s\n", e.what());
        return 1;
    }
    return 0;
}
