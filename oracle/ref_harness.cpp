/* ORACLE / TEST INFRASTRUCTURE — not product code.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load the library this file builds.
 *
 * This translation unit is linked together with the reference's OWN, UNMODIFIED
 * src/scene/WSTessendorf.cpp (compiled where it lies under /root/reference by oracle/Makefile;
 * no reference source is copied into this repository) into oracle/_ref/libwsref.so.
 *
 * It provides
 *   1. the six fftwf_* entry points the reference calls (reference: WSTessendorf.cpp:33,164,
 *      191-232,256-269,342-366) — FFTW 3.3.10 itself is a network-fetched dependency
 *      (reference: CMakeLists.txt:157-177) that is absent here.  Modes:
 *        0  float64 transform rounded once to fp32 (accuracy / parity mode, default)
 *        1  fp32 Stockham radix-4 (timing mode: "reference code + shim FFT, not FFTW")
 *        2  no-op (lets the harness read the pre-FFT spectra m_Height .. m_dzDisplacementZ)
 *        3  real libfftw3f.so.3 through dlopen when the box has one
 *   2. a C interface (wsref_*) over class WSTessendorf (reference: WSTessendorf.h:58-122) so
 *      Python can drive Prepare()/ComputeWaves(t) and read the private arrays
 *      (built with -fno-access-control).
 */
#include "pch.h"

#include <dlfcn.h>
#include <omp.h>

#include <cstdio>
#include <cstring>

#include "cpu_fft.h"
#include "scene/WSTessendorf.h"

// ---------------------------------------------------------------------------------------------
// FFTW-API shim
// ---------------------------------------------------------------------------------------------
struct wso_shim_plan_s {
    int n0, n1, sign;
    fftwf_complex* in;
    fftwf_complex* out;
    void* real_plan;                          // mode 3
    wso_cpu_fft::Plan2D<float, 16>* fast;     // mode 1 (twiddles + scratch planned once, like FFTW)
    wso_cpu_fft::Plan2D<double, 8>* exact;    // mode 0
};

static int g_fft_mode = 0;

namespace {
struct RealFftw {
    void* handle = nullptr;
    void* (*plan_dft_2d)(int, int, void*, void*, int, unsigned) = nullptr;
    void (*execute)(void*) = nullptr;
    void (*destroy_plan)(void*) = nullptr;
    bool tried = false;
    bool load() {
        if (tried) return handle != nullptr;
        tried = true;
        handle = dlopen("libfftw3f.so.3", RTLD_NOW | RTLD_LOCAL);
        if (!handle) return false;
        plan_dft_2d = (decltype(plan_dft_2d))dlsym(handle, "fftwf_plan_dft_2d");
        execute = (decltype(execute))dlsym(handle, "fftwf_execute");
        destroy_plan = (decltype(destroy_plan))dlsym(handle, "fftwf_destroy_plan");
        if (!plan_dft_2d || !execute || !destroy_plan) {
            dlclose(handle);
            handle = nullptr;
        }
        return handle != nullptr;
    }
} g_real;
}  // namespace

extern "C" {

fftwf_complex* fftwf_alloc_complex(size_t n) {
    void* p = nullptr;
    if (posix_memalign(&p, 64, n * sizeof(fftwf_complex)) != 0) return nullptr;
    std::memset(p, 0, n * sizeof(fftwf_complex));
    return (fftwf_complex*)p;
}

void fftwf_free(void* p) { free(p); }

fftwf_plan fftwf_plan_dft_2d(int n0, int n1, fftwf_complex* in, fftwf_complex* out, int sign,
                             unsigned flags) {
    wso_shim_plan_s* p = new wso_shim_plan_s{n0, n1, sign, in, out, nullptr, nullptr, nullptr};
    if (g_fft_mode == 3 && g_real.load())
        p->real_plan = g_real.plan_dft_2d(n0, n1, in, out, sign, flags);
    return p;
}

void fftwf_execute(const fftwf_plan p) {
    if (g_fft_mode == 2) return;
    if (p->real_plan) {
        g_real.execute(p->real_plan);
        return;
    }
    if (p->in != p->out)
        std::memcpy(p->out, p->in, sizeof(fftwf_complex) * (size_t)p->n0 * p->n1);
    std::complex<float>* d = reinterpret_cast<std::complex<float>*>(p->out);
    if (g_fft_mode == 1) {
        if (!p->fast) {
            p->fast = new wso_cpu_fft::Plan2D<float, 16>();
            p->fast->init(p->n0, p->n1, p->sign);
        }
        p->fast->execute(d);
    } else {
        if (!p->exact) {
            p->exact = new wso_cpu_fft::Plan2D<double, 8>();
            p->exact->init(p->n0, p->n1, p->sign);
        }
        p->exact->execute(d);
    }
}

void fftwf_destroy_plan(fftwf_plan p) {
    if (p && p->real_plan) g_real.destroy_plan(p->real_plan);
    if (p) {
        delete p->fast;
        delete p->exact;
    }
    delete p;
}

void fftwf_cleanup(void) {}

// ---------------------------------------------------------------------------------------------
// C interface over the reference class
// ---------------------------------------------------------------------------------------------

/* mode: see header comment.  Must be set BEFORE wsref_prepare (plans are made there). */
int wsref_set_fft_mode(int mode) {
    if (mode == 3 && !g_real.load()) return -1;
    g_fft_mode = mode;
    return 0;
}
int wsref_have_real_fftw(void) { return g_real.load() ? 1 : 0; }
int wsref_max_threads(void) { return omp_get_max_threads(); }
void wsref_set_threads(int n) { omp_set_num_threads(n); }

void* wsref_create(uint32_t tile_size, float tile_length) {
    return new WSTessendorf(tile_size, tile_length);
}
void wsref_destroy(void* h) { delete static_cast<WSTessendorf*>(h); }

void wsref_set_tile_size(void* h, uint32_t n) { static_cast<WSTessendorf*>(h)->SetTileSize(n); }
void wsref_set_tile_length(void* h, float l) { static_cast<WSTessendorf*>(h)->SetTileLength(l); }
void wsref_set_wind_direction(void* h, float x, float y) {
    static_cast<WSTessendorf*>(h)->SetWindDirection(glm::vec2(x, y));
}
void wsref_set_wind_speed(void* h, float v) { static_cast<WSTessendorf*>(h)->SetWindSpeed(v); }
void wsref_set_animation_period(void* h, float T) {
    static_cast<WSTessendorf*>(h)->SetAnimationPeriod(T);
}
void wsref_set_phillips_const(void* h, float A) {
    static_cast<WSTessendorf*>(h)->SetPhillipsConst(A);
}
void wsref_set_lambda(void* h, float l) { static_cast<WSTessendorf*>(h)->SetLambda(l); }
void wsref_set_damping(void* h, float d) { static_cast<WSTessendorf*>(h)->SetDamping(d); }

uint32_t wsref_get_tile_size(void* h) { return static_cast<WSTessendorf*>(h)->GetTileSize(); }
float wsref_get_tile_length(void* h) { return static_cast<WSTessendorf*>(h)->GetTileLength(); }
void wsref_get_wind_dir(void* h, float* xy) {
    auto w = static_cast<WSTessendorf*>(h)->GetWindDir();
    xy[0] = w.x;
    xy[1] = w.y;
}
float wsref_get_wind_speed(void* h) { return static_cast<WSTessendorf*>(h)->GetWindSpeed(); }
float wsref_get_animation_period(void* h) {
    return static_cast<WSTessendorf*>(h)->GetAnimationPeriod();
}
float wsref_get_base_freq(void* h) { return static_cast<WSTessendorf*>(h)->m_BaseFreq; }
float wsref_get_phillips_const(void* h) {
    return static_cast<WSTessendorf*>(h)->GetPhillipsConst();
}
float wsref_get_damping(void* h) { return static_cast<WSTessendorf*>(h)->GetDamping(); }
float wsref_get_lambda(void* h) { return static_cast<WSTessendorf*>(h)->GetDisplacementLambda(); }
float wsref_get_min_height(void* h) { return static_cast<WSTessendorf*>(h)->GetMinHeight(); }
float wsref_get_max_height(void* h) { return static_cast<WSTessendorf*>(h)->GetMaxHeight(); }

/* srand(seed) then the reference's Prepare() (reference: WSTessendorf.cpp:36-58); the app seeds
 * with the wall clock (reference: core/Application.cpp:21), the harness with a fixed seed. */
void wsref_prepare(void* h, unsigned seed) {
    std::srand(seed);
    static_cast<WSTessendorf*>(h)->Prepare();
}

/* The Gaussian array Prepare() would draw after srand(seed) (reference: WSTessendorf.cpp:87-103).
 * dst: N*N complex<float>. Leaves the model untouched. */
void wsref_gauss_array(void* h, unsigned seed, float* dst) {
    std::srand(seed);
    auto xi = static_cast<WSTessendorf*>(h)->ComputeGaussRandomArray();
    std::memcpy(dst, xi.data(), xi.size() * sizeof(xi[0]));
}

/* Prepare() with a caller-supplied Gaussian array instead of rand(): same calls in the same
 * order as reference Prepare() (WSTessendorf.cpp:43-57). */
void wsref_prepare_with_gauss(void* h, const float* xi) {
    WSTessendorf* w = static_cast<WSTessendorf*>(h);
    const uint32_t n = w->m_TileSize;
    w->m_WaveVectors = w->ComputeWaveVectors();
    std::vector<std::complex<float>> g((size_t)n * n);
    std::memcpy(g.data(), xi, g.size() * sizeof(g[0]));
    w->m_BaseWaveHeights = w->ComputeBaseWaveHeightField(g);
    w->m_Displacements.resize((size_t)n * n, WSTessendorf::Displacement(0.0));
    w->m_Normals.resize((size_t)n * n, WSTessendorf::Normal(0.0, 1.0, 0.0, 0.0));
    w->DestroyFFTW();
    w->SetupFFTW();
}

size_t wsref_h0_stride(void) { return sizeof(WSTessendorf::BaseWaveHeight); }

/* h0 in the reference's own 20-byte struct (reference: WSTessendorf.h:142-147). */
void wsref_export_h0(void* h, void* dst) {
    WSTessendorf* w = static_cast<WSTessendorf*>(h);
    std::memcpy(dst, w->m_BaseWaveHeights.data(),
                w->m_BaseWaveHeights.size() * sizeof(WSTessendorf::BaseWaveHeight));
}
void wsref_import_h0(void* h, const void* src) {
    WSTessendorf* w = static_cast<WSTessendorf*>(h);
    std::memcpy(w->m_BaseWaveHeights.data(), src,
                w->m_BaseWaveHeights.size() * sizeof(WSTessendorf::BaseWaveHeight));
}
/* (kx, kz, unit.x, unit.y) per point (reference: WSTessendorf.h:128-140). */
void wsref_export_wave_vectors(void* h, float* dst) {
    WSTessendorf* w = static_cast<WSTessendorf*>(h);
    std::memcpy(dst, w->m_WaveVectors.data(), w->m_WaveVectors.size() * 4 * sizeof(float));
}

float wsref_compute_waves(void* h, float t) { return static_cast<WSTessendorf*>(h)->ComputeWaves(t); }

void wsref_get_displacements(void* h, float* dst) {
    const auto& v = static_cast<WSTessendorf*>(h)->GetDisplacements();
    std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
}
void wsref_get_normals(void* h, float* dst) {
    const auto& v = static_cast<WSTessendorf*>(h)->GetNormals();
    std::memcpy(dst, v.data(), v.size() * sizeof(v[0]));
}
/* The 7 contiguous complex work arrays (reference: WSTessendorf.cpp:164-171), order
 * Height, SlopeX, SlopeZ, DisplacementX, DisplacementZ, dxDisplacementX, dzDisplacementZ. */
void wsref_export_work_arrays(void* h, float* dst) {
    WSTessendorf* w = static_cast<WSTessendorf*>(h);
    const size_t n2 = (size_t)w->m_TileSize * w->m_TileSize;
    std::memcpy(dst, w->m_Height, 7 * n2 * sizeof(std::complex<float>));
}

}  // extern "C"
