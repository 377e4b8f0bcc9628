/* ORACLE / TEST INFRASTRUCTURE — not product code.
 *
 * Plain C++ CPU restatement of the reference's wave-synthesis hot path
 * (kentril0/WaterSurfaceRendering, src/scene/WSTessendorf.{h,cpp}).  It is the checker the CUDA
 * path is compared with; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call it.  It is never linked into the product library.
 *
 * Pinning: the reference ships no tests or golden vectors for this path, so this
 * restatement is pinned against OUTPUTS OF THE REFERENCE ITSELF run in the build container
 * (oracle/_ref/libwsref.so = the reference's own WSTessendorf.cpp compiled verbatim) — see
 * tests/test_oracle.py — and against the fixtures generated from it under tests/golden/.
 *
 * Third-party arithmetic: the reference calls FFTW 3.3.10 (fftwf_plan_dft_2d, in-place, FFTW_BACKWARD,
 * unnormalised; reference: WSTessendorf.cpp:191-232, 342-366).  FFTW is not in /root/reference nor in
 * this image; its published definition  Y[j] = sum_k X[k] exp(+2*pi*i*j*k/n)  is restated in
 * oracle/cpu_fft.h and evaluated in float64 (rounded once to fp32) for parity, or in fp32 for timing.
 *
 * All element-wise arithmetic is IEEE fp32 in the reference's order of operations; build with
 * -ffp-contract=off.
 */
#include <omp.h>

#include <cfloat>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "cpu_fft.h"

namespace {

struct H0Rec {  // reference: WSTessendorf.h:142-147 (20 bytes)
    float re, im, re_c, im_c, omega;
};
static_assert(sizeof(H0Rec) == 20, "h0 record must be 20 bytes");

struct Cf {
    float re, im;
};
// complex product in the textbook order std::complex<float> uses for finite operands
inline Cf cmul(Cf a, Cf b) { return Cf{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
inline Cf cadd(Cf a, Cf b) { return Cf{a.re + b.re, a.im + b.im}; }

// reference: WSTessendorf.cpp:75-80 — float index arithmetic, double product with M_PI, narrowed once
inline float wave_number(int idx, int n, float tile_length) {
    const float centred = 2.0f * (float)idx - (float)n;
    return (float)(M_PI * centred / tile_length);
}

// reference: WSTessendorf.h:133-136 (glm::length / glm::normalize = v * (1/sqrt(dot)))
inline void unit_vector(float kx, float kz, float* ux, float* uz, float* len) {
    const float d = kx * kx + kz * kz;
    const float l = std::sqrt(d);
    *len = l;
    if (l > 0.00001f) {
        const float inv = 1.0f / std::sqrt(d);
        *ux = kx * inv;
        *uz = kz * inv;
    } else {
        *ux = 0.0f;
        *uz = 0.0f;
    }
}

struct Params {
    int n;
    float tile_length;
    float wind_x, wind_y;  // normalised (SetWindDirection, reference: WSTessendorf.cpp:476-479)
    float wind_speed;      // clamped  (SetWindSpeed,     reference: WSTessendorf.cpp:481-484)
    float phillips_a;
    float damping;
    float base_freq;       // (float)(2.0f*M_PI/T)        reference: WSTessendorf.cpp:486-490
};

const float kG = 9.81f;                                  // reference: WSTessendorf.h:230
const float kOneOverSqrt2 = 1.0f / std::sqrt(2.0f);      // reference: WSTessendorf.h:231

// reference: WSTessendorf.h:249-263
inline float phillips(const Params& p, float ux, float uz, float k) {
    const float k2 = k * k;
    const float k4 = k2 * k2;
    float cf = ux * p.wind_x + uz * p.wind_y;
    cf = cf * cf;
    const float L = p.wind_speed * p.wind_speed / kG;
    const float L2 = L * L;
    return p.phillips_a * std::exp(-1.0f / (k2 * L2)) / k4 * cf *
           std::exp(-k2 * p.damping * p.damping);
}

// reference: WSTessendorf.h:237-243
inline Cf base_amp(const Params& p, Cf xi, float ux, float uz, float k) {
    const float s = std::sqrt(phillips(p, ux, uz, k));
    return Cf{kOneOverSqrt2 * xi.re * s, kOneOverSqrt2 * xi.im * s};
}

// reference: WSTessendorf.h:284-297
inline float q_dispersion(const Params& p, float k) {
    return std::floor(std::sqrt(kG * k) / p.base_freq) * p.base_freq;
}

Params make_params(int n, float tile_length, float wind_x, float wind_y, float wind_speed,
                   float phillips_a, float damping, float anim_period) {
    Params p;
    p.n = n;
    p.tile_length = tile_length;
    const float inv = 1.0f / std::sqrt(wind_x * wind_x + wind_y * wind_y);
    p.wind_x = wind_x * inv;
    p.wind_y = wind_y * inv;
    p.wind_speed = wind_speed > 0.0001f ? wind_speed : 0.0001f;
    p.phillips_a = phillips_a;
    p.damping = damping;
    p.base_freq = (float)(2.0f * M_PI / anim_period);
    return p;
}

// glm::linearRand<float>(-1,1) over std::rand()  (reference dependency: libs/glm/glm/gtc/random.inl:13-86,177-183)
inline uint32_t rand_u8() { return (uint32_t)(std::rand() % 255); }
// `(hi << s) | (lo << 0)` on glm vec operands: GCC evaluates the RIGHT operand of the overloaded
// operator| first, so the low part draws from rand() before the high part (pinned against the
// reference build in tests/test_oracle.py).
inline uint32_t rand_u16() {
    const uint32_t lo = rand_u8();
    const uint32_t hi = rand_u8();
    return (hi << 8) | lo;
}
inline uint32_t rand_u32() {
    const uint32_t lo = rand_u16();
    const uint32_t hi = rand_u16();
    return (hi << 16) | lo;
}
inline float linear_rand_pm1() {
    return (float)rand_u32() / (float)UINT32_MAX * (1.0f - (-1.0f)) + (-1.0f);
}
// glm::gaussRand(0,1): Marsaglia polar (reference dependency: random.inl:218-232)
inline float gauss_rand01() {
    float w, x1, x2;
    do {
        x1 = linear_rand_pm1();
        x2 = linear_rand_pm1();
        w = x1 * x1 + x2 * x2;
    } while (w > 1.0f);
    return x2 * 1.0f * 1.0f * std::sqrt((-2.0f * std::log(w)) / w) + 0.0f;
}

}  // namespace

extern "C" {

/* a2 — reference: WSTessendorf.cpp:60-85.  out: N*N*(kx,kz,ux,uz) */
void wso_oracle_wave_vectors(int n, float tile_length, float* out) {
    for (int m = 0; m < n; ++m)
        for (int c = 0; c < n; ++c) {
            float* o = out + 4 * ((size_t)m * n + c);
            o[0] = wave_number(c, n, tile_length);
            o[1] = wave_number(m, n, tile_length);
            float len;
            unit_vector(o[0], o[1], &o[2], &o[3], &len);
        }
}

/* a3 — reference: WSTessendorf.cpp:87-103 after srand(seed).  xi: N*N complex<float> */
void wso_oracle_gauss_array(int n, unsigned seed, float* xi) {
    std::srand(seed);
    for (int m = 0; m < n; ++m)
        for (int c = 0; c < n; ++c) {
            const size_t i = (size_t)m * n + c;
            // Complex(gaussRand(), gaussRand()): GCC evaluates constructor arguments right-to-left,
            // so the IMAGINARY part is drawn first (pinned against the reference build).
            xi[2 * i + 1] = gauss_rand01();
            xi[2 * i] = gauss_rand01();
        }
}

/* a5-a8 — reference: WSTessendorf.cpp:105-148.  xi: N*N complex<float>; h0: N*N 20-byte records */
void wso_oracle_base_wave_heights(int n, float tile_length, float wind_x, float wind_y,
                                  float wind_speed, float phillips_a, float damping,
                                  float anim_period, const float* xi, void* h0_out) {
    const Params p =
        make_params(n, tile_length, wind_x, wind_y, wind_speed, phillips_a, damping, anim_period);
    H0Rec* h0 = static_cast<H0Rec*>(h0_out);
#pragma omp parallel for schedule(static)
    for (int m = 0; m < n; ++m)
        for (int c = 0; c < n; ++c) {
            const size_t i = (size_t)m * n + c;
            const float kx = wave_number(c, n, tile_length);
            const float kz = wave_number(m, n, tile_length);
            float ux, uz, k;
            unit_vector(kx, kz, &ux, &uz, &k);
            H0Rec r;
            if (k > 0.00001f) {
                const Cf g{xi[2 * i], xi[2 * i + 1]};
                const Cf a = base_amp(p, g, ux, uz, k);
                const Cf b = base_amp(p, g, -ux, -uz, k);
                r.re = a.re;
                r.im = a.im;
                r.re_c = b.re;
                r.im_c = -b.im;
                r.omega = q_dispersion(p, k);
            } else {
                r.re = r.im = r.re_c = 0.0f;
                r.im_c = -0.0f;
                r.omega = 0.0f;
            }
            h0[i] = r;
        }
}

/* Steps 1-3 of ComputeWaves — reference: WSTessendorf.cpp:294-336, WSTessendorf.h:265-275.
 * spectra: 7*N*N complex<float> in the reference's buffer order
 * (Height, SlopeX, SlopeZ, DisplacementX, DisplacementZ, dxDisplacementX, dzDisplacementZ). */
void wso_oracle_spectra(int n, float tile_length, const void* h0_in, float t, float* spectra) {
    const H0Rec* h0 = static_cast<const H0Rec*>(h0_in);
    const size_t n2 = (size_t)n * n;
    Cf* out = reinterpret_cast<Cf*>(spectra);
#pragma omp parallel for schedule(static)
    for (int m = 0; m < n; ++m)
        for (int c = 0; c < n; ++c) {
            const size_t i = (size_t)m * n + c;
            const float kx = wave_number(c, n, tile_length);
            const float kz = wave_number(m, n, tile_length);
            float ux, uz, len;
            unit_vector(kx, kz, &ux, &uz, &len);
            const float phase = h0[i].omega * t;
            const float pc = std::cos(phase);
            const float ps = std::sin(phase);
            const Cf h = cadd(cmul(Cf{h0[i].re, h0[i].im}, Cf{pc, ps}),
                              cmul(Cf{h0[i].re_c, h0[i].im_c}, Cf{pc, -ps}));
            const Cf dx = cmul(Cf{0.0f, -ux}, h);
            const Cf dz = cmul(Cf{0.0f, -uz}, h);
            out[0 * n2 + i] = h;
            out[1 * n2 + i] = cmul(Cf{0.0f, kx}, h);
            out[2 * n2 + i] = cmul(Cf{0.0f, kz}, h);
            out[3 * n2 + i] = dx;
            out[4 * n2 + i] = dz;
            out[5 * n2 + i] = cmul(Cf{0.0f, kx}, dx);
            out[6 * n2 + i] = cmul(Cf{0.0f, kz}, dz);
        }
}

/* ComputeWaves(t) — reference: WSTessendorf.cpp:284-455.
 * disp, norm: N*N*4 floats.  minmax_out: {min, max}.  fft_mode 0: float64 FFT, 1: fp32 FFT.
 * Returns the amplitude A. */
float wso_oracle_compute_waves(int n, float tile_length, float lambda, const void* h0_in, float t,
                               float* disp, float* norm, float* minmax_out, int fft_mode) {
    const size_t n2 = (size_t)n * n;
    std::vector<float> spec(2 * 7 * n2);
    wso_oracle_spectra(n, tile_length, h0_in, t, spec.data());
    std::complex<float>* f = reinterpret_cast<std::complex<float>*>(spec.data());
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < 7; ++i) {  // the reference's `omp sections` over 7 plans (cpp:338-378)
        if (fft_mode == 1)
            wso_cpu_fft::fft2d<float, 16>(f + i * n2, n, n, +1);
        else
            wso_cpu_fft::fft2d<double, 8>(f + i * n2, n, n, +1);
    }
    // reference: WSTessendorf.cpp:289-290, 380-412 — note max starts at FLT_MIN (smallest positive)
    float hmax = FLT_MIN, hmin = FLT_MAX;
#pragma omp parallel for schedule(static) reduction(max : hmax) reduction(min : hmin)
    for (int m = 0; m < n; ++m)
        for (int c = 0; c < n; ++c) {
            const size_t i = (size_t)m * n + c;
            const float sign = ((m + c) & 1) ? -1.0f : 1.0f;
            const float h = f[0 * n2 + i].real() * sign;
            hmax = h > hmax ? h : hmax;
            hmin = h < hmin ? h : hmin;
            disp[4 * i + 1] = h;
            disp[4 * i + 0] = sign * lambda * f[3 * n2 + i].real();
            disp[4 * i + 2] = sign * lambda * f[4 * n2 + i].real();
            disp[4 * i + 3] = 1.0f;
            norm[4 * i + 0] = sign * f[1 * n2 + i].real();
            norm[4 * i + 1] = sign * f[2 * n2 + i].real();
            norm[4 * i + 2] = sign * f[5 * n2 + i].real();
            norm[4 * i + 3] = sign * f[6 * n2 + i].real();
        }
    // reference: WSTessendorf.cpp:443-455
    const float a = std::fabs(hmin) > std::fabs(hmax) ? std::fabs(hmin) : std::fabs(hmax);
    const float inv = 1.0f / a;
    for (size_t i = 0; i < n2; ++i) disp[4 * i + 1] *= inv;
    if (minmax_out) {
        minmax_out[0] = hmin;
        minmax_out[1] = hmax;
    }
    return a;
}

int wso_oracle_max_threads(void) { return omp_get_max_threads(); }
void wso_oracle_set_threads(int n) { omp_set_num_threads(n); }

}  // extern "C"
