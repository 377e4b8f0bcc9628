/* ORACLE / TEST INFRASTRUCTURE — not product code.
 *
 * A small CPU FFT used (a) behind the FFTW-API shim that lets the reference's own
 * WSTessendorf.cpp run here (oracle/ref_harness.cpp) and (b) by the plain-C++ restatement's
 * float64 transform (oracle/ws_oracle.cpp).  FFTW 3.3.10 itself is not available in this image
 * (reference: CMakeLists.txt:157-177 fetches it from the network), so its published semantics are
 * restated: unnormalised DFT  Y[j] = sum_k X[k] exp(sign * 2*pi*i * j*k / n),  2-D = rows then columns.
 *
 * Algorithm: Stockham autosort, radix-4 with one radix-2 clean-up stage, over a batch of B lines
 * kept interleaved ([n][B] complex) so the inner loop vectorises.  Two instantiations are used:
 *   <double, *>  accuracy mode  (parity checks; float64 throughout, rounded once to fp32)
 *   <float, 16>  timing mode    (CPU baseline "reference code + shim FFT, not FFTW")
 */
#ifndef WSO_ORACLE_CPU_FFT_H_
#define WSO_ORACLE_CPU_FFT_H_

#include <cmath>
#include <complex>
#include <cstring>
#include <vector>

namespace wso_cpu_fft {

template <typename T>
struct Twiddles {
    int n = 0;
    int sign = +1;
    std::vector<T> cs;  // (cos, sin) of sign*2*pi*k/n for k in [0, n)
    void init(int n_, int sign_) {
        n = n_;
        sign = sign_;
        cs.resize(2 * (size_t)n);
        const long double two_pi = 6.283185307179586476925286766559005768L;
        for (int k = 0; k < n; ++k) {
            long double a = two_pi * (long double)k / (long double)n;
            cs[2 * k + 0] = (T)cosl(a);
            cs[2 * k + 1] = (T)(sign * sinl(a));
        }
    }
};

/* One batch: x and y are [n][B] complex (interleaved re,im), result returned in x or y
 * (pointer returned).  n must be a power of two. */
template <typename T, int B>
static inline T* fft_lines(T* x, T* y, int n, const Twiddles<T>& tw) {
    const T* cs = tw.cs.data();
    const T sgn = (T)tw.sign;
    int Ns = 1;
    // radix-2 first if log2(n) is odd
    int lg = 0;
    while ((1 << lg) < n) ++lg;
    if (lg & 1) {
        const int half = n / 2;
        for (int j = 0; j < half; ++j) {
            const T* a = x + 2 * (size_t)B * j;
            const T* b = x + 2 * (size_t)B * (j + half);
            T* o0 = y + 2 * (size_t)B * (2 * j);
            T* o1 = o0 + 2 * B;
            for (int i = 0; i < 2 * B; ++i) {
                o0[i] = a[i] + b[i];
                o1[i] = a[i] - b[i];
            }
        }
        T* t = x; x = y; y = t;
        Ns = 2;
    }
    for (; Ns < n; Ns *= 4) {
        const int q = n / 4;
        const int tstep = n / (4 * Ns);  // w_{4Ns}^k = w_n^{k*tstep}
        for (int j = 0; j < q; ++j) {
            const int k = j & (Ns - 1);
            const T w1r = cs[2 * (k * tstep)], w1i = cs[2 * (k * tstep) + 1];
            const T w2r = cs[2 * (2 * k * tstep)], w2i = cs[2 * (2 * k * tstep) + 1];
            const T w3r = cs[2 * (3 * k * tstep)], w3i = cs[2 * (3 * k * tstep) + 1];
            const T* p0 = x + 2 * (size_t)B * j;
            const T* p1 = p0 + 2 * (size_t)B * q;
            const T* p2 = p1 + 2 * (size_t)B * q;
            const T* p3 = p2 + 2 * (size_t)B * q;
            const int j0 = ((j - k) << 2) + k;
            T* o0 = y + 2 * (size_t)B * j0;
            T* o1 = o0 + 2 * (size_t)B * Ns;
            T* o2 = o1 + 2 * (size_t)B * Ns;
            T* o3 = o2 + 2 * (size_t)B * Ns;
            for (int i = 0; i < B; ++i) {
                const T ar = p0[2 * i], ai = p0[2 * i + 1];
                const T br = p1[2 * i] * w1r - p1[2 * i + 1] * w1i;
                const T bi = p1[2 * i] * w1i + p1[2 * i + 1] * w1r;
                const T cr = p2[2 * i] * w2r - p2[2 * i + 1] * w2i;
                const T ci = p2[2 * i] * w2i + p2[2 * i + 1] * w2r;
                const T dr = p3[2 * i] * w3r - p3[2 * i + 1] * w3i;
                const T di = p3[2 * i] * w3i + p3[2 * i + 1] * w3r;
                const T s0r = ar + cr, s0i = ai + ci;
                const T s1r = ar - cr, s1i = ai - ci;
                const T s2r = br + dr, s2i = bi + di;
                // (b - d) * (sign*i)
                const T s3r = -sgn * (bi - di), s3i = sgn * (br - dr);
                o0[2 * i] = s0r + s2r; o0[2 * i + 1] = s0i + s2i;
                o1[2 * i] = s1r + s3r; o1[2 * i + 1] = s1i + s3i;
                o2[2 * i] = s0r - s2r; o2[2 * i + 1] = s0i - s2i;
                o3[2 * i] = s1r - s3r; o3[2 * i + 1] = s1i - s3i;
            }
        }
        T* t = x; x = y; y = t;
    }
    return x;
}

/* Plan for an in-place 2-D transform of an n0 x n1 row-major complex<float> array, computed in type T.
 * For T=double every intermediate (including the array between the two passes) is float64 and the result
 * is rounded once to fp32; for T=float the data are transformed in place.  Twiddles and scratch live in
 * the plan (like an FFTW plan), so execute() allocates nothing. */
template <typename T, int B>
struct Plan2D {
    int n0 = 0, n1 = 0, sign = +1;
    Twiddles<T> tw0, tw1;
    std::vector<T> bx, by, work;

    void init(int n0_, int n1_, int sign_) {
        n0 = n0_; n1 = n1_; sign = sign_;
        tw0.init(n0, sign);
        tw1.init(n1, sign);
        const int nmax = n0 > n1 ? n0 : n1;
        bx.resize((size_t)2 * B * nmax);
        by.resize((size_t)2 * B * nmax);
        if (sizeof(T) != sizeof(float)) work.resize((size_t)2 * n0 * n1);
    }

    template <typename U>
    void passes(U* w) {
        // rows (transform along n1): B rows at a time, interleaved [n][b]
        for (int r0 = 0; r0 < n0; r0 += B) {
            const int nb = (n0 - r0) < B ? (n0 - r0) : B;
            for (int b = 0; b < B; ++b) {
                const U* src = w + 2 * (size_t)(r0 + (b < nb ? b : 0)) * n1;
                T* dst = bx.data() + 2 * b;
                for (int n = 0; n < n1; ++n) {
                    dst[2 * (size_t)n * B] = (T)src[2 * n];
                    dst[2 * (size_t)n * B + 1] = (T)src[2 * n + 1];
                }
            }
            const T* res = fft_lines<T, B>(bx.data(), by.data(), n1, tw1);
            for (int b = 0; b < nb; ++b) {
                U* dst = w + 2 * (size_t)(r0 + b) * n1;
                const T* src = res + 2 * b;
                for (int n = 0; n < n1; ++n) {
                    dst[2 * n] = (U)src[2 * (size_t)n * B];
                    dst[2 * n + 1] = (U)src[2 * (size_t)n * B + 1];
                }
            }
        }
        // columns (transform along n0): B adjacent columns are already interleaved in memory
        for (int c0 = 0; c0 < n1; c0 += B) {
            const int nb = (n1 - c0) < B ? (n1 - c0) : B;
            for (int m = 0; m < n0; ++m) {
                const U* src = w + 2 * ((size_t)m * n1 + c0);
                T* dst = bx.data() + 2 * (size_t)m * B;
                for (int i = 0; i < 2 * nb; ++i) dst[i] = (T)src[i];
                for (int i = 2 * nb; i < 2 * B; ++i) dst[i] = (T)0;
            }
            const T* res = fft_lines<T, B>(bx.data(), by.data(), n0, tw0);
            for (int m = 0; m < n0; ++m) {
                U* dst = w + 2 * ((size_t)m * n1 + c0);
                const T* src = res + 2 * (size_t)m * B;
                for (int i = 0; i < 2 * nb; ++i) dst[i] = (U)src[i];
            }
        }
    }

    void execute(std::complex<float>* data) {
        float* d = reinterpret_cast<float*>(data);
        if (sizeof(T) == sizeof(float)) {
            passes<float>(d);
        } else {
            const size_t cnt = (size_t)2 * n0 * n1;
            for (size_t i = 0; i < cnt; ++i) work[i] = (T)d[i];
            passes<T>(work.data());
            for (size_t i = 0; i < cnt; ++i) d[i] = (float)work[i];
        }
    }
};

/* One-shot convenience wrapper (plans on every call). */
template <typename T, int B>
static void fft2d(std::complex<float>* data, int n0, int n1, int sign) {
    Plan2D<T, B> p;
    p.init(n0, n1, sign);
    p.execute(data);
}

}  // namespace wso_cpu_fft
#endif
