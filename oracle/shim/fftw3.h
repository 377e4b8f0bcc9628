/* ORACLE BUILD SHIM (test infrastructure, not product code).
 *
 * FFTW 3.3.10 is a network-fetched dependency of the reference (reference: CMakeLists.txt:157-177)
 * and is absent from /root/reference and from this image.  This header declares exactly the six
 * fftwf_* entry points the hot path uses (reference: src/scene/WSTessendorf.cpp:33,164,191-232,
 * 256-269,342-366) with FFTW's documented semantics: 2-D row-major n0 x n1 complex transform,
 * in-place allowed, sign +1 == FFTW_BACKWARD == sum_k X[k] exp(+2*pi*i*j*k/n), unnormalised.
 * The implementation lives in oracle/ref_harness.cpp.
 */
#ifndef WSO_ORACLE_SHIM_FFTW3_H_
#define WSO_ORACLE_SHIM_FFTW3_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef float fftwf_complex[2];
typedef struct wso_shim_plan_s* fftwf_plan;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_ESTIMATE (1U << 6)

fftwf_complex* fftwf_alloc_complex(size_t n);
void fftwf_free(void* p);
fftwf_plan fftwf_plan_dft_2d(int n0, int n1, fftwf_complex* in, fftwf_complex* out, int sign,
                             unsigned flags);
void fftwf_execute(const fftwf_plan p);
void fftwf_destroy_plan(fftwf_plan p);
void fftwf_cleanup(void);

#ifdef __cplusplus
}
#endif
#endif
