/* wsocean.h — C ABI of the B200-native Tessendorf wave-synthesis library (libwsocean.so).
 *
 * Drop-in boundary for ONE hot path of kentril0/WaterSurfaceRendering: the CPU surface model
 * `class WSTessendorf` (reference: src/scene/WSTessendorf.h:33-321, src/scene/WSTessendorf.cpp).
 * Each entry point below names the reference interface it replaces.  Plain pointers and sizes only;
 * no C++ / CUDA / torch types cross this boundary.  Every function returns a wso_status (0 = ok) and
 * never throws or aborts; wso_last_error() gives the text of the last failure.
 *
 * Threading (reference: single frame-loop thread, model not re-entrant): one context = one CUDA
 * device + one stream; calls on a context must be serialised by the caller; contexts are independent.
 *
 * All kernels are hand-written sm_100a CUDA (no cuFFT).  There is no CPU fallback: on a machine
 * without a usable GPU wso_create() fails with WSO_ERR_CUDA.
 */
#ifndef WSOCEAN_H_
#define WSOCEAN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSO_API __attribute__((visibility("default")))

typedef struct wso_ctx wso_ctx;

typedef enum wso_status {
    WSO_OK = 0,
    WSO_ERR_INVALID_ARG = -1,
    WSO_ERR_BAD_TILE_SIZE = -2,   /* not a power of two in [16, 8192]; state unchanged
                                     (reference SetTileSize silently ignores non-pow2: WSTessendorf.cpp:459-468) */
    WSO_ERR_NOT_PREPARED = -3,    /* ComputeWaves before Prepare */
    WSO_ERR_CUDA = -4,
    WSO_ERR_OUT_OF_MEMORY = -5,
    WSO_ERR_H0_NOT_CONJUGATE = -6 /* imported h0 violates heightAmp_conj == conj(heightAmp), which every
                                     reference-built h0 satisfies (WSTessendorf.cpp:132-135) and the kernels rely on */
} wso_status;

/* The reference's tunables (setters: WSTessendorf.cpp:459-505; defaults: WSTessendorf.h:36-43,181). */
typedef struct wso_params {
    uint32_t tile_size;    /* N: points == waves per side, power of two      (SetTileSize)        */
    float tile_length;     /* L: world-space tile length                      (SetTileLength)      */
    float wind_dir_x;      /* any non-zero vector; normalised on use          (SetWindDirection)   */
    float wind_dir_y;
    float wind_speed;      /* clamped to >= 1e-4                              (SetWindSpeed)       */
    float anim_period;     /* T; base frequency w0 = 2*pi/T                   (SetAnimationPeriod) */
    float phillips_const;  /* A of the Phillips spectrum                      (SetPhillipsConst)   */
    float damping;         /* suppresses wave lengths below it                (SetDamping)         */
    float lambda;          /* choppy-displacement scale; no Prepare() needed  (SetLambda)          */
} wso_params;

/* The reference's 20-byte h0 record (struct BaseWaveHeight, WSTessendorf.h:142-147). */
typedef struct wso_h0_record {
    float amp_re, amp_im;            /* heightAmp       */
    float amp_conj_re, amp_conj_im;  /* heightAmp_conj  */
    float dispersion;                /* quantised omega */
} wso_h0_record;

enum { WSO_MAP_DISPLACEMENT = 0, WSO_MAP_NORMAL = 1 };

/* Fill *p with the reference defaults (WSTessendorf.h:36-43, lambda = -1: WSTessendorf.h:181). */
WSO_API int wso_default_params(wso_params* p);

/* Replaces: WSTessendorf::WSTessendorf(tileSize, tileLength) (WSTessendorf.cpp:13-26).
 * device: CUDA ordinal.  max_tiles: independent parameter sets / h0 fields sharing tile_size (>= 1).
 * max_slots: output map pairs kept resident on the device (>= 1; a batch writes one slot per tile-frame). */
WSO_API int wso_create(const wso_params* p, int device, uint32_t max_tiles, uint32_t max_slots,
                       wso_ctx** out);
/* Replaces: WSTessendorf::~WSTessendorf (WSTessendorf.cpp:28-34). */
WSO_API int wso_destroy(wso_ctx* ctx);

/* Replaces: the Set* family (WSTessendorf.cpp:459-505).  tile_size / length / wind / period / A / damping
 * take effect at the next wso_prepare* of that tile (as in the reference); lambda takes effect at the
 * next compute.  A tile_size different from the context's is accepted only while max_tiles == 1. */
WSO_API int wso_set_params(wso_ctx* ctx, uint32_t tile, const wso_params* p);
/* Replaces: the Get* family (WSTessendorf.h:82-89); wind direction is returned normalised. */
WSO_API int wso_get_params(const wso_ctx* ctx, uint32_t tile, wso_params* p);
WSO_API int wso_set_lambda(wso_ctx* ctx, uint32_t tile, float lambda);

/* Replaces: WSTessendorf::Prepare() (WSTessendorf.cpp:36-58) — wave vectors, Gaussian array drawn from
 * the C library rand() exactly like glm::gaussRand does (libs/glm/glm/gtc/random.inl:218-232), Phillips
 * h0(k) and dispersion; buffers (re)sized.  reseed != 0 calls srand(seed) first (the reference app seeds
 * once with the wall clock: core/Application.cpp:21). */
WSO_API int wso_prepare(wso_ctx* ctx, uint32_t tile, int reseed, unsigned seed);
/* Same, with a caller-supplied Gaussian array xi: N*N complex<float> (re,im), row-major [m][n]
 * (the array ComputeGaussRandomArray returns, WSTessendorf.cpp:87-103). */
WSO_API int wso_prepare_gauss(wso_ctx* ctx, uint32_t tile, const float* xi);
/* Import / export h0 in the reference's own layout: N*N records, row-major [m][n]
 * (m_BaseWaveHeights, WSTessendorf.h:197).  Import replaces the spectrum of a tile whose parameters
 * (tile_length in particular) were set beforehand; this is also the checkpoint/restore mechanism. */
WSO_API int wso_import_h0(wso_ctx* ctx, uint32_t tile, const wso_h0_record* h0);
WSO_API int wso_export_h0(const wso_ctx* ctx, uint32_t tile, wso_h0_record* h0);

/* Replaces: float WSTessendorf::ComputeWaves(float t) (WSTessendorf.cpp:284-441) for tile 0 -> slot 0.
 * Blocking.  On return both maps are in pinned host memory (wso_map_host) and on the device
 * (wso_map_device); *amplitude receives the return value A. */
WSO_API int wso_compute(wso_ctx* ctx, float t, float* amplitude);

/* Batched form: n tile-frames.  Item i evolves tile tiles[i] (NULL = all tile 0) to time t[i] and writes
 * device map slot first_slot + i.  Asynchronous on the context's stream; results stay on the device. */
WSO_API int wso_compute_batch(wso_ctx* ctx, uint32_t n, const uint32_t* tiles, const float* t,
                              uint32_t first_slot);
/* Batched form with host output: like wso_compute_batch, then streams each tile-frame's maps into
 * disp_host / norm_host (n * N*N*4 floats each, ideally pinned: wso_alloc_host) overlapping copies with
 * the next chunk's kernels; amplitude/min/max may be NULL.  Blocking. */
WSO_API int wso_compute_to_host(wso_ctx* ctx, uint32_t n, const uint32_t* tiles, const float* t,
                                float* disp_host, float* norm_host, float* amplitude,
                                float* min_height, float* max_height);
WSO_API int wso_sync(wso_ctx* ctx);

/* Replaces: GetMinHeight()/GetMaxHeight() (WSTessendorf.h:90-91) and the ComputeWaves return value,
 * for n slots starting at first_slot.  Blocking (small device->host read).  Any pointer may be NULL. */
WSO_API int wso_read_heights(wso_ctx* ctx, uint32_t first_slot, uint32_t n, float* amplitude,
                             float* min_height, float* max_height);

/* Replaces: GetDisplacements()/GetNormals() + Get*Count() (WSTessendorf.h:95-101): tightly packed
 * RGBA32F texels, row-major [m][n], N*N of them.
 *   displacement = (lambda*Dx, height/A, lambda*Dz, 1)      normal = (dh/dx, dh/dz, dDx/dx, dDz/dz)
 * wso_map_host: pinned host copy of slot 0 written by wso_compute (valid until the next compute/prepare).
 * wso_map_device: device pointer of any slot (valid until destroy / tile-size change).
 * wso_copy_map: blocking copy of one slot's map into caller memory. */
WSO_API int wso_map_host(wso_ctx* ctx, int which, const float** ptr, size_t* texels);
WSO_API int wso_map_device(wso_ctx* ctx, int which, uint32_t slot, void** dptr, size_t* texels);
WSO_API int wso_copy_map(wso_ctx* ctx, int which, uint32_t slot, float* dst_host);

/* Plumbing: run on a caller-owned CUDA stream (cudaStream_t as void*; NULL = the context's own), and
 * pinned host allocations for wso_compute_to_host. */
WSO_API int wso_set_stream(wso_ctx* ctx, void* cuda_stream);
WSO_API int wso_alloc_host(size_t bytes, void** ptr);
WSO_API int wso_free_host(void* ptr);

/* Introspection used by the benchmark: number of kernels launched so far by this context, the chunk
 * (tile-frames per launch) the batched calls use, and the CTA tiling of the two transform kernels. */
WSO_API int wso_get_stats(const wso_ctx* ctx, uint64_t* kernel_launches, uint32_t* chunk);

/* Opt-in per-kernel device timing (the analogue of the reference's VKP_PROFILE_SCOPE table,
 * core/Profile.h:16-32): with profiling on, every launch is bracketed by CUDA events on the compute
 * stream.  wso_get_profile synchronises and returns the accumulated milliseconds of
 * {K1 evolve+first transform, K2h height extrema, K2 second transform+pack}, the number of launches of each
 * and the tile-frames they covered since wso_set_profiling(ctx, 1). */
WSO_API int wso_set_profiling(wso_ctx* ctx, int on);
WSO_API int wso_get_profile(wso_ctx* ctx, double* kernel_ms3, uint64_t* launches, uint64_t* tile_frames);

WSO_API const char* wso_last_error(const wso_ctx* ctx); /* ctx may be NULL: last create failure */
WSO_API const char* wso_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WSOCEAN_H_ */
