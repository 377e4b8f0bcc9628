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
/* Replaces: the compile-time switch COMPUTE_JACOBIAN (WSTessendorf.cpp:158-175, 330-335, 368-377, 421-428; dead code in
 * the reference - it does not compile there - so this is an extension with no reference output to pin it to, SURVEY
 * row f-4).  on != 0: every following compute writes displacement.w = J = (1 + l*dxDx)(1 + l*dzDz) - (l*dxDz)(l*dzDx)
 * (l = lambda, all four derivatives already carrying the (-1)^(m+n) sign) instead of 1.0f; J < 0 marks folded
 * (foam) texels.  No extra transform: dxDz == dzDx rides in the empty real slot of the packed field that carries Dz.
 * Tile sizes up to 4096; not available on the slab path.  Default off = the reference's behaviour. */
WSO_API int wso_set_compute_jacobian(wso_ctx* ctx, int on);
WSO_API int wso_get_compute_jacobian(const wso_ctx* ctx, int* on);

/* Replaces: WSTessendorf::Prepare() (WSTessendorf.cpp:36-58) — wave vectors, Gaussian array drawn from
 * the C library rand() exactly like glm::gaussRand does (libs/glm/glm/gtc/random.inl:218-232), Phillips
 * h0(k) and dispersion; buffers (re)sized.  reseed != 0 calls srand(seed) first (the reference app seeds
 * once with the wall clock: core/Application.cpp:21). */
WSO_API int wso_prepare(wso_ctx* ctx, uint32_t tile, int reseed, unsigned seed);
/* Same, with a caller-supplied Gaussian array xi: N*N complex<float> (re,im), row-major [m][n]
 * (the array ComputeGaussRandomArray returns, WSTessendorf.cpp:87-103). */
WSO_API int wso_prepare_gauss(wso_ctx* ctx, uint32_t tile, const float* xi);
/* Prepare() ON THE DEVICE (replaces the CPU passes ComputeWaveVectors + ComputeBaseWaveHeightField,
 * WSTessendorf.cpp:60-85, 105-148, with PhillipsSpectrum / BaseWaveHeightFT / QDispersion, WSTessendorf.h:237-297):
 * one kernel builds the spectrum records straight in device memory from the uploaded Gaussian array xi (same
 * layout as above).  Dispersion, 1/|k| and layout are bit-identical to wso_prepare_gauss; amplitudes agree to
 * 1 ulp (the two exp() are evaluated in float64 on the device).  No host copy of h0 is kept: wso_export_h0 reads
 * the records back. */
WSO_API int wso_prepare_gauss_device(wso_ctx* ctx, uint32_t tile, const float* xi);
/* Same, the Gaussian array drawn in the kernel from the counter-based generator (seed, m*N+n) - a reproducible
 * stand-in for the reference's serial rand() stream (ComputeGaussRandomArray, WSTessendorf.cpp:87-103); the host
 * build of the same generator is wso_counter_h0. */
WSO_API int wso_prepare_counter(wso_ctx* ctx, uint32_t tile, uint64_t seed);
/* Import / export h0 in the reference's own layout: N*N records, row-major [m][n]
 * (m_BaseWaveHeights, WSTessendorf.h:197).  Import replaces the spectrum of a tile whose parameters
 * (tile_length in particular) were set beforehand; this is also the checkpoint/restore mechanism. */
WSO_API int wso_import_h0(wso_ctx* ctx, uint32_t tile, const wso_h0_record* h0);
WSO_API int wso_export_h0(const wso_ctx* ctx, uint32_t tile, wso_h0_record* h0);
/* Compact form, 12 bytes per wave vector (amp.re, amp.im, dispersion), row-major [m][n]: heightAmp_conj == conj(heightAmp)
 * in every reference-built h0 (WSTessendorf.cpp:132-135), so nothing is lost against the 20-byte BaseWaveHeight record. */
WSO_API int wso_import_h0_compact(wso_ctx* ctx, uint32_t tile, const float* h0_3f);
WSO_API int wso_export_h0_compact(const wso_ctx* ctx, uint32_t tile, float* h0_3f);

/* Replaces: float WSTessendorf::ComputeWaves(float t) (WSTessendorf.cpp:284-441) for tile 0 -> slot 0.
 * Blocking.  On return both maps are in pinned host memory (wso_map_host) and on the device
 * (wso_map_device); *amplitude receives the return value A. */
WSO_API int wso_compute(wso_ctx* ctx, float t, float* amplitude);
/* The same frame without the host copy and without blocking: tile 0 -> device slot 0 on the context's stream.  *event_out
 * is a cudaEvent_t (as void*) owned by the context and valid until the next call; it completes when both maps, A and
 * min/max of the frame are final on the device (wso_map_device, wso_read_heights).  A CUDA caller orders its own stream
 * behind it with cudaStreamWaitEvent; wso_wait_event blocks the host. */
WSO_API int wso_compute_async(wso_ctx* ctx, float t, void** event_out);
WSO_API int wso_wait_event(wso_ctx* ctx, void* event);

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
 *   Between Prepare() and the first compute every texel holds the reference's resize() defaults
 *   (WSTessendorf.cpp:48-54): displacement (0,0,0,0), normal (0,1,0,0) - on the host mirrors and in every device slot.
 * wso_map_device: device pointer of any slot (valid until destroy / tile-size change).
 * wso_copy_map: blocking copy of one slot's map into caller memory. */
WSO_API int wso_map_host(wso_ctx* ctx, int which, const float** ptr, size_t* texels);
WSO_API int wso_map_device(wso_ctx* ctx, int which, uint32_t slot, void** dptr, size_t* texels);
WSO_API int wso_copy_map(wso_ctx* ctx, int which, uint32_t slot, float* dst_host);

/* ------------------------------------------------------------------------------------------------------------
 * External-memory interop: the maps land in memory a Vulkan device can bind, so the renderer's
 * vkCmdCopyBufferToImage reads them where K2 wrote them.  Replaces: the two 16*N*N-byte memcpy calls into the mapped
 * staging buffer and the PCIe upload behind them (WaterSurfaceMesh.cpp:701-755, vulkan/Texture2D.cpp:175-226).
 * Both directions of VK_KHR_external_memory_fd are offered:
 *   wso_set_exportable + wso_export_fd : the library allocates the map arrays with the CUDA virtual-memory API
 *       (POSIX-fd shareable) and hands out a file descriptor; Vulkan imports it with VkImportMemoryFdInfoKHR
 *       { handleType = VK_EXTERNAL_MEMORY_HANDLE_TYPE_OPAQUE_FD_BIT } and binds a VkBuffer of *bytes to it.
 *       The caller owns the returned fd (close() it, or let the Vulkan import consume it).
 *   wso_import_external_fd : Vulkan allocated and exported the memory (vkGetMemoryFdKHR); the library maps
 *       [offset, offset + max_slots*16*N*N) of it as the output array of map `which`.  On success the fd belongs
 *       to CUDA.  A tile-size change (Prepare after SetTileSize) drops the import and returns to own memory.
 * Layout either way: [slot][N*N] RGBA32F texels, slot s at byte offset s*16*N*N (+ offset), row pitch 16*N.
 * Map contents are reset to zero by wso_set_exportable / wso_import_external_fd; Prepare() state is kept.
 * Ordering with the renderer: wso_import_semaphore_fd binds an exported VkSemaphore (binary: is_timeline = 0,
 * timeline: 1; `index` 0 or 1); wso_signal_semaphore / wso_wait_semaphore enqueue a signal / wait on the context's
 * stream after / before the work enqueued around them (value ignored for binary semaphores). */
WSO_API int wso_set_exportable(wso_ctx* ctx, int on);
WSO_API int wso_export_fd(wso_ctx* ctx, int which, int* fd, size_t* bytes);
WSO_API int wso_import_external_fd(wso_ctx* ctx, int which, int fd, size_t bytes, size_t offset);
WSO_API int wso_import_semaphore_fd(wso_ctx* ctx, int index, int fd, int is_timeline);
WSO_API int wso_signal_semaphore(wso_ctx* ctx, int index, uint64_t value);
WSO_API int wso_wait_semaphore(wso_ctx* ctx, int index, uint64_t value);

/* Plumbing: run on a caller-owned CUDA stream (cudaStream_t as void*; NULL = the context's own), and
 * pinned host allocations for wso_compute_to_host. */
WSO_API int wso_set_stream(wso_ctx* ctx, void* cuda_stream);
WSO_API int wso_alloc_host(size_t bytes, void** ptr);
WSO_API int wso_free_host(void* ptr);
/* Page-lock memory the CALLER owns (e.g. the storage of the std::vector<glm::vec4> the reference's accessors hand out,
 * WSTessendorf.h:95-101) so that wso_compute_to_host copies straight into it at full PCIe speed; unregister before
 * the memory is freed or reallocated.  Failing to register is not fatal: the copies then go through pageable memory. */
WSO_API int wso_register_host(void* ptr, size_t bytes);
WSO_API int wso_unregister_host(void* ptr);

/* Two implementations of the three hot-path kernels exist for 512^2, 1024^2 and 2048^2 batched launches: CTA-per-line
 * (Stockham stages through shared memory) and warp-per-line (radix-32 register stages, shuffle exchanges, bulk-copy line
 * pipeline).  They produce the same maps within the parity tolerance and share the intermediate layout.  mask bit 0 / 1 /
 * 2 puts K1 / K2h / K2 on the warp-per-line set; bits 4 / 6 (0x10 / 0x40) run K1 / K2 of the CTA-per-line set in their
 * persistent form (a fixed grid of CTAs walks the work items and requests the inputs of its next item ahead of the store
 * phase of the current one; bit-identical results); -1 restores the built-in choice (what measured faster per size).
 * Any other bit: WSO_ERR_INVALID_ARG.  Process-wide; meant for A/B measurements and for the parity tests, which run
 * every form. */
WSO_API int wso_select_kernels(int mask);

/* Introspection used by the benchmark: number of kernels launched so far by this context, the chunk
 * (tile-frames per launch) the batched calls use, and the CTA tiling of the two transform kernels. */
WSO_API int wso_get_stats(const wso_ctx* ctx, uint64_t* kernel_launches, uint32_t* chunk);

/* A call that computes ONE tile-frame (wso_compute, wso_compute_async, wso_compute_to_host or wso_compute_batch with
 * n == 1 - the reference's one ComputeWaves(t) per rendered frame, WaterSurfaceMesh.cpp:123-154) is enqueued as one
 * launch of an instantiated CUDA graph holding the frame's three kernels, whose parameters are rewritten per call,
 * instead of three kernel launches.  on = 1 (default): for tile sizes up to 512, where the launches cost more host time
 * than the kernels take on the device (12.3 instead of 15.6 us per 512^2 frame back to back on B200); larger tiles keep
 * plain launches, whose programmatic dependent launch overlaps consecutive frames.  on = 2: every size; on = 0: never
 * (also: environment WSO_FRAME_GRAPH=0|1|2).  The counters say how many frames went out as graph launches and how often
 * the graph was (re)captured. */
WSO_API int wso_set_frame_graph(wso_ctx* ctx, int on);
WSO_API int wso_get_frame_graph_stats(const wso_ctx* ctx, uint64_t* graph_launches, uint64_t* captures);

/* Opt-in per-kernel device timing (the analogue of the reference's VKP_PROFILE_SCOPE table,
 * core/Profile.h:16-32): with profiling on, every launch is bracketed by CUDA events on the compute
 * stream.  wso_get_profile synchronises and returns the accumulated milliseconds of
 * {K1 evolve+first transform, K2h height extrema, K2 second transform+pack}, the number of launches of each
 * and the tile-frames they covered since wso_set_profiling(ctx, 1). */
WSO_API int wso_set_profiling(wso_ctx* ctx, int on);
WSO_API int wso_get_profile(wso_ctx* ctx, double* kernel_ms3, uint64_t* launches, uint64_t* tile_frames);

/* ------------------------------------------------------------------------------------------------------------
 * Slab-decomposed path: ONE large grid (2048..16384 per side) spread over P = 1, 2, 4 or 8 devices, one process
 * and one wso_slab per device (BASELINE.json configs[4]; no reference counterpart - the reference caps the tile at
 * 1024^2, WaterSurfaceMesh.h:45-46 - the arithmetic is the same ComputeWaves, WSTessendorf.cpp:284-441).
 * Rank r evolves and transforms (along m) the column pairs (n, N-n) for n in [r*N/2P, (r+1)*N/2P) and, after the
 * exchange, transforms (along n) and packs the rows m' and N-m' for m' in the same range.  Per tile-frame:
 *     wso_slab_pass1(t)  ->  wso_slab_exchange  ->  wso_slab_heights  ->  all-reduce (min, max)  ->  wso_slab_pass2
 * pass1 stores what it separates where its threads hold it (coalesced); wso_slab_exchange is ONE transposing kernel
 * that carries the Hl x Hl blocks to the owners of the row items in 256-byte rows:
 *   fused   - straight into the owners' receive buffers over NVLink peer mappings (wso_slab_ipc_handle /
 *             wso_slab_open_peer / wso_slab_set_fused), followed by any stream-ordered inter-rank barrier of the caller's;
 *             the receive buffers alternate by frame parity, so frames may be enqueued back to back without host syncs;
 *   unfused - into the `world` equal blocks of the local send buffer, which the caller then moves with ONE all-to-all
 *             into the receive buffer (e.g. torch.distributed.all_to_all_single over NCCL) on the slab's stream.
 * The all-reduce (min over [0], max over [1] of the 2-float minmax buffer) is the caller's as well. */
typedef struct wso_slab wso_slab;
WSO_API int wso_slab_create(const wso_params* p, int device, uint32_t rank, uint32_t world, wso_slab** out);
WSO_API int wso_slab_destroy(wso_slab* s);
/* Spectrum of this rank's columns: from a full N*N array in the reference layout (small grids, tests), or generated
 * in place from a counter-based Gaussian source (seed, m*N+n) with the reference's Phillips/dispersion arithmetic
 * (WSTessendorf.cpp:105-148).  wso_counter_h0 evaluates the same records for rows [m0, m0+rows) - for checkers. */
WSO_API int wso_slab_import_h0(wso_slab* s, const wso_h0_record* h0_full);
WSO_API int wso_slab_prepare_counter(wso_slab* s, uint64_t seed);
/* the same spectrum built by the device Prepare kernel (this rank's column pairs only; no host pass) */
WSO_API int wso_slab_prepare_counter_device(wso_slab* s, uint64_t seed);
WSO_API int wso_counter_h0(const wso_params* p, uint64_t seed, uint32_t m0, uint32_t rows, wso_h0_record* out);
WSO_API int wso_slab_set_lambda(wso_slab* s, float lambda);
WSO_API int wso_slab_set_stream(wso_slab* s, void* cuda_stream);
/* Exchange buffers: send/recv = `world` blocks of block_bytes each (block d of send goes to rank d; block r of recv
 * came from rank r); minmax = 2 floats written by wso_slab_heights, to be all-reduced before wso_slab_pass2. */
WSO_API int wso_slab_buffers(wso_slab* s, void** send, void** recv, size_t* block_bytes, void** minmax);
/* Fused exchange: 64-byte CUDA IPC handle of this rank's receive buffer; map every peer's, then switch on. */
WSO_API int wso_slab_ipc_handle(wso_slab* s, void* handle64);
WSO_API int wso_slab_open_peer(wso_slab* s, uint32_t peer, const void* handle64);
WSO_API int wso_slab_set_fused(wso_slab* s, int on);
/* Use the two-CTA cluster variant of K2 (always used at 16384, where a line pair exceeds one SM's shared memory). */
WSO_API int wso_slab_force_pair(wso_slab* s, int on);
WSO_API int wso_slab_pass1(wso_slab* s, float t);
WSO_API int wso_slab_exchange(wso_slab* s);
/* Pipelined form of the two calls above: packed fields [field0, field0 + nfields) only (whole field groups,
 * wso_slab_fields_per_group), the exchange on `cuda_stream` (NULL = the slab's stream), so that the transfer of one field
 * overlaps the transform of the next.  wso_slab_pass1_fields with field0 == 0 starts a new tile-frame. */
WSO_API int wso_slab_pass1_fields(wso_slab* s, float t, int field0, int nfields);
WSO_API int wso_slab_exchange_fields(wso_slab* s, int field0, int nfields, void* cuda_stream);
WSO_API int wso_slab_fields_per_group(const wso_slab* s);
WSO_API int wso_slab_heights(wso_slab* s);
WSO_API int wso_slab_pass2(wso_slab* s);
WSO_API int wso_slab_sync(wso_slab* s);
WSO_API int wso_slab_read_heights(wso_slab* s, float* amplitude, float* min_height, float* max_height);
/* Output: this rank's 2*N/2P rows of each map, N RGBA32F texels per row; wso_slab_row_index gives the global row
 * of every local row (local row ml < N/2P is row m' = r*N/2P + ml, local row N/2P + ml its mirror N-m'; N/2 for m'=0). */
WSO_API int wso_slab_map_device(wso_slab* s, int which, void** dptr, uint32_t* rows);
WSO_API int wso_slab_copy_rows(wso_slab* s, int which, float* dst_host);
WSO_API int wso_slab_row_index(const wso_slab* s, uint32_t* rows);
WSO_API const char* wso_slab_last_error(const wso_slab* s);

WSO_API const char* wso_last_error(const wso_ctx* ctx); /* ctx may be NULL: last create failure */
WSO_API const char* wso_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WSOCEAN_H_ */
