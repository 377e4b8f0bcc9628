// wso_slab.cu — C ABI of the slab-decomposed path (include/wsocean.h, "slab" section): ONE large grid spread over
// P devices, one process per device (BASELINE config 5).
//
// Per tile-frame:  pass1 (K1 on this rank's column pairs; its stores are the transpose)  ->  exchange (peer stores
// over NVLink in fused mode, else one all-to-all of the send blocks, issued by the host language's collective
// library)  ->  heights (K2h on this rank's row items)  ->  all-reduce of (min, max)  ->  pass2 (K2).
// The collectives themselves are the caller's (torch.distributed / NCCL): this library owns kernels and buffers.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "wso_host_prepare.h"
#include "wso_kernels.cuh"
#include "wso_launch.h"
#include "wsocean.h"

using wso::TileDev;
using SlabArgs = wso::LaunchArgsT<1>;

struct wso_slab {
    int device = 0;
    uint32_t rank = 0, world = 1;
    int shift = 0;
    uint32_t n = 0, hl = 0;  // grid size, column pairs / row items per rank
    int logn = 0;
    wso_params params;
    bool prepared = false;
    bool force_pair = false;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    float4* d_h0 = nullptr;   // [hl][2][n]
    float4* d_hs = nullptr;   // [hl][n/2][2]
    float* d_kv = nullptr;    // [n]
    float2* d_tw = nullptr;   // [n]
    float2* d_stage = nullptr;  // K1's staged output [4][2][hl][n/2] (coalesced stores; hl >= 32)
    float2* d_send = nullptr;   // [world][hl][4][2][hl]: what the exchange kernel fills when a collective library moves the blocks
    float2* d_recv = nullptr;   // 2 x [world][hl][4][2][hl]: receive buffers by frame parity (direct peer stores), else the first
    uint64_t frame = 0;         // tile-frames started (wso_slab_pass1)
    float4* d_disp = nullptr; // [2*hl][n]
    float4* d_norm = nullptr;
    float* d_minmax = nullptr;  // [2]
    float* d_ampl = nullptr;    // [1]
    float* h_small = nullptr;   // pinned [3]
    float2* peers[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    bool fused = false;
    std::vector<void*> opened;  // IPC mappings to close
    TileDev td{};
    float last_t = 0.0f;
    std::string err;
};

namespace {

thread_local std::string g_slab_create_error;

int sfail(wso_slab* s, int code, const std::string& msg) {
    if (s) s->err = msg;
    else g_slab_create_error = msg;
    return code;
}
int sfail_cuda(wso_slab* s, cudaError_t e, const char* what) {
    return sfail(s, e == cudaErrorMemoryAllocation ? WSO_ERR_OUT_OF_MEMORY : WSO_ERR_CUDA,
                 std::string(what) + ": " + cudaGetErrorString(e));
}
#define SLAB_CUDA(s, call)                                        \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return sfail_cuda(s, e_, #call);   \
    } while (0)

size_t block_elems(const wso_slab* s) { return (size_t)s->hl * 4 * 2 * s->hl; }

// global column index of local column (jl, half)
inline uint32_t column_of(const wso_slab* s, uint32_t jl, int half) {
    const uint32_t j = s->rank * s->hl + jl;
    return half == 0 ? j : (j == 0 ? s->n / 2 : s->n - j);
}

// Build the device arrays from a functor giving the reference record of wave vector (m, n).
template <class RecordAt>
int upload_local(wso_slab* s, RecordAt&& record_at) {
    const uint32_t n = s->n, hl = s->hl, H = n / 2;
    std::vector<float> kv;
    wso::host_wave_numbers(n, s->params.tile_length, kv);
    const float omega0 = wso::derive_params(s->params).base_freq;
    std::vector<float4> rec((size_t)hl * 2 * n);
    std::vector<float4> recs((size_t)hl * H * 2, make_float4(0.f, 0.f, 0.f, 0.f));
    unsigned nthreads = std::thread::hardware_concurrency() / (s->world ? s->world : 1);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 16) nthreads = 16;
    if ((size_t)hl * n < (size_t)1 << 20) nthreads = 1;
    std::vector<int> jmax(nthreads, 0), bad(nthreads, 0);
    auto work = [&](unsigned tix) {
        for (uint32_t jl = tix; jl < hl; jl += nthreads) {
            for (int half = 0; half < 2; ++half) {
                const uint32_t col = column_of(s, jl, half);
                float4* dst = rec.data() + ((size_t)jl * 2 + half) * n;
                for (uint32_t m = 0; m < n; ++m) {
                    const wso_h0_record r = record_at(m, col);
                    if (!(r.amp_conj_re == r.amp_re && r.amp_conj_im == -r.amp_im)) bad[tix] |= 1;
                    // 1/|k| exactly as glm::normalize computes it (reference: WSTessendorf.h:133-136)
                    const float d = kv[col] * kv[col] + kv[m] * kv[m];
                    const float inv = std::sqrt(d) > 0.00001f ? 1.0f / std::sqrt(d) : 0.0f;
                    // the slab kernels always use the per-frame sincos table: omega must be an exact multiple of omega0
                    const float jf = std::nearbyint(r.dispersion / omega0);
                    if (!(jf >= 0.0f && jf < (float)wso::kMaxTable) || jf * omega0 != r.dispersion) bad[tix] |= 2;
                    const int ji = (int)jf;
                    if (ji > jmax[tix]) jmax[tix] = ji;
                    float wfield;
                    std::memcpy(&wfield, &ji, sizeof(float));
                    dst[m] = make_float4(r.amp_re, r.amp_im, inv, wfield);
                }
            }
            if (s->rank * hl + jl == 0) continue;  // column pair 0 goes through the per-point records
            const float4* cA = rec.data() + (size_t)jl * 2 * n;
            const float4* cB = cA + n;
            for (uint32_t i = 1; i < H; ++i) {
                const float4 a0 = cA[i], a3 = cB[n - i], a1 = cB[i], a2 = cA[n - i];
                if (std::memcmp(&a0.w, &a3.w, 4) != 0 || std::memcmp(&a1.w, &a2.w, 4) != 0) bad[tix] |= 4;
                recs[wso::hs_index((int)jl, (int)i, 0, (int)H)] = make_float4(a0.x + a3.x, a0.y + a3.y, a0.z, a0.w);
                recs[wso::hs_index((int)jl, (int)i, 1, (int)H)] = make_float4(a1.x + a2.x, a1.y + a2.y, a1.z, a1.w);
            }
        }
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (unsigned tix = 0; tix < nthreads; ++tix) th.emplace_back(work, tix);
        for (auto& t : th) t.join();
    }
    int jm = 0, flags = 0;
    for (unsigned tix = 0; tix < nthreads; ++tix) {
        if (jmax[tix] > jm) jm = jmax[tix];
        flags |= bad[tix];
    }
    if (flags & 1) return sfail(s, WSO_ERR_H0_NOT_CONJUGATE, "h0: heightAmp_conj != conj(heightAmp)");
    if (flags & 6) return sfail(s, WSO_ERR_INVALID_ARG, "slab path needs the reference's quantised, even dispersion");
    SLAB_CUDA(s, cudaSetDevice(s->device));
    SLAB_CUDA(s, cudaStreamSynchronize(s->stream));
    SLAB_CUDA(s, cudaMemcpy(s->d_h0, rec.data(), sizeof(float4) * rec.size(), cudaMemcpyHostToDevice));
    SLAB_CUDA(s, cudaMemcpy(s->d_hs, recs.data(), sizeof(float4) * recs.size(), cudaMemcpyHostToDevice));
    SLAB_CUDA(s, cudaMemcpy(s->d_kv, kv.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
    s->td.h0 = s->d_h0;
    s->td.hs = s->d_hs;
    s->td.kv = s->d_kv;
    s->td.lambda = s->params.lambda;
    s->td.omega0 = omega0;
    s->td.table_len = jm + 1;
    s->td.use_pairs = 1;
    s->td.j0 = (int)(s->rank * hl);
    s->prepared = true;
    return WSO_OK;
}

// Receive buffer of the current frame.  With direct peer stores a rank may already write frame f+1 into its peers while
// they still transform frame f (the last collective of a frame, the min/max all-reduce, precedes K2): the buffers
// alternate by frame parity.  A frame f+2 store cannot overtake K2 of frame f: the writer has passed the all-reduce of
// frame f+1 by then, which the reader only enters after its K2h(f+1), stream-ordered behind its K2(f).
inline size_t recv_offset(const wso_slab* s) { return s->fused ? (size_t)(s->frame & 1) * block_elems(s) * s->world : 0; }
inline bool staged(const wso_slab* s) { return s->hl >= 32; }

// where this rank's separated half-spectra go, per owner d of the row items
void exchange_dst(const wso_slab* s, float2* (&dst)[8]) {
    const size_t blk = block_elems(s);
    for (uint32_t d = 0; d < 8; ++d) dst[d] = nullptr;
    for (uint32_t d = 0; d < s->world; ++d)
        dst[d] = s->fused ? s->peers[d] + recv_offset(s) + (size_t)s->rank * blk
               : (s->world == 1 ? s->d_recv : s->d_send + (size_t)d * blk);  // one rank: nothing to exchange
}

void fill_args(const wso_slab* s, SlabArgs& a, float t) {
    std::memset(&a, 0, sizeof(a));
    a.tw = s->d_tw;
    a.W = s->d_recv + recv_offset(s);
    a.slab_stage = staged(s) ? s->d_stage : nullptr;
    a.disp = s->d_disp;
    a.norm = s->d_norm;
    a.minmax = s->d_minmax;
    a.amp_out = s->d_ampl;
    a.slab_shift = s->shift;
    a.slab_rank = (int)s->rank;
    exchange_dst(s, a.Wdst);
    a.items[0].tile = 0;
    a.items[0].slot = 0;
    a.items[0].t = t;
    a.td[0] = s->td;
    a.td[0].lambda = s->params.lambda;
}

}  // namespace

extern "C" {

int wso_slab_create(const wso_params* p, int device, uint32_t rank, uint32_t world, wso_slab** out) {
    if (!p || !out || world == 0 || world > 8 || (world & (world - 1)) != 0 || rank >= world)
        return sfail(nullptr, WSO_ERR_INVALID_ARG, "world must be 1, 2, 4 or 8 and rank < world");
    *out = nullptr;
    const uint32_t n = p->tile_size;
    int logn = 0;
    while ((1u << logn) < n) ++logn;
    if (n == 0 || (1u << logn) != n || !wso::slab_size_supported(logn))
        return sfail(nullptr, WSO_ERR_BAD_TILE_SIZE, "slab path: tile size must be 64, 256 or 2048..16384");
    if (!(p->tile_length > 0.0f) || !(p->anim_period > 0.0f) || (p->wind_dir_x == 0.0f && p->wind_dir_y == 0.0f))
        return sfail(nullptr, WSO_ERR_INVALID_ARG, "invalid parameters");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) return sfail(nullptr, WSO_ERR_CUDA, "no usable CUDA device (no CPU fallback exists)");
    if (device < 0 || device >= ndev) return sfail(nullptr, WSO_ERR_INVALID_ARG, "device ordinal out of range");
    wso_slab* s = new (std::nothrow) wso_slab();
    if (!s) return sfail(nullptr, WSO_ERR_OUT_OF_MEMORY, "host allocation failed");
    s->device = device;
    s->rank = rank;
    s->world = world;
    while ((1u << s->shift) < world) ++s->shift;
    s->n = n;
    s->logn = logn;
    s->hl = (n / 2) / world;
    s->params = *p;
    wso::normalise_like_setters(s->params, nullptr);
    int rc = WSO_OK;
    auto alloc = [&]() -> int {
        SLAB_CUDA(s, cudaSetDevice(device));
        SLAB_CUDA(s, cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking));
        s->stream = s->own_stream;
        const size_t blk = block_elems(s);
        SLAB_CUDA(s, cudaMalloc(&s->d_h0, sizeof(float4) * (size_t)s->hl * 2 * n));
        SLAB_CUDA(s, cudaMalloc(&s->d_hs, sizeof(float4) * (size_t)s->hl * (n / 2) * 2));
        SLAB_CUDA(s, cudaMalloc(&s->d_kv, sizeof(float) * n));
        SLAB_CUDA(s, cudaMalloc(&s->d_tw, sizeof(float2) * n));
        SLAB_CUDA(s, cudaMalloc(&s->d_stage, sizeof(float2) * blk * world));
        SLAB_CUDA(s, cudaMalloc(&s->d_send, sizeof(float2) * blk * world));
        SLAB_CUDA(s, cudaMalloc(&s->d_recv, sizeof(float2) * blk * world * 2));
        SLAB_CUDA(s, cudaMalloc(&s->d_disp, sizeof(float4) * (size_t)2 * s->hl * n));
        SLAB_CUDA(s, cudaMalloc(&s->d_norm, sizeof(float4) * (size_t)2 * s->hl * n));
        SLAB_CUDA(s, cudaMalloc(&s->d_minmax, sizeof(float) * 2));
        SLAB_CUDA(s, cudaMalloc(&s->d_ampl, sizeof(float)));
        SLAB_CUDA(s, cudaMallocHost(&s->h_small, sizeof(float) * 3));
        std::vector<float2> tw(n);
        for (uint32_t k = 0; k < n; ++k) {
            const double a = 2.0 * 3.14159265358979323846 * (double)k / (double)n;
            tw[k] = make_float2((float)std::cos(a), (float)std::sin(a));
        }
        SLAB_CUDA(s, cudaMemcpy(s->d_tw, tw.data(), sizeof(float2) * n, cudaMemcpyHostToDevice));
        return WSO_OK;
    };
    rc = alloc();
    if (rc != WSO_OK) {
        g_slab_create_error = s->err;
        wso_slab_destroy(s);
        return rc;
    }
    *out = s;
    return WSO_OK;
}

int wso_slab_destroy(wso_slab* s) {
    if (!s) return WSO_OK;
    cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (void* p : s->opened) cudaIpcCloseMemHandle(p);
    cudaFree(s->d_h0); cudaFree(s->d_hs); cudaFree(s->d_kv); cudaFree(s->d_tw);
    cudaFree(s->d_stage); cudaFree(s->d_send); cudaFree(s->d_recv); cudaFree(s->d_disp); cudaFree(s->d_norm);
    cudaFree(s->d_minmax); cudaFree(s->d_ampl);
    cudaFreeHost(s->h_small);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    delete s;
    return WSO_OK;
}

int wso_slab_import_h0(wso_slab* s, const wso_h0_record* h0) {
    if (!s || !h0) return WSO_ERR_INVALID_ARG;
    const uint32_t n = s->n;
    return upload_local(s, [&](uint32_t m, uint32_t col) { return h0[(size_t)m * n + col]; });
}

int wso_slab_prepare_counter(wso_slab* s, uint64_t seed) {
    if (!s) return WSO_ERR_INVALID_ARG;
    const uint32_t n = s->n;
    const wso::H0Builder hb(s->params);
    std::vector<float> kv;
    wso::host_wave_numbers(n, s->params.tile_length, kv);
    return upload_local(s, [&](uint32_t m, uint32_t col) {
        float re, im;
        wso::counter_gauss(seed, (uint64_t)m * n + col, &re, &im);
        return hb.at(kv[col], kv[m], re, im);
    });
}

int wso_slab_prepare_counter_device(wso_slab* s, uint64_t seed) {
    if (!s) return WSO_ERR_INVALID_ARG;
    const uint32_t n = s->n;
    const wso::DerivedParams d = wso::derive_params(s->params);
    std::vector<float> kv;
    wso::host_wave_numbers(n, s->params.tile_length, kv);
    const float kmax = std::sqrt(kv[0] * kv[0] + kv[0] * kv[0]);
    const float jf = std::floor(std::sqrt(9.81f * kmax) / d.base_freq);
    if (!(d.base_freq > 0.0f) || !(jf >= 0.0f && jf < (float)wso::kMaxTable))
        return sfail(s, WSO_ERR_INVALID_ARG, "slab path needs a dispersion table of at most 1024 entries");
    wso::PrepareArgs a;
    a.h0 = s->d_h0;
    a.hs = s->d_hs;
    a.kv = s->d_kv;
    a.xi = nullptr;
    a.seed_mixed = wso::counter_seed_mix(seed);
    a.n = (int)n;
    a.j0 = (int)(s->rank * s->hl);
    a.wind_x = d.wind_x;
    a.wind_y = d.wind_y;
    a.omega0 = d.base_freq;
    a.phillips_const = s->params.phillips_const;
    a.damping = s->params.damping;
    a.inv_sqrt2 = 1.0f / std::sqrt(2.0f);
    const float Lw = d.wind_speed * d.wind_speed / 9.81f;
    a.Lw2 = Lw * Lw;
    SLAB_CUDA(s, cudaSetDevice(s->device));
    SLAB_CUDA(s, cudaMemcpyAsync(s->d_kv, kv.data(), sizeof(float) * n, cudaMemcpyHostToDevice, s->stream));
    SLAB_CUDA(s, wso::launch_prepare(a, (int)s->hl, s->stream));
    SLAB_CUDA(s, cudaStreamSynchronize(s->stream));
    s->td.h0 = s->d_h0;
    s->td.hs = s->d_hs;
    s->td.kv = s->d_kv;
    s->td.lambda = s->params.lambda;
    s->td.omega0 = d.base_freq;
    s->td.table_len = (int)jf + 1;
    s->td.use_pairs = 1;
    s->td.j0 = (int)(s->rank * s->hl);
    s->prepared = true;
    return WSO_OK;
}

int wso_counter_h0(const wso_params* p, uint64_t seed, uint32_t m0, uint32_t rows, wso_h0_record* out) {
    if (!p || !out) return WSO_ERR_INVALID_ARG;
    wso_params q = *p;
    wso::normalise_like_setters(q, nullptr);
    const uint32_t n = q.tile_size;
    const wso::H0Builder hb(q);
    std::vector<float> kv;
    wso::host_wave_numbers(n, q.tile_length, kv);
    for (uint32_t m = m0; m < m0 + rows && m < n; ++m)
        for (uint32_t c = 0; c < n; ++c) {
            float re, im;
            wso::counter_gauss(seed, (uint64_t)m * n + c, &re, &im);
            out[(size_t)(m - m0) * n + c] = hb.at(kv[c], kv[m], re, im);
        }
    return WSO_OK;
}

int wso_slab_set_lambda(wso_slab* s, float lambda) {
    if (!s || !std::isfinite(lambda)) return WSO_ERR_INVALID_ARG;
    s->params.lambda = lambda;
    return WSO_OK;
}

int wso_slab_set_stream(wso_slab* s, void* cuda_stream) {
    if (!s) return WSO_ERR_INVALID_ARG;
    SLAB_CUDA(s, cudaSetDevice(s->device));
    SLAB_CUDA(s, cudaStreamSynchronize(s->stream));
    s->stream = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->own_stream;
    return WSO_OK;
}

int wso_slab_force_pair(wso_slab* s, int on) {
    if (!s) return WSO_ERR_INVALID_ARG;
    s->force_pair = on != 0;
    return WSO_OK;
}

int wso_slab_buffers(wso_slab* s, void** send, void** recv, size_t* block_bytes, void** minmax) {
    if (!s) return WSO_ERR_INVALID_ARG;
    if (send) *send = s->d_send;
    if (recv) *recv = s->d_recv;
    if (block_bytes) *block_bytes = block_elems(s) * sizeof(float2);
    if (minmax) *minmax = s->d_minmax;
    return WSO_OK;
}

int wso_slab_ipc_handle(wso_slab* s, void* handle64) {
    if (!s || !handle64) return WSO_ERR_INVALID_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    SLAB_CUDA(s, cudaSetDevice(s->device));
    SLAB_CUDA(s, cudaIpcGetMemHandle(&h, s->d_recv));
    std::memcpy(handle64, &h, 64);
    return WSO_OK;
}

int wso_slab_open_peer(wso_slab* s, uint32_t peer, const void* handle64) {
    if (!s || peer >= s->world) return WSO_ERR_INVALID_ARG;
    SLAB_CUDA(s, cudaSetDevice(s->device));
    if (peer == s->rank) {
        s->peers[peer] = s->d_recv;
        return WSO_OK;
    }
    if (!handle64) return WSO_ERR_INVALID_ARG;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle64, 64);
    void* p = nullptr;
    SLAB_CUDA(s, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    s->opened.push_back(p);
    s->peers[peer] = static_cast<float2*>(p);
    return WSO_OK;
}

int wso_slab_set_fused(wso_slab* s, int on) {
    if (!s) return WSO_ERR_INVALID_ARG;
    if (on)
        for (uint32_t d = 0; d < s->world; ++d)
            if (!s->peers[d]) return sfail(s, WSO_ERR_INVALID_ARG, "fused mode needs every peer buffer (wso_slab_open_peer)");
    s->fused = on != 0;
    return WSO_OK;
}

static int slab_phase(wso_slab* s, int phase, int field0 = 0, int nfields = 0) {
    if (!s) return WSO_ERR_INVALID_ARG;
    if (!s->prepared) return sfail(s, WSO_ERR_NOT_PREPARED, "slab: no spectrum (wso_slab_prepare_counter / wso_slab_import_h0)");
    SLAB_CUDA(s, cudaSetDevice(s->device));
    SlabArgs a;
    fill_args(s, a, s->last_t);
    const int nf = wso::slab_fields_per_group(s->logn);
    if (nfields > 0 && (field0 % nf != 0 || nfields % nf != 0)) return sfail(s, WSO_ERR_INVALID_ARG, "slab: field range must cover whole field groups");
    a.slab_field0 = field0 / nf;
    cudaError_t e = wso::launch_slab_phase(s->logn, phase, a, s->force_pair, nfields, s->stream);
    if (e != cudaSuccess) return sfail_cuda(s, e, "slab kernel launch");
    return WSO_OK;
}

int wso_slab_pass1(wso_slab* s, float t) {
    if (!s) return WSO_ERR_INVALID_ARG;
    s->last_t = t;
    s->frame += 1;
    return slab_phase(s, 0);
}

// The exchange step: one transposing kernel carries K1's staged output to the owners of the row items - straight into
// their receive buffers over NVLink peer memory (wso_slab_set_fused), or into the blocks of the local send buffer that a
// collective library then moves.  A no-op for grids too small to stage (K1 has stored in the exchange layout itself).
int wso_slab_exchange(wso_slab* s) { return wso_slab_exchange_fields(s, 0, 4, nullptr); }

// Pipelined form: pass1 and the exchange field by field, so that the transfer of field f overlaps the transform of field
// f+1 - the caller enqueues wso_slab_pass1_fields(t, f, k) on the slab's stream and wso_slab_exchange_fields(f, k, xs) on
// a second stream xs that it has ordered behind the pass1 launch (event).  Fields must come as whole field groups
// (wso_slab_fields_per_group: 1 for the production sizes).
int wso_slab_pass1_fields(wso_slab* s, float t, int field0, int nfields) {
    if (!s || field0 < 0 || nfields < 1 || field0 + nfields > 4) return WSO_ERR_INVALID_ARG;
    if (field0 == 0) {
        s->last_t = t;
        s->frame += 1;
    }
    return slab_phase(s, 0, field0, nfields);
}

int wso_slab_fields_per_group(const wso_slab* s) { return s ? wso::slab_fields_per_group(s->logn) : 4; }

int wso_slab_exchange_fields(wso_slab* s, int field0, int nfields, void* cuda_stream) {
    if (!s || field0 < 0 || nfields < 1 || field0 + nfields > 4) return WSO_ERR_INVALID_ARG;
    if (!s->prepared) return sfail(s, WSO_ERR_NOT_PREPARED, "slab: no spectrum (wso_slab_prepare_counter / wso_slab_import_h0)");
    if (!staged(s)) return WSO_OK;
    SLAB_CUDA(s, cudaSetDevice(s->device));
    wso::XposeDst dst;
    exchange_dst(s, dst.p);
    int hl_log = 0;
    while ((1u << hl_log) < s->hl) ++hl_log;
    cudaStream_t st = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : s->stream;
    cudaError_t e = wso::launch_slab_exchange(s->d_stage, dst, (int)s->world, (int)s->rank, hl_log, s->logn - 1, field0, nfields, st);
    if (e != cudaSuccess) return sfail_cuda(s, e, "slab exchange kernel launch");
    return WSO_OK;
}
int wso_slab_heights(wso_slab* s) { return slab_phase(s, 1); }
int wso_slab_pass2(wso_slab* s) { return slab_phase(s, 2); }

int wso_slab_sync(wso_slab* s) {
    if (!s) return WSO_ERR_INVALID_ARG;
    SLAB_CUDA(s, cudaSetDevice(s->device));
    SLAB_CUDA(s, cudaStreamSynchronize(s->stream));
    return WSO_OK;
}

int wso_slab_read_heights(wso_slab* s, float* amplitude, float* min_height, float* max_height) {
    if (!s) return WSO_ERR_INVALID_ARG;
    SLAB_CUDA(s, cudaSetDevice(s->device));
    SLAB_CUDA(s, cudaMemcpyAsync(s->h_small, s->d_ampl, sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    SLAB_CUDA(s, cudaMemcpyAsync(s->h_small + 1, s->d_minmax, 2 * sizeof(float), cudaMemcpyDeviceToHost, s->stream));
    SLAB_CUDA(s, cudaStreamSynchronize(s->stream));
    if (amplitude) *amplitude = s->h_small[0];
    if (min_height) *min_height = s->h_small[1];
    if (max_height) *max_height = s->h_small[2];
    return WSO_OK;
}

int wso_slab_map_device(wso_slab* s, int which, void** dptr, uint32_t* rows) {
    if (!s || !dptr) return WSO_ERR_INVALID_ARG;
    if (which != WSO_MAP_DISPLACEMENT && which != WSO_MAP_NORMAL) return sfail(s, WSO_ERR_INVALID_ARG, "bad map id");
    *dptr = which == WSO_MAP_DISPLACEMENT ? s->d_disp : s->d_norm;
    if (rows) *rows = 2 * s->hl;
    return WSO_OK;
}

int wso_slab_copy_rows(wso_slab* s, int which, float* dst_host) {
    void* src = nullptr;
    uint32_t rows = 0;
    int rc = wso_slab_map_device(s, which, &src, &rows);
    if (rc != WSO_OK) return rc;
    if (!dst_host) return WSO_ERR_INVALID_ARG;
    SLAB_CUDA(s, cudaSetDevice(s->device));
    SLAB_CUDA(s, cudaMemcpyAsync(dst_host, src, sizeof(float4) * (size_t)rows * s->n, cudaMemcpyDeviceToHost, s->stream));
    SLAB_CUDA(s, cudaStreamSynchronize(s->stream));
    return WSO_OK;
}

int wso_slab_row_index(const wso_slab* s, uint32_t* rows) {
    if (!s || !rows) return WSO_ERR_INVALID_ARG;
    for (uint32_t ml = 0; ml < s->hl; ++ml) {
        const uint32_t mp = s->rank * s->hl + ml;
        rows[ml] = mp;
        rows[s->hl + ml] = mp == 0 ? s->n / 2 : s->n - mp;
    }
    return WSO_OK;
}

const char* wso_slab_last_error(const wso_slab* s) { return s ? s->err.c_str() : g_slab_create_error.c_str(); }

}  // extern "C"
