// wso_kernels2.cuh — the warp-per-line transform core (tile sizes N = 32*L, L = 16, 32, 64: 512^2, 1024^2, 2048^2).
//
// Same algorithm, data layout and results as wso_kernels.cuh (four packed real spectra, two real columns
// per complex transform, Hermitian half W[m'][f][j][half]); what changes is how a 1-D line transform is carried out:
//
//   * a line of N = 32*L points belongs to L threads that hold 32 values each IN REGISTERS;
//   * transform = radix-32 in registers -> twiddle -> ONE shared-memory transpose -> radix-L in registers (L = 64: radix
//     32 in registers plus one radix-2 butterfly between neighbouring lanes by warp shuffle): 2 shared-memory operations
//     per value instead of 6, no CTA-wide barrier inside a transform;
//   * the Hermitian mirror partner of everything a lane holds (row N-m of row m, column N-n of column n) lives in ONE
//     other lane of the same warp, so the packing of the real spectra (K1 front end), the two-for-one separation (K1 back
//     end) and the unfolding of the paired W layout (K2 front end) are lane-to-lane shuffles;
//   * K2 / K2h fetch their W lines (8*N contiguous bytes each) with 1-D bulk copies (cp.async.bulk, completion on an
//     mbarrier), double buffered per line group: the next line is in flight while the current one is transformed, and the
//     load latency never sits in a register dependency chain;
//   * K2 / K2h are persistent (a fixed number of CTAs walks the row items), K2 packs both output rows of a row item from
//     registers - 512-byte coalesced map stores.
//
// Replaces reference WSTessendorf::ComputeWaves (src/scene/WSTessendorf.cpp:284-441), i.e. the spectrum evolution
// (cpp:292-336), the seven fftwf_execute calls (cpp:338-378), sign/lambda/packing (cpp:385-437) and NormalizeHeights
// (cpp:443-455).  Written against a context type (wso_simt.cuh) so that tests/emu can step the same bodies on the CPU.
#pragma once

#include "wso_kernels.cuh"
#include "wso_simt.cuh"

namespace wso {
namespace v2 {

// ---------------------------------------------------------------------------------------------------------
// compile-time roots of unity
// ---------------------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
#define WSO_CX __host__ __device__ constexpr
#else
#define WSO_CX constexpr
#endif
constexpr double kPi = 3.14159265358979323846264338327950288;
WSO_CX double cx_cos_taylor(double x) {  // |x| <= pi
    double term = 1.0, sum = 1.0;
    for (int k = 1; k <= 16; ++k) {
        term *= -x * x / (double)((2 * k - 1) * (2 * k));
        sum += term;
    }
    return sum;
}
WSO_CX double cx_sin_taylor(double x) {
    double term = x, sum = x;
    for (int k = 1; k <= 16; ++k) {
        term *= -x * x / (double)((2 * k) * (2 * k + 1));
        sum += term;
    }
    return sum;
}
// cos / sin of 2*pi*k/n with the argument folded into [-pi, pi]
WSO_CX double root_cos(int k, int n) {
    k = ((k % n) + n) % n;
    if (4 * k == n || 4 * k == 3 * n) return 0.0;
    const int kk = (2 * k > n) ? k - n : k;
    return cx_cos_taylor(2.0 * kPi * (double)kk / (double)n);
}
WSO_CX double root_sin(int k, int n) {
    k = ((k % n) + n) % n;
    if (k == 0 || 2 * k == n) return 0.0;
    const int kk = (2 * k > n) ? k - n : k;
    return cx_sin_taylor(2.0 * kPi * (double)kk / (double)n);
}

// a * exp(+2*pi*i*E/M), E and M compile-time
template <int E, int M>
WSO_HD float2 mul_root(float2 a) {
    constexpr int e = ((E % M) + M) % M;
    if (e == 0) return a;
    if (4 * e == M) return make_float2(-a.y, a.x);
    if (2 * e == M) return make_float2(-a.x, -a.y);
    if (4 * e == 3 * M) return make_float2(a.y, -a.x);
    constexpr float wc = (float)root_cos(e, M), ws = (float)root_sin(e, M);
    return cmul(a, make_float2(wc, ws));
}

// 32-point backward DFT in registers, natural order in and out: radix 2 over two 16-point transforms
WSO_HD void dft32(float2* v) {
    float2 e[16], o[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        e[k] = v[2 * k];
        o[k] = v[2 * k + 1];
    }
    Dft<16>::run(e);
    Dft<16>::run(o);
    static_for<0, 16>([&](auto kc) {
        constexpr int K = decltype(kc)::value;
        const float2 t = mul_root<K, 32>(o[K]);
        v[K] = cadd(e[K], t);
        v[K + 16] = csub(e[K], t);
    });
}

// v[c] *= w^c for c = 1..31 given w^1, w^2, w^4, w^8, w^16 (table values: no error growth in the bases); every other
// power is a product of at most four table values.
WSO_HD void apply_twiddle_powers(float2* v, const float2 (&wb)[5]) {
    float2 p[32];
    p[1] = wb[0];
    p[2] = wb[1];
    p[4] = wb[2];
    p[8] = wb[3];
    p[16] = wb[4];
    static_for<1, 32>([&](auto cc) {
        constexpr int C = decltype(cc)::value;
        constexpr int low = C & (-C);  // lowest set bit
        if constexpr (low != C) p[C] = cmul(p[C - low], p[low]);
        v[C] = cmul(v[C], p[C]);
    });
}

// ---------------------------------------------------------------------------------------------------------
// geometry of a line of N = 32*L points held by L threads
// ---------------------------------------------------------------------------------------------------------
template <int LOGN>
struct Geo {
    static_assert(LOGN >= 9 && LOGN <= 11, "warp-per-line core: 512, 1024, 2048");
    static constexpr int N = 1 << LOGN;
    static constexpr int H = N / 2;
    static constexpr int L = N / 32;  // threads per line
    // Shared-memory exchange layout between the two register stages: element (c, b) - output index c of the first stage
    // of input residue b - at c*S + b.  First-stage writes are unit stride in b; second-stage reads are stride S across
    // lanes (S odd), or stride S with two interleaved residues per c (L = 64: S = 2 mod 16): no bank conflicts for
    // 64-bit accesses.
    static constexpr int S = (L == 64) ? 66 : L + 1;
    static constexpr int LINE0 = ((32 * S + 15) / 16) * 16;  // float2 per line buffer, before the per-kernel offset
};

// Input side: thread tl in [0, L) of a line group holds the points L*a + b, a = 0..31, of residue b = b_of(tl).  The
// mirror N - (L*a + b) = L*(31-a) + (L-b) belongs to residue L-b, which is held by lane partner_lane() of the SAME warp
// (L = 64: the residues are dealt to the two warps of a line in mirror-closed sets).  Residue 0 is its own mirror with
// the register index shifted by one (N - L*a = L*(32-a)); residue L/2 is its own mirror in place.
template <int L>
struct InMap;
template <>
struct InMap<32> {
    static WSO_HD int b_of(int tl) { return tl; }
    static WSO_HD int partner_lane(int lane, int) { return (32 - lane) & 31; }
};
template <>
struct InMap<16> {
    static WSO_HD int b_of(int tl) { return tl; }
    static WSO_HD int partner_lane(int lane, int) { return (lane & 16) | ((16 - (lane & 15)) & 15); }
};
template <>
struct InMap<64> {
    // warp 0 of the line: residues 1..16 and 48..63;  warp 1: residue 0 and 17..47
    static WSO_HD int b_of(int tl) {
        const int lane = tl & 31;
        return tl < 32 ? (lane < 16 ? lane + 1 : lane + 32) : (lane ? lane + 16 : 0);
    }
    static WSO_HD int partner_lane(int lane, int tl) { return tl < 32 ? 31 - lane : (32 - lane) & 31; }
};

// v[0..15] hold this thread's own values for a = 0..15 and B[a] what it computed for the mirror of point L*a + b.
// Afterwards v[0..31] is the thread's full column of 32 inputs.
template <class Ctx>
WSO_HD void mirror_exchange(Ctx& cx, float2* v, const float2* B, int partner, bool b0) {
    static_for<0, 16>([&](auto ac) {
        constexpr int A = decltype(ac)::value;
        v[31 - A] = cx.shfl(B[A], partner);
    });
    // residue 0: the mirror of point L*a is point L*(32-a); a = 0 pairs point 0 with point N/2 = L*16
    static_for<0, 15>([&](auto kc) {
        constexpr int K = 31 - decltype(kc)::value;  // 31 .. 17
        if (b0) v[K] = v[K - 1];
    });
    if (b0) v[16] = B[0];
}

// first register stage + twiddle + store into the exchange layout
template <int LOGN>
WSO_HD void stage1_store(float2* v, const float2 (&wb)[5], float2* line, int b) {
    constexpr int S = Geo<LOGN>::S;
    dft32(v);
    apply_twiddle_powers(v, wb);
#pragma unroll
    for (int c = 0; c < 32; ++c) line[c * S + b] = v[c];
}

// twiddle bases w_N^(b * 2^k), k = 0..4, from the table tw[k] = exp(+2*pi*i*k/N)
template <int LOGN>
WSO_HD void load_twiddle_bases(const float2* __restrict__ tw, int b, float2 (&wb)[5]) {
#pragma unroll
    for (int k = 0; k < 5; ++k) wb[k] = tw[(b << k) & (Geo<LOGN>::N - 1)];
}

// ---------------------------------------------------------------------------------------------------------
// K1: evolve + Hermitian packing + transform along m + two-for-one separation -> W
// ---------------------------------------------------------------------------------------------------------
// CTA = CP consecutive column pairs (one line group of L threads each), NF packed fields processed one after the other
// by the same threads: h~ is evaluated ONCE per wave-vector pair and kept in registers for all NF fields.
// The second register stage is dealt to the threads of the CTA so that CP adjacent lanes hold the same output rows of
// the CP adjacent column pairs (16*CP contiguous bytes of W per row), and so that the mirror row N-m' sits in a lane of
// the same warp.  Requires every item to have the sincos table and the pair-summed records (any Prepare()-built h0).
template <int LOGN, int CP, int NF>
struct Pass1W {
    using G = Geo<LOGN>;
    static constexpr int N = G::N, H = G::H, L = G::L, S = G::S;
    static constexpr int T = CP * L;
    static constexpr int NQ = 32 / CP;  // lanes of a warp per column pair in the second stage
    // line buffers 16/CP (64-bit) banks apart: the CP lines a second-stage half-warp reads never collide
    static constexpr int LINE = G::LINE0 + (CP > 1 ? 16 / CP : 0);
    static constexpr int SMEM_BYTES = (CP * LINE + kMaxTable) * (int)sizeof(float2) + N * (int)sizeof(float);
    static_assert(CP == 1 || CP == 2 || CP == 4 || CP == 8 || CP == 16, "CP must divide 16");
    static_assert(T % 32 == 0 && T <= 1024, "bad CTA size");
    static_assert(4 % NF == 0, "bad field grouping");
    static_assert(L != 64 || CP <= 8, "2048: at most 8 column pairs per CTA");
    using P1 = Pass1<LOGN, 1, 4, false, true>;  // general_points(): the per-point record path of the index-0 / N/2 lines

    template <int F>
    static WSO_HD void interior_all(const float* s0, const float* s1, const float* inv, const float* kz, float kxA,
                                    float kxB, float2* v, float2* B) {
        static_for<0, 16>([&](auto ac) {
            constexpr int A = decltype(ac)::value;
            pack_interior<F>(s0[A], kxA, kz[A], inv[A], s1[A], kxB, kz[A], inv[A], &v[A], &B[A]);
        });
    }

    static WSO_HD void general_item(const TileDev& td, const float2* table, float t, int f, int i, int jl, float2* a,
                                    float2* b) {
        Point pt[4];
        const int mask = P1::template general_points<true>(td, table, t, i, jl, pt);
        if (f == 0) pack_general<0>(pt, mask, a, b);
        else if (f == 1) pack_general<1>(pt, mask, a, b);
        else if (f == 2) pack_general<2>(pt, mask, a, b);
        else pack_general<3>(pt, mask, a, b);
    }

    // (Wa, Wb) of one output row from the transformed pair line: c1 = C[m'], c2 = C[N-m']
    static WSO_HD float4 separate(float2 c1, float2 c2) {
        const float2 wa = cscale(0.5f, cadd_conj(c1, c2));
        const float2 d = csub_conj(c1, c2);
        return make_float4(wa.x, wa.y, 0.5f * d.y, -0.5f * d.x);
    }

    // bx: column-pair group, by: field group, bz: item within the chunk
    template <class Ctx, class Args>
    static WSO_HD void run(Ctx& cx, float2* smem, int bx, int by, int bz, const Args& args) {
        const int tid = cx.tid;
        const BatchItem item = args.items[bz];
        const TileDev& td = args.td[bz];
        const float t = item.t;
        const int cp = tid / L, tl = tid % L;
        const int jl = bx * CP + cp, j = td.j0 + jl;
        const int b = InMap<L>::b_of(tl);
        const bool b0 = (b == 0);
        const int partner = InMap<L>::partner_lane(tid & 31, tl);
        float2* table = smem + CP * LINE;
        float* kvs = reinterpret_cast<float*>(table + kMaxTable);
        float2* line = smem + cp * LINE;

        // ---- pair-summed records of this thread's 16 row pairs i = L*a + b (first half requested right away)
        const float4* rec = td.hs + ((size_t)jl * H + b) * 2;
        float4 q0[8], q1[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            q0[a] = ld_ro(rec + (size_t)a * L * 2);
            q1[a] = ld_ro(rec + (size_t)a * L * 2 + 1);
        }
        // ---- per-frame (cos,sin)(omega_j * t) table and the wave numbers -> shared memory
        for (int jj = tid; jj < td.table_len; jj += T) {
            float s, c;
            sincos_acc(rmul(rmul((float)jj, td.omega0), t), &s, &c);
            table[jj] = make_float2(c, s);
        }
        for (int i = tid; i < N; i += T) kvs[i] = td.kv[i];
        float2 wb[5];
        load_twiddle_bases<LOGN>(args.tw, b, wb);
        cx.cta_sync();

        // ---- evolve: s = (h~(k) + h~(-k)) / 2 at the two wave vectors of every row pair, 1/|k|, kz
        float s0[16], s1[16], inv[16], kz[16];
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            s0[a] = 0.5f * eval_height<true>(q0[a], table, t);
            s1[a] = 0.5f * eval_height<true>(q1[a], table, t);
            inv[a] = q0[a].z;
            kz[a] = kvs[L * a + b];
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            q0[a] = ld_ro(rec + (size_t)(a + 8) * L * 2);
            q1[a] = ld_ro(rec + (size_t)(a + 8) * L * 2 + 1);
        }
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            s0[a + 8] = 0.5f * eval_height<true>(q0[a], table, t);
            s1[a + 8] = 0.5f * eval_height<true>(q1[a], table, t);
            inv[a + 8] = q0[a].z;
            kz[a + 8] = kvs[L * (a + 8) + b];
        }
        const float kxA = kvs[j & (N - 1)], kxB = kvs[(N - j) & (N - 1)];

        // second-stage role of this thread (see the struct comment)
        const int lane = tid & 31, w = tid >> 5;
        const int cp2 = lane % CP, q = lane / CP;
        const float2* line2 = smem + cp2 * LINE;
        float4* Wit = reinterpret_cast<float4*>(args.W + (size_t)bz * ((size_t)H * 4 * N));
        const int jcol = bx * CP + cp2;
        bool waited = false;

#pragma unroll 1
        for (int fi = 0; fi < NF; ++fi) {
            const int f = by * NF + fi;
            float2 v[32], B[16];
            if (f == 0) interior_all<0>(s0, s1, inv, kz, kxA, kxB, v, B);
            else if (f == 1) interior_all<1>(s0, s1, inv, kz, kxA, kxB, v, B);
            else if (f == 2) interior_all<2>(s0, s1, inv, kz, kxA, kxB, v, B);
            else interior_all<3>(s0, s1, inv, kz, kxA, kxB, v, B);
            // the row pair (0, N/2) keeps its wave number along n under the mirror: per-point records
            if (b0 && j != 0) general_item(td, table, t, f, 0, jl, &v[0], &B[0]);
            mirror_exchange(cx, v, B, partner, b0);
            // column pair (0, N/2): every row pair through the per-point records, staged in natural order in the line
            // buffer (one line group of the whole grid; the CTA-uniform test keeps the barriers convergent)
            if (bx == 0 && td.j0 == 0) {
                if (j == 0) {
                    for (int a = 0; a < 16; ++a) {
                        const int i = L * a + b;
                        float2 xa, xb;
                        general_item(td, table, t, f, i, jl, &xa, &xb);
                        line[i] = xa;
                        line[i == 0 ? H : N - i] = xb;
                    }
                }
                cx.cta_sync();
                if (j == 0) {
#pragma unroll
                    for (int a = 0; a < 32; ++a) v[a] = line[L * a + b];
                }
                cx.cta_sync();
            }
            stage1_store<LOGN>(v, wb, line, b);
            cx.cta_sync();

            // everything above touched only per-tile constants and shared memory: it may overlap the tail of the
            // previous kernel in the stream, which still reads W and the slot's min/max
            if (!waited) {
                cx.pdl_wait();
                cx.pdl_release();
                waited = true;
                if (bx == 0 && by == 0 && tid == 0) {
                    args.minmax[2 * item.slot + 0] = kInitMin;
                    args.minmax[2 * item.slot + 1] = kInitMax;
                }
            }

            if constexpr (L == 32) {
                constexpr int R = NQ / 2;
                int c = q < R ? R * w + 1 + q : 32 - (R * w + 1 + (NQ - 1 - q));
                if (c == 16 && q >= R) c = 0;
                const bool self = (c == 0 || c == 16);
                const int plane = self ? lane : CP * (NQ - 1 - q) + cp2;
                const float2* src = line2 + c * S;
#pragma unroll
                for (int bb = 0; bb < 32; ++bb) v[bb] = src[bb];
                dft32(v);  // v[d] = C[c + 32 d]
                static_for<0, 16>([&](auto dc) {
                    constexpr int D = decltype(dc)::value;
                    // C[N - m'], m' = c + 32 D: register 31-D of the lane that owns 32-c (c = 0: own register 32-D)
                    const float2 send = (c == 0) ? v[D == 0 ? 16 : 32 - D] : v[31 - D];
                    const float2 c2 = cx.shfl(send, plane);
                    float4 o = separate(v[D], c2);
                    if (D == 0 && c == 0) o = make_float4(v[0].x, v[16].x, v[0].y, v[16].y);  // rows 0 and N/2 are real
                    Wit[((size_t)(c + 32 * D) * 4 + f) * H + jcol] = o;
                });
            } else if constexpr (L == 16) {
                const int p = NQ * w + q;  // mirror-closed pair of first-stage outputs {p, 32-p} ({0, 16} for p = 0)
                const int c1 = p, c2i = p ? 32 - p : 16;
                const float2* srcA = line2 + c1 * S;
                const float2* srcB = line2 + c2i * S;
#pragma unroll
                for (int bb = 0; bb < 16; ++bb) {
                    v[bb] = srcA[bb];
                    v[16 + bb] = srcB[bb];
                }
                Dft<16>::run(v);       // v[d]      = C[c1 + 32 d]
                Dft<16>::run(v + 16);  // v[16 + d] = C[c2 + 32 d]
                static_for<0, 8>([&](auto dc) {
                    constexpr int D = decltype(dc)::value;
                    const float2 mA = p ? v[16 + 15 - D] : v[(16 - D) & 15];
                    const float2 mB = p ? v[15 - D] : v[16 + 15 - D];
                    float4 oA = separate(v[D], mA);
                    if (D == 0 && p == 0) oA = make_float4(v[0].x, v[8].x, v[0].y, v[8].y);
                    const float4 oB = separate(v[16 + D], mB);
                    Wit[((size_t)(c1 + 32 * D) * 4 + f) * H + jcol] = oA;
                    Wit[((size_t)(c2i + 32 * D) * 4 + f) * H + jcol] = oB;
                });
            } else {
                // L == 64: q = 2*ci + h; the two halves h of first-stage output c sit in neighbouring q
                constexpr int R = NQ / 4;
                const int ci = q >> 1, h = q & 1;
                int c = ci < R ? R * w + 1 + ci : 32 - (R * w + 1 + (2 * R - 1 - ci));
                if (c == 16 && ci >= R) c = 0;
                const bool self = (c == 0 || c == 16);
                const int plane = self ? (lane ^ CP) : CP * (NQ - 1 - q) + cp2;
                const float2* src = line2 + c * S + h;
#pragma unroll
                for (int bb = 0; bb < 32; ++bb) v[bb] = src[2 * bb];
                dft32(v);  // partial transform over the residues of parity h
                const float sgn = h ? -1.0f : 1.0f;
                static_for<0, 32>([&](auto dc) {
                    constexpr int D = decltype(dc)::value;
                    const float2 other = cx.shfl_xor(v[D], CP);
                    const float2 p0 = h ? other : v[D], p1 = h ? v[D] : other;
                    const float2 tw = mul_root<D, 64>(p1);
                    v[D] = cadd(p0, cscale(sgn, tw));  // C[c + 32 D + 1024 h]
                });
                static_for<0, 32>([&](auto dc) {
                    constexpr int D = decltype(dc)::value;
                    // C[N - m'], m' = c + 32 D (< N/2, held by h = 0): the h = 1 thread of 32-c, register 31-D
                    // (c = 0: register (32-D) & 31, where D = 0 fetches C[N/2])
                    const float2 send = (c == 0) ? v[(32 - D) & 31] : v[31 - D];
                    const float2 c2 = cx.shfl(send, plane);
                    float4 o = separate(v[D], c2);
                    if (D == 0 && c == 0) o = make_float4(v[0].x, c2.x, v[0].y, c2.y);
                    if (h == 0) Wit[((size_t)(c + 32 * D) * 4 + f) * H + jcol] = o;
                });
            }
            if (fi + 1 < NF) cx.cta_sync();  // the line buffers are rewritten by the next field
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// K2 (maps) and K2h (height extrema): persistent, one line group (L threads) per row item
// ---------------------------------------------------------------------------------------------------------
// MODE 0: displacement map (packed fields 1 then 0), 1: normal map (fields 3 then 2), 2: K2h (field 0, min/max only)
template <int LOGN, int GPC>
struct Pass2W {
    using G = Geo<LOGN>;
    static constexpr int N = G::N, H = G::H, L = G::L, S = G::S;
    static constexpr int LINE = G::LINE0;
    static constexpr int T = GPC * L;
    static constexpr int LINE_BYTES = N * (int)sizeof(float2);
    static constexpr int SMEM_BYTES = GPC * 2 * LINE * (int)sizeof(float2) + GPC * 2 * (int)sizeof(uint64_t);
    static_assert(T % 32 == 0 && T <= 1024, "bad CTA size");
    static_assert(LINE >= N, "a raw line must fit the exchange buffer");

    template <class Ctx>
    static WSO_HD void group_sync(Ctx& cx, int g) {
        if constexpr (L <= 32) cx.syncwarp();
        else cx.bar(1 + g, L);
    }

    // column of register r of this thread after the second stage
    template <int R>
    static WSO_HD int column_of(int tl) {
        if constexpr (L == 32) return tl + 32 * R;
        else if constexpr (L == 16) return tl + 16 * (R >> 4) + 32 * (R & 15);
        else return 16 * (tl >> 5) + ((tl & 31) >> 1) + 32 * R + 1024 * (tl & 1);
    }

    // raw line (paired layout [j][half]) in buf -> transformed line in v[] (register r = column_of<r>)
    template <class Ctx>
    static WSO_HD void transform_line(Ctx& cx, float2* buf, int g, int tl, int b, bool b0, int partner,
                                      const float2 (&wb)[5], float2* v) {
        const float4* raw = reinterpret_cast<const float4*>(buf);
        float2 B[16];
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            const float4 qq = raw[L * a + b];  // columns n = j and N - j of pair j = L*a + b
            v[a] = make_float2(qq.x, qq.y);
            B[a] = make_float2(qq.z, qq.w);
        }
        group_sync(cx, g);  // the raw line is dead: the exchange layout goes on top of it
        mirror_exchange(cx, v, B, partner, b0);
        stage1_store<LOGN>(v, wb, buf, b);
        group_sync(cx, g);
        if constexpr (L == 32) {
            const float2* src = buf + tl * S;
#pragma unroll
            for (int bb = 0; bb < 32; ++bb) v[bb] = src[bb];
        } else if constexpr (L == 16) {
            const float2* srcA = buf + tl * S;
            const float2* srcB = buf + (tl + 16) * S;
#pragma unroll
            for (int bb = 0; bb < 16; ++bb) {
                v[bb] = srcA[bb];
                v[16 + bb] = srcB[bb];
            }
        } else {
            const int c = 16 * (tl >> 5) + ((tl & 31) >> 1), h = tl & 1;
            const float2* src = buf + c * S + h;
#pragma unroll
            for (int bb = 0; bb < 32; ++bb) v[bb] = src[2 * bb];
        }
        // every thread of the group has its inputs: the buffer may be refilled by the async proxy (the fence orders this
        // thread's generic-proxy accesses to the buffer before the bulk copy issued behind the barrier)
        cx.fence_async_smem();
        group_sync(cx, g);
    }
    template <class Ctx>
    static WSO_HD void finish_line(Ctx& cx, int tl, float2* v) {
        if constexpr (L == 32) {
            dft32(v);
        } else if constexpr (L == 16) {
            Dft<16>::run(v);
            Dft<16>::run(v + 16);
        } else {
            const int h = tl & 1;
            dft32(v);
            const float sgn = h ? -1.0f : 1.0f;
            static_for<0, 32>([&](auto dc) {
                constexpr int D = decltype(dc)::value;
                const float2 other = cx.shfl_xor(v[D], 1);
                const float2 p0 = h ? other : v[D], p1 = h ? v[D] : other;
                v[D] = cadd(p0, cscale(sgn, mul_root<D, 64>(p1)));
            });
        }
    }

    // nbx: CTAs walking the row items (gridDim.x); n_items: tile-frames of this launch
    template <int MODE, class Ctx, class Args>
    static WSO_HD void run(Ctx& cx, float2* smem, int bx, int nbx, int n_items, const Args& args) {
        constexpr int LPU = (MODE == 2) ? 1 : 2;  // lines per unit (row item)
        const int tid = cx.tid;
        const int g = tid / L, tl = tid % L;
        const int b = InMap<L>::b_of(tl);
        const bool b0 = (b == 0);
        const int partner = InMap<L>::partner_lane(tid & 31, tl);
        float2* bufs = smem + (size_t)g * 2 * LINE;
        uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + (size_t)GPC * 2 * LINE) + g * 2;
        float2 wb[5];
        load_twiddle_bases<LOGN>(args.tw, b, wb);

        const int U = n_items * H;
        const int GT = nbx * GPC;
        const int gg = bx * GPC + g;
        const int nlines = ((U + GT - 1) / GT) * LPU;

        if (tl == 0) {
            cx.mbar_init(&mbar[0], 1);
            cx.mbar_init(&mbar[1], 1);
            cx.mbar_init_fence();
        }
        cx.cta_sync();
        // K2h consumes what K1 (its predecessor in the stream) wrote: wait first.  K2 is released by K2h only after
        // K2h's own wait, i.e. K1 is complete when K2 starts: K2 transforms W right away and waits (for K2h's min/max)
        // only before its first pack.
        if constexpr (MODE == 2) {
            cx.pdl_wait();
            cx.pdl_release();
        }
        bool waited = (MODE == 2);

        auto field_of = [&](int li) { return MODE == 2 ? 0 : (MODE == 0 ? (li == 0 ? 1 : 0) : (li == 0 ? 3 : 2)); };
        auto issue = [&](int k) {
            const int u = gg + (k / LPU) * GT;
            if (k < nlines && u < U && tl == 0) {
                const int it = u / H, mp = u - it * H;
                const float2* src = args.W + (size_t)it * ((size_t)H * 4 * N) + ((size_t)mp * 4 + field_of(k % LPU)) * N;
                cx.bulk_g2s(bufs + (k & 1) * LINE, src, (unsigned)LINE_BYTES, &mbar[k & 1]);
            }
        };
        issue(0);
        issue(1);

        float2 keep[32];  // first line of the unit (MODE 0 uses .y and, for rows 0 / N/2, nothing else)
        float run_mn = kInitMin, run_mx = kInitMax;
        int run_item = -1;
        auto commit_minmax = [&]() {
            // fold over the warp, one atomic pair per warp
            float mn = run_mn, mx = run_mx;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float a = cx.shfl_xor(mn, o), c = cx.shfl_xor(mx, o);
                mn = a < mn ? a : mn;
                mx = c > mx ? c : mx;
            }
            if ((tid & 31) == 0) cx.atomic_minmax(args.minmax + 2 * args.items[run_item].slot, mn, mx);
        };

#pragma unroll 1
        for (int k = 0; k < nlines; ++k) {
            const int u = gg + (k / LPU) * GT;
            if (u >= U) break;  // uniform over the line group and, for L = 16, over the warp (U and GT are even)
            const int li = k % LPU;
            const int it = u / H, mp = u - it * H;
            const bool special = (mp == 0);  // rows 0 and N/2 ride one complex line: separated through shared memory
            // L = 16: two row items share a warp; the neighbour of the special one follows its barriers
            const bool sync_special = (L == 16) ? ((mp >> 1) == 0) : special;
            float2* buf = bufs + (k & 1) * LINE;
            float2 v[32];
            cx.mbar_wait(&mbar[k & 1], (unsigned)((k >> 1) & 1));
            transform_line(cx, buf, g, tl, b, b0, partner, wb, v);
            if (!sync_special) issue(k + 2);
            finish_line(cx, tl, v);
            if (special) {
                // natural order back into the line's own buffer; refills wait until the unit is done
                static_for<0, 32>([&](auto rc) {
                    constexpr int R = decltype(rc)::value;
                    buf[column_of<R>(tl)] = v[R];
                });
            }
            if constexpr (MODE == 2) {
                if (it != run_item) {
                    if (run_item >= 0) commit_minmax();
                    run_item = it;
                    run_mn = kInitMin;
                    run_mx = kInitMax;
                }
                if (!special) {
                    const float s = ((mp + column_of<0>(tl)) & 1) ? -1.0f : 1.0f;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float h = rmul(v[r].x, s);
                        run_mn = h < run_mn ? h : run_mn;
                        run_mx = h > run_mx ? h : run_mx;
                    }
                }
                if (sync_special) {
                    group_sync(cx, g);
                    if (special) {
                        for (int c = tl; c < N; c += L) {
                            const float s = (c & 1) ? -1.0f : 1.0f;
                            const float2 a = buf[c], m = buf[(N - c) & (N - 1)];
                            const float hA = rmul(0.5f * (a.x + m.x), s), hB = rmul(0.5f * (a.y + m.y), s);
                            run_mn = hA < run_mn ? hA : run_mn;
                            run_mx = hA > run_mx ? hA : run_mx;
                            run_mn = hB < run_mn ? hB : run_mn;
                            run_mx = hB > run_mx ? hB : run_mx;
                        }
                    }
                    cx.fence_async_smem();
                    group_sync(cx, g);
                    issue(k + 2);
                }
            } else {
                if (li == 0) {
#pragma unroll
                    for (int r = 0; r < 32; ++r) keep[r] = v[r];
                    continue;
                }
                if (!waited) {
                    cx.pdl_wait();
                    cx.pdl_release();
                    waited = true;
                }
                const BatchItem item = args.items[it];
                const float lambda = args.td[it].lambda;
                const float amp = amplitude_of(args.minmax[2 * item.slot], args.minmax[2 * item.slot + 1]);
                const float inv_amp = rdiv(1.0f, amp);
                if (MODE == 0 && mp == 0 && tl == 0) args.amp_out[item.slot] = amp;
                float4* out = (MODE == 0 ? args.disp : args.norm) + (size_t)item.slot * ((size_t)N * N);
                float4* outA = out + (size_t)mp * N;
                float4* outB = out + (size_t)(mp == 0 ? H : N - mp) * N;
                if (!special) {
                    // reference: WSTessendorf.cpp:385-437 (sign, lambda, packing) and :443-455 (normalisation)
                    const float s = ((mp + column_of<0>(tl)) & 1) ? -1.0f : 1.0f;
                    const float sl = rmul(s, lambda);
                    static_for<0, 32>([&](auto rc) {
                        constexpr int R = decltype(rc)::value;
                        const int c = column_of<R>(tl);
                        const int cm = (N - c) & (N - 1);  // row N-m' is the conjugate mirror of row m'
                        if (MODE == 0) {
                            const float y = rmul(rmul(v[R].x, s), inv_amp);
                            const float x = rmul(sl, v[R].y), z = rmul(sl, keep[R].y);
                            st_stream(&outA[c], make_float4(x, y, z, 1.0f));
                            st_stream(&outB[cm], make_float4(-x, y, -z, 1.0f));
                        } else {
                            const float4 ta = make_float4(s * v[R].y, s * keep[R].y, s * v[R].x, s * keep[R].x);
                            st_stream(&outA[c], ta);
                            st_stream(&outB[cm], make_float4(-ta.x, -ta.y, ta.z, ta.w));
                        }
                    });
                }
                if (sync_special) {
                    group_sync(cx, g);
                    const float2* l0 = buf;                             // field 0 / 2 (this line)
                    const float2* l1 = bufs + ((k - 1) & 1) * LINE;     // field 1 / 3 (the unit's first line)
                    for (int c = tl; special && c < N; c += L) {
                        const float s = (c & 1) ? -1.0f : 1.0f;
                        const float sl = rmul(s, lambda);
                        const int cm = (N - c) & (N - 1);
                        const float2 p0 = l0[c], p1 = l1[c], m0 = l0[cm], m1 = l1[cm];
                        const float2 a0 = make_float2(0.5f * (p0.x + m0.x), 0.5f * (p0.y - m0.y));
                        const float2 a1 = make_float2(0.5f * (p1.x + m1.x), 0.5f * (p1.y - m1.y));
                        const float2 b0v = make_float2(0.5f * (p0.y + m0.y), -0.5f * (p0.x - m0.x));
                        const float2 b1v = make_float2(0.5f * (p1.y + m1.y), -0.5f * (p1.x - m1.x));
                        if (MODE == 0) {
                            st_stream(&outA[c], make_float4(rmul(sl, a0.y), rmul(rmul(a0.x, s), inv_amp), rmul(sl, a1.y), 1.0f));
                            st_stream(&outB[c], make_float4(rmul(sl, b0v.y), rmul(rmul(b0v.x, s), inv_amp), rmul(sl, b1v.y), 1.0f));
                        } else {
                            st_stream(&outA[c], make_float4(s * a0.y, s * a1.y, s * a0.x, s * a1.x));
                            st_stream(&outB[c], make_float4(s * b0v.y, s * b1v.y, s * b0v.x, s * b1v.x));
                        }
                    }
                    cx.fence_async_smem();
                    group_sync(cx, g);
                    issue(k + 1);
                    issue(k + 2);
                }
            }
        }
        if constexpr (MODE == 2) {
            if (run_item >= 0) commit_minmax();
        }
    }
};

}  // namespace v2
}  // namespace wso
