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
    // residue 0 is its own partner: the mirror of point L*a is point L*(32-a), and a = 0 pairs point 0 with point
    // N/2 = L*16 - so that lane feeds itself B rotated by one and the receive side needs no special case
    static_for<0, 16>([&](auto ac) {
        constexpr int A = decltype(ac)::value;
        const float2 own = B[A], rot = B[(A + 1) & 15];
        const float2 send = make_float2(b0 ? rot.x : own.x, b0 ? rot.y : own.y);
        v[31 - A] = cx.shfl(send, partner);
    });
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
// CTA = CP consecutive column pairs x NF packed fields; every (column pair, field) line has its own group of L threads,
// so each thread carries exactly one line of 32 values through the two register stages.
//   * h~ is evaluated ONCE per wave-vector pair: the NF line groups of a column pair split its 16 row-pair records
//     between them and publish (s0, s1, 1/|k|, kz) through shared memory; every group then packs its own field from it.
//   * The second register stage is dealt to the CP*L threads of a FIELD group so that CP adjacent lanes hold the same
//     output rows of the CP adjacent column pairs (16*CP contiguous bytes of W per row) and the mirror row N-m' sits in
//     a lane of the same warp; the field groups synchronise among themselves only (named barrier per field).
// Requires every item to have the sincos table and the pair-summed records (any Prepare()-built h0).
template <int LOGN, int CP, int NF>
struct Pass1W {
    using G = Geo<LOGN>;
    static constexpr int N = G::N, H = G::H, L = G::L, S = G::S;
    static constexpr int GT = CP * L;   // threads of one field group
    static constexpr int T = NF * GT;
    static constexpr int NQ = 32 / CP;  // lanes of a warp per column pair in the second stage
    static constexpr int RPG = 16 / NF; // row-pair records evaluated per line group
    // line buffers 16/CP (64-bit) banks apart: the CP lines a second-stage half-warp reads never collide
    static constexpr int LINE = G::LINE0 + (CP > 1 ? 16 / CP : 0);
    static constexpr int STATE_F4 = CP * 16 * L;  // float4 (s0, s1, 1/|k|, kz) per row pair and thread of a column pair
    static constexpr int SMEM_BYTES = (NF * CP * LINE + kMaxTable + CP * 4) * (int)sizeof(float2) +
                                      STATE_F4 * (int)sizeof(float4) + N * (int)sizeof(float);
    static_assert(CP == 1 || CP == 2 || CP == 4 || CP == 8 || CP == 16, "CP must divide 16");
    static_assert(NF == 1 || NF == 2 || NF == 4, "bad field grouping");
    static_assert(GT % 32 == 0 && T <= 1024, "bad CTA size");
    static_assert(L != 64 || CP <= 8, "2048: at most 8 column pairs per CTA");
    using P1 = Pass1<LOGN, 1, 4, false, true>;  // general_points(): the per-point record path of the index-0 / N/2 lines

    // this thread's 16 row pairs of packed field F from the published evolve state
    template <int F>
    static WSO_HD void pack_all(const float4* state, int tl, float kxA, float kxB, float2* v, float2* B) {
#pragma unroll
        for (int a = 0; a < 16; ++a) {
            const float4 st = state[a * L + tl];  // (s0, s1, 1/|k|, kz)
            pack_interior<F>(st.x, kxA, st.w, st.z, st.y, kxB, st.w, st.z, &v[a], &B[a]);
        }
    }
    // the row pair (0, N/2) of a column pair j >= 1: rows keep their wave number along n under the mirror, so the four
    // wave vectors (rows 0, N/2 x columns j, N-j) are packed one by one; g = (h~, 1/|k|) of each, published by the first
    // line group of the column pair
    template <int F>
    static WSO_HD void pack_row0(const float2* g, float kxA, float kxB, float kz0, float kzH, float2* a, float2* b) {
        Point pt[4];
        pt[0] = Point{g[0].x, kxA, kz0, rmul(kxA, g[0].y), rmul(kz0, g[0].y)};
        pt[1] = Point{g[1].x, kxB, kz0, rmul(kxB, g[1].y), rmul(kz0, g[1].y)};
        pt[2] = Point{g[2].x, kxA, kzH, rmul(kxA, g[2].y), rmul(kzH, g[2].y)};
        pt[3] = Point{g[3].x, kxB, kzH, rmul(kxB, g[3].y), rmul(kzH, g[3].y)};
        pack_general<F>(pt, 1, a, b);
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
        const int fg = tid / GT, r = tid - fg * GT;  // field group, thread within it
        const int cp = r / L, tl = r % L;
        const int f = by * NF + fg;
        const int jl = bx * CP + cp, j = td.j0 + jl;
        const int b = InMap<L>::b_of(tl);
        const bool b0 = (b == 0);
        const int partner = InMap<L>::partner_lane(tid & 31, tl);
        float2* lines = smem + fg * CP * LINE;  // the CP line buffers of this field group
        float2* line = lines + cp * LINE;
        float4* state = reinterpret_cast<float4*>(smem + NF * CP * LINE) + cp * 16 * L;
        float2* table = reinterpret_cast<float2*>(reinterpret_cast<float4*>(smem + NF * CP * LINE) + STATE_F4);
        float2* grow0 = table + kMaxTable + cp * 4;  // (h~, 1/|k|) of the four wave vectors of the row pair (0, N/2)
        float* kvs = reinterpret_cast<float*>(table + kMaxTable + CP * 4);

        // ---- pair-summed records of this group's share of the 16 row pairs i = L*a + b of the column pair
        float4 q0[RPG], q1[RPG];
#pragma unroll
        for (int k = 0; k < RPG; ++k) {
            const int a = fg * RPG + k;
            q0[k] = ld_ro(td.hs + hs_index(jl, a * L + b, 0, H));
            q1[k] = ld_ro(td.hs + hs_index(jl, a * L + b, 1, H));
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

        // ---- evolve: s = (h~(k) + h~(-k)) / 2 at the two wave vectors of every row pair
#pragma unroll
        for (int k = 0; k < RPG; ++k) {
            const int a = fg * RPG + k;
            const float s0 = 0.5f * eval_height<true>(q0[k], table, t);
            const float s1 = 0.5f * eval_height<true>(q1[k], table, t);
            state[a * L + tl] = make_float4(s0, s1, q0[k].z, kvs[L * a + b]);
        }
        // the four wave vectors of the row pair (0, N/2): one thread each of the column pair's first line group
        if (fg == 0 && tl < 4) {
            const float4 g = ld_ro(td.h0 + (size_t)jl * 2 * N + (tl & 1) * N + (tl >> 1) * H);
            grow0[tl] = make_float2(eval_height<true>(g, table, t), g.z);
        }
        const float kxA = kvs[j & (N - 1)], kxB = kvs[(N - j) & (N - 1)];
        float2 v[32], B[16];
        cx.cta_sync();

        // ---- this group's packed field: Z = even(R) - odd(V) under the mirror
        if (f == 0) pack_all<0>(state, tl, kxA, kxB, v, B);
        else if (f == 1) pack_all<1>(state, tl, kxA, kxB, v, B);
        else if (f == 2) pack_all<2>(state, tl, kxA, kxB, v, B);
        else pack_all<3>(state, tl, kxA, kxB, v, B);
        if (b0 && j != 0) {
            const float kz0 = kvs[0], kzH = kvs[H];
            if (f == 0) pack_row0<0>(grow0, kxA, kxB, kz0, kzH, &v[0], &B[0]);
            else if (f == 1) pack_row0<1>(grow0, kxA, kxB, kz0, kzH, &v[0], &B[0]);
            else if (f == 2) pack_row0<2>(grow0, kxA, kxB, kz0, kzH, &v[0], &B[0]);
            else pack_row0<3>(grow0, kxA, kxB, kz0, kzH, &v[0], &B[0]);
        }
        mirror_exchange(cx, v, B, partner, b0);
        // column pair (0, N/2): every row pair through the per-point records, staged in natural order in the line
        // buffer (one column pair of the whole grid; the CTA-uniform test keeps the barriers convergent)
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
        // the second stage reads the CP lines of this field group only
        if constexpr (NF == 1) cx.cta_sync();
        else cx.bar(1 + fg, GT);

        // everything above touched only per-tile constants and shared memory: it may overlap the tail of the previous
        // kernel in the stream, which still reads W and the slot's min/max
        cx.pdl_wait();
        cx.pdl_release();
        if (bx == 0 && by == 0 && tid == 0) {
            args.minmax[2 * item.slot + 0] = kInitMin;
            args.minmax[2 * item.slot + 1] = kInitMax;
        }

        // second-stage role of this thread (see the struct comment)
        const int lane = tid & 31, w = r >> 5;
        const int cp2 = lane % CP, q = lane / CP;
        const float2* line2 = lines + cp2 * LINE;
        float4* Wit = reinterpret_cast<float4*>(args.W + (size_t)bz * ((size_t)H * 4 * N));
        const int jcol = bx * CP + cp2;

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
    }
};

// ---------------------------------------------------------------------------------------------------------
// K2 (maps) and K2h (height extrema): persistent, one line group (L threads) per row item
// ---------------------------------------------------------------------------------------------------------
// MODE 0: displacement map (packed fields 1 then 0), 1: normal map (fields 3 then 2), 2: K2h (field 0, min/max only)
// NBUF: line buffers per group - 2: the next line is in flight while the current one is transformed (K2: several row
// items per group); 1: the refill is issued as soon as the second stage has its inputs (K2h: a launch has about one line
// per resident group, and the smaller footprint buys more resident groups).
template <int LOGN, int GPC, int NBUF = 2>
struct Pass2W {
    using G = Geo<LOGN>;
    static constexpr int N = G::N, H = G::H, L = G::L, S = G::S;
    static constexpr int LINE = G::LINE0;
    static constexpr int T = GPC * L;
    static constexpr int LINE_BYTES = N * (int)sizeof(float2);
    static constexpr int SMEM_BYTES = GPC * NBUF * LINE * (int)sizeof(float2) + GPC * 2 * (int)sizeof(uint64_t);
    static_assert(NBUF == 1 || NBUF == 2, "one or two line buffers per group");
    static_assert(T % 32 == 0 && T <= 1024, "bad CTA size");
    static_assert(LINE >= N, "a raw line must fit the exchange buffer");
    static_assert(GPC + GPC / 2 <= 15, "named barriers: one per line group and one per row item");

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

    // Both output rows of the register half HALF (registers 16*HALF .. 16*HALF+15 = half of this thread's columns) of a
    // regular row item.  x: transformed line of packed field 0 / 2, y: of field 1 / 3 (same columns, same thread index).
    // reference: WSTessendorf.cpp:385-437 (sign, lambda, packing) and :443-455 (normalisation)
    template <int MODE, int HALF>
    static WSO_HD void pack_half(const float2* x, const float2* y, int tl, int mp, float lambda, float inv_amp,
                                 float4* outA, float4* outB) {
        const float s = ((mp + column_of<0>(tl)) & 1) ? -1.0f : 1.0f;
        const float sl = rmul(s, lambda);
        static_for<0, 16>([&](auto rc) {
            constexpr int Q = decltype(rc)::value;
            constexpr int R = 16 * HALF + Q;
            const int c = column_of<R>(tl);
            const int cm = (N - c) & (N - 1);  // row N-m' is the conjugate mirror of row m'
            if (MODE == 0) {
                const float yy = rmul(rmul(x[Q].x, s), inv_amp);
                const float xx = rmul(sl, x[Q].y), zz = rmul(sl, y[Q].y);
                st_stream(&outA[c], make_float4(xx, yy, zz, 1.0f));
                st_stream(&outB[cm], make_float4(-xx, yy, -zz, 1.0f));
            } else {
                const float4 ta = make_float4(s * x[Q].y, s * y[Q].y, s * x[Q].x, s * y[Q].x);
                st_stream(&outA[c], ta);
                st_stream(&outB[cm], make_float4(-ta.x, -ta.y, ta.z, ta.w));
            }
        });
    }

    // One row item of a map = TWO line groups: role 0 transforms packed field 0 / 2, role 1 field 1 / 3 (same thread
    // index = same columns).  Each hands the half of its result the other one packs through its own line buffer, so a
    // thread never holds more than one transformed line plus half of another.
    // nbx: CTAs walking the row items (gridDim.x); n_items: tile-frames of this launch
    template <int MODE, class Ctx, class Args>
    static WSO_HD void run(Ctx& cx, float2* smem, int bx, int nbx, int n_items, const Args& args) {
        constexpr bool MAPS = (MODE != 2);
        constexpr int GPU = MAPS ? 2 : 1;  // line groups per unit (row item)
        constexpr int UPC = GPC / GPU;     // units in flight per CTA
        static_assert(GPC % GPU == 0, "the map kernels need an even number of line groups per CTA");
        const int tid = cx.tid;
        const int g = tid / L, tl = tid % L;
        const int role = MAPS ? (g & 1) : 0;
        const int ug = MAPS ? (g >> 1) : g;
        const int b = InMap<L>::b_of(tl);
        const bool b0 = (b == 0);
        const int partner = InMap<L>::partner_lane(tid & 31, tl);
        float2* bufs = smem + (size_t)g * NBUF * LINE;
        float2* pbufs = smem + (size_t)(g ^ 1) * NBUF * LINE;  // the other line group of the row item (MAPS)
        uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + (size_t)GPC * NBUF * LINE) + g * 2;
        float2 wb[5];
        load_twiddle_bases<LOGN>(args.tw, b, wb);

        const int U = n_items * H;
        const int UT = nbx * UPC;
        const int uu = bx * UPC + ug;
        const int nlines = (U + UT - 1) / UT;
        const int field = MODE == 2 ? 0 : (MODE == 0 ? role : 2 + role);

        // barrier over the line groups of one row item
        auto unit_sync = [&]() {
            if constexpr (!MAPS) group_sync(cx, g);
            else if constexpr (2 * L <= 32) cx.syncwarp();
            else cx.bar(1 + GPC + ug, 2 * L);  // ids 1..GPC belong to the line groups (group_sync)
        };

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

        auto issue = [&](int k) {
            const int u = uu + k * UT;
            if (k < nlines && u < U && tl == 0) {
                const int it = u / H, mp = u - it * H;
                const float2* src = args.W + (size_t)it * ((size_t)H * 4 * N) + ((size_t)mp * 4 + field) * N;
                cx.bulk_g2s(bufs + (k % NBUF) * LINE, src, (unsigned)LINE_BYTES, &mbar[k % NBUF]);
            }
        };
        issue(0);
        if (NBUF == 2) issue(1);

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
            const int u = uu + k * UT;
            if (u >= U) break;  // uniform over the unit and, for K2h at L = 16, over the warp (U and UT are even)
            const int it = u / H, mp = u - it * H;
            const bool special = (mp == 0);  // rows 0 and N/2 ride one complex line: separated through shared memory
            // K2h, L = 16: two row items share a warp; the neighbour of the special one follows its barriers
            const bool sync_special = (!MAPS && L == 16) ? ((mp >> 1) == 0) : special;
            float2* buf = bufs + (k % NBUF) * LINE;
            float2 v[32];
            cx.mbar_wait(&mbar[k % NBUF], (unsigned)((k / NBUF) & 1));
            transform_line(cx, buf, g, tl, b, b0, partner, wb, v);
            if (!MAPS && !sync_special) issue(k + NBUF);
            finish_line(cx, tl, v);
            if (special) {
                // natural order back into the line's own buffer
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
                    issue(k + NBUF);
                }
            } else {
                // hand over the half the other group packs: role 0 keeps registers 0..15, role 1 keeps 16..31
                if (!special) {
                    if (role == 0) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) buf[q * L + tl] = v[16 + q];
                    } else {
#pragma unroll
                        for (int q = 0; q < 16; ++q) buf[q * L + tl] = v[q];
                    }
                }
                unit_sync();
                if (!waited) {
                    // only the displacement map needs K2h's result (disp.y is written already divided by A): the
                    // groups of the normal map pack right away and fill the device while K2h drains
                    if (MODE == 0) cx.pdl_wait();
                    cx.pdl_release();
                    waited = true;
                }
                const BatchItem item = args.items[it];
                const float lambda = args.td[it].lambda;
                const float amp = MODE == 0 ? amplitude_of(args.minmax[2 * item.slot], args.minmax[2 * item.slot + 1]) : 1.0f;
                const float inv_amp = rdiv(1.0f, amp);
                if (MODE == 0 && mp == 0 && role == 0 && tl == 0) args.amp_out[item.slot] = amp;
                float4* out = (MODE == 0 ? args.disp : args.norm) + (size_t)item.slot * ((size_t)N * N);
                float4* outA = out + (size_t)mp * N;
                float4* outB = out + (size_t)(mp == 0 ? H : N - mp) * N;
                const float2* pbuf = pbufs + (k % NBUF) * LINE;
                if (!special) {
                    float2 got[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) got[q] = pbuf[q * L + tl];
                    cx.fence_async_smem();
                    unit_sync();  // both groups have what they need: the line buffers may be refilled
                    issue(k + NBUF);
                    if (role == 0) pack_half<MODE, 0>(v, got, tl, mp, lambda, inv_amp, outA, outB);
                    else pack_half<MODE, 1>(got, v + 16, tl, mp, lambda, inv_amp, outA, outB);
                } else {
                    const float2* l0 = role == 0 ? buf : pbuf;  // field 0 / 2
                    const float2* l1 = role == 0 ? pbuf : buf;  // field 1 / 3
                    for (int c = tl + L * role; c < N; c += 2 * L) {
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
                    unit_sync();
                    issue(k + NBUF);
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
