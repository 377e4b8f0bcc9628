// wso_kernels.cuh — CTA bodies of the three hot-path kernels (see DESIGN.md §3 for the derivation).
//
//   K1  Pass1  : spectrum evolve h0 -> h~(k,t) -> 4 real packed spectra Z_f, fused with the first 1-D
//                transform (along m, two real columns per complex FFT) -> Hermitian half W[m'][f][slot]
//   K2  Pass2  : second 1-D transform (along n) fused with sign fix, lambda and packing into the
//                RGBA32F displacement / normal maps + height min/max
//   K3  Normalize : disp.y *= 1/A,  A = max(|min|,|max|)       (reference: WSTessendorf.cpp:443-455)
//
// Replaces reference WSTessendorf::ComputeWaves (src/scene/WSTessendorf.cpp:284-441) including the
// seven fftwf_execute calls (cpp:338-378).
#pragma once

#include <cstring>

#include "wso_device.cuh"

namespace wso {

static constexpr int kMaxChunk = 64;   // batch items (tile-frames) per kernel launch
static constexpr int kMaxTable = 1024;  // entries of the per-frame sincos table kept in shared memory

// Per-tile constants, device pointers (written once per Prepare()).
struct TileDev {
    // [n][m] (TRANSPOSED) one 16-byte record per wave vector (m,n):
    //   x,y = heightAmp re,im (reference h0 record)          z = 1/|k| (0 where |k| <= 1e-5: WSTessendorf.h:135)
    //   w   = quantised dispersion omega (WSTessendorf.h:284-287), or - when table_len > 0 - its integer
    //         multiple j of the base frequency (omega == fl(float(j)*omega0) exactly, checked at Prepare)
    const float4* h0;
    const float* kv;     // [N]: kv[i] = (float)(M_PI*(2.0f*i-N)/L)  (reference: WSTessendorf.cpp:75-80)
    float lambda;        // displacement scale (reference: WSTessendorf.h:181)
    float omega0;        // base frequency (float)(2*pi/T)
    int table_len;       // j_max+1 when the per-frame sincos table is usable, else 0
    int pad_;
};

struct BatchItem {
    uint32_t tile;  // which h0 / parameter set
    uint32_t slot;  // which output map slot
    float t;        // time
};

struct LaunchArgs {
    const TileDev* tiles;  // device array
    const float2* tw;      // [N] exp(+2*pi*i*k/N)
    float2* W;             // chunk scratch: [item][N/2][4][N]
    float4* disp;          // [slot][N*N]
    float4* norm;          // [slot][N*N]
    float* minmax;         // [slot][2]
    float* amp_out;        // [slot] amplitude A
    BatchItem items[kMaxChunk];
};

// reference: WSTessendorf.cpp:289-290 — max starts at FLT_MIN (smallest positive), min at FLT_MAX
static constexpr float kInitMax = 1.17549435e-38f;
static constexpr float kInitMin = 3.402823466e+38f;

struct Point {
    float H, kx, kz, ux, uz;
};

// h~(k,t) and the wave-vector factors for wave vector (m,n).
// reference: WaveHeightFT, WSTessendorf.h:265-275 with heightAmp_conj == conj(heightAmp) (validated at
// import): h~ = 2*(a*cos(wt) - b*sin(wt)), imaginary part exactly 0.  Unit vector: WSTessendorf.h:133-136.
// The phase w*t is the same single fp32 product as the reference's; with TABLE the (cos,sin) pair comes
// from the per-frame table over j (bit-identical: the table entry is sincosf(fl(fl(j*omega0)*t))).
template <bool TABLE>
WSO_HD Point eval_point(const TileDev& td, const float2* table, int N, int m, int n, float t) {
    const float4 q = td.h0[n * N + m];
    float s, c;
    if (TABLE) {
#if defined(__CUDA_ARCH__)
        const float2 cs = table[__float_as_int(q.w)];
#else
        int j; std::memcpy(&j, &q.w, 4);
        const float2 cs = table[j];
#endif
        c = cs.x;
        s = cs.y;
    } else {
        sincos_acc(rmul(q.w, t), &s, &c);
    }
    const float x = rsub(rmul(q.x, c), rmul(q.y, s));
    Point p;
    p.H = radd(x, x);
    p.kx = td.kv[n];
    p.kz = td.kv[m];
    p.ux = rmul(p.kx, q.z);
    p.uz = rmul(p.kz, q.z);
    return p;
}

// Even-type (real spectrum) and odd-type (imaginary spectrum i*V) member of packed field F.
//   F=0: (height, Dx)   F=1: (none, Dz)   F=2: (dxDx, slopeX)   F=3: (dzDz, slopeZ)
// reference: WSTessendorf.cpp:303-336 (same products in the same order).
template <int F>
WSO_HD void field_values(const Point& p, float* R, float* V) {
    if (F == 0) { *R = p.H;                              *V = rmul(-p.ux, p.H); }
    if (F == 1) { *R = 0.0f;                             *V = rmul(-p.uz, p.H); }
    if (F == 2) { *R = rmul(p.kx, rmul(p.ux, p.H));      *V = rmul(p.kx, p.H); }
    if (F == 3) { *R = rmul(p.kz, rmul(p.uz, p.H));      *V = rmul(p.kz, p.H); }
}

// Z = even(R) - odd(V) under DFT-index reflection, for the 4 points (mA|mB) x (nA|nB).
// MASK bit1: rows are mirrored into each other (i >= 1); bit0: columns are (j >= 1).
template <int F, int MASK>
WSO_HD void pack_field(const Point (&pt)[4], float2* outA, float2* outB) {
    float R[4], V[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) field_values<F>(pt[q], &R[q], &V[q]);
    float z[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int qm = q ^ MASK;
        z[q] = 0.5f * (R[q] + R[qm]) - 0.5f * (V[q] - V[qm]);
    }
    // q = 2*a + b : a = row (A,B), b = column (A,B).  complex sample = Z(m, nA) + i Z(m, nB)
    *outA = make_float2(z[0], z[1]);
    *outB = make_float2(z[2], z[3]);
}

// -------------------------------------------------------------------------------------------------
// K1
// -------------------------------------------------------------------------------------------------
template <int LOGN, int CP, int NF>
struct Pass1 {
    static constexpr int N = 1 << LOGN;
    static constexpr int H = N / 2;
    static constexpr int B = CP * NF;                 // FFT lines per CTA
    static constexpr int T = B * N / kValsPerThread;  // threads per CTA
    static constexpr int LS = LineStride<N>::value;
    static constexpr int SMEM_BYTES = (B * LS + kMaxTable) * (int)sizeof(float2);
    static_assert(T >= 1 && T <= 1024, "bad CTA size");
    static_assert(H % CP == 0 && 4 % NF == 0, "bad tiling");

    template <int MASK, bool TABLE>
    static WSO_HD void evolve_item(const TileDev& td, const float2* table, float t, int fg, float2* smem,
                                   int cp, int mA, int mB, int nA, int nB) {
        Point pt[4];
        pt[0] = eval_point<TABLE>(td, table, N, mA, nA, t);
        pt[1] = eval_point<TABLE>(td, table, N, mA, nB, t);
        pt[2] = eval_point<TABLE>(td, table, N, mB, nA, t);
        pt[3] = eval_point<TABLE>(td, table, N, mB, nB, t);
        const int eA = pad_idx(mA), eB = pad_idx(mB);
        float2 a, b;
        if (NF == 4) {
            pack_field<0, MASK>(pt, &a, &b); smem[(0 * CP + cp) * LS + eA] = a; smem[(0 * CP + cp) * LS + eB] = b;
            pack_field<1, MASK>(pt, &a, &b); smem[(1 * CP + cp) * LS + eA] = a; smem[(1 * CP + cp) * LS + eB] = b;
            pack_field<2, MASK>(pt, &a, &b); smem[(2 * CP + cp) * LS + eA] = a; smem[(2 * CP + cp) * LS + eB] = b;
            pack_field<3, MASK>(pt, &a, &b); smem[(3 * CP + cp) * LS + eA] = a; smem[(3 * CP + cp) * LS + eB] = b;
        } else if (NF == 2) {
            if (fg == 0) {
                pack_field<0, MASK>(pt, &a, &b); smem[(0 * CP + cp) * LS + eA] = a; smem[(0 * CP + cp) * LS + eB] = b;
                pack_field<1, MASK>(pt, &a, &b); smem[(1 * CP + cp) * LS + eA] = a; smem[(1 * CP + cp) * LS + eB] = b;
            } else {
                pack_field<2, MASK>(pt, &a, &b); smem[(0 * CP + cp) * LS + eA] = a; smem[(0 * CP + cp) * LS + eB] = b;
                pack_field<3, MASK>(pt, &a, &b); smem[(1 * CP + cp) * LS + eA] = a; smem[(1 * CP + cp) * LS + eB] = b;
            }
        } else {
            if (fg == 0) pack_field<0, MASK>(pt, &a, &b);
            if (fg == 1) pack_field<1, MASK>(pt, &a, &b);
            if (fg == 2) pack_field<2, MASK>(pt, &a, &b);
            if (fg == 3) pack_field<3, MASK>(pt, &a, &b);
            smem[cp * LS + eA] = a;
            smem[cp * LS + eB] = b;
        }
    }

    // bx: column-pair group, by: field group, bz: item within the chunk
    template <class Exec>
    static WSO_HD void run(Exec& ex, float2* smem, int bx, int by, int bz, const LaunchArgs& args) {
        const BatchItem item = args.items[bz];
        const TileDev td = args.tiles[item.tile];
        const float t = item.t;

        // the height min/max accumulators of this item's slot are reset here, ahead of K2
        if (bx == 0 && by == 0) {
            ex.each([&](int tid, ThreadState&) {
                if (tid == 0) {
                    args.minmax[2 * item.slot + 0] = kInitMin;
                    args.minmax[2 * item.slot + 1] = kInitMax;
                }
            });
        }

        // ---- per-frame (cos,sin)(omega_j * t) table: omega takes few distinct values j*omega0 ------
        float2* table = smem + B * LS;
        const bool use_table = td.table_len > 0;
        if (use_table) {
            ex.each([&](int tid, ThreadState&) {
                for (int j = tid; j < td.table_len; j += T) {
                    float s, c;
                    sincos_acc(rmul(rmul((float)j, td.omega0), t), &s, &c);
                    table[j] = make_float2(c, s);
                }
            });
            ex.sync();
        }

        // ---- evolve: 4 points per work item (rows mA,mB x columns nA,nB), all NF fields ----------
        ex.each([&](int tid, ThreadState&) {
            for (int it = tid; it < CP * H; it += T) {
                const int cp = it / H;
                const int i = it % H;
                const int j = bx * CP + cp;
                const int mA = i, mB = (i == 0) ? H : N - i;
                const int nA = j, nB = (j == 0) ? H : N - j;
                const int mask = ((i != 0) ? 2 : 0) | ((j != 0) ? 1 : 0);
                if (use_table) {
                    if (mask == 3) evolve_item<3, true>(td, table, t, by, smem, cp, mA, mB, nA, nB);
                    else if (mask == 2) evolve_item<2, true>(td, table, t, by, smem, cp, mA, mB, nA, nB);
                    else if (mask == 1) evolve_item<1, true>(td, table, t, by, smem, cp, mA, mB, nA, nB);
                    else evolve_item<0, true>(td, table, t, by, smem, cp, mA, mB, nA, nB);
                } else {
                    if (mask == 3) evolve_item<3, false>(td, table, t, by, smem, cp, mA, mB, nA, nB);
                    else if (mask == 2) evolve_item<2, false>(td, table, t, by, smem, cp, mA, mB, nA, nB);
                    else if (mask == 1) evolve_item<1, false>(td, table, t, by, smem, cp, mA, mB, nA, nB);
                    else evolve_item<0, false>(td, table, t, by, smem, cp, mA, mB, nA, nB);
                }
            }
        });
        ex.sync();

        // ---- B complex FFTs of length N along m ---------------------------------------------------
        RunStages<LOGN, B, 0, 1, Exec>::run(ex, smem, args.tw);
        ex.sync();  // the split below reads CP lines per work item

        // ---- split the two real columns, keep m' in [0, N/2), store W[m'][f][slot] ---------------
        float2* Wit = args.W + (size_t)bz * ((size_t)H * 4 * N);
        ex.each([&](int tid, ThreadState&) {
            for (int it = tid; it < NF * H * CP; it += T) {
                const int cp = it % CP;
                const int rest = it / CP;
                const int mp = rest % H;
                const int fl = rest / H;
                const float2* line = smem + (fl * CP + cp) * LS;
                const float2 c1 = line[pad_idx(mp)];
                const float2 c2 = line[pad_idx((N - mp) & (N - 1))];
                float2 wa = make_float2(0.5f * (c1.x + c2.x), 0.5f * (c1.y - c2.y));
                float2 wb = make_float2(0.5f * (c1.y + c2.y), -0.5f * (c1.x - c2.x));
                if (mp == 0) {  // pack the (real) Nyquist bin m'=N/2 into the imaginary part of m'=0
                    const float2 ch = line[pad_idx(H)];
                    wa.y = ch.x;
                    wb.y = ch.y;
                }
                const int f = by * NF + fl;
                const int j = bx * CP + cp;
                float2* dst = Wit + ((size_t)mp * 4 + f) * N;
                dst[j] = wa;
                dst[H + j] = wb;
            }
        });
    }
};

// -------------------------------------------------------------------------------------------------
// K2
// -------------------------------------------------------------------------------------------------
WSO_HD int slot_of_column(int n, int N) {
    const int H = N >> 1;
    return (n < H) ? n : ((n == H) ? H : H + (N - n));
}

template <int LOGN, int RI>
struct Pass2 {
    static constexpr int N = 1 << LOGN;
    static constexpr int H = N / 2;
    static constexpr int B = RI * 2;
    static constexpr int T = B * N / kValsPerThread;
    static constexpr int LS = LineStride<N>::value;
    static constexpr int SMEM_BYTES = B * LS * (int)sizeof(float2);
    static_assert(T >= 1 && T <= 1024, "bad CTA size");
    static_assert(H % RI == 0, "bad tiling");

    // bx: row-item group, by: 0 = displacement map (fields 0,1), 1 = normal map (fields 2,3), bz: item
    template <class Exec>
    static WSO_HD void run(Exec& ex, float2* smem, int bx, int by, int bz, const LaunchArgs& args) {
        const BatchItem item = args.items[bz];
        const float lambda = args.tiles[item.tile].lambda;
        const float2* Wit = args.W + (size_t)bz * ((size_t)H * 4 * N);

        // ---- first stage straight from global memory (W rows are contiguous) ---------------------
        {
            constexpr int R = Plan<LOGN>::R[0];
            using St = Stage<N, B, R, 1>;
            ex.each([&](int tid, ThreadState& st) {
#pragma unroll
                for (int i = 0; i < St::NB; ++i) {
                    const int line = tid / St::G;
                    const int j = tid % St::G + St::G * i;
                    const int mp = bx * RI + (line >> 1);
                    const int f = by * 2 + (line & 1);
                    const float2* src = Wit + ((size_t)mp * 4 + f) * N;
#pragma unroll
                    for (int r = 0; r < R; ++r) st.v[i * R + r] = src[slot_of_column(j + r * St::JN, N)];
                }
                St::twiddle_dft(args.tw, tid, st);
                St::store(smem, tid, st);
            });
            ex.template sync_group<St::G, T>(1);
            if constexpr (Plan<LOGN>::S > 1) RunStages<LOGN, B, 1, R, Exec>::run(ex, smem, args.tw);
        }
        // the pack phase of a row item reads both of its lines: barrier over that pair of line groups
        constexpr int G2 = 2 * (N / kValsPerThread);
        ex.template sync_group<G2, T>(1 + ((N / kValsPerThread) > 32 ? B : 0));

        // ---- pack: each transformed line pair yields output rows m' and N-m' ---------------------
        float4* out = (by == 0 ? args.disp : args.norm) + (size_t)item.slot * ((size_t)N * N);
        ex.each([&](int tid, ThreadState& st) {
            float mn = kInitMin, mx = kInitMax;
            const int ri = tid / G2;                  // the row item this thread's line pair belongs to
            for (int c = tid % G2; c < N; c += G2) {  // output column n'
                const int cm = (N - c) & (N - 1);     // mirrored column
                const int mp = bx * RI + ri;
                const float2* l0 = smem + (ri * 2 + 0) * LS;
                const float2* l1 = smem + (ri * 2 + 1) * LS;
                float2 a0 = l0[pad_idx(c)], a1 = l1[pad_idx(c)];  // F at (rowA, c)
                float2 b0, b1;                                    // F at (rowB, colB)
                int rowA, rowB, colB;
                if (mp == 0) {
                    // rows 0 and N/2 were transformed as one complex line: separate them
                    const float2 m0 = l0[pad_idx(cm)], m1 = l1[pad_idx(cm)];
                    b0 = make_float2(0.5f * (a0.y + m0.y), -0.5f * (a0.x - m0.x));
                    b1 = make_float2(0.5f * (a1.y + m1.y), -0.5f * (a1.x - m1.x));
                    a0 = make_float2(0.5f * (a0.x + m0.x), 0.5f * (a0.y - m0.y));
                    a1 = make_float2(0.5f * (a1.x + m1.x), 0.5f * (a1.y - m1.y));
                    rowA = 0; rowB = H; colB = c;
                } else {
                    b0 = cconj(a0);
                    b1 = cconj(a1);
                    rowA = mp; rowB = N - mp; colB = cm;
                }
                // reference: WSTessendorf.cpp:385-437 — sign = (-1)^(m+n)
                const float sA = ((rowA + c) & 1) ? -1.0f : 1.0f;
                const float sB = ((rowB + colB) & 1) ? -1.0f : 1.0f;
                float4 ta, tb;
                if (by == 0) {
                    const float hA = rmul(a0.x, sA), hB = rmul(b0.x, sB);
                    mn = hA < mn ? hA : mn; mx = hA > mx ? hA : mx;
                    mn = hB < mn ? hB : mn; mx = hB > mx ? hB : mx;
                    ta = make_float4(rmul(rmul(sA, lambda), a0.y), hA, rmul(rmul(sA, lambda), a1.y), 1.0f);
                    tb = make_float4(rmul(rmul(sB, lambda), b0.y), hB, rmul(rmul(sB, lambda), b1.y), 1.0f);
                } else {
                    ta = make_float4(sA * a0.y, sA * a1.y, sA * a0.x, sA * a1.x);
                    tb = make_float4(sB * b0.y, sB * b1.y, sB * b0.x, sB * b1.x);
                }
                out[(size_t)rowA * N + c] = ta;
                out[(size_t)rowB * N + colB] = tb;
            }
            st.v[0] = make_float2(mn, mx);
        });
        if (by == 0) ex.commit_minmax(args.minmax + 2 * item.slot);
    }
};

// -------------------------------------------------------------------------------------------------
// K3 — reference: NormalizeHeights, WSTessendorf.cpp:443-455
// -------------------------------------------------------------------------------------------------
WSO_HD float amplitude_of(float mn, float mx) {
    const float a = mn < 0.0f ? -mn : mn;
    const float b = mx < 0.0f ? -mx : mx;
    return a > b ? a : b;
}

}  // namespace wso
