// chain.cuh -- field evaluation and VJP for "chain" fields: a Flux Chain of up to 8 Dense layers without time
// input, optionally preceded by an elementwise tanh -- the Latent-ODE generator dynamics
//   Chain(x -> tanh.(x), Dense(20,50,tanh), Dense(50,20,tanh), ... x8)      /root/reference/experiments/latent_ode.jl:109-121
// evaluated as re(p)(u) because the node is built with time_dep = false (src/models/neural_ode.jl:57,88,119,155).
// All weights (8 280 floats for the reference's shape) live in shared memory for the whole solve; one CTA owns a tile
// of NP columns and one thread one (output row, column) element of a layer.
// Canonical arithmetic (oracle/rnde_oracle.c chain_column): one fma chain over the inputs in ascending order starting
// from 0, then + bias, then the activation.
#pragma once
#include <algorithm>
#include <cstdlib>
#include "common.cuh"

namespace rnde {

struct ChainView {
    int L, D, NP, hrows;
    const int* w;      // widths
    const int* a;      // activations
    int pre;
    const float* sW;   // all parameters, Flux.destructure order
    float* sA; float* sB;   // ping-pong activations, maxw x NP each
    float* sH;              // backward only: this record's activations a_0..a_{L-2} (hrows x NP), staged once per VJP
};

// Quad-parallel small dense product.  Work item = (block of the contraction index, output row, group of 4 columns);
// the 4 items of a row sit in adjacent lanes, each runs one fma chain over its contiguous quarter of the contraction
// index for 4 columns, and two xor-shuffles add the quarters as (q0 + q1) + (q2 + q3) -- the canonical order of
// oracle chain_column.  Lane `blk` of the quad then finalises column 4*cg + blk through fin(row, column, value).
//   TRANS = false: value[a][n] = sum_c W[M*c + a] * in[c][n]   (a < M rows, c < K)      forward layer
//   TRANS = true : value[a][n] = sum_c W[M*a + c] * in[c][n]   (a < K rows, c < M)      W^T g in the VJP
template <int NP, int NT, bool TRANS, class Fin>
__device__ __forceinline__ void quad_dense(const float* __restrict__ W, const int M, const int K, const float* __restrict__ in, Fin fin) {
    const int A = TRANS ? K : M, Cn = TRANS ? M : K;
    const int kb = (Cn + 3) >> 2;
    const int total = A * 4 * (NP / 4);
    for (int base = 0; base < total; base += NT) {
        const int item = base + (int)threadIdx.x;
        const bool valid = item < total;
        const int blk = item & 3, rest = item >> 2;
        int a, cg;
        if constexpr (NP == 4) { a = valid ? rest : 0; cg = 0; }        // one column group: no index division
        else { a = valid ? rest % A : 0; cg = valid ? rest / A : 0; }
        const int c0 = blk * kb, c1 = valid ? min(c0 + kb, Cn) : c0;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        const float* wp = TRANS ? W + M * a + c0 : W + M * c0 + a;
        const int wstep = TRANS ? 1 : M;
        const float* ip = in + c0 * NP + cg * 4;
#pragma unroll 4
        for (int c = c0; c < c1; ++c) {
            const float w = *wp; wp += wstep;
            const float4 x = *reinterpret_cast<const float4*>(ip); ip += NP;
            acc0 = rn_fmaf(w, x.x, acc0); acc1 = rn_fmaf(w, x.y, acc1); acc2 = rn_fmaf(w, x.z, acc2); acc3 = rn_fmaf(w, x.w, acc3);
        }
        // (q0 + q1), (q2 + q3), then their sum: identical in all 4 lanes (IEEE addition commutes)
        acc0 = acc0 + __shfl_xor_sync(0xffffffffu, acc0, 1); acc1 = acc1 + __shfl_xor_sync(0xffffffffu, acc1, 1);
        acc2 = acc2 + __shfl_xor_sync(0xffffffffu, acc2, 1); acc3 = acc3 + __shfl_xor_sync(0xffffffffu, acc3, 1);
        acc0 = acc0 + __shfl_xor_sync(0xffffffffu, acc0, 2); acc1 = acc1 + __shfl_xor_sync(0xffffffffu, acc1, 2);
        acc2 = acc2 + __shfl_xor_sync(0xffffffffu, acc2, 2); acc3 = acc3 + __shfl_xor_sync(0xffffffffu, acc3, 2);
        const float v = blk == 0 ? acc0 : (blk == 1 ? acc1 : (blk == 2 ? acc2 : acc3));
        if (valid) fin(a, cg * 4 + blk, v);
    }
}

// The same product with compile-time layer shape (MC x KC) for 4-column tiles: the quarter loops unroll completely and every
// offset is a constant -- these layers are tiny (5 or 13 fma per lane) and otherwise dominated by loop and index overhead.
template <int NT, bool TRANS, int MC, int KC, class Fin>
__device__ __forceinline__ void quad_dense_fixed(const float* __restrict__ W, const float* __restrict__ in, Fin fin) {
    constexpr int NP = 4;
    constexpr int A = TRANS ? KC : MC, Cn = TRANS ? MC : KC;
    constexpr int kb = (Cn + 3) / 4;
    constexpr int total = A * 4;
    static_assert(total <= NT, "one pass");
    const int item = (int)threadIdx.x;
    const bool valid = item < total;
    const int blk = item & 3, a = valid ? (item >> 2) : 0;
    const int c0 = blk * kb;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    const float* wp = TRANS ? W + MC * a + c0 : W + MC * c0 + a;
    const float* ip = in + c0 * NP;
#pragma unroll
    for (int j = 0; j < kb; ++j) {
        if (valid && c0 + j < Cn) {
            const float w = TRANS ? wp[j] : wp[j * MC];
            const float4 x = *reinterpret_cast<const float4*>(ip + j * NP);
            acc0 = rn_fmaf(w, x.x, acc0); acc1 = rn_fmaf(w, x.y, acc1); acc2 = rn_fmaf(w, x.z, acc2); acc3 = rn_fmaf(w, x.w, acc3);
        }
    }
    acc0 = acc0 + __shfl_xor_sync(0xffffffffu, acc0, 1); acc1 = acc1 + __shfl_xor_sync(0xffffffffu, acc1, 1);
    acc2 = acc2 + __shfl_xor_sync(0xffffffffu, acc2, 1); acc3 = acc3 + __shfl_xor_sync(0xffffffffu, acc3, 1);
    acc0 = acc0 + __shfl_xor_sync(0xffffffffu, acc0, 2); acc1 = acc1 + __shfl_xor_sync(0xffffffffu, acc1, 2);
    acc2 = acc2 + __shfl_xor_sync(0xffffffffu, acc2, 2); acc3 = acc3 + __shfl_xor_sync(0xffffffffu, acc3, 2);
    const float v = blk == 0 ? acc0 : (blk == 1 ? acc1 : (blk == 2 ? acc2 : acc3));
    if (valid) fin(a, blk, v);
}

// dispatch: the Latent-ODE layer shapes (20 <-> 50, latent_ode.jl:109-121) get the unrolled form
template <int NP, int NT, bool TRANS, class Fin>
__device__ __forceinline__ void quad_dense_any(const float* __restrict__ W, const int M, const int K, const float* __restrict__ in, Fin fin) {
    if constexpr (NP == 4) {
        if (M == 50 && K == 20) { quad_dense_fixed<NT, TRANS, 50, 20>(W, in, fin); return; }
        if (M == 20 && K == 50) { quad_dense_fixed<NT, TRANS, 20, 50>(W, in, fin); return; }
    }
    quad_dense<NP, NT, TRANS>(W, M, K, in, fin);
}

// sOut = f(sIn).  rec >= 0: record z, a_0..a_{L-2} and k on the tape ([rec][tile][row][NP]).
template <int NP, int NT>
__device__ __forceinline__ void chain_rhs(const KParams& P, const ChainView& c, const float* sIn, float* sOut, const int rec, const int q) {
    const int tid = threadIdx.x;
    const int D = c.D;
    float* cur = c.sA; float* nxt = c.sB;
    const size_t hbase = ((size_t)max(rec, 0) * P.Q + q) * c.hrows * NP;
    const size_t dbase = ((size_t)max(rec, 0) * P.Q + q) * D * NP;
    for (int e = tid; e < D * NP; e += NT) {
        const float z = sIn[e];
        const float a0 = c.pre == RNDE_ACT_TANH ? canon_tanhf(z) : z;
        cur[e] = a0;
        if (rec >= 0) { P.tapeZ[dbase + e] = z; P.tapeH[hbase + e] = a0; }
    }
    __syncthreads();
    const float* W = c.sW;
    int K = D, hoff = D;
    for (int l = 0; l < c.L; ++l) {
        const int M = c.w[l];
        const float* b = W + M * K;
        const bool last = (l == c.L - 1);
        float* dst = last ? sOut : nxt;
        const int act = c.a[l];
        float* tp = rec < 0 ? nullptr : (last ? P.tapeK + dbase : P.tapeH + hbase + (size_t)hoff * NP);
        quad_dense_any<NP, NT, false>(W, M, K, cur, [&](const int o, const int n, const float s) {
            float v = s + b[o];
            if (act == RNDE_ACT_TANH) v = canon_tanhf(v);
            dst[o * NP + n] = v;
            if (tp) tp[o * NP + n] = v;
        });
        __syncthreads();
        if (!last) { float* t = cur; cur = nxt; nxt = t; hoff += M; }
        W = b + M; K = M;
    }
}

// VJP of record `rec`: on entry sKbar holds kbar (D x NP); delta_{L-1} replaces k on the tape, delta_l (l < L-1) goes
// to tapeD1 at the row offset of a_{l+1}; the input cotangent ends up in shared memory (D x NP), returned.
// kalt != nullptr (a6.cuh, second VJP of record 0): k is read from that copy and the deltas are ADDED to the record's.
template <int NP, int NT>
__device__ __forceinline__ const float* chain_vjp(const KParams& P, const ChainView& c, float* sKbar, const int rec, const int q, const float* kalt = nullptr) {
    const int tid = threadIdx.x;
    const int D = c.D;
    const size_t hbase = ((size_t)rec * P.Q + q) * c.hrows * NP;
    const size_t dbase = ((size_t)rec * P.Q + q) * D * NP;
    // parameter / activation-row offsets of every layer
    int poff[9], hoff[9];
    poff[0] = 0; hoff[0] = 0;
    {
        int K = D;
        for (int l = 0; l < c.L; ++l) { poff[l + 1] = poff[l] + c.w[l] * K + c.w[l]; hoff[l + 1] = hoff[l] + K; K = c.w[l]; }
    }
    float* g = c.sB; float* gn = c.sA;
    // one batch of tape loads per VJP (a single L2/HBM latency) instead of one dependent load per layer
    for (int e = tid; e < c.hrows * NP; e += NT) c.sH[e] = __ldcg(P.tapeH + hbase + e);
    for (int e = tid; e < D * NP; e += NT) {
        float d = sKbar[e];
        if (c.a[c.L - 1] == RNDE_ACT_TANH) { const float kv = kalt ? __ldcg(kalt + e) : __ldcg(P.tapeK + dbase + e); d = d * (1.f - kv * kv); }
        g[e] = d;
        if (kalt) P.tapeK[dbase + e] += d;
        else P.tapeK[dbase + e] = d;
    }
    __syncthreads();
    for (int l = c.L - 1; l >= 0; --l) {
        const int K = l ? c.w[l - 1] : D, M = c.w[l];
        // derivative of the activation that produced this layer's input (layer l-1's, or the pre-activation)
        const int actin = l ? c.a[l - 1] : c.pre;
        const float* ain = c.sH + hoff[l] * NP;
        float* dout = l ? P.tapeD1 + hbase + (size_t)hoff[l] * NP : nullptr;
        quad_dense_any<NP, NT, true>(c.sW + poff[l], M, K, g, [&](const int i, const int n, float v) {
            if (actin == RNDE_ACT_TANH) { const float av = ain[i * NP + n]; v = v * (1.f - av * av); }
            gn[i * NP + n] = v;
            if (dout) { if (kalt) dout[i * NP + n] += v; else dout[i * NP + n] = v; }
        });
        __syncthreads();
        float* t = g; g = gn; gn = t;
    }
    return g;
}

// Parameter gradients of Dense layers from a tape: for every layer dW = sum over (record, column) of delta a^T and
// db = sum delta.  Tapes are [record][tile][row][NP]; a layer is described by where its delta and its input rows live.
// grid = (splits, layers); a CTA walks a contiguous range of (record, tile) pairs 64 columns (a "stage") at a time and every
// thread owns one 4 x 4 tile of (output, input) pairs.  Every product is a Float64 fma (the regulariser cotangents cancel
// between records: DESIGN.md section 5; a 64-column Float32 partial made the parameter gradient of the regularised toy /
// chain cases 2-2.5x noisier than the CPU Float32 adjoint).  What round 2 measured on the FFJORD step and changed
// (profiles/r2zz_next_rows_ncu.txt):
//   * converting float -> double in the inner loop put 32 F2F per 64 DFMA on the XU pipe (16 lanes/clk/SM against 64 for DFMA): the
//     XU pipe sat at 100 %.  The operands are now converted ONCE per stage; shared memory holds doubles.
//   * the float rows 272 B apart made the delta reads 4-way bank conflicted.  A tile's four outputs are now MT rows apart, the
//     lanes of a warp read consecutive rows, CW_LD = 66 doubles = 16 B mod 128 B apart: conflict-free LDS.128.
//   * the loads of a stage were exposed (26 % of the warp samples waited on them, 29 % at the barrier behind them): the next stage's
//     rows now travel as cp.async into a raw Float32 buffer while this stage's products run, and are converted after them.
//   * 512 threads with one tile each instead of 256 with up to two (the second pass occupied 3 of 8 warps).
#ifndef RNDE_CW_NT
#define RNDE_CW_NT 512
#endif
constexpr int CW_NT = RNDE_CW_NT, CW_COLS = 64, CW_LD = 66;   // CW_LD: row stride in doubles
constexpr int CW_MAXTILES = 512;                               // 4 x 4 output tiles per pseudo-layer (>= 10 x 44: the 40 x 176 GRU layers)
constexpr int CW_TPT = CW_MAXTILES / CW_NT;
struct WgLayer {
    const float* dptr; const float* aptr;     // delta tape, input-activation tape
    int dstride, astride;                     // floats per (record, tile) block of each tape
    int doff, aoff;                           // first row of this layer's delta / input inside a block
    int M, K, poff;                           // out, in, offset of W in the parameter vector (bias follows W)
    int nobias;                               // 1: no bias row (a second outer product into the same W: the FFJORD transposed chain)
};
struct WgDesc { WgLayer l[16]; int nl; int cta0[17]; };      // cta0: first CTA of every layer (filled by launch_dense_wgrad)

__device__ __forceinline__ void cw_cp16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
// Item enumeration of a stage, shared by the asynchronous copy and the conversion: NPt is a power of two (4..64), so a
// (record, tile) block of a layer is per = rows * NPt / 4 consecutive float4; items are dealt as (tile of the stage, float4 of
// the block) with the block length padded to a power of two (shifts, no divisions).  visit(isA, tl, e, idx): idx = compact index.
struct WgGeom { int lv4, ltiles, perD, perA, lpD, lpA, nD2, n2; };
__device__ __forceinline__ WgGeom wg_geom(const WgLayer& Ld, const int NPt) {
    WgGeom g;
    g.lv4 = __ffs(NPt) - 3; g.ltiles = 6 - (g.lv4 + 2);
    g.perD = Ld.M << g.lv4; g.perA = Ld.K << g.lv4;
    g.lpD = 32 - __clz(max(g.perD, 1) - 1); g.lpA = 32 - __clz(max(g.perA, 1) - 1);
    g.nD2 = 1 << (g.lpD + g.ltiles); g.n2 = g.nD2 + (1 << (g.lpA + g.ltiles));
    return g;
}
template <class V>
__device__ __forceinline__ void wg_items(const WgGeom& g, V visit) {
    for (int e2 = threadIdx.x; e2 < g.n2; e2 += CW_NT) {
        const bool isA = e2 >= g.nD2;
        const int u = isA ? e2 - g.nD2 : e2, lp = isA ? g.lpA : g.lpD, per = isA ? g.perA : g.perD;
        const int tl = u >> lp, e = u & ((1 << lp) - 1);
        if (e < per) visit(isA, tl, e, (isA ? (g.perD << g.ltiles) : 0) + tl * per + e);
    }
}

__global__ void __launch_bounds__(CW_NT) dense_wgrad_kernel(const WgDesc desc, const int NPt, const long long ntile, double* __restrict__ acc_out) {
    extern __shared__ __align__(16) double csm[];
    const int tid = threadIdx.x;
    int layer = 0;
    while (layer + 1 < desc.nl && (int)blockIdx.x >= desc.cta0[layer + 1]) ++layer;
    const WgLayer& Ld = desc.l[layer];
    const int split = (int)blockIdx.x - desc.cta0[layer], nsplit = desc.cta0[layer + 1] - desc.cta0[layer];
    const int M = Ld.M, K = Ld.K;
    const int MT = (M + 3) / 4, IT = (K + 1 + 3) / 4;        // 4 x 4 output tiles: outputs {mt, mt+MT, mt+2MT, mt+3MT} x inputs 4it..4it+3 (input K = bias)
    const int ntiles_out = MT * IT;
    double* sDel = csm;                                      // 4*MT x CW_LD
    double* sAct = csm + MT * 4 * CW_LD;                     // 4*IT x CW_LD, row K = 1 (bias)
    float* sRaw = reinterpret_cast<float*>(csm + (MT * 4 + IT * 4) * CW_LD);      // (M + K) rows x 64 columns as they lie on the tapes
    for (int e = tid; e < (MT * 4 + IT * 4) * CW_LD; e += CW_NT) csm[e] = 0.0;      // padding rows stay zero
    const WgGeom g = wg_geom(Ld, NPt);
    const int tiles_per_stage = CW_COLS / NPt;
    const long long nstage = (ntile + tiles_per_stage - 1) / tiles_per_stage;
    const long long s0 = nstage * split / nsplit, s1 = nstage * (split + 1) / nsplit;
    auto issue = [&](const long long s) {
        const long long t0 = s * tiles_per_stage;
        wg_items(g, [&](const bool isA, const int tl, const int e, const int idx) {
            const long long tile = t0 + tl;
            float* dst = sRaw + (size_t)idx * 4;
            if (tile < ntile) {
                const float* src = isA ? Ld.aptr + (size_t)tile * Ld.astride + (size_t)Ld.aoff * NPt : Ld.dptr + (size_t)tile * Ld.dstride + (size_t)Ld.doff * NPt;
                cw_cp16(dst, src + (size_t)e * 4);
            } else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        });
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto convert = [&](const long long s) {
        const long long t0 = s * tiles_per_stage;
        wg_items(g, [&](const bool isA, const int tl, const int e, const int idx) {
            const float4 x = *reinterpret_cast<const float4*>(sRaw + (size_t)idx * 4);
            const int row = e >> g.lv4;
            double2* d2 = reinterpret_cast<double2*>((isA ? sAct : sDel) + row * CW_LD + tl * NPt + ((e - (row << g.lv4)) << 2));
            d2[0] = make_double2((double)x.x, (double)x.y);
            d2[1] = make_double2((double)x.z, (double)x.w);
        });
        if (!Ld.nobias)
            for (int c = tid; c < CW_COLS; c += CW_NT) sAct[K * CW_LD + c] = (t0 + c / NPt < ntile) ? 1.0 : 0.0;
    };
    double acc[CW_TPT][16];
#pragma unroll
    for (int t = 0; t < CW_TPT; ++t)
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[t][r] = 0.0;
    if (s0 < s1) issue(s0);
    for (long long s = s0; s < s1; ++s) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();                 // stage s has landed in sRaw; every thread is done with the doubles of stage s - 1
        convert(s);
        __syncthreads();
        if (s + 1 < s1) issue(s + 1);    // travels while the products of stage s run
#pragma unroll
        for (int t = 0; t < CW_TPT; ++t) {
            const int ot = tid + t * CW_NT;
            if (ot < ntiles_out) {
                const int it = ot / MT, mt = ot - it * MT;
                const double* dp = sDel + mt * CW_LD;
                const double* ap = sAct + it * 4 * CW_LD;
#pragma unroll 4
                for (int c2 = 0; c2 < CW_COLS / 2; ++c2) {
                    double2 d[4], a[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        d[r] = *reinterpret_cast<const double2*>(dp + r * MT * CW_LD + c2 * 2);
                        a[r] = *reinterpret_cast<const double2*>(ap + r * CW_LD + c2 * 2);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int o = 0; o < 4; ++o) {
                            double v = acc[t][i * 4 + o];
                            v = fma(d[o].x, a[i].x, v); v = fma(d[o].y, a[i].y, v);
                            acc[t][i * 4 + o] = v;
                        }
                }
            }
        }
    }
#pragma unroll
    for (int t = 0; t < CW_TPT; ++t) {
        const int ot = tid + t * CW_NT;
        if (ot < ntiles_out) {
            const int it = ot / MT, mt = ot - it * MT;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int in = it * 4 + i, on = mt + o * MT;
                    if (in < K + (Ld.nobias ? 0 : 1) && on < M) atomicAdd(acc_out + Ld.poff + in * M + on, acc[t][i * 4 + o]);   // Flux order: W column-major, bias (in == K) behind it
                }
        }
    }
}

__global__ void wgrad_finish_kernel(const double* __restrict__ acc, float* __restrict__ dp, const int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dp[i] = (float)acc[i];
}

// host helper: launch the contraction for all layers of `desc` and convert the FP64 sums to dp (overwritten)
inline size_t dense_wgrad_smem(int M, int K) {
    const size_t rows = (size_t)(M + 3) / 4 * 4 + (size_t)(K + 1 + 3) / 4 * 4;
    return sizeof(double) * rows * CW_LD + sizeof(float) * (size_t)(M + K) * CW_COLS;
}
constexpr size_t CW_SMEM_MAX = 227 * 1024;
inline cudaError_t launch_dense_wgrad(const WgDesc& desc, int NPt, long long ntile, int np, int num_sms, double* acc, float* dp,
                                      cudaStream_t st, int64_t* launches) {
    if (NPt < 4 || NPt > CW_COLS || (NPt & (NPt - 1)) != 0) return cudaErrorInvalidValue;      // wg_items shifts by log2(NPt)
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * np, st);
    if (e != cudaSuccess) return e;
    size_t smem = 0;
    for (int l = 0; l < desc.nl; ++l) {
        smem = std::max(smem, dense_wgrad_smem(desc.l[l].M, desc.l[l].K));
        if (((desc.l[l].M + 3) / 4) * ((desc.l[l].K + 1 + 3) / 4) > CW_MAXTILES) return cudaErrorInvalidValue;
    }
    if (smem > CW_SMEM_MAX) return cudaErrorInvalidValue;
    // CTAs per layer: the layers differ 10x in work per stage (output tiles) and a CTA fills the SM's shared memory, so the grid is
    // several CTAs per SM and layer for the block scheduler to balance.  Measured on the FFJORD step (profiles/r2zz_next_rows_ncu.txt):
    // 2 / 3 / 4 / 6 / 8 CTAs per SM in total: 2.72 / 2.30 / 2.21 / 2.07 / 2.12 ms; one wave of CTAs dealt to the layers by a cost model
    // (tiles + a fixed share): 3.6 - 4.5 ms.
    const long long nstage = (ntile * NPt + CW_COLS - 1) / CW_COLS;
    WgDesc d = desc;
    static const int oversub = [] { const char* v = getenv("RNDE_CW_OVERSUB"); const int n = v ? atoi(v) : 0; return n > 0 ? n : 6; }();
    long long splits = (long long)oversub * num_sms / d.nl;
    if (splits > nstage) splits = nstage;
    if (splits < 1) splits = 1;
    for (int l = 0; l <= d.nl; ++l) d.cta0[l] = (int)(l * splits);
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(dense_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dense_wgrad_kernel<<<d.cta0[d.nl], CW_NT, smem, st>>>(d, NPt, ntile, acc);
    wgrad_finish_kernel<<<(np + 255) / 256, 256, 0, st>>>(acc, dp, np);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace rnde
