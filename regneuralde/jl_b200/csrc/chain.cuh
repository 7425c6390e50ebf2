// chain.cuh -- field evaluation and VJP for "chain" fields: a Flux Chain of up to 8 Dense layers without time
// input, optionally preceded by an elementwise tanh -- the Latent-ODE generator dynamics
//   Chain(x -> tanh.(x), Dense(20,50,tanh), Dense(50,20,tanh), ... x8)      /root/reference/experiments/latent_ode.jl:109-121
// evaluated as re(p)(u) because the node is built with time_dep = false (src/models/neural_ode.jl:57,88,119,155).
// All weights (8 280 floats for the reference's shape) live in shared memory for the whole solve; one CTA owns a tile
// of NP columns and one thread one (output row, column) element of a layer.
// Canonical arithmetic (oracle/rnde_oracle.c chain_column): one fma chain over the inputs in ascending order starting
// from 0, then + bias, then the activation.
#pragma once
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
// grid = (splits, layers, output chunks); a CTA walks a contiguous range of (record, tile) pairs, stages 64 columns at
// a time and every thread owns up to CW_OUT (output, input) pairs of its chunk; partial sums are FP32 over one stage
// and FP64 across stages (the regulariser cotangents cancel between records: DESIGN.md section 5).
constexpr int CW_NT = 256, CW_COLS = 64, CW_LD = 68;   // CW_LD: padded row stride (conflict-free LDS.128)
constexpr int CW_TPT = 2;                              // 4 x 4 output tiles per thread: 512 tiles >= 10 x 44 (the 40 x 176 GRU layers)
struct WgLayer {
    const float* dptr; const float* aptr;     // delta tape, input-activation tape
    int dstride, astride;                     // floats per (record, tile) block of each tape
    int doff, aoff;                           // first row of this layer's delta / input inside a block
    int M, K, poff;                           // out, in, offset of W in the parameter vector (bias follows W)
    int nobias;                               // 1: no bias row (a second outer product into the same W: the FFJORD transposed chain)
};
struct WgDesc { WgLayer l[12]; int nl; };

// rows x 64 columns of a tape into shared memory as float4 (tile-major items: consecutive threads read consecutive rows of one
// (record, tile) block = contiguous memory); `ones` appends a row of 1 (the bias input)
__device__ __forceinline__ void wg_stage(float* dst, const float* __restrict__ src, const int stride, const int row0, const int rows, const bool ones,
                                         const int NPt, const long long t0, const long long ntile) {
    const int v4 = NPt / 4, tiles = CW_COLS / NPt;
    const int per_tile = rows * v4;
    for (int item = threadIdx.x; item < tiles * per_tile; item += CW_NT) {
        const int tl = item / per_tile, rem = item - tl * per_tile;
        const int row = rem / v4, v = rem - row * v4;
        const long long tile = t0 + tl;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tile < ntile) x = __ldcg(reinterpret_cast<const float4*>(src + (size_t)tile * stride + (size_t)(row0 + row) * NPt + v * 4));
        *reinterpret_cast<float4*>(dst + row * CW_LD + tl * NPt + v * 4) = x;
    }
    if (ones) {
        for (int c = threadIdx.x; c < CW_COLS; c += CW_NT) dst[rows * CW_LD + c] = (t0 + c / NPt < ntile) ? 1.f : 0.f;
    }
}

__global__ void __launch_bounds__(CW_NT) dense_wgrad_kernel(const WgDesc desc, const int NPt, const long long ntile, double* __restrict__ acc_out) {
    extern __shared__ __align__(16) float csm[];
    const int tid = threadIdx.x;
    const WgLayer& Ld = desc.l[blockIdx.y];
    const int M = Ld.M, K = Ld.K;
    const int MT = (M + 3) / 4, IT = (K + 1 + 3) / 4;        // 4 x 4 output tiles: 4 outputs x 4 inputs (input K = bias)
    const int ntiles_out = MT * IT;
    float* sDel = csm;                                       // round_up(M,4) x CW_LD
    float* sAct = csm + MT * 4 * CW_LD;                      // round_up(K+1,4) x CW_LD, row K = 1 (bias)
    for (int e = tid; e < (MT * 4 + IT * 4) * CW_LD; e += CW_NT) csm[e] = 0.f;      // padding rows stay zero
    const int tiles_per_stage = CW_COLS / NPt;
    const long long nstage = (ntile + tiles_per_stage - 1) / tiles_per_stage;
    const long long s0 = nstage * blockIdx.x / gridDim.x, s1 = nstage * (blockIdx.x + 1) / gridDim.x;
    double acc[CW_TPT][16];
#pragma unroll
    for (int t = 0; t < CW_TPT; ++t)
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[t][r] = 0.0;
    for (long long s = s0; s < s1; ++s) {
        const long long t0 = s * tiles_per_stage;
        __syncthreads();
        wg_stage(sDel, Ld.dptr, Ld.dstride, Ld.doff, M, false, NPt, t0, ntile);
        wg_stage(sAct, Ld.aptr, Ld.astride, Ld.aoff, K, !Ld.nobias, NPt, t0, ntile);
        __syncthreads();
#pragma unroll
        for (int t = 0; t < CW_TPT; ++t) {
            const int ot = tid + t * CW_NT;
            if (ot < ntiles_out) {
                const int it = ot / MT, mt = ot - it * MT;
                const float* dp = sDel + mt * 4 * CW_LD;
                const float* ap = sAct + it * 4 * CW_LD;
                // Float32 partial sums over FOUR columns only, then Float64 (round 2: a 64-column Float32 partial made the parameter
                // gradient of the regularised toy / chain cases 2-2.5x noisier than the CPU Float32 adjoint; the products of the
                // cancelling O(10) cotangents must not pile up in Float32)
#pragma unroll 4
                for (int c4 = 0; c4 < CW_COLS / 4; ++c4) {
                    float4 d[4], a[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) { d[r] = *reinterpret_cast<const float4*>(dp + r * CW_LD + c4 * 4); a[r] = *reinterpret_cast<const float4*>(ap + r * CW_LD + c4 * 4); }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int o = 0; o < 4; ++o) {
                            double v = acc[t][i * 4 + o];
                            v = fma((double)d[o].x, (double)a[i].x, v); v = fma((double)d[o].y, (double)a[i].y, v);
                            v = fma((double)d[o].z, (double)a[i].z, v); v = fma((double)d[o].w, (double)a[i].w, v);
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
                    const int in = it * 4 + i, on = mt * 4 + o;
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
inline cudaError_t launch_dense_wgrad(const WgDesc& desc, int NPt, long long ntile, int np, int num_sms, double* acc, float* dp,
                                      cudaStream_t st, int64_t* launches) {
    cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double) * np, st);
    if (e != cudaSuccess) return e;
    int maxrows = 0;
    for (int l = 0; l < desc.nl; ++l) {
        const int rows = (desc.l[l].M + 3) / 4 * 4 + (desc.l[l].K + 1 + 3) / 4 * 4;
        maxrows = maxrows > rows ? maxrows : rows;
        if (((desc.l[l].M + 3) / 4) * ((desc.l[l].K + 1 + 3) / 4) > CW_TPT * CW_NT) return cudaErrorInvalidValue;
    }
    const long long nstage = (ntile * NPt + CW_COLS - 1) / CW_COLS;
    long long splits = 2LL * num_sms / desc.nl;
    if (splits > nstage) splits = nstage;
    if (splits < 1) splits = 1;
    const size_t smem = sizeof(float) * (size_t)maxrows * CW_LD;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(dense_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    dense_wgrad_kernel<<<dim3((unsigned)splits, desc.nl), CW_NT, smem, st>>>(desc, NPt, ntile, acc);
    wgrad_finish_kernel<<<(np + 255) / 256, 256, 0, st>>>(acc, dp, np);
    if (launches) *launches += 2;
    return cudaGetLastError();
}

}  // namespace rnde
