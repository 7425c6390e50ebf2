// wgrad_kernel.cuh -- parameter gradients of the field as two batched contractions.
//
//   dW1aug[h][d] = sum_{rec, col} delta1[rec][h][col] * [Z; t; 1][rec][d][col]      (H x (D+2))
//   dW2aug[d][h] = sum_{rec, col} delta2[rec][d][col] * [Hact; t; 1][rec][h][col]   (D x (H+2))
// where rec runs over every field evaluation on the tape (1 + 6*naccept records) and col over the
// batch: a contraction of length nrec*B (~1e5 for the MNIST config), i.e. a real dense GEMM
// instead of ~200 rank-16 updates inside the latency-bound reverse sweep.  The "t" and "1" rows
// give the time-column and bias gradients for free.
// Replaces the weight-gradient part of Tracker's back-propagation through Flux.Dense inside dudt_
// (/root/reference/src/models/neural_ode.jl:120, experiments/mnist_node.jl:51-54).
//
// Layout: both operands are stored [tile][row][NP] (NP contiguous contraction entries per row).
// Kernel: 128x64 output tile per CTA, 64 contraction entries per pipeline stage copied straight
// into shared memory with cp.async (double buffered, no transposition: threads own interleaved
// rows so the k-contiguous float4 reads are bank-conflict free), 8x4 accumulators per thread.
// Accuracy: the regulariser part of the gradient is a sum of large cancelling terms (adjacent
// stages of one step carry +-O(10) cotangents, DESIGN.md "gradient conditioning"), so
//   * the contraction runs step by step and, inside a step, column tile by column tile with the six
//     records of the step back to back (tile_of): cancelling terms meet inside one FP32 chunk,
//   * split-K assigns CONTIGUOUS ranges of that order to a split,
//   * FP32 accumulators are flushed into FP64 every stage (64 entries), partials are FP64, the
//     final fixed-order reduction over splits is FP64 (deterministic, no atomics).
#pragma once
#include "common.cuh"

namespace rnde {

constexpr int WG_TM = 128;           // output rows per CTA (operand A)
constexpr int WG_TN = 64;            // output columns per CTA (operand B)
constexpr int WG_KC = 64;            // contraction entries per pipeline stage
constexpr int WG_LD = WG_KC + 4;     // padded row stride (floats): k-contiguous float4 reads are conflict free
constexpr int WG_SPLITS = 20;
constexpr int WG_FLUSH = 1;          // stages between FP32 -> FP64 flushes (round 2: every stage of 64 entries; 4 was measurably noisier on the toy shapes)

__host__ inline size_t wgrad_workspace_floats(int D, int H) {
    const size_t a = (size_t)H * (D + 2), b = (size_t)D * (H + 2);
    // FFMA path: WG_SPLITS double partials; tcgen05 path: 21 splits x 16 chunk slots of floats
    return (size_t)(21 * 16 + 8) * (a > b ? a : b);
}

__device__ __forceinline__ float rec_time(const StepRec* steps, float t0, int rec) {
    if (rec == 0) return t0;
    const int s = (rec - 1) / 6, i = (rec - 1) % 6 + 2;
    const StepRec sr = steps[s];
    return stage_time(sr.t, sr.dt, i);
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Contraction order.  Logical tile index tt -> (rec, q): record 0 first, then for every step all
// column tiles, and for every column tile the 6 records of that step back to back -- so the
// cotangents of adjacent stages of one step, which nearly cancel, are summed within one FP32 chunk.
__device__ __forceinline__ void tile_of(int tt, int Q, int& rec, int& q) {
    if (tt < Q) { rec = 0; q = tt; return; }
    const int u = tt - Q, per_step = 6 * Q;
    const int s = u / per_step, v = u - s * per_step;
    q = v / 6;
    rec = 1 + 6 * s + (v - q * 6);
}

// A: [ntiles][M][NP], Bm: [ntiles][Nrows][NP] (physical tile = rec*Q + q); out partial[split][M][Naug] (double)
template <int NP>
__global__ void __launch_bounds__(256, 1) wgrad_gemm_kernel(const float* __restrict__ A, int M, const float* __restrict__ Bm, int Nrows,
                                                           int ntiles, int Q, const StepRec* __restrict__ steps, float t0, int td,
                                                           double* __restrict__ partial) {
    extern __shared__ __align__(16) float wsm[];
    constexpr int STAGE_FLOATS = (WG_TM + WG_TN) * WG_LD;
    float* As[2] = {wsm, wsm + STAGE_FLOATS};
    float* Bs[2] = {wsm + WG_TM * WG_LD, wsm + STAGE_FLOATS + WG_TM * WG_LD};
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;       // thread owns A rows ty+16i (i<8), B rows tx+16j (j<4)
    const int m_base = blockIdx.x * WG_TM, n_base = blockIdx.y * WG_TN;
    const int Naug = Nrows + 2;
    constexpr int tps = WG_KC / NP;               // tiles per stage
    constexpr int vpr = NP / 4;                   // float4 per row per tile
    const int per = (ntiles + gridDim.z - 1) / gridDim.z;
    const int tile0 = blockIdx.z * per, tile1 = min(ntiles, tile0 + per);
    const int nstage = (tile1 - tile0 + tps - 1) / tps;

    auto issue = [&](int stage, int buf) {
#pragma unroll
        for (int sub = 0; sub < tps; ++sub) {
            const int tt = tile0 + stage * tps + sub;
            const bool live = tt < tile1;
            int rec = 0, q = 0;
            if (live) tile_of(tt, Q, rec, q);
            const size_t phys = (size_t)rec * Q + q;
            constexpr int NVA = WG_TM * vpr, NVB = WG_TN * vpr;
#pragma unroll
            for (int e = tid; e < NVA; e += 256) {
                const int row = e / vpr, k4 = e - row * vpr;
                float* da = As[buf] + row * WG_LD + sub * NP + k4 * 4;
                const int m = m_base + row;
                if (live && m < M) cp_async16(da, A + (phys * M + m) * NP + k4 * 4);
                else *reinterpret_cast<float4*>(da) = make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int e = tid; e < NVB; e += 256) {
                const int row = e / vpr, k4 = e - row * vpr;
                float* db = Bs[buf] + row * WG_LD + sub * NP + k4 * 4;
                const int n = n_base + row;
                if (live && n < Nrows) cp_async16(db, Bm + (phys * Nrows + n) * NP + k4 * 4);
                else {
                    float v = 0.f;
                    if (live && n == Nrows) v = td ? rec_time(steps, t0, rec) : 0.f;
                    else if (live && n == Nrows + 1) v = 1.f;
                    *reinterpret_cast<float4*>(db) = make_float4(v, v, v, v);
                }
            }
        }
        cp_async_commit();
    };

    const bool all_double = (long long)M * Nrows <= 16384;      // uniform: small fields afford Float64 products
    double accd[8][4];
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { accd[i][j] = 0.0; acc[i][j] = 0.f; }

    if (nstage > 0) issue(0, 0);
    for (int s = 0; s < nstage; ++s) {
        const int buf = s & 1;
        if (s + 1 < nstage) { issue(s + 1, buf ^ 1); cp_async_wait<1>(); }
        else cp_async_wait<0>();
        __syncthreads();
        const float* ap = As[buf] + ty * WG_LD;
        const float* bp = Bs[buf] + tx * WG_LD;
#pragma unroll 2
        for (int k = 0; k < WG_KC; k += 4) {
            float4 a4[8], b4[4];
#pragma unroll
            for (int i = 0; i < 8; ++i) a4[i] = *reinterpret_cast<const float4*>(ap + i * 16 * WG_LD + k);
#pragma unroll
            for (int j = 0; j < 4; ++j) b4[j] = *reinterpret_cast<const float4*>(bp + j * 16 * WG_LD + k);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (all_double) {      // small fields: every product in Float64 (the toy shapes' stiffness-estimate gradient needs it to stay within
                        // 1.5x of a CPU Float32 adjoint: 1.2x, against 2.5x with Float32 products and 7x with a Float32 partial per stage)
                        accd[i][j] = fma((double)a4[i].x, (double)b4[j].x, accd[i][j]); accd[i][j] = fma((double)a4[i].y, (double)b4[j].y, accd[i][j]);
                        accd[i][j] = fma((double)a4[i].z, (double)b4[j].z, accd[i][j]); accd[i][j] = fma((double)a4[i].w, (double)b4[j].w, accd[i][j]);
                    } else {
                        // four entries of ONE record (no cancellation among them) in Float32, then Float64: a Float32 partial over a whole
                        // stage mixes the records of a step, whose +-O(10) cotangents cancel (found at the end of round 2, tools/stiff_probe.py)
                        accd[i][j] += (double)rn_fmaf(a4[i].w, b4[j].w, rn_fmaf(a4[i].z, b4[j].z, rn_fmaf(a4[i].y, b4[j].y, a4[i].x * b4[j].x)));
                    }
                }
        }
        if ((s % WG_FLUSH) == WG_FLUSH - 1 || s == nstage - 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { accd[i][j] += (double)acc[i][j]; acc[i][j] = 0.f; }
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m_base + ty + 16 * i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n_base + tx + 16 * j;
            if (n < Naug) partial[((size_t)blockIdx.z * M + m) * Naug + n] = accd[i][j];
        }
    }
}

// fixed-order FP64 sum over splits; scatter into Flux.destructure layout:
//   outW[m + M*n] for n < Nrows (+ the time column n == Nrows when td), outb[m] from n == Nrows+1
__global__ void wgrad_reduce_kernel(const double* __restrict__ partial, int nsplit, int M, int Nrows, int td, float* __restrict__ outW,
                                    float* __restrict__ outb) {
    const int Naug = Nrows + 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * Naug) return;
    const int n = idx / M, m = idx - n * M;     // m fastest: coalesced writes of the column-major W
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[((size_t)sp * M + m) * Naug + n];
    if (n < Nrows) outW[(size_t)M * n + m] = (float)s;
    else if (n == Nrows) { if (td) outW[(size_t)M * Nrows + m] = (float)s; }
    else outb[m] = (float)s;
}

constexpr size_t WG_SMEM = sizeof(float) * 2 * (WG_TM + WG_TN) * WG_LD;

// dp layout: W1 (H x (D+td)), b1 (H), W2 (D x (H+td)), b2 (D)
static int launch_wgrad(int D, int H, int td, int nrec, int Q, int NP, int B, const float* tapeZ, const float* tapeD2, const float* tapeH,
                        const float* tapeD1, const StepRec* steps, float t0, float* ws, float* dp, cudaStream_t st, int64_t* launches) {
    (void)B;
    typedef void (*gemm_t)(const float*, int, const float*, int, int, int, const StepRec*, float, int, double*);
    gemm_t gemm = NP == 16 ? wgrad_gemm_kernel<16> : (NP == 32 ? wgrad_gemm_kernel<32> : wgrad_gemm_kernel<4>);
    cudaFuncSetAttribute(gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM);
    const int ntiles = nrec * Q;
    const int tps = WG_KC / NP;
    int nsplit = WG_SPLITS;
    if (ntiles < nsplit * tps) nsplit = (ntiles + tps - 1) / tps;
    if (nsplit < 1) nsplit = 1;
    double* wsd = reinterpret_cast<double*>(ws);
    float* dW1 = dp;
    float* db1 = dW1 + (size_t)H * (D + td);
    float* dW2 = db1 + H;
    float* db2 = dW2 + (size_t)D * (H + td);
    {   // dW1aug = delta1 . [Z; t; 1]^T
        dim3 grid((H + WG_TM - 1) / WG_TM, (D + 2 + WG_TN - 1) / WG_TN, nsplit);
        gemm<<<grid, 256, WG_SMEM, st>>>(tapeD1, H, tapeZ, D, ntiles, Q, steps, t0, td, wsd);
        const int tot = H * (D + 2);
        wgrad_reduce_kernel<<<(tot + 255) / 256, 256, 0, st>>>(wsd, nsplit, H, D, td, dW1, db1);
    }
    {   // dW2aug = delta2 . [Hact; t; 1]^T
        dim3 grid((D + WG_TM - 1) / WG_TM, (H + 2 + WG_TN - 1) / WG_TN, nsplit);
        gemm<<<grid, 256, WG_SMEM, st>>>(tapeD2, D, tapeH, H, ntiles, Q, steps, t0, td, wsd);
        const int tot = D * (H + 2);
        wgrad_reduce_kernel<<<(tot + 255) / 256, 256, 0, st>>>(wsd, nsplit, D, H, td, dW2, db2);
    }
    if (launches) *launches += 4;
    return (int)cudaGetLastError();
}

}  // namespace rnde
