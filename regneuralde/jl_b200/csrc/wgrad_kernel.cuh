// wgrad_kernel.cuh -- parameter gradients of the field as two batched contractions.
//
//   dW1aug[h][d] = sum_{rec, col} delta1[rec][h][col] * [Z; t; 1][rec][d][col]      (H x (D+2))
//   dW2aug[d][h] = sum_{rec, col} delta2[rec][d][col] * [Hact; t; 1][rec][h][col]   (D x (H+2))
// where rec runs over every field evaluation on the tape (1 + 6*naccept records) and col
// over the batch: a contraction of length nrec*B (~1e5 for the MNIST config), i.e. a real
// dense GEMM instead of ~200 rank-32 updates inside the latency-bound reverse sweep.
// The "t" and "1" rows give the time-column and bias gradients for free.
// Replaces the weight-gradient part of Tracker's back-propagation through
// Flux.Dense inside dudt_ (/root/reference/src/models/neural_ode.jl:120,
// experiments/mnist_node.jl:51-54).
//
// Both operands are stored [tile][row][NP] (NP contiguous contraction entries per row), so a
// 64-row x NP tile is one contiguous block.  Split-K over tiles with a deterministic two-pass
// reduction (no atomics): partials[split][M][Naug] then a fixed-order sum.
#pragma once
#include "common.cuh"

namespace rnde {

constexpr int WG_TILE = 64;
constexpr int WG_SPLITS = 24;

__host__ inline size_t wgrad_workspace_floats(int D, int H) {
    const size_t a = (size_t)H * (D + 2), b = (size_t)D * (H + 2);
    return (size_t)WG_SPLITS * (a > b ? a : b);
}

__device__ __forceinline__ float rec_time(const StepRec* steps, float t0, int rec) {
    if (rec == 0) return t0;
    const int s = (rec - 1) / 6, i = (rec - 1) % 6 + 2;
    const StepRec sr = steps[s];
    return stage_time(sr.t, sr.dt, i);
}

// A: [ntiles][M][NP], Bm: [ntiles][Nrows][NP]; out partial[split][M][Naug], Naug = Nrows + 2
__global__ void __launch_bounds__(256) wgrad_gemm_kernel(const float* __restrict__ A, int M, const float* __restrict__ Bm, int Nrows, int NP,
                                                        int ntiles, int Q, const StepRec* __restrict__ steps, float t0, int td,
                                                        float* __restrict__ partial) {
    constexpr int LD = WG_TILE + 4;
    __shared__ __align__(16) float As[32 * LD];
    __shared__ __align__(16) float Bs[32 * LD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m_base = blockIdx.x * WG_TILE, n_base = blockIdx.y * WG_TILE;
    const int Naug = Nrows + 2;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int vec_per_row = NP / 4;
    const int nvec = WG_TILE * vec_per_row;
    for (int tt = blockIdx.z; tt < ntiles; tt += gridDim.z) {
        const float trec = td ? rec_time(steps, t0, tt / Q) : 0.f;
        for (int e = tid; e < nvec; e += 256) {
            const int row = e / vec_per_row, k4 = e - row * vec_per_row;
            const int m = m_base + row;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < M) v = __ldg(reinterpret_cast<const float4*>(A + ((size_t)tt * M + m) * NP + k4 * 4));
            As[(k4 * 4 + 0) * LD + row] = v.x; As[(k4 * 4 + 1) * LD + row] = v.y;
            As[(k4 * 4 + 2) * LD + row] = v.z; As[(k4 * 4 + 3) * LD + row] = v.w;
            const int n = n_base + row;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < Nrows) w = __ldg(reinterpret_cast<const float4*>(Bm + ((size_t)tt * Nrows + n) * NP + k4 * 4));
            else if (n == Nrows) w = make_float4(trec, trec, trec, trec);
            else if (n == Nrows + 1) w = make_float4(1.f, 1.f, 1.f, 1.f);
            Bs[(k4 * 4 + 0) * LD + row] = w.x; Bs[(k4 * 4 + 1) * LD + row] = w.y;
            Bs[(k4 * 4 + 2) * LD + row] = w.z; Bs[(k4 * 4 + 3) * LD + row] = w.w;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < NP; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(As + k * LD + ty * 4);
            const float4 b4 = *reinterpret_cast<const float4*>(Bs + k * LD + tx * 4);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = rn_fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m_base + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n_base + tx * 4 + j;
            if (n < Naug) partial[((size_t)blockIdx.z * M + m) * Naug + n] = acc[i][j];
        }
    }
}

// fixed-order sum over splits; scatter into Flux.destructure layout:
//   outW[m + M*n] for n < Nrows (+ the time column n == Nrows when td), outb[m] from n == Nrows+1
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, int M, int Nrows, int td, float* __restrict__ outW,
                                    float* __restrict__ outb) {
    const int Naug = Nrows + 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * Naug) return;
    const int n = idx / M, m = idx - n * M;     // m fastest: coalesced writes of the column-major W
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += partial[((size_t)sp * M + m) * Naug + n];
    if (n < Nrows) outW[(size_t)M * n + m] = s;
    else if (n == Nrows) { if (td) outW[(size_t)M * Nrows + m] = s; }
    else outb[m] = s;
}

// dp layout: W1 (H x (D+td)), b1 (H), W2 (D x (H+td)), b2 (D)
static int launch_wgrad(int D, int H, int td, int nrec, int Q, int NP, int B, const float* tapeZ, const float* tapeD2, const float* tapeH,
                        const float* tapeD1, const StepRec* steps, float t0, float* ws, float* dp, cudaStream_t st, int64_t* launches) {
    (void)B;
    const int ntiles = nrec * Q;
    const int nsplit = ntiles < WG_SPLITS ? ntiles : WG_SPLITS;
    float* dW1 = dp;
    float* db1 = dW1 + (size_t)H * (D + td);
    float* dW2 = db1 + H;
    float* db2 = dW2 + (size_t)D * (H + td);
    {   // dW1aug = delta1 . [Z; t; 1]^T
        dim3 grid((H + WG_TILE - 1) / WG_TILE, (D + 2 + WG_TILE - 1) / WG_TILE, nsplit);
        wgrad_gemm_kernel<<<grid, 256, 0, st>>>(tapeD1, H, tapeZ, D, NP, ntiles, Q, steps, t0, td, ws);
        const int tot = H * (D + 2);
        wgrad_reduce_kernel<<<(tot + 255) / 256, 256, 0, st>>>(ws, nsplit, H, D, td, dW1, db1);
    }
    {   // dW2aug = delta2 . [Hact; t; 1]^T
        dim3 grid((D + WG_TILE - 1) / WG_TILE, (H + 2 + WG_TILE - 1) / WG_TILE, nsplit);
        wgrad_gemm_kernel<<<grid, 256, 0, st>>>(tapeD2, D, tapeH, H, NP, ntiles, Q, steps, t0, td, ws);
        const int tot = D * (H + 2);
        wgrad_reduce_kernel<<<(tot + 255) / 256, 256, 0, st>>>(ws, nsplit, D, H, td, dW2, db2);
    }
    if (launches) *launches += 4;
    return (int)cudaGetLastError();
}

}  // namespace rnde
