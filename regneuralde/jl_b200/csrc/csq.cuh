// csq.cuh -- the FFJORD vector field: MLPDynamics of three ConcatSquash layers with softplus
// between them, evaluated together with the transposed-Jacobian product e^T J for a fixed Hutchinson noise column
//   /root/reference/experiments/ffjord_tabular.jl:47-105 (ConcatSquashLinear, MLPDynamics, forw_n_back)
//   /root/reference/src/models/ffjord.jl:53-66           (_ffjord: rows [f; -sum(eJ .* e) (; ||f||^2; ||eJ||^2)])
// Runs inside the generic Tsit5 stepper (fwd_kernel<1,4|8,1,true,NT,FIELD=1>) and behind the test hook rnde_test_csq_rhs; bit-identical
// to the C oracle on a B200 (tests/test_gpu_ffjord.py).  Its reverse pass is csq_bwd.cuh.
// Canonical arithmetic = oracle/rnde_oracle.c csq_column: all six products through quad_dense (contraction index in four
// contiguous quarters, (q0+q1)+(q2+q3)); layer r = fma(W x + B, g, fma(bW, t, bB)), g = canon_sigmoidf(G*t); the transposed
// chain multiplies by the gate first; row sums are ascending fma chains from 0.
#pragma once
#include "common.cuh"
#include "chain.cuh"

namespace rnde {

struct CsqLayer { const float *W, *B, *bW, *bB, *G; int M, K; };

__host__ __device__ inline const float* csq_take(const float* p, int M, int K, CsqLayer& L) {
    L.M = M; L.K = K; L.W = p; L.B = p + (size_t)M * K; L.bW = L.B + M; L.bB = L.bW + M; L.G = L.bB + M;
    return L.G + M;
}
__host__ __device__ inline int csq_num_params(int Dz, int H) { return (H * Dz + 4 * H) + (H * H + 4 * H) + (Dz * H + 4 * Dz); }

// shared-memory tiles of one evaluation, [row][NP] each
struct CsqTiles {
    float* g;      // gates: H + H + Dz
    float* r1;     // H x NP   pre-activations of layer 1; overwritten in place by the gated cotangent g1 .* sigmoid(r1) .* v
    float* r2;     // H x NP   likewise for layer 2
    float* a;      // H x NP   softplus outputs (input of the next layer)
    float* u;      // Dz x NP  g3 .* e
    float* eJ;     // Dz x NP
};
__host__ __device__ inline int csq_tile_floats(int Dz, int H, int NP) {
    return round_up(2 * H + Dz, 4) + (3 * H + 2 * Dz) * NP;
}
__device__ inline CsqTiles csq_carve(float* s, int Dz, int H, int NP) {
    CsqTiles T;
    T.g = s; s += round_up(2 * H + Dz, 4);
    T.r1 = s; s += H * NP; T.r2 = s; s += H * NP; T.a = s; s += H * NP; T.u = s; s += Dz * NP; T.eJ = s;
    return T;
}

// sOut (D x NP, D = Dz + X) = field(sZ (first Dz rows used), t) for the noise tile sE (Dz x NP).  p: the parameters
// (global or shared memory), Flux.destructure order.  All threads of the block call it; ends with a barrier.
template <int NP, int NT>
__device__ __forceinline__ void csq_rhs(const float* __restrict__ p, const int Dz, const int H, const int X, const float t,
                                        const float* __restrict__ sZ, const float* __restrict__ sE, float* __restrict__ sOut, const CsqTiles T) {
    const int tid = threadIdx.x;
    CsqLayer L1, L2, L3;
    const float* q = csq_take(p, H, Dz, L1); q = csq_take(q, H, H, L2); csq_take(q, Dz, H, L3);
    float* g1 = T.g; float* g2 = T.g + H; float* g3 = T.g + 2 * H;
    for (int o = tid; o < 2 * H + Dz; o += NT) {
        const float G = o < H ? L1.G[o] : (o < 2 * H ? L2.G[o - H] : L3.G[o - 2 * H]);
        T.g[o] = canon_sigmoidf(G * t);
    }
    __syncthreads();
    // forward chain
    quad_dense<NP, NT, false>(L1.W, H, Dz, sZ, [&](const int o, const int n, const float s) {
        const float r = rn_fmaf(s + L1.B[o], g1[o], rn_fmaf(L1.bW[o], t, L1.bB[o]));
        T.r1[o * NP + n] = r; T.a[o * NP + n] = canon_softplusf(r);
    });
    __syncthreads();
    quad_dense<NP, NT, false>(L2.W, H, H, T.a, [&](const int o, const int n, const float s) {
        T.r2[o * NP + n] = rn_fmaf(s + L2.B[o], g2[o], rn_fmaf(L2.bW[o], t, L2.bB[o]));
    });
    __syncthreads();
    for (int e = tid; e < H * NP; e += NT) T.a[e] = canon_softplusf(T.r2[e]);
    for (int e = tid; e < Dz * NP; e += NT) T.u[e] = g3[e / NP] * sE[e];            // u3 = g3 .* e
    __syncthreads();
    quad_dense<NP, NT, false>(L3.W, Dz, H, T.a, [&](const int o, const int n, const float s) {
        sOut[o * NP + n] = rn_fmaf(s + L3.B[o], g3[o], rn_fmaf(L3.bW[o], t, L3.bB[o]));
    });
    // transposed chain: v = W^T u, then w = sigmoid(r) .* v and the next gate.  Element (k, n) of r is read and replaced by
    // the one lane that finalises (k, n), so the in-place update needs no barrier of its own (and the layer-3 product
    // above reads T.a only: the two products may overlap)
    quad_dense<NP, NT, true>(L3.W, Dz, H, T.u, [&](const int k, const int n, const float v) {
        const float w = canon_sigmoidf(T.r2[k * NP + n]) * v;
        T.r2[k * NP + n] = g2[k] * w;
    });
    __syncthreads();
    quad_dense<NP, NT, true>(L2.W, H, H, T.r2, [&](const int k, const int n, const float v) {
        const float w = canon_sigmoidf(T.r1[k * NP + n]) * v;
        T.r1[k * NP + n] = g1[k] * w;
    });
    __syncthreads();
    quad_dense<NP, NT, true>(L1.W, H, Dz, T.r1, [&](const int k, const int n, const float v) { T.eJ[k * NP + n] = v; });
    __syncthreads();
    // augmented rows: one thread per (row kind, column), ascending fma chains
    if (tid < 3 * NP) {
        const int kind = tid / NP, n = tid - kind * NP;
        if (kind == 0 || X == 3) {
            float s = 0.f;
            for (int i = 0; i < Dz; ++i) {
                const float a = kind == 1 ? sOut[i * NP + n] : T.eJ[i * NP + n];
                const float b = kind == 0 ? sE[i * NP + n] : a;
                s = rn_fmaf(a, b, s);
            }
            sOut[(Dz + kind) * NP + n] = kind == 0 ? -s : s;
        }
    }
    __syncthreads();
}

// Test hook kernel: one CTA per tile of 4 columns, parameters read from global memory.
constexpr int CSQ_TEST_NP = 4, CSQ_TEST_NT = 256;
__global__ void __launch_bounds__(CSQ_TEST_NT) csq_rhs_test_kernel(const float* __restrict__ p, int Dz, int H, int X, int B, float t,
                                                                   const float* __restrict__ z, const float* __restrict__ e, float* __restrict__ k) {
    constexpr int NP = CSQ_TEST_NP, NT = CSQ_TEST_NT;
    extern __shared__ __align__(16) float csq_smem[];
    const int D = Dz + X, tid = threadIdx.x, c0 = blockIdx.x * NP;
    float* sZ = csq_smem; float* sE = sZ + D * NP; float* sOut = sE + Dz * NP;
    const CsqTiles T = csq_carve(sOut + D * NP, Dz, H, NP);
    for (int idx = tid; idx < D * NP; idx += NT) { const int i = idx / NP, n = idx - i * NP; sZ[idx] = (c0 + n < B) ? z[(size_t)D * (c0 + n) + i] : 0.f; }
    for (int idx = tid; idx < Dz * NP; idx += NT) { const int i = idx / NP, n = idx - i * NP; sE[idx] = (c0 + n < B) ? e[(size_t)Dz * (c0 + n) + i] : 0.f; }
    __syncthreads();
    csq_rhs<NP, NT>(p, Dz, H, X, t, sZ, sE, sOut, T);
    for (int idx = tid; idx < D * NP; idx += NT) { const int i = idx / NP, n = idx - i * NP; if (c0 + n < B) k[(size_t)D * (c0 + n) + i] = sOut[idx]; }
}

}  // namespace rnde
