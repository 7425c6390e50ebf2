// gru_kernel.cuh -- the Latent-ODE recognition RNN as one persistent kernel per direction.
//
// Replaces (p::LatentGRU)(x) -- 49 sequential single_run calls, each 6 CUBLAS GEMMs + ~20 broadcast kernels in the
// reference (/root/reference/experiments/latent_ode.jl:39-99) -- by ONE launch: all 29 320 weights of the three
// two-layer gate networks live in shared memory (117 KB), a CTA owns a tile of 4 batch columns and walks the
// sequence backwards in time with the recurrent state in shared memory.  The backward kernel is the exact reverse
// sweep (BPTT) over a tape of the per-step activations; it leaves the per-layer deltas on a second tape, and the six
// weight gradients are contractions over (time, column) done by dense_wgrad_kernel (chain.cuh).
//
//   y_concat = [y_mean; y_std; x_t]                        C = 2L + X rows,  X = 2*in_dim + 1
//   u = sigmoid(Wu2 tanh(Wu1 y_concat + bu1) + bu2)        update gate     (C -> H -> L)
//   r = sigmoid(Wr2 tanh(Wr1 y_concat + br1) + br2)        reset gate
//   n = Wn2 tanh(Wn1 [y_mean.*r; y_std.*r; x_t] + bn1) + bn2               (C -> H -> 2L)
//   y' = (1-u).*n + u.*y   where any of x_t[X÷2+1:end] sums > 0, else y
#pragma once
#include "common.cuh"
#include "chain.cuh"

namespace rnde {

struct GruParams {
    int I, H, L, X, C, T, B, Q;        // Q = column tiles
    const float* x;                    // X x T x B, column-major
    const float* p;
    float* out;                        // 2L x B
    float* tapeA; float* tapeD;        // [step][tile][rows][NP]
    int arows, drows, need_tape;
    const float* dout; float* dx;      // backward
};

struct GruOffsets {
    // parameter offsets
    int Wu1, bu1, Wu2, bu2, Wr1, br1, Wr2, br2, Wn1, bn1, Wn2, bn2, np;
    // activation-tape rows
    int a_yc, a_hu, a_hr, a_u, a_r, a_cc, a_hn, a_ns, a_mask, arows;
    // delta-tape rows
    int d_hu, d_hr, d_u, d_r, d_hn, d_ns, drows;
};

__host__ __device__ inline GruOffsets gru_offsets(int I, int H, int L) {
    GruOffsets o;
    const int C = 2 * L + 2 * I + 1;
    int q = 0;
    o.Wu1 = q; q += H * C; o.bu1 = q; q += H; o.Wu2 = q; q += L * H; o.bu2 = q; q += L;
    o.Wr1 = q; q += H * C; o.br1 = q; q += H; o.Wr2 = q; q += L * H; o.br2 = q; q += L;
    o.Wn1 = q; q += H * C; o.bn1 = q; q += H; o.Wn2 = q; q += 2 * L * H; o.bn2 = q; q += 2 * L;
    o.np = q;
    int a = 0;
    o.a_yc = a; a += C; o.a_hu = a; a += H; o.a_hr = a; a += H; o.a_u = a; a += L; o.a_r = a; a += L;
    o.a_cc = a; a += C; o.a_hn = a; a += H; o.a_ns = a; a += 2 * L; o.a_mask = a; a += 1;
    o.arows = a;
    int d = 0;
    o.d_hu = d; d += H; o.d_hr = d; d += H; o.d_u = d; d += L; o.d_r = d; d += L; o.d_hn = d; d += H; o.d_ns = d; d += 2 * L;
    o.drows = d;
    return o;
}

constexpr int GRU_NT = 256, GRU_NP = 4;

__device__ __forceinline__ float gru_sigmoid(const float v) { return rn_fmaf(0.5f, canon_tanhf(0.5f * v), 0.5f); }

// out[o][n] = act(sum_i W[o,i] in[i][n] + b[o]) for o < M; W column-major M x K in shared memory
template <class Act>
__device__ __forceinline__ void gru_dense(const float* W, const float* b, const float* in, int K, int M, float* out, int e_first, int e_count, Act act) {
    for (int ee = (int)threadIdx.x - e_first; ee < M * GRU_NP; ee += e_count) {
        if (ee < 0) continue;
        const int o = ee / GRU_NP, n = ee - o * GRU_NP;
        float acc = 0.f;
#pragma unroll 5
        for (int i = 0; i < K; ++i) acc = fmaf(W[M * i + o], in[i * GRU_NP + n], acc);
        out[ee] = act(acc + b[o]);
    }
}

__global__ void __launch_bounds__(GRU_NT, 1) gru_fwd_kernel(const GruParams P) {
    extern __shared__ __align__(16) float gsm[];
    const int tid = threadIdx.x, q = blockIdx.x;
    const int H = P.H, L = P.L, X = P.X, C = P.C;
    const GruOffsets O = gru_offsets(P.I, H, L);
    float* sW = gsm;
    float* sYc = sW + round_up(O.np, 4);          // C x NP: y_mean, y_std, x_t
    float* sCc = sYc + C * GRU_NP;                // C x NP: y_mean.*r, y_std.*r, x_t
    float* sHu = sCc + C * GRU_NP;                // H
    float* sHr = sHu + H * GRU_NP;
    float* sHn = sHr + H * GRU_NP;
    float* sU = sHn + H * GRU_NP;                 // L
    float* sR = sU + L * GRU_NP;
    float* sNs = sR + L * GRU_NP;                 // 2L
    float* sMask = sNs + 2 * L * GRU_NP;          // NP
    const int c0 = q * GRU_NP;
    for (int e = tid; e < O.np; e += GRU_NT) sW[e] = __ldg(P.p + e);
    for (int e = tid; e < 2 * L * GRU_NP; e += GRU_NT) sYc[e] = 0.f;
    __syncthreads();
    auto tanh_act = [](float v) { return canon_tanhf(v); };
    auto sig_act = [](float v) { return gru_sigmoid(v); };
    auto id_act = [](float v) { return v; };
    for (int step = 0; step < P.T; ++step) {
        const int t = P.T - 1 - step;            // for t = size(x, 2):-1:1
        for (int e = tid; e < X * GRU_NP; e += GRU_NT) {
            const int n = e / X, f = e - n * X;   // f fastest: coalesced
            const float v = (c0 + n < P.B) ? __ldg(P.x + (size_t)X * (t + (size_t)P.T * (c0 + n)) + f) : 0.f;
            sYc[(2 * L + f) * GRU_NP + n] = v;
            sCc[(2 * L + f) * GRU_NP + n] = v;
        }
        __syncthreads();
        if (tid < GRU_NP) {                       // mask = sum(x[(size(x,1) ÷ 2 + 1):end, :]) > 0
            float s = 0.f;
            for (int f = X / 2; f < X; ++f) s += sYc[(2 * L + f) * GRU_NP + tid];
            sMask[tid] = s > 0.f ? 1.f : 0.f;
        }
        // hidden layers of the two gates (thread ranges side by side)
        gru_dense(sW + O.Wu1, sW + O.bu1, sYc, C, H, sHu, 0, GRU_NT, tanh_act);
        gru_dense(sW + O.Wr1, sW + O.br1, sYc, C, H, sHr, (H * GRU_NP) % GRU_NT, GRU_NT, tanh_act);
        __syncthreads();
        gru_dense(sW + O.Wu2, sW + O.bu2, sHu, H, L, sU, 0, GRU_NT, sig_act);
        gru_dense(sW + O.Wr2, sW + O.br2, sHr, H, L, sR, (L * GRU_NP) % GRU_NT, GRU_NT, sig_act);
        __syncthreads();
        for (int e = tid; e < 2 * L * GRU_NP; e += GRU_NT) {
            const int j = e / GRU_NP, n = e - j * GRU_NP;
            sCc[e] = sYc[e] * sR[(j % L) * GRU_NP + n];
        }
        __syncthreads();
        gru_dense(sW + O.Wn1, sW + O.bn1, sCc, C, H, sHn, 0, GRU_NT, tanh_act);
        __syncthreads();
        gru_dense(sW + O.Wn2, sW + O.bn2, sHn, H, 2 * L, sNs, 0, GRU_NT, id_act);
        __syncthreads();
        if (P.need_tape) {
            float* A = P.tapeA + ((size_t)step * P.Q + q) * P.arows * GRU_NP;
            for (int e = tid; e < C * GRU_NP; e += GRU_NT) { A[O.a_yc * GRU_NP + e] = sYc[e]; A[O.a_cc * GRU_NP + e] = sCc[e]; }
            for (int e = tid; e < H * GRU_NP; e += GRU_NT) { A[O.a_hu * GRU_NP + e] = sHu[e]; A[O.a_hr * GRU_NP + e] = sHr[e]; A[O.a_hn * GRU_NP + e] = sHn[e]; }
            for (int e = tid; e < L * GRU_NP; e += GRU_NT) { A[O.a_u * GRU_NP + e] = sU[e]; A[O.a_r * GRU_NP + e] = sR[e]; }
            for (int e = tid; e < 2 * L * GRU_NP; e += GRU_NT) A[O.a_ns * GRU_NP + e] = sNs[e];
            if (tid < GRU_NP) A[O.a_mask * GRU_NP + tid] = sMask[tid];
        }
        for (int e = tid; e < 2 * L * GRU_NP; e += GRU_NT) {
            const int j = e / GRU_NP, n = e - j * GRU_NP;
            const float u = sU[(j % L) * GRU_NP + n], y = sYc[e];
            const float ny = (1.f - u) * sNs[e] + u * y;
            const float m = sMask[n];
            sYc[e] = m * ny + (1.f - m) * y;
        }
        __syncthreads();
    }
    for (int e = tid; e < 2 * L * GRU_NP; e += GRU_NT) {
        const int n = e / (2 * L), j = e - n * (2 * L);
        if (c0 + n < P.B) P.out[(size_t)2 * L * (c0 + n) + j] = sYc[j * GRU_NP + n];
    }
}

// in-place accumulate: dst[i][n] += sum_o W[o,i] g[o][n] for i < Kuse (W column-major M x K)
__device__ __forceinline__ void gru_dense_T(const float* W, const float* g, int M, int Kuse, float* dst, bool accumulate) {
    for (int e = threadIdx.x; e < Kuse * GRU_NP; e += GRU_NT) {
        const int i = e / GRU_NP, n = e - i * GRU_NP;
        float acc = 0.f;
#pragma unroll 5
        for (int o = 0; o < M; ++o) acc = fmaf(W[M * i + o], g[o * GRU_NP + n], acc);
        dst[e] = accumulate ? dst[e] + acc : acc;
    }
}

__global__ void __launch_bounds__(GRU_NT, 1) gru_bwd_kernel(const GruParams P) {
    extern __shared__ __align__(16) float gsm[];
    const int tid = threadIdx.x, q = blockIdx.x;
    const int H = P.H, L = P.L, C = P.C;
    const GruOffsets O = gru_offsets(P.I, H, L);
    float* sW = gsm;
    float* sYb = sW + round_up(O.np, 4);          // 2L x NP: cotangent of the state after this step
    float* sYp = sYb + 2 * L * GRU_NP;            // 2L x NP: cotangent of the state before this step
    float* sDns = sYp + 2 * L * GRU_NP;           // 2L
    float* sDu = sDns + 2 * L * GRU_NP;           // L
    float* sDr = sDu + L * GRU_NP;                // L
    float* sDhn = sDr + L * GRU_NP;               // H
    float* sDhu = sDhn + H * GRU_NP;
    float* sDhr = sDhu + H * GRU_NP;
    float* sCb = sDhr + H * GRU_NP;               // 2L: cotangent of [y_mean.*r; y_std.*r]
    const int c0 = q * GRU_NP;
    for (int e = tid; e < O.np; e += GRU_NT) sW[e] = __ldg(P.p + e);
    for (int e = tid; e < 2 * L * GRU_NP; e += GRU_NT) {
        const int n = e / (2 * L), j = e - n * (2 * L);
        sYb[j * GRU_NP + n] = (c0 + n < P.B) ? __ldg(P.dout + (size_t)2 * L * (c0 + n) + j) : 0.f;
    }
    __syncthreads();
    for (int step = P.T - 1; step >= 0; --step) {
        const float* A = P.tapeA + ((size_t)step * P.Q + q) * P.arows * GRU_NP;
        float* Dl = P.tapeD + ((size_t)step * P.Q + q) * P.drows * GRU_NP;
        // gates and the convex combination
        for (int e = tid; e < L * GRU_NP; e += GRU_NT) {
            const int n = e % GRU_NP;
            const float m = A[O.a_mask * GRU_NP + n];
            const float u = A[O.a_u * GRU_NP + e];
            const float ym = A[O.a_yc * GRU_NP + e], ys = A[(O.a_yc + L) * GRU_NP + e];
            const float nm = A[O.a_ns * GRU_NP + e], nsd = A[(O.a_ns + L) * GRU_NP + e];
            const float gm = sYb[e] * m, gs = sYb[L * GRU_NP + e] * m;      // cotangents of new_y_mean / new_y_std
            sDns[e] = (1.f - u) * gm;
            sDns[L * GRU_NP + e] = (1.f - u) * gs;
            const float ub = (ym - nm) * gm + (ys - nsd) * gs;
            sDu[e] = ub * u * (1.f - u);
            sYp[e] = sYb[e] * (1.f - m) + u * gm;
            sYp[L * GRU_NP + e] = sYb[L * GRU_NP + e] * (1.f - m) + u * gs;
        }
        __syncthreads();
        for (int e = tid; e < 2 * L * GRU_NP; e += GRU_NT) Dl[O.d_ns * GRU_NP + e] = sDns[e];
        for (int e = tid; e < L * GRU_NP; e += GRU_NT) Dl[O.d_u * GRU_NP + e] = sDu[e];
        // new_state network
        gru_dense_T(sW + O.Wn2, sDns, 2 * L, H, sDhn, false);
        __syncthreads();
        for (int e = tid; e < H * GRU_NP; e += GRU_NT) { const float hv = A[O.a_hn * GRU_NP + e]; const float d = sDhn[e] * (1.f - hv * hv); sDhn[e] = d; Dl[O.d_hn * GRU_NP + e] = d; }
        __syncthreads();
        gru_dense_T(sW + O.Wn1, sDhn, H, 2 * L, sCb, false);
        __syncthreads();
        for (int e = tid; e < L * GRU_NP; e += GRU_NT) {
            const float r = A[O.a_r * GRU_NP + e];
            const float ym = A[O.a_yc * GRU_NP + e], ys = A[(O.a_yc + L) * GRU_NP + e];
            const float cm = sCb[e], cs = sCb[L * GRU_NP + e];
            sYp[e] += cm * r;
            sYp[L * GRU_NP + e] += cs * r;
            const float d = (cm * ym + cs * ys) * r * (1.f - r);
            sDr[e] = d; Dl[O.d_r * GRU_NP + e] = d;
        }
        __syncthreads();
        // gate networks
        gru_dense_T(sW + O.Wr2, sDr, L, H, sDhr, false);
        gru_dense_T(sW + O.Wu2, sDu, L, H, sDhu, false);
        __syncthreads();
        for (int e = tid; e < H * GRU_NP; e += GRU_NT) {
            const float hr = A[O.a_hr * GRU_NP + e], hu = A[O.a_hu * GRU_NP + e];
            const float dr = sDhr[e] * (1.f - hr * hr), du = sDhu[e] * (1.f - hu * hu);
            sDhr[e] = dr; sDhu[e] = du;
            Dl[O.d_hr * GRU_NP + e] = dr; Dl[O.d_hu * GRU_NP + e] = du;
        }
        __syncthreads();
        gru_dense_T(sW + O.Wr1, sDhr, H, 2 * L, sYp, true);
        __syncthreads();
        gru_dense_T(sW + O.Wu1, sDhu, H, 2 * L, sYp, true);
        __syncthreads();
        { float* tmp = sYb; sYb = sYp; sYp = tmp; }
    }
    (void)C;
}

}  // namespace rnde
