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

// Up to two small dense products side by side (e.g. the hidden layers of the update and reset gates), each split
// over lane quads like chain.cuh quad_dense: item = (quarter of the contraction index, row), the four quarters of a row
// sit in adjacent lanes and are added by two xor-shuffles; lane q of the quad finalises column q (GRU_NP == 4).
//   TRANS = false: value[a][n] = sum_c W[M*c + a] * in[c][n]  (a < M, c < K);   TRANS = true: sum_c W[M*a + c] * in[c][n]  (a < K, c < M)
struct GruJob { const float* W; const float* in; int M, K; };
template <bool TRANS, class Fin>
__device__ __forceinline__ void gru_quad(const GruJob j0, const GruJob j1, const int njobs, Fin fin) {
    const int A0 = TRANS ? j0.K : j0.M, A1 = njobs > 1 ? (TRANS ? j1.K : j1.M) : 0;
    const int total = 4 * (A0 + A1);
    for (int base = 0; base < total; base += GRU_NT) {
        const int item = base + (int)threadIdx.x;
        const bool valid = item < total;
        const int blk = item & 3, rest = item >> 2;
        const int job = (valid && rest >= A0) ? 1 : 0;
        const GruJob& J = job ? j1 : j0;
        const int a = valid ? rest - (job ? A0 : 0) : 0;
        const int Cn = TRANS ? J.M : J.K, kb = (Cn + 3) >> 2;
        const int c0 = blk * kb, c1 = valid ? min(c0 + kb, Cn) : c0;
        float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
        const float* wp = TRANS ? J.W + J.M * a + c0 : J.W + J.M * c0 + a;
        const int wstep = TRANS ? 1 : J.M;
        const float* ip = J.in + c0 * GRU_NP;
#pragma unroll 4
        for (int c = c0; c < c1; ++c) {
            const float w = *wp; wp += wstep;
            const float4 x = *reinterpret_cast<const float4*>(ip); ip += GRU_NP;
            acc0 = fmaf(w, x.x, acc0); acc1 = fmaf(w, x.y, acc1); acc2 = fmaf(w, x.z, acc2); acc3 = fmaf(w, x.w, acc3);
        }
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 1); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 1);
        acc2 += __shfl_xor_sync(0xffffffffu, acc2, 1); acc3 += __shfl_xor_sync(0xffffffffu, acc3, 1);
        acc0 += __shfl_xor_sync(0xffffffffu, acc0, 2); acc1 += __shfl_xor_sync(0xffffffffu, acc1, 2);
        acc2 += __shfl_xor_sync(0xffffffffu, acc2, 2); acc3 += __shfl_xor_sync(0xffffffffu, acc3, 2);
        const float v = blk == 0 ? acc0 : (blk == 1 ? acc1 : (blk == 2 ? acc2 : acc3));
        if (valid) fin(job, a, blk, v);
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
        // hidden layers of the two gates side by side, then their sigmoid outputs
        gru_quad<false>(GruJob{sW + O.Wu1, sYc, H, C}, GruJob{sW + O.Wr1, sYc, H, C}, 2, [&](int job, int o, int n, float v) {
            (job ? sHr : sHu)[o * GRU_NP + n] = canon_tanhf(v + sW[(job ? O.br1 : O.bu1) + o]);
        });
        __syncthreads();
        gru_quad<false>(GruJob{sW + O.Wu2, sHu, L, H}, GruJob{sW + O.Wr2, sHr, L, H}, 2, [&](int job, int o, int n, float v) {
            (job ? sR : sU)[o * GRU_NP + n] = gru_sigmoid(v + sW[(job ? O.br2 : O.bu2) + o]);
        });
        __syncthreads();
        for (int e = tid; e < 2 * L * GRU_NP; e += GRU_NT) {
            const int j = e / GRU_NP, n = e - j * GRU_NP;
            sCc[e] = sYc[e] * sR[(j % L) * GRU_NP + n];
        }
        __syncthreads();
        gru_quad<false>(GruJob{sW + O.Wn1, sCc, H, C}, GruJob{nullptr, nullptr, 0, 0}, 1, [&](int, int o, int n, float v) {
            sHn[o * GRU_NP + n] = canon_tanhf(v + sW[O.bn1 + o]);
        });
        __syncthreads();
        gru_quad<false>(GruJob{sW + O.Wn2, sHn, 2 * L, H}, GruJob{nullptr, nullptr, 0, 0}, 1, [&](int, int o, int n, float v) {
            sNs[o * GRU_NP + n] = v + sW[O.bn2 + o];
        });
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
        gru_quad<true>(GruJob{sW + O.Wn2, sDns, 2 * L, H}, GruJob{nullptr, nullptr, 0, 0}, 1, [&](int, int i, int n, float v) {
            const float hv = A[(O.a_hn + i) * GRU_NP + n];
            const float d = v * (1.f - hv * hv);
            sDhn[i * GRU_NP + n] = d; Dl[(O.d_hn + i) * GRU_NP + n] = d;
        });
        __syncthreads();
        // only the first 2L inputs of the new-state network carry a cotangent (the rest is data): K is cut to 2L rows
        gru_quad<true>(GruJob{sW + O.Wn1, sDhn, H, 2 * L}, GruJob{nullptr, nullptr, 0, 0}, 1, [&](int, int j, int n, float v) { sCb[j * GRU_NP + n] = v; });
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
        gru_quad<true>(GruJob{sW + O.Wu2, sDu, L, H}, GruJob{sW + O.Wr2, sDr, L, H}, 2, [&](int job, int i, int n, float v) {
            const float hv = A[((job ? O.a_hr : O.a_hu) + i) * GRU_NP + n];
            const float d = v * (1.f - hv * hv);
            (job ? sDhr : sDhu)[i * GRU_NP + n] = d;
            Dl[((job ? O.d_hr : O.d_hu) + i) * GRU_NP + n] = d;
        });
        __syncthreads();
        gru_quad<true>(GruJob{sW + O.Wr1, sDhr, H, 2 * L}, GruJob{nullptr, nullptr, 0, 0}, 1, [&](int, int j, int n, float v) { sYp[j * GRU_NP + n] += v; });
        __syncthreads();
        gru_quad<true>(GruJob{sW + O.Wu1, sDhu, H, 2 * L}, GruJob{nullptr, nullptr, 0, 0}, 1, [&](int, int j, int n, float v) { sYp[j * GRU_NP + n] += v; });
        __syncthreads();
        { float* tmp = sYb; sYb = sYp; sYp = tmp; }
    }
    (void)C;
}

}  // namespace rnde
