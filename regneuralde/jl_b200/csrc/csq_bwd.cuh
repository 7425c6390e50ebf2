// csq_bwd.cuh -- reverse mode through one evaluation of the FFJORD field of csq.cuh.
//   /root/reference/src/models/ffjord.jl:53-66 (_ffjord), experiments/ffjord_tabular.jl:47-105 (ConcatSquashLinear, forw_n_back):
//   the reference lets Tracker differentiate forw_n_back; here the evaluation graph
//       r1 -> a1 -> r2 -> a2 -> r3 = f;   v3 = W3^T (g3 .* e) -> w2 = sigma(r2) .* v3 -> v2 = W2^T (g2 .* w2) -> w1 = sigma(r1) .* v2
//       -> eJ = W1^T (g1 .* w1);   rows [f; -sum(eJ .* e) (; sum f^2; sum eJ^2)]
//   is reversed by hand (first-order reverse mode; the second-order character is in the graph: sigma' couples the two chains).
// Same math as the oracle (oracle/rnde_oracle_bwd.inc, VJP_FN with csq_extra > 0), which is checked against torch autograd
// (tests/test_ffjord_oracle.py).  The forward values are recomputed from the taped stage input; the products reuse quad_dense
// (chain.cuh).  Parameter gradients are NOT formed here: every record leaves the vectors of its six outer products and of its
// bias / gate sums on the tape (rows below), and dense_wgrad_kernel contracts them over (record, column) in Float64 afterwards.
#pragma once
#include "csq.cuh"

namespace rnde {

// rows of one (record, tile) block of the vector tape, in units of rows x NP
struct CsqTapeRows {
    int linb1, linb2, linb3;      // cotangents of W x + B per layer                  (H, H, Dz)
    int a1, a2;                   // inputs of layers 2, 3 (layer 1's input z is on the state tape)   (H, H)
    int u1, u2, u3;               // left factors of the transposed-chain outer products (H, H, Dz)
    int vb1, vb2, vb3;            // their right factors: cotangents of eJ, v2, v3    (Dz, H, H)
    int vec1, vec2, vec3;         // [rbar * t | rbar | gbar g (1 - g) t] per layer -> gradients of bias_W, bias_B, gate_W   (3H, 3H, 3Dz)
    int total;
};
__host__ __device__ inline CsqTapeRows csq_tape_rows(int Dz, int H) {
    CsqTapeRows R; int o = 0;
    R.linb1 = o; o += H; R.linb2 = o; o += H; R.linb3 = o; o += Dz;
    R.a1 = o; o += H; R.a2 = o; o += H;
    R.u1 = o; o += H; R.u2 = o; o += H; R.u3 = o; o += Dz;
    R.vb1 = o; o += Dz; R.vb2 = o; o += H; R.vb3 = o; o += H;
    R.vec1 = o; o += 3 * H; R.vec2 = o; o += 3 * H; R.vec3 = o; o += 3 * Dz;
    R.total = o;
    return R;
}

// shared-memory tiles of the reverse pass, [row][NP] each
struct CsqBwdTiles {
    float* g;                                  // gates: H + H + Dz
    float *lin1, *s1, *a1, *lin2, *s2, *a2;    // H rows each
    float *lin3, *r3, *u3, *eJ;                // Dz rows each
    float *v3, *v2, *u2, *u1;                  // H rows each
    float *rb1, *rb2, *vbH, *gbH;              // H rows each (vbH: v2bar / v3bar in turn, gbH: gate cotangent of the layer in hand)
    float *rb3, *vb1, *zb;                     // Dz rows each (zb: the result, D rows with the augmented rows zero)
};
__host__ __device__ inline int csq_bwd_tile_floats(int Dz, int H, int X, int NP) {
    return round_up(2 * H + Dz, 4) + (14 * H + 6 * Dz + (Dz + X)) * NP;
}
__device__ inline CsqBwdTiles csq_bwd_carve(float* s, int Dz, int H, int NP) {
    CsqBwdTiles T;
    T.g = s; s += round_up(2 * H + Dz, 4);
    float** hs[] = {&T.lin1, &T.s1, &T.a1, &T.lin2, &T.s2, &T.a2, &T.v3, &T.v2, &T.u2, &T.u1, &T.rb1, &T.rb2, &T.vbH, &T.gbH};
    for (int i = 0; i < 14; ++i) { *hs[i] = s; s += H * NP; }
    float** ds[] = {&T.lin3, &T.r3, &T.u3, &T.eJ, &T.rb3, &T.vb1};
    for (int i = 0; i < 6; ++i) { *ds[i] = s; s += Dz * NP; }
    T.zb = s;
    return T;
}

// VJP of one evaluation.  sZ: the stage input (first Dz rows used), sE: noise tile, sKbar: cotangent of the D = Dz + X output rows.
// tape: this (record, tile) block of the vector tape (global memory, rows x NP).  Returns the input cotangent (D x NP, shared
// memory).  All threads of the block call it; ends with a barrier.
template <int NP, int NT>
__device__ __forceinline__ const float* csq_vjp(const float* __restrict__ p, const int Dz, const int H, const int X, const float t,
                                                const float* __restrict__ sZ, const float* __restrict__ sE, const float* __restrict__ sKbar,
                                                float* __restrict__ tape, const CsqBwdTiles T) {
    const int tid = threadIdx.x;
    CsqLayer L1, L2, L3;
    const float* q = csq_take(p, H, Dz, L1); q = csq_take(q, H, H, L2); csq_take(q, Dz, H, L3);
    const CsqTapeRows R = csq_tape_rows(Dz, H);
    float* g1 = T.g; float* g2 = T.g + H; float* g3 = T.g + 2 * H;
    for (int o = tid; o < 2 * H + Dz; o += NT) {
        const float G = o < H ? L1.G[o] : (o < 2 * H ? L2.G[o - H] : L3.G[o - 2 * H]);
        T.g[o] = canon_sigmoidf(G * t);
    }
    __syncthreads();
    // ---- forward values (the arithmetic of csq_rhs) ----
    quad_dense<NP, NT, false>(L1.W, H, Dz, sZ, [&](const int o, const int n, const float s) {
        const float lin = s + L1.B[o];
        const float r = rn_fmaf(lin, g1[o], rn_fmaf(L1.bW[o], t, L1.bB[o]));
        T.lin1[o * NP + n] = lin; T.s1[o * NP + n] = canon_sigmoidf(r);
        const float a = canon_softplusf(r);
        T.a1[o * NP + n] = a; tape[(R.a1 + o) * NP + n] = a;
    });
    __syncthreads();
    quad_dense<NP, NT, false>(L2.W, H, H, T.a1, [&](const int o, const int n, const float s) {
        const float lin = s + L2.B[o];
        const float r = rn_fmaf(lin, g2[o], rn_fmaf(L2.bW[o], t, L2.bB[o]));
        T.lin2[o * NP + n] = lin; T.s2[o * NP + n] = canon_sigmoidf(r);
        const float a = canon_softplusf(r);
        T.a2[o * NP + n] = a; tape[(R.a2 + o) * NP + n] = a;
    });
    for (int e = tid; e < Dz * NP; e += NT) { const float u = g3[e / NP] * sE[e]; T.u3[e] = u; tape[R.u3 * NP + e] = u; }
    __syncthreads();
    quad_dense<NP, NT, false>(L3.W, Dz, H, T.a2, [&](const int o, const int n, const float s) {
        const float lin = s + L3.B[o];
        T.lin3[o * NP + n] = lin; T.r3[o * NP + n] = rn_fmaf(lin, g3[o], rn_fmaf(L3.bW[o], t, L3.bB[o]));
    });
    quad_dense<NP, NT, true>(L3.W, Dz, H, T.u3, [&](const int k, const int n, const float v) {
        T.v3[k * NP + n] = v;
        const float u = g2[k] * (T.s2[k * NP + n] * v);
        T.u2[k * NP + n] = u; tape[(R.u2 + k) * NP + n] = u;
    });
    __syncthreads();
    quad_dense<NP, NT, true>(L2.W, H, H, T.u2, [&](const int k, const int n, const float v) {
        T.v2[k * NP + n] = v;
        const float u = g1[k] * (T.s1[k * NP + n] * v);
        T.u1[k * NP + n] = u; tape[(R.u1 + k) * NP + n] = u;
    });
    __syncthreads();
    quad_dense<NP, NT, true>(L1.W, H, Dz, T.u1, [&](const int k, const int n, const float v) { T.eJ[k * NP + n] = v; });
    __syncthreads();
    // ---- cotangents of the rows [f; -sum(eJ .* e); sum f^2; sum eJ^2] ----
    for (int e = tid; e < Dz * NP; e += NT) {
        const int n = e % NP;
        const float kl = sKbar[Dz * NP + n];
        float rb = sKbar[e], vb = -kl * sE[e];
        if (X == 3) { rb = rn_fmaf(2.f * sKbar[(Dz + 1) * NP + n], T.r3[e], rb); vb = rn_fmaf(2.f * sKbar[(Dz + 2) * NP + n], T.eJ[e], vb); }
        T.rb3[e] = rb; T.vb1[e] = vb; tape[R.vb1 * NP + e] = vb;
    }
    __syncthreads();
    // ---- the transposed chain in reverse: v = W^T u  =>  dW += u vbar^T (outer product, later), ubar = W vbar ----
    quad_dense<NP, NT, false>(L1.W, H, Dz, T.vb1, [&](const int m, const int n, const float ub) {      // eJ = W1^T u1, u1 = g1 .* s1 .* v2
        const int e = m * NP + n;
        const float s1 = T.s1[e], v2 = T.v2[e];
        T.gbH[e] = ub * (s1 * v2);                     // gate cotangent, layer 1 (completed in the forward-chain pass)
        const float wb = ub * g1[m];
        T.rb1[e] = wb * v2 * s1 * (1.f - s1);
        const float vb = wb * s1;                      // v2 bar
        T.vbH[e] = vb; tape[(R.vb2 + m) * NP + n] = vb;
    });
    __syncthreads();
    // gb1 is needed again below; keep it where lin-free space exists: reuse T.u1 (its tape copy is written)
    for (int e = tid; e < H * NP; e += NT) T.u1[e] = T.gbH[e];
    __syncthreads();
    quad_dense<NP, NT, false>(L2.W, H, H, T.vbH, [&](const int m, const int n, const float ub) {       // v2 = W2^T u2, u2 = g2 .* s2 .* v3
        const int e = m * NP + n;
        const float s2 = T.s2[e], v3 = T.v3[e];
        T.gbH[e] = ub * (s2 * v3);                     // gate cotangent, layer 2
        const float wb = ub * g2[m];
        T.rb2[e] = wb * v3 * s2 * (1.f - s2);
        T.v2[e] = wb * s2;                             // v3 bar (v2 itself is no longer needed)
    });
    __syncthreads();
    for (int e = tid; e < H * NP; e += NT) { T.u2[e] = T.gbH[e]; tape[R.vb3 * NP + e] = T.v2[e]; }      // gb2 parked in T.u2
    __syncthreads();
    quad_dense<NP, NT, false>(L3.W, Dz, H, T.v2, [&](const int m, const int n, const float ub) {        // v3 = W3^T u3, u3 = g3 .* e
        T.eJ[m * NP + n] = ub * sE[m * NP + n];        // gate cotangent, layer 3 (eJ itself is no longer needed)
    });
    __syncthreads();
    // ---- the forward chain in reverse: r = lin .* g + bW t + bB, lin = W x + B ----
    for (int e = tid; e < Dz * NP; e += NT) {
        const int o = e / NP;
        const float rb = T.rb3[e], g = g3[o];
        const float lb = rb * g;
        const float gb = rn_fmaf(rb, T.lin3[e], T.eJ[e]);
        T.rb3[e] = lb;                                 // linb3
        tape[R.linb3 * NP + e] = lb;
        tape[(R.vec3 + o) * NP + e % NP] = rb * t; tape[(R.vec3 + Dz + o) * NP + e % NP] = rb; tape[(R.vec3 + 2 * Dz + o) * NP + e % NP] = gb * g * (1.f - g) * t;
    }
    __syncthreads();
    quad_dense<NP, NT, true>(L3.W, Dz, H, T.rb3, [&](const int k, const int n, const float xb) {       // a2 = softplus(r2)
        T.rb2[k * NP + n] = rn_fmaf(xb, T.s2[k * NP + n], T.rb2[k * NP + n]);
    });
    __syncthreads();
    for (int e = tid; e < H * NP; e += NT) {
        const int o = e / NP, n = e % NP;
        const float rb = T.rb2[e], g = g2[o];
        const float lb = rb * g;
        const float gb = rn_fmaf(rb, T.lin2[e], T.u2[e]);
        T.rb2[e] = lb;                                 // linb2
        tape[(R.linb2 + o) * NP + n] = lb;
        tape[(R.vec2 + o) * NP + n] = rb * t; tape[(R.vec2 + H + o) * NP + n] = rb; tape[(R.vec2 + 2 * H + o) * NP + n] = gb * g * (1.f - g) * t;
    }
    __syncthreads();
    quad_dense<NP, NT, true>(L2.W, H, H, T.rb2, [&](const int k, const int n, const float xb) {        // a1 = softplus(r1)
        T.rb1[k * NP + n] = rn_fmaf(xb, T.s1[k * NP + n], T.rb1[k * NP + n]);
    });
    __syncthreads();
    for (int e = tid; e < H * NP; e += NT) {
        const int o = e / NP, n = e % NP;
        const float rb = T.rb1[e], g = g1[o];
        const float lb = rb * g;
        const float gb = rn_fmaf(rb, T.lin1[e], T.u1[e]);
        T.rb1[e] = lb;                                 // linb1
        tape[(R.linb1 + o) * NP + n] = lb;
        tape[(R.vec1 + o) * NP + n] = rb * t; tape[(R.vec1 + H + o) * NP + n] = rb; tape[(R.vec1 + 2 * H + o) * NP + n] = gb * g * (1.f - g) * t;
    }
    for (int e = tid; e < X * NP; e += NT) T.zb[Dz * NP + e] = 0.f;      // the field does not read the augmented rows
    __syncthreads();
    quad_dense<NP, NT, true>(L1.W, H, Dz, T.rb1, [&](const int k, const int n, const float xb) { T.zb[k * NP + n] = xb; });
    __syncthreads();
    return T.zb;
}

}  // namespace rnde
