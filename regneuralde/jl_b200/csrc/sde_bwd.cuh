// sde_bwd.cuh -- reverse sweep of the Neural-SDE solve: the discrete adjoint of the accepted
// SOSRI / SOSRI2 steps with the step sizes and the Wiener increments frozen -- what Tracker.gradient differentiates through
// solve(SDEProblem, SOSRI(); sensealg = SensitivityADPassThrough()) (src/models/neural_sde.jl:84-146,
// experiments/mnist_nsde.jl:201-204): the proposed dt are detached as in the ODE path, the increments come from the untracked RNG.
// The forward kernel tapes per accepted step the state at its start and the increments (dW, dZ) of its tile plus (dt, EEst);
// the sweep recomputes the four-stage Roessler step from them (the arithmetic of sde_kernel.cuh) and reverses it: cotangents of
// the new state and of the saved value EEst * dt flow through the error estimate (delta E1 + E2) / (atol + max(|u|, |u'|) rtol),
// the update formula, and the eight field evaluations in reverse stage order.
// Decomposition as in the forward kernel: one CTA per tile of NP columns, all 5 248 parameters AND their gradient accumulators
// in shared memory (rank-NP updates per VJP); the per-CTA gradients are added over the tiles in a fixed order by
// sde_grad_reduce_kernel (Float64).  No grid-wide dependency: EEst comes from the tape.
// Checked against oracle/sde_oracle.py replay_torch (torch autograd through the same steps, Float64).
#pragma once
#include "sde_kernel.cuh"

namespace rnde {

struct SdeBwdParams {
    int D, H, B, Q, alg, reg_kind, tape_cap;
    float abstol, reltol;
    const float* p; const float* tape; const float* tape_steps;      // tape: [step][Q][3][D * NP] (u at step start, dW, dZ); steps: [step][4] (dt, EEst, rms(k4 - k3), rms(H03 - H02))
    const SdeStats* stats;                                           // naccept of the forward solve, read on the device
    const float* du; const float* dsaveval;                          // cotangents of the final state (D x B) and of the saved values
    float* dx; float* gpart;                                         // dx: D x B; gpart: [Q][np] per-tile parameter gradients
};

__host__ __device__ inline int sde_bwd_smem_floats(int D, int H, int np, int NP) {
    return 2 * ((np + 3) / 4 * 4) + 30 * D * NP + 5 * H * NP + 16;
}

template <int NP>
__global__ void __launch_bounds__(SDE_NT) sde_bwd_kernel(const SdeBwdParams P) {
    constexpr int NT = SDE_NT;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int D = P.D, H = P.H, T = D * NP, HT = H * NP;
    const int q = blockIdx.x, c0 = q * NP;
    const int Nloc = max(0, min(NP, P.B - c0));
    const int np = H * D + H + D * H + D + D * D + D, npp = (np + 3) / 4 * 4;
    float* sP = smem; float* sGr = smem + npp;
    float* base = sGr + npp;
    float* sU = base; float* sUn = sU + T; float* sdW = sUn + T; float* sdZ = sdW + T;
    float* sK[4]; float* sG[4]; float* sH0[3]; float* sH1[3]; float* sKb[4]; float* sGb[4]; float* sHid[4];
    float* cur = sdZ + T;
    for (int i = 0; i < 4; ++i) { sK[i] = cur; cur += T; }
    for (int i = 0; i < 4; ++i) { sG[i] = cur; cur += T; }
    for (int i = 0; i < 3; ++i) { sH0[i] = cur; cur += T; }
    for (int i = 0; i < 3; ++i) { sH1[i] = cur; cur += T; }
    for (int i = 0; i < 4; ++i) { sKb[i] = cur; cur += T; }
    for (int i = 0; i < 4; ++i) { sGb[i] = cur; cur += T; }
    float* sUb = cur; cur += T;       // cotangent of the state at the start of the step in hand (on entry: of its end)
    float* sA = cur; cur += T;        // input cotangent of a drift VJP
    float* sBv = cur; cur += T;       // input cotangent of a diffusion VJP
    float* sE2 = cur; cur += T;
    for (int i = 0; i < 4; ++i) { sHid[i] = cur; cur += HT; }
    float* sHb = cur; cur += HT;
    const float* W1 = sP; const float* b1 = W1 + H * D; const float* W2 = b1 + H; const float* b2 = W2 + D * H;
    const float* Wg = b2 + D; const float* bg = Wg + D * D;
    float* gW1 = sGr; float* gb1 = gW1 + H * D; float* gW2 = gb1 + H; float* gb2 = gW2 + D * H; float* gWg = gb2 + D; float* gbg = gWg + D * D;
    const SriTableau& tb = c_SRI[P.alg == 1 ? 1 : 0];

    for (int e = tid; e < np; e += NT) { sP[e] = __ldg(P.p + e); sGr[e] = 0.f; }
    for (int e = tid; e < T; e += NT) {
        const int r = e / NP, n = e - r * NP;
        sUb[e] = (P.du && n < Nloc) ? __ldg(P.du + (size_t)D * (c0 + n) + r) : 0.f;
    }
    __syncthreads();
    const int nsteps = min(P.stats->naccept, P.tape_cap);
    const float atol = P.abstol, rtol = P.reltol;
    const float cnt = (float)D * (float)P.B;
    const float delta = (float)(1.0 / 26.0);

    // forward: k = f(in) with its hidden activations kept, g = g(in2)
    auto eval_fg = [&](const float* in, float* outk, float* hid, const float* in2, float* outg) {
        for (int e = tid; e < HT; e += NT) {
            const int h = e / NP, n = e - h * NP;
            float s = 0.f;
            for (int k = 0; k < D; ++k) s = rn_fmaf(W1[h + H * k], in[k * NP + n], s);
            hid[e] = tanhf(s + b1[h]);
        }
        __syncthreads();
        for (int e = tid; e < 2 * T; e += NT) {
            const bool dif = e >= T;
            const int ee = dif ? e - T : e;
            const int r = ee / NP, n = ee - r * NP;
            if (!dif) {
                float s = 0.f;
                for (int k = 0; k < H; ++k) s = rn_fmaf(W2[r + D * k], hid[k * NP + n], s);
                outk[ee] = s + b2[r];
            } else {
                float s = 0.f;
                for (int k = 0; k < D; ++k) s = rn_fmaf(Wg[r + D * k], in2[k * NP + n], s);
                outg[ee] = s + bg[r];
            }
        }
        __syncthreads();
    };
    // VJPs of one stage: drift at `in` (hidden activations hid) with cotangent kb -> sA; diffusion at in2 with cotangent gb -> sBv;
    // parameter gradients accumulate in shared memory (every entry is owned by one thread per loop: no atomics)
    auto vjp_fg = [&](const float* in, const float* hid, const float* kb, const float* in2, const float* gb) {
        for (int e = tid; e < HT; e += NT) {                  // hidden cotangent: W2^T kb, through tanh
            const int h = e / NP, n = e - h * NP;
            float s = 0.f;
            for (int r = 0; r < D; ++r) s = rn_fmaf(W2[r + D * h], kb[r * NP + n], s);
            const float hv = hid[e];
            sHb[e] = s * (1.f - hv * hv);
        }
        for (int e = tid; e < D * H; e += NT) {               // dW2[r, h] += sum_n kb[r, n] hid[h, n]
            const int r = e % D, h = e / D;
            float s = 0.f;
#pragma unroll
            for (int n = 0; n < NP; ++n) s = rn_fmaf(kb[r * NP + n], hid[h * NP + n], s);
            gW2[e] += s;
        }
        for (int e = tid; e < D * D; e += NT) {               // dWg[r, k] += sum_n gb[r, n] in2[k, n]
            const int r = e % D, k = e / D;
            float s = 0.f;
#pragma unroll
            for (int n = 0; n < NP; ++n) s = rn_fmaf(gb[r * NP + n], in2[k * NP + n], s);
            gWg[e] += s;
        }
        for (int r = tid; r < D; r += NT) {
            float s2 = 0.f, sg = 0.f;
#pragma unroll
            for (int n = 0; n < NP; ++n) { s2 += kb[r * NP + n]; sg += gb[r * NP + n]; }
            gb2[r] += s2; gbg[r] += sg;
        }
        __syncthreads();
        for (int e = tid; e < 2 * T; e += NT) {               // input cotangents: W1^T delta1 and Wg^T gb
            const bool dif = e >= T;
            const int ee = dif ? e - T : e;
            const int k = ee / NP, n = ee - k * NP;
            float s = 0.f;
            if (!dif) { for (int h = 0; h < H; ++h) s = rn_fmaf(W1[h + H * k], sHb[h * NP + n], s); sA[ee] = s; }
            else { for (int r = 0; r < D; ++r) s = rn_fmaf(Wg[r + D * k], gb[r * NP + n], s); sBv[ee] = s; }
        }
        for (int e = tid; e < H * D; e += NT) {               // dW1[h, k] += sum_n delta1[h, n] in[k, n]
            const int h = e % H, k = e / H;
            float s = 0.f;
#pragma unroll
            for (int n = 0; n < NP; ++n) s = rn_fmaf(sHb[h * NP + n], in[k * NP + n], s);
            gW1[e] += s;
        }
        for (int h = tid; h < H; h += NT) {
            float s = 0.f;
#pragma unroll
            for (int n = 0; n < NP; ++n) s += sHb[h * NP + n];
            gb1[h] += s;
        }
        __syncthreads();
    };

    for (int s = nsteps - 1; s >= 0; --s) {
        const float* tp = P.tape + (((size_t)s * P.Q + q) * 3) * T;
        for (int e = tid; e < T; e += NT) { sU[e] = __ldcg(tp + e); sdW[e] = __ldcg(tp + T + e); sdZ[e] = __ldcg(tp + 2 * T + e); }
        const float dtc = __ldg(P.tape_steps + 4 * s), EEst = __ldg(P.tape_steps + 4 * s + 1);
        const float n1 = __ldg(P.tape_steps + 4 * s + 2), n2 = __ldg(P.tape_steps + 4 * s + 3);
        const float sqdt = (float)sqrt((double)dtc), sqrt3 = (float)sqrt(3.0);
        __syncthreads();
        auto chi1 = [&](int e) { const float w = sdW[e]; return (w * w - dtc) / (2.f * sqdt); };
        auto chi2 = [&](int e) { return (sdW[e] + sdZ[e] / sqrt3) / 2.f; };
        auto chi3 = [&](int e) { const float w = sdW[e]; return (w * w * w - 3.f * w * dtc) / (6.f * dtc); };
        // ---- the step again (sde_kernel.cuh), keeping every stage input and hidden activation ----
        eval_fg(sU, sK[0], sHid[0], sU, sG[0]);
        for (int e = tid; e < T; e += NT) {
            const float k1 = sK[0][e], g1 = sG[0][e], u = sU[e];
            sH0[0][e] = u + dtc * tb.a021 * k1 + tb.b021 * chi2(e) * g1;
            sH1[0][e] = u + dtc * tb.a121 * k1 + sqdt * tb.b121 * g1;
        }
        __syncthreads();
        eval_fg(sH0[0], sK[1], sHid[1], sH1[0], sG[1]);
        for (int e = tid; e < T; e += NT) {
            const float k1 = sK[0][e], k2 = sK[1][e], g1 = sG[0][e], g2 = sG[1][e], u = sU[e];
            sH0[1][e] = u + dtc * (tb.a031 * k1 + tb.a032 * k2) + chi2(e) * (tb.b031 * g1 + tb.b032 * g2);
            sH1[1][e] = u + dtc * (tb.a131 * k1 + tb.a132 * k2) + sqdt * (tb.b131 * g1 + tb.b132 * g2);
        }
        __syncthreads();
        eval_fg(sH0[1], sK[2], sHid[2], sH1[1], sG[2]);
        for (int e = tid; e < T; e += NT) {
            const float k1 = sK[0][e], k2 = sK[1][e], k3 = sK[2][e], g1 = sG[0][e], g2 = sG[1][e], g3 = sG[2][e], u = sU[e];
            sH0[2][e] = u + dtc * (tb.a041 * k1 + tb.a042 * k2 + tb.a043 * k3) + chi2(e) * (tb.b041 * g1 + tb.b042 * g2 + tb.b043 * g3);
            sH1[2][e] = u + dtc * (tb.a141 * k1 + tb.a142 * k2 + tb.a143 * k3) + sqdt * (tb.b141 * g1 + tb.b142 * g2 + tb.b143 * g3);
        }
        __syncthreads();
        eval_fg(sH0[2], sK[3], sHid[3], sH1[2], sG[3]);
        // ---- cotangents of the update formula and of the error estimate ----
        const float sbar = (P.dsaveval && P.reg_kind == RNDE_REG_ERR_DT) ? __ldg(P.dsaveval + s + 1) : 0.f;
        // stiffness-estimate regulariser of AutoSOSRI2 (mnist_nsde.jl:52-56): saved = |eig| / 10.6, eig = rms(k4 - k3) / rms(H03 - H02)
        // -> cotangents on k4, k3 and directly on the stage inputs H03, H02 (added where those inputs' cotangents are distributed)
        float cK = 0.f, cH = 0.f;
        if (P.dsaveval && P.reg_kind == RNDE_REG_STIFF_SCALED && P.alg == 1 && n1 > 0.f && n2 > 0.f) {
            const float eigbar = __ldg(P.dsaveval + s + 1) * (1.f / 10.6f);       // eig = n1 / n2 >= 0
            cK = (eigbar / n2) / (cnt * n1);                                      // multiplies (k4 - k3)
            cH = (-eigbar * n1 / (n2 * n2)) / (cnt * n2);                         // multiplies (H03 - H02)
        }
        const float gE = (sbar != 0.f && EEst > 0.f) ? sbar * dtc / (cnt * EEst) : 0.f;      // d(EEst dt)/d(resid_e) = dt resid_e / (cnt EEst)
        for (int e = tid; e < T; e += NT) {
            const int n = e % NP;
            const float k1 = sK[0][e], k2 = sK[1][e], k3 = sK[2][e], k4 = sK[3][e];
            const float g1 = sG[0][e], g2 = sG[1][e], g3 = sG[2][e], g4 = sG[3][e], u = sU[e];
            const float c1 = chi1(e), c2 = chi2(e), c3 = chi3(e), w = sdW[e];
            const float E2 = c2 * (tb.be31 * g1 + tb.be32 * g2 + tb.be33 * g3 + tb.be34 * g4) + c3 * (tb.be41 * g1 + tb.be42 * g2 + tb.be43 * g3 + tb.be44 * g4);
            const float un = u + dtc * (tb.al1 * k1 + tb.al2 * k2 + tb.al3 * k3 + tb.al4 * k4) + E2 +
                             w * (tb.be11 * g1 + tb.be12 * g2 + tb.be13 * g3 + tb.be14 * g4) + c1 * (tb.be21 * g1 + tb.be22 * g2 + tb.be23 * g3 + tb.be24 * g4);
            float Eb = 0.f, denb = 0.f;
            const float au = fabsf(u), an = fabsf(un);
            if (gE != 0.f && n < Nloc) {
                const float den = atol + fmaxf(au, an) * rtol;
                const float resid = (delta * dtc * (k1 + k2 + k3 + k4) + E2) / den;
                const float rb = gE * resid;
                Eb = rb / den; denb = -rb * resid / den;
            }
            const float sgu = u > 0.f ? 1.f : (u < 0.f ? -1.f : 0.f), sgn = un > 0.f ? 1.f : (un < 0.f ? -1.f : 0.f);
            const float wn = an > au ? 1.f : (an < au ? 0.f : 0.5f);      // max(|u|, |u'|): the larger branch, a tie splits
            const float unb = sUb[e] + wn * denb * rtol * sgn;
            sUb[e] = unb + (1.f - wn) * denb * rtol * sgu;                 // direct path u' = u + ... and the |u| branch
            const float kcommon = dtc * delta * Eb;
            sKb[0][e] = dtc * tb.al1 * unb + kcommon; sKb[1][e] = dtc * tb.al2 * unb + kcommon;
            const float dk = (n < Nloc) ? cK * (k4 - k3) : 0.f;
            sKb[2][e] = dtc * tb.al3 * unb + kcommon - dk; sKb[3][e] = dtc * tb.al4 * unb + kcommon + dk;
            const float ue = unb + Eb;                                     // E2 sits in u' and in the residual
            sGb[0][e] = (w * tb.be11 + c1 * tb.be21) * unb + (c2 * tb.be31 + c3 * tb.be41) * ue;
            sGb[1][e] = (w * tb.be12 + c1 * tb.be22) * unb + (c2 * tb.be32 + c3 * tb.be42) * ue;
            sGb[2][e] = (w * tb.be13 + c1 * tb.be23) * unb + (c2 * tb.be33 + c3 * tb.be43) * ue;
            sGb[3][e] = (w * tb.be14 + c1 * tb.be24) * unb + (c2 * tb.be34 + c3 * tb.be44) * ue;
        }
        __syncthreads();
        // ---- the stages in reverse ----
        vjp_fg(sH0[2], sHid[3], sKb[3], sH1[2], sGb[3]);      // k4 = f(H03), g4 = g(H13)
        for (int e = tid; e < T; e += NT) {
            const float dh = ((e % NP) < Nloc) ? cH * (sH0[2][e] - sH0[1][e]) : 0.f;
            const float a = sA[e] + dh, b = sBv[e], c2 = chi2(e);
            sUb[e] += a + b;
            sKb[0][e] += dtc * (tb.a041 * a + tb.a141 * b); sKb[1][e] += dtc * (tb.a042 * a + tb.a142 * b); sKb[2][e] += dtc * (tb.a043 * a + tb.a143 * b);
            sGb[0][e] += c2 * tb.b041 * a + sqdt * tb.b141 * b; sGb[1][e] += c2 * tb.b042 * a + sqdt * tb.b142 * b; sGb[2][e] += c2 * tb.b043 * a + sqdt * tb.b143 * b;
        }
        __syncthreads();
        vjp_fg(sH0[1], sHid[2], sKb[2], sH1[1], sGb[2]);      // k3 = f(H02), g3 = g(H12)
        for (int e = tid; e < T; e += NT) {
            const float dh = ((e % NP) < Nloc) ? cH * (sH0[2][e] - sH0[1][e]) : 0.f;
            const float a = sA[e] - dh, b = sBv[e], c2 = chi2(e);
            sUb[e] += a + b;
            sKb[0][e] += dtc * (tb.a031 * a + tb.a131 * b); sKb[1][e] += dtc * (tb.a032 * a + tb.a132 * b);
            sGb[0][e] += c2 * tb.b031 * a + sqdt * tb.b131 * b; sGb[1][e] += c2 * tb.b032 * a + sqdt * tb.b132 * b;
        }
        __syncthreads();
        vjp_fg(sH0[0], sHid[1], sKb[1], sH1[0], sGb[1]);      // k2 = f(H01), g2 = g(H11)
        for (int e = tid; e < T; e += NT) {
            const float a = sA[e], b = sBv[e];
            sUb[e] += a + b;
            sKb[0][e] += dtc * (tb.a021 * a + tb.a121 * b);
            sGb[0][e] += chi2(e) * tb.b021 * a + sqdt * tb.b121 * b;
        }
        __syncthreads();
        vjp_fg(sU, sHid[0], sKb[0], sU, sGb[0]);              // k1 = f(u), g1 = g(u)
        for (int e = tid; e < T; e += NT) sUb[e] += sA[e] + sBv[e];
        __syncthreads();
    }
    for (int e = tid; e < T; e += NT) {
        const int r = e / NP, n = e - r * NP;
        if (P.dx && n < Nloc) P.dx[(size_t)D * (c0 + n) + r] = sUb[e];
    }
    for (int e = tid; e < np; e += NT) P.gpart[(size_t)q * np + e] = sGr[e];
}

// fixed-order sum of the per-tile gradients (Float64)
__global__ void sde_grad_reduce_kernel(const float* __restrict__ gpart, int Q, int np, float* __restrict__ dp) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    double s = 0.0;
    for (int q = 0; q < Q; ++q) s += (double)gpart[(size_t)q * np + i];
    dp[i] = (float)s;
}

}  // namespace rnde
