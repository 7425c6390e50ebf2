// bwd_kernel.cuh -- persistent reverse sweep over the recorded accepted steps.
//
// This is the discrete adjoint of the exact step sequence the forward kernel took
// (what Tracker.gradient through solve(...; sensealg=SensitivityADPassThrough()) computes:
// /root/reference/src/models/neural_ode.jl:134, experiments/mnist_node.jl:229-232,
// SURVEY.md 3.2 / Appendix A.6 with the step sizes frozen).  Cotangents enter from the
// final state and from every saved value EEst_j*dt_j / eigen_est_j (through norm ->
// residual -> utilde, u0, u1).  The same cluster decomposition as the forward kernel is
// used: each CTA owns R state rows of NP columns; per stage it needs one H x NP exchange
// (W2^T delta2 is a split-K over the CTA's rows).  The per-stage deltas are written back
// over the tape so that the parameter gradients become two large batched contractions
// (wgrad_kernel.cuh) instead of rank-NP updates inside this latency-bound sweep.
#pragma once
#include "common.cuh"
#include "chain.cuh"
#include "a6.cuh"
#include "csq_bwd.cuh"
#include "wgrad_kernel.cuh"      // rec_time

namespace rnde {

struct BwdLayout {
    int HP, RP;
    int oW2T, oW1T, oUb, oUpb, oKb, oPart, oD1, total;
};

__host__ __device__ inline BwdLayout make_bwd_layout(int G, int NP, bool WS, int D, int H, int R, int HS) {
    BwdLayout L;
    L.HP = round_up(H, 4);
    L.RP = round_up(R, 4);
    int o = 0;
    L.oW2T = o; o += WS ? R * L.HP : 0;     // [k = local row][m = hidden]
    L.oW1T = o; o += WS ? H * L.RP : 0;     // [k = hidden][m = local row]
    L.oUb = o; o += L.RP * NP;
    L.oUpb = o; o += L.RP * NP;
    L.oKb = o; o += 7 * L.RP * NP;
    L.oPart = o; o += G * HS * NP;
    L.oD1 = o; o += L.HP * NP;
    L.total = o;
    return L;
}

// FIELD: 0 = the 2-layer time-concatenated field or a chain field; 1 = the FFJORD field (csq_bwd.cuh)
template <int G, int NP, int TM, bool WS, int NT, int FIELD = 0>
__global__ void __launch_bounds__(NT, 1) bwd_kernel(const KParams P) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    constexpr int LN = NP / 4;
    constexpr int LM = 32 / LN;
    constexpr int TMW = LM * TM;
    const int rank = (G > 1) ? (int)cluster_ctarank() : 0;
    const int q = blockIdx.x / G;
    const int D = P.D, H = P.H, R = P.R, HS = P.HS, td = P.td;
    const int r0 = rank * R;
    const int Rloc = max(0, min(R, D - r0));
    const int c0 = q * NP;
    const int Nloc = max(0, min(NP, P.B - c0));
    const int HSloc = max(0, min(HS, H - rank * HS));
    const BwdLayout L = make_bwd_layout(G, NP, WS, D, H, R, HS);
    const int HP = L.HP, RP = L.RP;
    float* sW2T = smem + L.oW2T; float* sW1T = smem + L.oW1T;
    float* sUb = smem + L.oUb; float* sUpb = smem + L.oUpb;
    float* sPart = smem + L.oPart; float* sD1 = smem + L.oD1;
    float* const sKb = smem + L.oKb;
    const int kstride = RP * NP;
    int flipK = 0;
    auto Kb = [&](const int j) -> float* {
        int slot = j - 1;
        if (j == 1) slot = flipK ? 6 : 0;
        if (j == 7) slot = flipK ? 0 : 6;
        return sKb + slot * kstride;
    };
    const float* gW1 = P.p;
    const float* gW2 = gW1 + (size_t)H * (D + td) + H;
    const size_t tileD = (size_t)D * NP, tileH = (size_t)H * NP;
    auto tile_off_D = [&](int rec) -> size_t { return ((size_t)rec * P.Q + q) * tileD + (size_t)r0 * NP; };
    auto tile_off_H = [&](int rec) -> size_t { return ((size_t)rec * P.Q + q) * tileH; };

    const bool chain = (G == 1) && P.n_layers > 0;
    ChainView cv;
    cv.L = P.n_layers; cv.D = D; cv.NP = NP; cv.hrows = P.hrows; cv.w = P.lw; cv.a = P.la; cv.pre = P.pre_act;
    cv.sW = smem + P.oCW; cv.sA = smem + P.oCA; cv.sB = smem + P.oCB; cv.sH = smem + P.oCH;
    if (chain) for (int e = tid; e < P.chain_np; e += NT) smem[P.oCW + e] = __ldg(P.p + e);
    if constexpr (FIELD == 1) {      // noise tile of this CTA's columns, [row][NP] like the state (as in fwd_kernel)
        const int Dz = D - P.csq_extra;
        float* sE = smem + P.oCS;
        for (int e = tid; e < Dz * NP; e += NT) {
            const int n = e / Dz, i = e - n * Dz;
            sE[i * NP + n] = (n < Nloc) ? __ldg(P.noise + (size_t)Dz * (c0 + n) + i) : 0.f;
        }
        if (P.oCSP > 0) {
            const int npar = csq_num_params(Dz, H);
            for (int e = tid; e < npar; e += NT) smem[P.oCSP + e] = __ldg(P.p + e);
        }
    }
    if constexpr (WS && FIELD == 0) if (!chain) {
        for (int e = tid; e < R * HP; e += NT) {       // W2T[k][m] = W2[r0+k, m]
            const int k = e / HP, m = e - k * HP;
            sW2T[e] = (k < Rloc && m < H) ? __ldg(gW2 + (size_t)D * m + r0 + k) : 0.f;
        }
        for (int e = tid; e < H * RP; e += NT) {       // W1T[k][m] = W1[k, r0+m]
            const int k = e / RP, m = e - k * RP;
            sW1T[e] = (m < Rloc) ? __ldg(gW1 + (size_t)H * (r0 + m) + k) : 0.f;
        }
    }
    for (int e = tid; e < RP * NP; e += NT) {
        const int n = e / RP, m = e - n * RP;
        sUb[m * NP + n] = (P.du && m < Rloc && n < Nloc) ? __ldg(P.du + (size_t)D * (c0 + n) + r0 + m) : 0.f;
        sUpb[m * NP + n] = 0.f;
    }
    for (int e = tid; e < 7 * kstride; e += NT) sKb[e] = 0.f;
    __syncthreads();
    if constexpr (G > 1) cluster_sync_all();

    const int ln = lane % LN, lm = lane / LN, n0 = ln * 4;

    // VJP of one field evaluation (record `rec`).  On entry Kb(i) holds kbar_i; it is
    // turned into delta2 in place.  `epi(m, n0, zbar[4])` consumes zbar = W1^T delta1.
    // Appendix A.6 (a6.cuh): this thread's part of dL/d(dt_1); `wt` weighs the time cotangent of the evaluation,
    // `accum` adds the deltas to the record instead of replacing k (second VJP of record 0, k taken from the copy a6_f0)
    double dacc = 0.0;
    auto vjp = [&](float* sKbar, const int rec, const float wt, const bool accum, auto epi) {
        if constexpr (FIELD == 1) {      // FFJORD field: recompute the evaluation from the taped stage input and reverse it
            const int Dz = D - P.csq_extra;
            float* sE = smem + P.oCS;
            float* sZc = sE + Dz * NP;
            const size_t base = ((size_t)rec * P.Q + q) * (size_t)D * NP;
            for (int e = tid; e < D * NP; e += NT) sZc[e] = __ldcg(P.tapeZ + base + e);
            __syncthreads();
            const float* zb = csq_vjp<NP, NT>(P.oCSP > 0 ? smem + P.oCSP : P.p, Dz, H, P.csq_extra, rec_time(P.steps, P.t0, rec), sZc, sE, sKbar,
                                              P.tapeH + ((size_t)rec * P.Q + q) * (size_t)P.hrows * NP, csq_bwd_carve(sZc + D * NP, Dz, H, NP));
            for (int e = tid; e < Rloc * (NP / 4); e += NT) {
                const int m = e / (NP / 4), nn = (e - m * (NP / 4)) * 4;
                epi(m, nn, zb + m * NP + nn);
            }
            __syncthreads();
            return;
        }
        if constexpr (G == 1 && WS) {
            if (chain) {
                const float* zb = chain_vjp<NP, NT>(P, cv, sKbar, rec, q, accum ? P.a6_f0 + (size_t)q * tileD : nullptr);
                for (int e = tid; e < Rloc * (NP / 4); e += NT) {
                    const int m = e / (NP / 4), nn = (e - m * (NP / 4)) * 4;
                    epi(m, nn, zb + m * NP + nn);
                }
                __syncthreads();
                return;
            }
        }
        const size_t offD = tile_off_D(rec), offH = tile_off_H(rec);
        const bool want_t = (wt != 0.f) && td;
        for (int e = tid; e < Rloc * NP; e += NT) {
            const float kb = sKbar[e];
            float d2 = kb;
            if (P.act2 == RNDE_ACT_TANH) {
                const float kv = accum ? __ldcg(P.a6_f0 + (size_t)q * tileD + (size_t)r0 * NP + e) : __ldcg(P.tapeK + offD + e);
                d2 = kb * (1.f - kv * kv);
            }
            sKbar[e] = d2;
            if (accum) P.tapeK[offD + e] += d2;
            else P.tapeK[offD + e] = d2;       // delta2 replaces k in the tape (consumed by wgrad)
            if (want_t) dacc += (double)(wt * d2 * __ldg(gW2 + (size_t)D * H + r0 + e / NP));
        }
        __syncthreads();
        // phase A': hbar partial = W2[rows, :H]^T delta2   (split-K over this CTA's rows)
        for (int mt = warp; mt * TMW < H; mt += NW) {
            const int m0 = mt * TMW + lm * TM;
            if (m0 < H) {
                float acc[TM][4];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
                for (int k = 0; k < Rloc; ++k) {
                    float w[TM];
                    if constexpr (TM == 4) {
                        if constexpr (WS) {
                            const float4 w4 = *reinterpret_cast<const float4*>(sW2T + k * HP + m0);
                            w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
                        } else {
#pragma unroll
                            for (int i = 0; i < TM; ++i) w[i] = (m0 + i < H) ? __ldg(gW2 + (size_t)D * (m0 + i) + r0 + k) : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < TM; ++i) {
                            if constexpr (WS) w[i] = sW2T[k * HP + min(m0 + i, HP - 1)];
                            else w[i] = (m0 + i < H) ? __ldg(gW2 + (size_t)D * (m0 + i) + r0 + k) : 0.f;
                        }
                    }
                    const float4 x4 = *reinterpret_cast<const float4*>(sKbar + k * NP + n0);
                    const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = rn_fmaf(w[i], xv[j], acc[i][j]);
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const int m = m0 + i;
                    if (m < H) {
                        const float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                        if constexpr (G > 1) {
                            const int d = m / HS, ml = m - d * HS;
                            st_cluster_f4(mapa_u32(smem_u32(sPart + (rank * HS + ml) * NP + n0), d), v);
                        } else {
                            *reinterpret_cast<float4*>(sPart + m * NP + n0) = v;
                        }
                    }
                }
            }
        }
        group_sync<G>();
        // phase B': reduce over the cluster, delta1 = hbar * act1'(h), broadcast
        for (int e = tid; e < HSloc * NP; e += NT) {
            const int ml = e / NP, n = e - ml * NP;
            const int m = rank * HS + ml;
            float s = sPart[ml * NP + n];
#pragma unroll
            for (int c = 1; c < G; ++c) s = s + sPart[(c * HS + ml) * NP + n];
            float d1 = s;
            if (P.act1 == RNDE_ACT_TANH) { const float hv = __ldcg(P.tapeH + offH + (size_t)m * NP + n); d1 = s * (1.f - hv * hv); }
            if constexpr (G > 1) {
                const uint32_t a = smem_u32(sD1 + m * NP + n);
#pragma unroll
                for (int d = 0; d < G; ++d) st_cluster_f32(mapa_u32(a, d), d1);
            } else {
                sD1[m * NP + n] = d1;
            }
            if (accum) P.tapeD1[offH + (size_t)m * NP + n] += d1;
            else P.tapeD1[offH + (size_t)m * NP + n] = d1;
            if (want_t) dacc += (double)(wt * d1 * __ldg(gW1 + (size_t)H * D + m));
        }
        group_sync<G>();
        // phase C': zbar = W1[:, rows]^T delta1  for this CTA's rows
        for (int mt = warp; mt * TMW < Rloc; mt += NW) {
            const int m0 = mt * TMW + lm * TM;
            if (m0 < Rloc) {
                float acc[TM][4];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
                for (int k = 0; k < H; ++k) {
                    float w[TM];
                    if constexpr (TM == 4) {
                        if constexpr (WS) {
                            const float4 w4 = *reinterpret_cast<const float4*>(sW1T + k * RP + m0);
                            w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
                        } else {
#pragma unroll
                            for (int i = 0; i < TM; ++i) w[i] = (m0 + i < Rloc) ? __ldg(gW1 + (size_t)H * (r0 + m0 + i) + k) : 0.f;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < TM; ++i) {
                            if constexpr (WS) w[i] = sW1T[k * RP + min(m0 + i, RP - 1)];
                            else w[i] = (m0 + i < Rloc) ? __ldg(gW1 + (size_t)H * (r0 + m0 + i) + k) : 0.f;
                        }
                    }
                    const float4 x4 = *reinterpret_cast<const float4*>(sD1 + k * NP + n0);
                    const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = rn_fmaf(w[i], xv[j], acc[i][j]);
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const int m = m0 + i;
                    if (m < Rloc) epi(m, n0, acc[i]);
                }
            }
        }
        __syncthreads();
    };

    const float atol = P.abstol, rtol = P.reltol;
    const float cntf = (float)P.norm_count;
    const float stab = rn_divf(1.0f, (float)TS_STABILITY_SIZE);

    // ---- Appendix A.6: the two evaluations of the initial-dt heuristic (launched after the sweep, a6.cuh) ----------------
    if (P.a6_mode != 0) {
        const A6Scal sc = a6_scalars(P, P.a6_sum[0]);
        const float d0 = P.initdt[0], d1 = P.initdt[1], d2 = P.initdt[2], dt0 = P.initdt[3];
        const float* f0t = P.a6_f0 + (size_t)q * tileD + (size_t)r0 * NP;
        const size_t boff = (size_t)q * tileD + (size_t)r0 * NP, bstr = (size_t)P.Q * tileD;     // three buffers: u1bar, f1bar, w * wbar
        const size_t off0 = tile_off_D(0);
        double part = 0.0;
        if (P.a6_mode == 1) {
            // cotangent of f1 through d2 = rms((f1 - f0) / sk) / dt0
            const float r2 = d2 * dt0;
            const float coef = (sc.d2bar != 0.f && r2 > 0.f) ? (sc.d2bar / dt0) / (cntf * r2) : 0.f;
            const size_t offx = tile_off_D(P.rec_x);
            for (int e = tid; e < Rloc * NP; e += NT) {
                const float u0 = __ldcg(P.tapeZ + off0 + e), f0 = __ldcg(f0t + e), f1 = __ldcg(P.tapeK + offx + e);
                const float sk = rn_fmaf(fabsf(u0), rtol, atol);
                const float wv = (f1 - f0) / sk;
                const float wb = ((e % NP) < Nloc) ? coef * wv : 0.f;
                Kb(7)[e] = wb / sk;
                P.a6_u1bar[bstr + boff + e] = wb / sk;
                P.a6_u1bar[2 * bstr + boff + e] = wb * wv / sk;
            }
            __syncthreads();
            vjp(Kb(7), P.rec_x, 1.f, false, [&](const int m, const int nn, const float* zb) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = m * NP + nn + j;
                    P.a6_u1bar[boff + e] = zb[j];
                    part += (double)(zb[j] * __ldcg(f0t + e));       // u1 = u0 + dt0 f0: <u1bar, f0> goes to dt0
                }
            });
            part += dacc;      // the evaluation's time t0 + dt0
            if (blockIdx.x == 0 && tid == 0) {      // the pseudo-step whose stage 7 (time t + dt) is this record, for the wgrad time row
                StepRec sr; sr.t = P.t0; sr.dt = dt0; sr.eest = 0.f; sr.eig = 0.f; sr.n1 = 0.f; sr.n2 = 0.f; sr.pad0 = 0.f; sr.pad1 = 0.f;
                P.steps[P.nsteps] = sr;
            }
        } else {
            float dt0bar = sc.dt0bar + (float)P.a6_sum[1], d0bar = 0.f, d1bar = sc.d1bar;
            if (sc.dt0_free) { d0bar = dt0bar * dt0 / d0; d1bar -= dt0bar * dt0 / d1; }
            const float c1 = (d1bar != 0.f && d1 > 0.f) ? d1bar / (cntf * d1) : 0.f;
            const float c0f = (d0bar != 0.f && d0 > 0.f) ? d0bar / (cntf * d0) : 0.f;
            for (int e = tid; e < Rloc * NP; e += NT) {
                float f0bar = 0.f, u0bar = 0.f;
                if ((e % NP) < Nloc) {
                    const float u0 = __ldcg(P.tapeZ + off0 + e), f0 = __ldcg(f0t + e);
                    const float sk = rn_fmaf(fabsf(u0), rtol, atol);
                    const float u1b = P.a6_u1bar[boff + e], f1b = P.a6_u1bar[bstr + boff + e], wwb = P.a6_u1bar[2 * bstr + boff + e];
                    const float v = f0 / sk, y = u0 / sk;
                    const float vb = c1 * v, yb = c0f * y;
                    f0bar = rn_fmaf(dt0, u1b, vb / sk - f1b);
                    const float skbar = -wwb - (vb * v + yb * y) / sk;
                    u0bar = u1b + yb / sk + skbar * rtol * (u0 > 0.f ? 1.f : (u0 < 0.f ? -1.f : 0.f));
                }
                Kb(7)[e] = f0bar;
                sUb[e] = u0bar;
            }
            __syncthreads();
            vjp(Kb(7), 0, 0.f, true, [&](const int m, const int nn, const float* zb) {
                if (P.dx) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int n = nn + j;
                        if (n < Nloc) P.dx[(size_t)D * (c0 + n) + r0 + m] += sUb[m * NP + n] + zb[j];
                    }
                }
            });
        }
        a6_block_sum<NT>(part, sUpb, P.a6_part);
        if constexpr (G > 1) cluster_sync_all();
        return;
    }
    const float clampN = P.a6 ? P.initdt[5] : 0.f;

    // saveat cotangents are consumed newest first; times beyond the last accepted step were never saved
    int sidx = P.dusave ? P.n_saveat - 1 : -1;
    if (sidx >= 0) {
        const float tend = P.nsteps > 0 ? P.steps[P.nsteps - 1].t + P.steps[P.nsteps - 1].dt : P.t0;
        while (sidx >= 0 && __ldg(P.saveat + sidx) > tend) --sidx;
    }
    auto gsave = [&](const int si, const int e) -> float {
        const int m = e / NP, n = e - m * NP;
        return (n < Nloc) ? __ldg(P.dusave + (size_t)D * (si + (size_t)P.n_saveat * (c0 + n)) + r0 + m) : 0.f;
    };

    for (int s = P.nsteps - 1; s >= 0; --s) {
        const StepRec sr = P.steps[s];
        const float dt = sr.dt, EEst = sr.eest, eig = sr.eig, n1 = sr.n1, n2 = sr.n2;
        // cotangents of this step's saved value
        const float sbar = P.dsaveval ? __ldg(P.dsaveval + s + 1) : 0.f;
        float eestbar = 0.f, eigbar = 0.f;
        if (sbar != 0.f) {
            switch (P.reg_kind) {
                case RNDE_REG_ERR_DT: eestbar = sbar * dt; break;
                case RNDE_REG_STIFF_DT_ABS: eigbar = sbar * ((eig * dt) >= 0.f ? 1.f : -1.f) * dt; break;
                case RNDE_REG_STIFF_SCALED: { const float a = fabsf(eig); if (!(a == 0.f || a != a)) eigbar = sbar * stab * (eig >= 0.f ? 1.f : -1.f); break; }
                case RNDE_REG_ERR_PLUS_STIFF: {
                    const float e = EEst * dt;
                    if (!(e == 0.f || e != e)) eestbar = sbar * dt;
                    if (!(eig == 0.f || eig != eig)) eigbar = sbar * (0.1f * stab);
                    break;
                }
                default: break;
            }
        }
        if (P.alg != RNDE_ALG_AUTO_TSIT5) eigbar = 0.f;
        const bool use_eest = (eestbar != 0.f) && (EEst > 0.f);
        const bool use_eig = (eigbar != 0.f) && (n1 > 0.f) && (n2 > 0.f);
        const float gE = use_eest ? eestbar / (cntf * EEst) : 0.f;
        const float n1b = use_eig ? eigbar / n2 : 0.f;
        const float n2b = use_eig ? -eigbar * n1 / (n2 * n2) : 0.f;
        const float gA = use_eig ? n1b / (cntf * n1) : 0.f;    // multiplies (k7-k6)
        const float gB = use_eig ? n2b / (cntf * n2) : 0.f;    // multiplies (u-g6)
        const int recU0 = 6 * s, recU1 = 6 * s + 6, recG6 = 6 * s + 5;
        // weight of this step's explicit dt in dL/d(dt_1): the first step takes dt_1, the last one shrinks by it when it was cut
        const float wdir = P.a6 ? ((s == 0 ? 1.f : 0.f) - (s == P.nsteps - 1 ? clampN : 0.f)) : 0.f;
        const float wshift = (P.a6 && s >= 1) ? 1.f : 0.f;      // ... and every later step starts dt_1 later
        if (wdir != 0.f && sbar != 0.f && blockIdx.x == 0 && tid == 0 && P.a6_scalar) {
            if (P.reg_kind == RNDE_REG_ERR_DT) dacc += (double)(wdir * sbar * EEst);
            else if (P.reg_kind == RNDE_REG_STIFF_DT_ABS && P.alg == RNDE_ALG_AUTO_TSIT5) dacc += wdir * sbar * ((eig * dt) >= 0.f ? 1.f : -1.f) * eig;
            else if (P.reg_kind == RNDE_REG_ERR_PLUS_STIFF) { const float e = EEst * dt; if (!(e == 0.f || e != e)) dacc += wdir * sbar * EEst; }
        }

        // reset the per-step cotangents (Kb(7) carries k7bar from the following step)
        for (int e = tid; e < Rloc * NP; e += NT) {
            sUpb[e] = 0.f;
#pragma unroll
            for (int j = 1; j <= 6; ++j) Kb(j)[e] = 0.f;
        }
        __syncthreads();
        // states saved inside this step: u(theta) = uprev + dt * sum_j b_j(theta) k_j   (theta frozen)
        while (sidx >= 0 && __ldg(P.saveat + sidx) > sr.t) {
            const float tau = __ldg(P.saveat + sidx);
            if (tau == sr.t + dt) {
                for (int e = tid; e < Rloc * NP; e += NT) sUb[e] += gsave(sidx, e);
            } else {
                float bw[8];
                interp_weights(rn_divf(tau - sr.t, dt), bw);
                for (int e = tid; e < Rloc * NP; e += NT) {
                    const float g = gsave(sidx, e);
                    sUpb[e] += g;
#pragma unroll
                    for (int j = 1; j <= 7; ++j) Kb(j)[e] += (dt * bw[j]) * g;
                    if (wdir != 0.f) {
                        float ssum = 0.f;
                        for (int j = 1; j <= 7; ++j) ssum = rn_fmaf(bw[j], __ldcg(P.tapeK + tile_off_D(6 * s + j - 1) + e), ssum);
                        dacc += (double)(wdir * g * ssum);
                    }
                }
            }
            --sidx;
        }
        if (use_eest || use_eig) {
            for (int e = tid; e < Rloc * NP; e += NT) {
                const int n = e % NP;
                if (n >= Nloc) continue;
                const float up = __ldcg(P.tapeZ + tile_off_D(recU0) + e);
                const float un = __ldcg(P.tapeZ + tile_off_D(recU1) + e);
                float kv[8];
#pragma unroll
                for (int j = 1; j <= 7; ++j) kv[j] = __ldcg(P.tapeK + tile_off_D(6 * s + j - 1) + e);
                if (use_eest) {
                    float ssum = ts_bt(1) * kv[1];
#pragma unroll
                    for (int j = 2; j <= 7; ++j) ssum = rn_fmaf(ts_bt(j), kv[j], ssum);
                    const float ut = dt * ssum;
                    const float a0 = fabsf(up), a1 = fabsf(un);
                    const float mx = a0 > a1 ? a0 : a1;
                    const float den = rn_fmaf(mx, rtol, atol);
                    const float at = ut / den;
                    const float ab = gE * at;
                    const float utb = ab / den;
                    const float mb = (-ab * at / den) * rtol;
                    if (a0 > a1) sUpb[e] += mb * (up >= 0.f ? 1.f : -1.f);
                    else if (a1 > a0) sUb[e] += mb * (un >= 0.f ? 1.f : -1.f);
                    else {
                        sUpb[e] += 0.5f * mb * (up > 0.f ? 1.f : (up < 0.f ? -1.f : 0.f));
                        sUb[e] += 0.5f * mb * (un > 0.f ? 1.f : (un < 0.f ? -1.f : 0.f));
                    }
#pragma unroll
                    for (int j = 1; j <= 7; ++j) Kb(j)[e] += dt * ts_bt(j) * utb;
                    dacc += (double)(wdir * utb * ssum);
                }
                if (use_eig) {
                    const float g6 = __ldcg(P.tapeZ + tile_off_D(recG6) + e);
                    const float ga = gA * (kv[7] - kv[6]);
                    const float gb = gB * (un - g6);
                    Kb(7)[e] += ga; Kb(6)[e] -= ga;
                    sUb[e] += gb;            // the matching -gb on g6 is applied in stage 6's epilogue
                }
            }
            __syncthreads();
        }
        for (int i = 7; i >= 2; --i) {
            const int rec = 6 * s + i - 1;
            vjp(Kb(i), rec, wshift + wdir * ts_c(i), false, [&](const int m, const int nn, const float* zb) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = m * NP + nn + j;
                    float g = zb[j];
                    if (i == 7) g += sUb[e];
                    if (i == 6 && use_eig && (nn + j) < Nloc) {
                        const float un = __ldcg(P.tapeZ + tile_off_D(recU1) + e);
                        const float g6 = __ldcg(P.tapeZ + tile_off_D(recG6) + e);
                        g -= gB * (un - g6);
                    }
                    for (int jj = 1; jj < i; ++jj) Kb(jj)[e] = rn_fmaf(dt * ts_a(i, jj), g, Kb(jj)[e]);
                    sUpb[e] += g;
                    if (wdir != 0.f && (nn + j) < Nloc) {       // z_i = u + dt * sum_j a_ij k_j (records of the later stages already hold deltas)
                        float ssum = 0.f;
                        for (int jj = 1; jj < i; ++jj) ssum = rn_fmaf(ts_a(i, jj), __ldcg(P.tapeK + tile_off_D(6 * s + jj - 1) + e), ssum);
                        dacc += (double)(wdir * g * ssum);
                    }
                }
            });
        }
        // hand over: u_new(prev step) = uprev, k7(prev step) = k1
        { float* tmp = sUb; sUb = sUpb; sUpb = tmp; }
        flipK ^= 1;
        __syncthreads();
    }
    // initial fsalfirst = f(u0, t0): record 0, cotangent carried in Kb(7)
    vjp(Kb(7), 0, 0.f, false, [&](const int m, const int nn, const float* zb) {
        if (P.dx) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = nn + j;
                if (n < Nloc) {
                    float g = sUb[m * NP + n] + zb[j];
                    for (int si = sidx; si >= 0; --si) g += gsave(si, m * NP + n);     // saves at t0 are the input itself
                    P.dx[(size_t)D * (c0 + n) + r0 + m] = g;
                }
            }
        }
    });
    if (P.a6) a6_block_sum<NT>(dacc, sUpb, P.a6_part);
    if constexpr (G > 1) cluster_sync_all();
}

}  // namespace rnde
