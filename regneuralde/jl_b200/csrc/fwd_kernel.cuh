// fwd_kernel.cuh -- persistent fused adaptive-Tsit5 stepper (forward solve).
//
// One launch integrates the whole batch from t0 to t1 with no host round trip:
// 7-stage Tsit5 with the stage combinations fused, the 2-layer time-concatenated
// MLP field evaluated on chip, embedded error norm, PI controller, accept/reject,
// regulariser saved values, NFE accounting and the backward tape, all on device.
// Replaces: solve(prob, Tsit5()|AutoTsit5(Tsit5()); callback=SavingCallback(func, sv), ...)
// at /root/reference/src/models/neural_ode.jl:131-137 and everything it calls
// (OrdinaryDiffEq 5.50.0 perform_step!/loopfooter!).
//
// Work decomposition (template parameters):
//   G   CTAs per cluster.  The cluster owns a tile of NP batch columns; CTA `rank`
//       owns state rows [rank*R, rank*R+R) of every column of the tile and the
//       matching slices of W1 (all hidden units x its R input rows, a split-K of
//       layer 1) and W2 (its R output rows x all hidden units).  The only traffic
//       between CTAs per field evaluation is the H x NP hidden pre-activation:
//       partial sums are pushed through distributed shared memory to the CTA that
//       reduces that hidden slice (fixed order = canonical), the activated slice is
//       broadcast back.  State, stage values and weights never leave shared memory.
//   NP  columns per tile (multiple of 4).
//   TM  rows per thread tile in the register-tiled FFMA GEMMs (TM x 4 accumulators).
//   WS  weights resident in shared memory (true) or streamed from global/L2 (false).
// The only grid-wide dependency is the once-per-step RMS norm: per-column sums of
// squares go through a small global buffer and a grid barrier, and every CTA folds
// them in the same fixed order, so all CTAs take identical controller decisions.
#pragma once
#include "common.cuh"
#include "chain.cuh"
#include "csq.cuh"

namespace rnde {

struct SmemLayout {
    int HP, RP, nbl, ngmax;
    int oW1, oW1t, ob1, oW2, oW2t, ob2, oU, oZ, oK, oPart, oH, oRed, oCP, oTot, oCtl, total;
};

struct Ctl {
    float t, dt, dtpropose, qold, q11, eig_prev, EEst, eig, dt_init, dt_last;
    float qold_pow, qold_pow_next;     // qold^beta2 of the current / of the candidate next qold (computed by a second thread)
    int accept, accept_prev, done, iter, nf, naccept, nreject, n_saved, retcode, as_count, as_stiff;
};

__host__ __device__ inline SmemLayout make_layout(int G, int NP, bool WS, int D, int H, int R, int HS, int kblock) {
    SmemLayout L;
    L.HP = round_up(H, 4);
    L.RP = round_up(R, 4);
    L.nbl = (G > 1) ? 1 : (D + kblock - 1) / kblock;
    L.ngmax = (kblock + 3) / 4;
    int o = 0;
    L.oW1 = o; o += WS ? R * L.HP : 0;
    L.oW1t = o; o += L.HP;
    L.ob1 = o; o += L.HP;
    L.oW2 = o; o += WS ? H * L.RP : 0;
    L.oW2t = o; o += L.RP;
    L.ob2 = o; o += L.RP;
    L.oU = o; o += L.RP * NP;
    L.oZ = o; o += L.RP * NP;
    L.oK = o; o += 7 * L.RP * NP;
    // the norm scratch (sRed) is only live between field evaluations, the exchange buffers
    // (sPart, sH) only inside one: they share storage.
    const int xchg = G * round_up(HS, 1) * NP + L.HP * NP;
    const int red = 3 * L.nbl * L.ngmax * NP;
    L.oPart = o;
    L.oH = o + G * round_up(HS, 1) * NP;
    L.oRed = o;
    o += (xchg > red ? xchg : red);
    L.oCP = o; o += 3 * G * NP;
    L.oTot = o; o += 4;
    L.oCtl = o; o += 32;
    L.total = o;
    return L;
}

// Canonical combination of per-block partials: adjacent blocks paired first, pairs accumulated
// left to right (oracle/rnde_oracle.c combine_blocks).
template <class F>
__device__ __forceinline__ float combine_blocks(int n, F part) {
    float tot = 0.f;
    for (int b = 0; b < n; b += 2) {
        const float pair = (b + 1 < n) ? part(b) + part(b + 1) : part(b);
        tot = (b == 0) ? pair : tot + pair;
    }
    return tot;
}

// Canonical RMS norms of up to NV fields at once.
//   val(ml, n, out[NV]) gives the field values at local row ml, column n.
// Order: per column, per kblock: groups of 4 consecutive rows are fma chains, group sums are
// added in order, blocks (== CTAs of the cluster when G>1) combined by combine_blocks; columns
// folded by 32 interleaved chains + xor butterfly.  Identical to oracle col_sumsq/cols_total.
template <int G, int NP, int NV, int NT, class F>
__device__ __forceinline__ void grid_rms(const KParams& P, const SmemLayout& L, float* smem, int rank, int q, int Rloc, int Nloc,
                                         unsigned& norm_seq, unsigned& bar_gen, F val, float* out, unsigned xseq_base = 0) {
    const int tid = threadIdx.x;
    float* sRed = smem + L.oRed;
    float* sCP = smem + L.oCP;
    float* sTot = smem + L.oTot;
    const int nbl = L.nbl, ngmax = L.ngmax;
    constexpr int NL = NT / NP;       // group lanes per column
    {
        const int n = tid % NP, l = tid / NP;
        for (int b = 0; b < nbl; ++b) {
            const int rb0 = (G > 1) ? 0 : b * P.kblock;
            const int rb1 = (G > 1) ? Rloc : min(rb0 + P.kblock, Rloc);
            for (int g = l; rb0 + 4 * g < rb1; g += NL) {
                const int g0 = rb0 + 4 * g, g1 = min(g0 + 4, rb1);
                float acc[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[v] = 0.f;
                for (int r = g0; r < g1; ++r) {
                    float vv[NV];
                    val(r, n, vv);
#pragma unroll
                    for (int v = 0; v < NV; ++v) acc[v] = rn_fmaf(vv[v], vv[v], acc[v]);
                }
#pragma unroll
                for (int v = 0; v < NV; ++v) sRed[((v * nbl + b) * ngmax + g) * NP + n] = acc[v];
            }
        }
    }
    __syncthreads();
    const unsigned slot = norm_seq & 1u;
    float* gcol = P.colsum + (size_t)slot * 3 * P.colsum_stride;
    if (tid < NP * NV) {
        const int n = tid % NP, v = tid / NP;
        const float tot = combine_blocks(nbl, [&](int b) {
            const int rb0 = (G > 1) ? 0 : b * P.kblock;
            const int rb1 = (G > 1) ? Rloc : min(rb0 + P.kblock, Rloc);
            const int ng = (rb1 - rb0 + 3) / 4;
            const float* rp = sRed + ((v * nbl + b) * ngmax) * NP + n;
            float s = 0.f;
            for (int g = 0; g < ng; ++g) s = (g == 0) ? rp[0] : s + rp[g * NP];
            return s;
        });
        if constexpr (G > 1) {
            st_cluster_f32(mapa_u32(smem_u32(sCP + (v * G + rank) * NP + n), 0), tot);
        } else {
            if (n < Nloc) publish_colsum(P, (size_t)slot * 3 * P.colsum_stride + (size_t)v * P.colsum_stride + P.col_offset + q * NP + n, tot);
            if (P.nranks > 1) __threadfence_system();
        }
    }
    if constexpr (G > 1) {
        cluster_sync_all();
        if (rank == 0 && tid < NP * NV) {
            const int n = tid % NP, v = tid / NP;
            const float tot = combine_blocks(G, [&](int c) { return sCP[(v * G + c) * NP + n]; });
            if (n < Nloc) publish_colsum(P, (size_t)slot * 3 * P.colsum_stride + (size_t)v * P.colsum_stride + P.col_offset + q * NP + n, tot);
            if (P.nranks > 1) __threadfence_system();
        }
    }
    grid_barrier(P.bar, gridDim.x, bar_gen);
    xrank_barrier(P, xseq_base + norm_seq + 1u);
    const int warp = tid >> 5, lane = tid & 31;
    if (warp < NV) {
        const float* g = gcol + (size_t)warp * P.colsum_stride;
        float s = 0.f;
        for (int j0 = lane; j0 < P.Bglobal; j0 += 32 * 16) {      // 16 loads in flight, added in the canonical order
            float v[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) v[u] = (j0 + 32 * u < P.Bglobal) ? __ldcg(g + j0 + 32 * u) : 0.f;
#pragma unroll
            for (int u = 0; u < 16; ++u) if (j0 + 32 * u < P.Bglobal) s = s + v[u];
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == 0) sTot[warp] = rn_sqrtf(rn_divf(s, (float)P.norm_count));
    }
    __syncthreads();
#pragma unroll
    for (int v = 0; v < NV; ++v) out[v] = sTot[v];
    norm_seq += 1;
    __syncthreads();
}

__device__ __forceinline__ float saved_value(int kind, float EEst, float eig, float dt) {
    const float stab = rn_divf(1.0f, (float)TS_STABILITY_SIZE);
    switch (kind) {
        case RNDE_REG_ERR_DT: return EEst * dt;
        case RNDE_REG_STIFF_DT_ABS: return fabsf(eig * dt);
        case RNDE_REG_STIFF_SCALED: { float s = fabsf(eig); return stab * ((s == 0.f || s != s) ? 0.f : s); }
        case RNDE_REG_ERR_PLUS_STIFF: {
            float e = EEst * dt;
            float a = (e == 0.f || e != e) ? 0.f : e;
            float b = (eig == 0.f || eig != eig) ? 0.f : eig;
            return (a + (0.1f * stab) * b) * 1.0f;
        }
        default: return 0.f;
    }
}

// FIELD: 0 = the 2-layer time-concatenated field or a chain field (runtime P.n_layers); 1 = the FFJORD field of csq.cuh
template <int G, int NP, int TM, bool WS, int NT, int FIELD = 0>
__global__ void __launch_bounds__(NT, 1) fwd_kernel(const KParams P) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NT / 32;
    constexpr int LN = NP / 4;        // lanes along the column dimension
    constexpr int LM = 32 / LN;       // lanes along the row dimension
    constexpr int TMW = LM * TM;      // rows per warp tile
    const int rank = (G > 1) ? (int)cluster_ctarank() : 0;
    const int q = blockIdx.x / G;
    const int D = P.D, H = P.H, R = P.R, HS = P.HS, td = P.td;
    const int r0 = rank * R;
    const int Rloc = max(0, min(R, D - r0));
    const int c0 = q * NP;
    const int Nloc = max(0, min(NP, P.B - c0));
    const int HSloc = max(0, min(HS, H - rank * HS));
    const SmemLayout L = make_layout(G, NP, WS, D, H, R, HS, P.kblock);
    const int HP = L.HP, RP = L.RP;
    float* sW1 = smem + L.oW1; float* sW1t = smem + L.oW1t; float* sb1 = smem + L.ob1;
    float* sW2 = smem + L.oW2; float* sW2t = smem + L.oW2t; float* sb2 = smem + L.ob2;
    float* sU = smem + L.oU; float* sZ = smem + L.oZ;
    float* sPart = smem + L.oPart; float* sH = smem + L.oH;
    Ctl* ctl = reinterpret_cast<Ctl*>(smem + L.oCtl);
    // k1..k7 live in 7 fixed buffers; FSAL (k1 <- k7) and u <- u_new flip two bits instead of copying.
    float* const sKbase = smem + L.oK;
    const int kstride = RP * NP;
    int flipK = 0;
    auto K = [&](const int j) -> float* {
        int slot = j - 1;
        if (j == 1) slot = flipK ? 6 : 0;
        if (j == 7) slot = flipK ? 0 : 6;
        return sKbase + slot * kstride;
    };

    const float* gW1 = P.p;
    const float* gb1 = gW1 + (size_t)H * (D + td);
    const float* gW2 = gb1 + H;
    const float* gb2 = gW2 + (size_t)D * (H + td);

    // ---- stage weights / biases / initial state into shared memory -------
    const bool chain = (G == 1) && P.n_layers > 0;
    if constexpr (WS && FIELD == 0) if (!chain) {
        for (int e = tid; e < R * HP; e += NT) {
            const int k = e / HP, m = e - k * HP;
            sW1[e] = (k < Rloc && m < H) ? __ldg(gW1 + (size_t)(r0 + k) * H + m) : 0.f;
        }
        for (int e = tid; e < H * RP; e += NT) {
            const int k = e / RP, m = e - k * RP;
            sW2[e] = (m < Rloc) ? __ldg(gW2 + (size_t)D * k + r0 + m) : 0.f;
        }
    }
    if constexpr (FIELD == 1) {      // noise tile of this CTA's columns, [row][NP] like the state
        static_assert(G == 1 && WS, "the FFJORD field runs on single-CTA tiles");
        const int Dz = D - P.csq_extra;
        float* sE = smem + P.oCS;
        for (int e = tid; e < Dz * NP; e += NT) {
            const int n = e / Dz, i = e - n * Dz;
            sE[i * NP + n] = (n < Nloc) ? __ldg(P.noise + (size_t)Dz * (c0 + n) + i) : 0.f;
        }
        if (P.oCSP > 0) {      // all parameters (78 KB for MLPDynamics(43, 100)) stay in shared memory for the whole solve
            const int npar = csq_num_params(Dz, H);
            for (int e = tid; e < npar; e += NT) smem[P.oCSP + e] = __ldg(P.p + e);
        }
    } else if (!chain) {
        for (int m = tid; m < HP; m += NT) {
            sW1t[m] = (td && m < H) ? __ldg(gW1 + (size_t)H * D + m) : 0.f;
            sb1[m] = (m < H) ? __ldg(gb1 + m) : 0.f;
        }
        for (int m = tid; m < RP; m += NT) {
            sW2t[m] = (td && m < Rloc) ? __ldg(gW2 + (size_t)D * H + r0 + m) : 0.f;
            sb2[m] = (m < Rloc) ? __ldg(gb2 + r0 + m) : 0.f;
        }
    } else {
        for (int e = tid; e < P.chain_np; e += NT) smem[P.oCW + e] = __ldg(P.p + e);
    }
    ChainView cv;
    cv.L = P.n_layers; cv.D = D; cv.NP = NP; cv.hrows = P.hrows; cv.w = P.lw; cv.a = P.la; cv.pre = P.pre_act;
    cv.sW = smem + P.oCW; cv.sA = smem + P.oCA; cv.sB = smem + P.oCB; cv.sH = nullptr;
    for (int e = tid; e < RP * NP; e += NT) {
        const int n = e / RP, m = e - n * RP;   // m fastest: coalesced read of column-major x
        sU[m * NP + n] = (m < Rloc && n < Nloc) ? __ldg(P.x + (size_t)D * (c0 + n) + r0 + m) : 0.f;
    }
    if (tid == 0) {
        Ctl c;
        c.t = P.t0; c.dt = 0.f; c.dtpropose = 0.f; c.qold = (float)1e-4; c.q11 = 1.f; c.eig_prev = 1.f; c.EEst = 1.f; c.eig = 1.f;
        c.dt_init = 0.f; c.dt_last = 0.f;
        c.accept = 0; c.accept_prev = 1; c.done = 0; c.iter = 0; c.nf = 0; c.naccept = 0; c.nreject = 0; c.n_saved = 0;
        c.retcode = RNDE_OK; c.as_count = 0; c.as_stiff = 0;
        if (P.reg_kind != RNDE_REG_NONE) {
            // SavingCallback initial entry: EEst = 1, dt = 0, eigen_est = 1 (Appendix A.7)
            if (blockIdx.x == 0 && P.saveval) P.saveval[0] = saved_value(P.reg_kind, 1.f, 1.f, 0.f);
            c.n_saved = 1;
        }
        *ctl = c;
    }
    __syncthreads();
    if constexpr (G > 1) cluster_sync_all();   // peers' shared memory is live before any DSMEM store

    // saveat: times <= t0 save the initial state (save_start when tspan[1] is in saveat)
    int save_idx = 0;
    auto save_out = [&](const int sidx, auto val) {
        for (int e = tid; e < RP * NP; e += NT) {
            const int n = e / RP, m = e - n * RP;
            if (m < Rloc && n < Nloc) P.usave[(size_t)D * (sidx + (size_t)P.n_saveat * (c0 + n)) + r0 + m] = val(m * NP + n);
        }
    };
    while (save_idx < P.n_saveat && __ldg(P.saveat + save_idx) <= P.t0) {
        save_out(save_idx, [&](int e) { return sU[e]; });
        ++save_idx;
    }

    unsigned norm_seq = 0, bar_gen = 0;
    unsigned* xseq_ptr = reinterpret_cast<unsigned*>(P.peers[P.rank]) + P.flag_off + 32;
    const unsigned xseq_base = (P.nranks > 1) ? *xseq_ptr : 0u;

    // ---- one evaluation of the vector field: sOut = f(sIn, tstage) --------
    auto rhs = [&](const float* sIn, float* sOut, const float tstage, const int rec) {
        if constexpr (FIELD == 1) {
            const int Dz = D - P.csq_extra;
            float* sE = smem + P.oCS;
            csq_rhs<NP, NT>(P.oCSP > 0 ? smem + P.oCSP : P.p, Dz, H, P.csq_extra, P.csq_reverse ? (P.t0 + P.t1) - tstage : tstage, sIn, sE, sOut, csq_carve(sE + Dz * NP, Dz, H, NP));
            if (P.csq_reverse) {      // the flow backwards (`sample`, ffjord.jl:160-167)
                for (int e = tid; e < D * NP; e += NT) sOut[e] = -sOut[e];
                __syncthreads();
            }
            if (rec >= 0) {      // tape: stage input and k (the reverse pass recomputes everything else, csq_bwd.cuh)
                const size_t base = ((size_t)rec * P.Q + q) * D * NP;
                for (int e = tid; e < D * NP; e += NT) { P.tapeZ[base + e] = sIn[e]; P.tapeK[base + e] = sOut[e]; }
            }
            return;
        }
        if constexpr (G == 1 && WS) {
            if (chain) { chain_rhs<NP, NT>(P, cv, sIn, sOut, rec, q); return; }
        }
        // phase A: partial hidden pre-activations over this CTA's input rows
        const int ln = lane % LN, lm = lane / LN, n0 = ln * 4;
        for (int mt = warp; mt * TMW < H; mt += NW) {
            const int m0 = mt * TMW + lm * TM;
            float tot[TM][4];
            const bool active = m0 < H;
            if (active) {
                int blk = 0;
                float pairv[TM][4];
                for (int kb0 = 0; kb0 < Rloc; kb0 += P.kblock, ++blk) {
                    const int kb1 = min(kb0 + P.kblock, Rloc);
                    float acc[TM][4];
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
                    for (int k = kb0; k < kb1; ++k) {
                        float w[TM];
                        if constexpr (TM == 4) {
                            float4 w4;
                            if constexpr (WS) w4 = *reinterpret_cast<const float4*>(sW1 + k * HP + m0);
                            else w4 = ldg4(gW1 + (size_t)(r0 + k) * H + m0, H - m0);
                            w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
                        } else {
#pragma unroll
                            for (int i = 0; i < TM; ++i) {
                                if constexpr (WS) w[i] = sW1[k * HP + min(m0 + i, HP - 1)];
                                else w[i] = (m0 + i < H) ? __ldg(gW1 + (size_t)(r0 + k) * H + m0 + i) : 0.f;
                            }
                        }
                        const float4 x4 = *reinterpret_cast<const float4*>(sIn + k * NP + n0);
                        const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                        for (int i = 0; i < TM; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = rn_fmaf(w[i], xv[j], acc[i][j]);
                    }
                    // canonical block combination: adjacent blocks paired, pairs accumulated in order
                    const bool odd = (blk & 1) != 0, flush = odd || kb1 == Rloc;
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            pairv[i][j] = odd ? pairv[i][j] + acc[i][j] : acc[i][j];
                            if (flush) tot[i][j] = (blk < 2) ? pairv[i][j] : tot[i][j] + pairv[i][j];
                        }
                }
                if (blk == 0) {
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) tot[i][j] = 0.f;
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const int m = m0 + i;
                    if (m < H) {
                        const float4 v = make_float4(tot[i][0], tot[i][1], tot[i][2], tot[i][3]);
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
        // phase B: fixed-order reduction over the cluster, time column, bias, activation
        for (int e = tid; e < HSloc * NP; e += NT) {
            const int ml = e / NP, n = e - ml * NP;
            const int m = rank * HS + ml;
            float s = combine_blocks(G, [&](int c) { return sPart[(c * HS + ml) * NP + n]; });
            if (td) s = rn_fmaf(sW1t[m], tstage, s);
            s = s + sb1[m];
            const float hv = act_apply(P.act1, s);
            if constexpr (G > 1) {
                const uint32_t a = smem_u32(sH + m * NP + n);
#pragma unroll
                for (int d = 0; d < G; ++d) st_cluster_f32(mapa_u32(a, d), hv);
            } else {
                sH[m * NP + n] = hv;
            }
            if (rec >= 0) P.tapeH[((size_t)rec * P.Q + q) * H * NP + (size_t)m * NP + n] = hv;
        }
        group_sync<G>();
        // phase C: layer 2 for this CTA's output rows (full K = H chain)
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
                        float4 w4;
                        if constexpr (WS) w4 = *reinterpret_cast<const float4*>(sW2 + k * RP + m0);
                        else w4 = ldg4(gW2 + (size_t)D * k + r0 + m0, Rloc - m0);
                        w[0] = w4.x; w[1] = w4.y; w[2] = w4.z; w[3] = w4.w;
                    } else {
#pragma unroll
                        for (int i = 0; i < TM; ++i) {
                            if constexpr (WS) w[i] = sW2[k * RP + min(m0 + i, RP - 1)];
                            else w[i] = (m0 + i < Rloc) ? __ldg(gW2 + (size_t)D * k + r0 + m0 + i) : 0.f;
                        }
                    }
                    const float4 x4 = *reinterpret_cast<const float4*>(sH + k * NP + n0);
                    const float xv[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = rn_fmaf(w[i], xv[j], acc[i][j]);
                }
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const int m = m0 + i;
                    if (m < Rloc) {
                        float o[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float v = acc[i][j];
                            if (td) v = rn_fmaf(sW2t[m], tstage, v);
                            v = v + sb2[m];
                            o[j] = act_apply(P.act2, v);
                        }
                        const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<float4*>(sOut + m * NP + n0) = o4;
                        if (rec >= 0) {
                            const size_t off = ((size_t)rec * P.Q + q) * D * NP + (size_t)(r0 + m) * NP + n0;
                            *reinterpret_cast<float4*>(P.tapeK + off) = o4;
                            *reinterpret_cast<float4*>(P.tapeZ + off) = *reinterpret_cast<const float4*>(sIn + m * NP + n0);
                        }
                    }
                }
            }
        }
        __syncthreads();
    };

    // z_i = uprev + dt * sum_j a_ij k_j   (canonical association, see oracle stage_combo)
    auto combo_val = [&](const int i, const float dt, const float a2, const int e) -> float {
        if (i == 2) return rn_fmaf(a2, K(1)[e], sU[e]);
        float s = ts_a(i, 1) * K(1)[e];
        for (int j = 2; j < i; ++j) s = rn_fmaf(ts_a(i, j), K(j)[e], s);
        return rn_fmaf(dt, s, sU[e]);
    };

    // ---- initialize!: fsalfirst = f(u0, t0) --------------------------------
    rhs(sU, K(1), P.t0, P.need_tape ? 0 : -1);
    const float dtmax = P.t1 - P.t0;
    // ---- initial dt (Hairer-Wanner, Appendix A.5) ---------------------------
    {
        float d01[2];
        grid_rms<G, NP, 2, NT>(P, L, smem, rank, q, Rloc, Nloc, norm_seq, bar_gen,
            [&](int r, int n, float* o) {
                const float u = sU[r * NP + n];
                const float sk = rn_fmaf(fabsf(u), P.reltol, P.abstol);
                o[0] = rn_divf(u, sk);
                o[1] = rn_divf(K(1)[r * NP + n], sk);
            }, d01, xseq_base);
        const float d0 = d01[0], d1 = d01[1];
        float dt0;
        if (d0 < (float)1e-5 || d1 < (float)1e-5) dt0 = (float)1e-6;
        else dt0 = rn_divf(rn_divf(d0, d1), 100.f);
        if (dt0 > dtmax) dt0 = dtmax;
        for (int e = tid; e < Rloc * NP; e += NT) sZ[e] = rn_fmaf(dt0, K(1)[e], sU[e]);
        __syncthreads();
        rhs(sZ, K(2), P.t0 + dt0, (P.need_tape && P.a6) ? P.rec_init : -1);      // Appendix A.6: this evaluation stays on the tape
        float d2v[1];
        grid_rms<G, NP, 1, NT>(P, L, smem, rank, q, Rloc, Nloc, norm_seq, bar_gen,
            [&](int r, int n, float* o) {
                const float u = sU[r * NP + n];
                const float sk = rn_fmaf(fabsf(u), P.reltol, P.abstol);
                o[0] = rn_divf(K(2)[r * NP + n] - K(1)[r * NP + n], sk);
            }, d2v, xseq_base);
        if (tid == 0) {
            const float d2 = rn_divf(d2v[0], dt0);
            const float md = d1 > d2 ? d1 : d2;
            float dt1;
            if (md <= (float)1e-15) {
                const float a = dt0 * (float)1e-3;
                dt1 = a > (float)1e-6 ? a : (float)1e-6;
            } else {
                const float l10 = canon_log10f(md);
                const float ex = rn_divf(-(2.0f + l10), 5.0f);
                dt1 = (float)canon_exp10((double)ex);
            }
            float dt = 100.f * dt0;
            if (dt1 < dt) dt = dt1;
            if (dtmax < dt) dt = dtmax;
            if (dt < P.dtmin) dt = P.dtmin;
            ctl->dt = dt; ctl->dtpropose = dt; ctl->dt_init = dt; ctl->nf = 3;
            if (blockIdx.x == 0 && P.initdt) { P.initdt[0] = d0; P.initdt[1] = d1; P.initdt[2] = d2; P.initdt[3] = dt0; P.initdt[4] = dt1; }
        }
        __syncthreads();
    }

    const float gamma = (float)(9.0 / 10.0), qmin = (float)(1.0 / 5.0), qmax = 10.f;
    const float beta1 = (float)(7.0 / 50.0), beta2 = (float)(2.0 / 25.0), qoldinit = (float)1e-4;
    const bool limited = P.need_tape || P.reg_kind != RNDE_REG_NONE;

#ifdef RNDE_TIMELINE      // phase accumulators for tools/latent_timeline.py; compiled out of the product build
    long long tl_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tl_t = clock64();
#define TL(k) do { const long long _n = clock64(); tl_acc[k] += _n - tl_t; tl_t = _n; } while (0)
#else
#define TL(k) do { } while (0)
#endif
    // ---- solve!: the hot loop ---------------------------------------------
    while (true) {
        TL(7);
        if (tid == 0) {   // loopheader!
            Ctl& c = *ctl;
            if (!(c.t < P.t1)) c.done = 1;
            else if (c.iter >= P.max_steps) { c.retcode = RNDE_ERR_MAXITERS; c.done = 1; }
            else if (limited && c.naccept >= P.tape_cap) { c.retcode = RNDE_ERR_TAPE_FULL; c.done = 1; }
            else {
                float dt = c.dt;
                if (c.iter > 0) {
                    if (c.accept_prev) dt = c.dtpropose;
                    else {
                        const float f = rn_divf(c.q11, gamma), lim = rn_divf(1.f, qmin);
                        dt = rn_divf(dt, (lim < f ? lim : f));
                    }
                }
                c.iter += 1;
                if (P.alg == RNDE_ALG_AUTO_TSIT5) {   // AutoSwitch (Appendix A.8)
                    const float stiffness = fabsf(rn_divf(c.eig_prev * dt, (float)TS_STABILITY_SIZE));
                    const bool stiff = stiffness > (float)(9.0 / 10.0);
                    c.as_count = stiff ? (c.as_count < 0 ? 1 : c.as_count + 1) : (c.as_count > 0 ? -1 : c.as_count - 1);
                    if (!c.as_stiff && c.as_count > 10) { dt = dt * 2.f; c.as_stiff = 1; c.nf += 1; }
                    else if (c.as_stiff && c.as_count < -3) { dt = rn_divf(dt, 2.f); c.as_stiff = 0; c.nf += 1; }
                }
                if (dt > dtmax) dt = dtmax;
                if (dt < P.dtmin) dt = P.dtmin;
                const float rem = P.t1 - c.t;
                if (blockIdx.x == 0 && P.initdt) P.initdt[5] = rem < dt ? 1.f : 0.f;
                if (rem < dt) dt = rem;
                c.dt = dt;
            }
        }
        __syncthreads();
        if (ctl->done) break;
        TL(0);
        const float t = ctl->t, dt = ctl->dt;
        const int srec = P.need_tape ? 1 + 6 * ctl->naccept : -1;
        // (ctl is next written by thread 0 in loopfooter!, after the barriers inside grid_rms)
        const float a2 = dt * (float)TS_A21;
        // stages 2..7
        for (int i = 2; i <= 7; ++i) {
            for (int e = tid; e < Rloc * NP; e += NT) sZ[e] = combo_val(i, dt, a2, e);
            __syncthreads();
            TL(1);
            rhs(sZ, K(i), stage_time(t, dt, i), srec >= 0 ? srec + (i - 2) : -1);
            TL(2);
        }
        // embedded error estimate (+ eigen_est for the composite algorithm)
        float EEst, eig = 1.f, en1 = 0.f, en2 = 0.f;
        auto atmp_val = [&](int r, int n) -> float {
            const int e = r * NP + n;
            float s = ts_bt(1) * K(1)[e];
#pragma unroll
            for (int j = 2; j <= 7; ++j) s = rn_fmaf(ts_bt(j), K(j)[e], s);
            const float ut = dt * s;
            const float a0 = fabsf(sU[e]), a1 = fabsf(sZ[e]);
            const float m = a0 > a1 ? a0 : a1;
            return rn_divf(ut, rn_fmaf(m, P.reltol, P.abstol));
        };
        if (P.alg == RNDE_ALG_AUTO_TSIT5) {
            float o3[3];
            grid_rms<G, NP, 3, NT>(P, L, smem, rank, q, Rloc, Nloc, norm_seq, bar_gen,
                [&](int r, int n, float* o) {
                    const int e = r * NP + n;
                    o[0] = K(7)[e] - K(6)[e];
                    o[1] = sZ[e] - combo_val(6, dt, a2, e);
                    o[2] = atmp_val(r, n);
                }, o3, xseq_base);
            eig = rn_divf(o3[0], o3[1]); en1 = o3[0]; en2 = o3[1];
            EEst = o3[2];
        } else {
            float o1[1];
            grid_rms<G, NP, 1, NT>(P, L, smem, rank, q, Rloc, Nloc, norm_seq, bar_gen,
                [&](int r, int n, float* o) { o[0] = atmp_val(r, n); }, o1, xseq_base);
            EEst = o1[0];
        }
        TL(3);
        if (tid == 0) {   // loopfooter!
            Ctl& c = *ctl;
            c.nf += 6;
            c.EEst = EEst; c.eig = eig;
            if (EEst != EEst) { c.retcode = RNDE_ERR_NAN; c.done = 1; c.accept = 0; }
            else {
                float qv;
                if (EEst == 0.f) qv = rn_divf(1.f, qmax);
                else {
                    c.q11 = canon_powf(EEst, beta1);
                    qv = rn_divf(c.q11, canon_powf(c.qold, beta2));
                    float qq = rn_divf(qv, gamma);
                    const float hi = rn_divf(1.f, qmin), lo = rn_divf(1.f, qmax);
                    qq = hi < qq ? hi : qq;
                    qv = lo > qq ? lo : qq;
                }
                const int accept = EEst <= 1.f;
                if (P.alg == RNDE_ALG_AUTO_TSIT5) c.eig_prev = eig;
                if (accept) {
                    if (blockIdx.x == 0) {
                        if (c.naccept < P.tape_cap) { StepRec sr; sr.t = c.t; sr.dt = dt; sr.eest = EEst; sr.eig = eig; sr.n1 = en1; sr.n2 = en2; sr.pad0 = 0.f; sr.pad1 = 0.f; P.steps[c.naccept] = sr; }
                        if (P.reg_kind != RNDE_REG_NONE && P.saveval) P.saveval[c.n_saved] = saved_value(P.reg_kind, EEst, eig, dt);
                    }
                    if (P.reg_kind != RNDE_REG_NONE) c.n_saved += 1;
                    c.naccept += 1;
                    c.qold = EEst > qoldinit ? EEst : qoldinit;
                    const float dtnew = rn_divf(dt, qv);
                    c.t = c.t + dt;
                    float dp = dtnew < dtmax ? dtnew : dtmax;
                    if (dp < P.dtmin) dp = P.dtmin;
                    c.dtpropose = dp;
                    c.dt_last = dt;
                } else {
                    c.nreject += 1;
                    if (dt <= P.dtmin) { c.retcode = RNDE_ERR_DTMIN; c.done = 1; }
                }
                c.accept = accept;
                c.accept_prev = accept;
            }
        }
        __syncthreads();
        const int accepted = ctl->accept, finished = ctl->done;
        __syncthreads();   // everyone has read ctl before thread 0 starts the next loopheader!
        TL(4);
        if (accepted && save_idx < P.n_saveat) {   // savevalues!: every pending saveat time <= t_new
            const float tnew = t + dt;
            while (save_idx < P.n_saveat) {
                const float tau = __ldg(P.saveat + save_idx);
                if (!(tau <= tnew)) break;
                if (tau == tnew) save_out(save_idx, [&](int e) { return sZ[e]; });
                else {
                    float bw[8];
                    interp_weights(rn_divf(tau - t, dt), bw);
                    save_out(save_idx, [&](int e) {
                        float sacc = bw[1] * K(1)[e];
#pragma unroll
                        for (int j = 2; j <= 7; ++j) sacc = rn_fmaf(bw[j], K(j)[e], sacc);
                        return rn_fmaf(dt, sacc, sU[e]);
                    });
                }
                ++save_idx;
            }
        }
        if (accepted) {    // apply_step!: u <- u_new, fsalfirst <- fsallast
            float* tmp = sU; sU = sZ; sZ = tmp;
            flipK ^= 1;
        }
        TL(5);
        if (finished) break;
    }
#ifdef RNDE_TIMELINE
    if (P.dbg && blockIdx.x == 0 && tid == 0) for (int k = 0; k < 8; ++k) P.dbg[k] = tl_acc[k];
#endif

    // ---- write back ----------------------------------------------------------
    for (int e = tid; e < RP * NP; e += NT) {
        const int n = e / RP, m = e - n * RP;
        if (P.u_out && m < Rloc && n < Nloc) P.u_out[(size_t)D * (c0 + n) + r0 + m] = sU[m * NP + n];
    }
    if (blockIdx.x == 0 && tid == 0) {
        DevStats s;
        s.nf = ctl->nf; s.naccept = ctl->naccept; s.nreject = ctl->nreject; s.n_saved = ctl->n_saved; s.retcode = ctl->retcode;
        s.t_final = ctl->t; s.dt_last = ctl->dt_last; s.dt_init = ctl->dt_init;
        *P.stats = s;
        if (P.nranks > 1) *xseq_ptr = xseq_base + norm_seq;
    }
    if constexpr (G > 1) cluster_sync_all();   // no CTA exits while peers may still address its shared memory
}

}  // namespace rnde
