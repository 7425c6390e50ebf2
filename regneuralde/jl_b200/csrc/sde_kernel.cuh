// sde_kernel.cuh -- persistent adaptive SOSRI / SOSRI2 stepper for the reference's Neural SDE:
//   du = f(u) dt + g(u) dW,  f = Chain(Dense(D,H,tanh), Dense(H,D)),  g = Dense(D,D)  (diagonal noise),
// experiments/mnist_nsde.jl:73-80 behind TrackedNeuralDSDE (src/models/neural_sde.jl:84-146).  One launch integrates the whole
// batch: the four-stage Roessler SRI step (StochasticDiffEq FourStageSRIConstantCache), the embedded error estimate
// (delta*E1 + E2), the PI controller, accept / reject, the RSwM3 noise bookkeeping (futures / re-use stacks with Brownian
// bridging on rejection), the two NFE counters and the regulariser's saved values all run on the device.
//
// Decomposition: the same scaffolding as the chain stepper -- one CTA owns a tile of 4 batch columns, all 5 248 parameters
// sit in shared memory, the state and every stage value of the tile too (D x 4 floats each; 8- or 16-column tiles when the batch
// exceeds the co-resident capacity: 1776 / 3552 / 7104 columns), the RSwM3 stacks of the tile in HBM (L2-resident); the only grid-wide dependency
// is the RMS norm of the error estimate (per-CTA partial sums of squares in Float64 -> global -> grid barrier -> every CTA
// adds the Q partials in the same order, so all CTAs take identical controller decisions).
// Noise is INJECTED: normals[draw][row][column] holds standard normals; every request of the solver (dW, dZ of a fresh step,
// the bridge of a rejected one) consumes the next draw, in the order of oracle/sde_oracle.py -- the reference's random
// stream is not reproducible, parity is "with supplied noise".
// Arithmetic: Float32 like the reference (Float32 parameters and state), elementwise operations in the oracle's order; the
// step-size controller in Float64 like the oracle's Python scalars.  Parity with the oracle is held to 1e-5, not to the bit
// (its matrix products run through BLAS).
#pragma once
#include "common.cuh"

namespace rnde {

constexpr int SDE_NT = 256, SDE_MAXS = 32;

struct SdeStats {
    int nfe1, nfe2, naccept, nreject, n_saved, draws, retcode, pad;
    float t_final, dt_init, dt_last, pad2;
};

struct SdeParams {
    int D, H, B, Q, alg, reg_kind, max_steps, max_saved, n_draws;
    float t0, t1, abstol, reltol;
    const float* x; const float* p; const float* normals; float* u_out; float* saveval;
    double* partial;      // [2 slots][3][Q]
    float* stacks;        // RSwM3 stacks in HBM: [Q][4 (S1 dW, S1 dZ, S2 dW, S2 dZ)][SDE_MAXS][D * NP]
    unsigned* bar;
    SdeStats* stats;
    float* log; int log_cap;      // per attempt: dt, EEst, accepted
    // tape of the accepted steps for the reverse sweep (sde_bwd.cuh): [step][Q][3][D * NP] = state at the step's start, dW, dZ; [step][4] = dt, EEst, rms(k4 - k3), rms(H03 - H02)
    float* tape; float* tape_steps; int tape_cap;
};

// the 51 coefficients of a four-stage SRI tableau (filled by the host from the same Float64 literals as the oracle)
struct SriTableau {
    float a021, a031, a032, a041, a042, a043, a121, a131, a132, a141, a142, a143;
    float b021, b031, b032, b041, b042, b043, b121, b131, b132, b141, b142, b143;
    float al1, al2, al3, al4;
    float be11, be12, be13, be14, be21, be22, be23, be24, be31, be32, be33, be34, be41, be42, be43, be44;
};
__constant__ SriTableau c_SRI[2];

__host__ __device__ inline int sde_smem_floats(int D, int H, int np, int NP) {
    const int T = D * NP;
    return ((np + 3) / 4 * 4) + 16 * T + H * NP + 2 * SDE_MAXS * 2 + 64;
}

template <int NP>
__global__ void __launch_bounds__(SDE_NT) sde_kernel(const SdeParams P) {
    constexpr int NT = SDE_NT, MAXS = SDE_MAXS;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int D = P.D, H = P.H, T = D * NP;
    const int q = blockIdx.x, c0 = q * NP;
    const int Nloc = max(0, min(NP, P.B - c0));
    const int np = H * D + H + D * H + D + D * D + D;
    float* sP = smem;
    float* base = smem + (np + 3) / 4 * 4;
    float* sU = base; float* sUn = sU + T; float* sH0 = sUn + T; float* sH1 = sH0 + T;
    float* sK[4]; float* sG[4];
    for (int i = 0; i < 4; ++i) { sK[i] = sH1 + T + i * T; sG[i] = sH1 + T + (4 + i) * T; }
    float* sdW = sH1 + 9 * T; float* sdZ = sdW + T; float* sE2 = sdZ + T; float* sTmp = sE2 + T;      // 16 tiles in all
    float* sHid = sTmp + T;
    float* stW1 = P.stacks + (size_t)q * 4 * MAXS * T; float* stZ1 = stW1 + MAXS * T; float* stW2 = stZ1 + MAXS * T; float* stZ2 = stW2 + MAXS * T;
    double* stL1 = reinterpret_cast<double*>(sHid + H * NP);          // piece lengths of S1, then of S2
    double* stL2 = stL1 + MAXS;
    double* sRed = stL2 + MAXS;                                       // 8 warp partials + result
    const float* W1 = sP; const float* b1 = W1 + H * D; const float* W2 = b1 + H; const float* b2 = W2 + D * H;
    const float* Wg = b2 + D; const float* bg = Wg + D * D;
    const SriTableau& tb = c_SRI[P.alg == 1 ? 1 : 0];

    for (int e = tid; e < np; e += NT) sP[e] = __ldg(P.p + e);
    for (int e = tid; e < T; e += NT) {
        const int r = e / NP, n = e - r * NP;
        sU[e] = (n < Nloc) ? __ldg(P.x + (size_t)D * (c0 + n) + r) : 0.f;
    }
    __syncthreads();

    unsigned bar_gen = 0, norm_seq = 0;
    int draw = 0, retcode = RNDE_OK;
    int nfe1 = 0, nfe2 = 0;

    // drift f(in) -> outk and diffusion g(in2) -> outg of one stage (either may be skipped)
    auto eval_fg = [&](const float* in, float* outk, const float* in2, float* outg) {
        if (outk) {
            for (int e = tid; e < H * NP; e += NT) {
                const int h = e / NP, n = e - h * NP;
                float s = 0.f;
                for (int k = 0; k < D; ++k) s = rn_fmaf(W1[h + H * k], in[k * NP + n], s);
                sHid[e] = tanhf(s + b1[h]);
            }
        }
        __syncthreads();
        // drift layer 2 on the first T threads, the diffusion on the next T (T = 128 for the experiment's D = 32)
        for (int e = tid; e < 2 * T; e += NT) {
            const bool dif = e >= T;
            const int ee = dif ? e - T : e;
            const int r = ee / NP, n = ee - r * NP;
            if (!dif && outk) {
                float s = 0.f;
                for (int k = 0; k < H; ++k) s = rn_fmaf(W2[r + D * k], sHid[k * NP + n], s);
                outk[ee] = s + b2[r];
            } else if (dif && outg) {
                float s = 0.f;
                for (int k = 0; k < D; ++k) s = rn_fmaf(Wg[r + D * k], in2[k * NP + n], s);
                outg[ee] = s + bg[r];
            }
        }
        __syncthreads();
        if (outk) nfe1 += 1;
        if (outg) nfe2 += 1;
    };

    // RMS norms of up to 3 fields over the whole D x B array: val(e, out[NV]) per element of the tile
    auto norms = [&](auto val, auto nv_tag, double* result) {
        constexpr int NV = decltype(nv_tag)::value;
        double acc[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) acc[v] = 0.0;
        for (int e = tid; e < T; e += NT) {
            if ((e % NP) < Nloc) {
                float vv[NV];
                val(e, vv);
#pragma unroll
                for (int v = 0; v < NV; ++v) acc[v] += (double)vv[v] * (double)vv[v];
            }
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) acc[v] += __shfl_xor_sync(0xffffffffu, acc[v], off);
            if ((tid & 31) == 0) sRed[v * 8 + (tid >> 5)] = acc[v];
        }
        __syncthreads();
        const unsigned slot = norm_seq & 1u;
        if (tid < NV) {
            double s = 0.0;
            for (int w = 0; w < NT / 32; ++w) s += sRed[tid * 8 + w];
            P.partial[((size_t)slot * 3 + tid) * P.Q + q] = s;
        }
        grid_barrier(P.bar, gridDim.x, bar_gen);
        if (tid < NV) {
            const double* g = P.partial + ((size_t)slot * 3 + tid) * P.Q;
            double s = 0.0;
            for (int j = 0; j < P.Q; ++j) s += __ldcg(g + j);
            sRed[24 + tid] = sqrt(s / ((double)D * (double)P.B));
        }
        __syncthreads();
#pragma unroll
        for (int v = 0; v < NV; ++v) result[v] = sRed[24 + v];
        norm_seq += 1;
        __syncthreads();
    };

    auto normal = [&](int k, int e) -> float {      // standard normal of draw k for tile element e
        const int r = e / NP, n = e - r * NP;
        return (n < Nloc) ? __ldg(P.normals + ((size_t)k * D + r) * P.B + c0 + n) : 0.f;
    };
    // two consecutive draws (dW then dZ) scaled by sq: fresh increments of length h
    auto fresh = [&](double h, float* oW, float* oZ, bool add) {
        if (draw + 2 > P.n_draws) { retcode = RNDE_ERR_ARG; return; }
        const float sq = (float)sqrt(fabs(h));
        for (int e = tid; e < T; e += NT) {
            const float w = sq * normal(draw, e), z = sq * normal(draw + 1, e);
            oW[e] = add ? oW[e] + w : w; oZ[e] = add ? oZ[e] + z : z;
        }
        draw += 2;
    };
    // bridge inside a piece (L2, L3) of length h at fraction qf: out = qf * piece + sqrt((1-qf) qf h) * xi
    auto bridge = [&](double qf, double h, const float* L2, const float* L3, float* oW, float* oZ) {
        if (draw + 2 > P.n_draws) { retcode = RNDE_ERR_ARG; return; }
        const float sq = (float)sqrt((1.0 - qf) * qf * fabs(h)), qd = (float)qf;
        for (int e = tid; e < T; e += NT) {
            oW[e] = qd * L2[e] + sq * normal(draw, e);
            oZ[e] = qd * L3[e] + sq * normal(draw + 1, e);
        }
        draw += 2;
    };
    int n1 = 0, n2 = 0;       // stack depths (identical in every thread: all control flow below is uniform)
    double cur_dt = 0.0;
    auto push1 = [&](double L, const float* w, const float* z, float sw, const float* w2, const float* z2) {      // S1 <- (L, w - sw*w2, z - sw*z2)
        if (n1 >= MAXS) { retcode = RNDE_ERR_TAPE_FULL; return; }
        for (int e = tid; e < T; e += NT) {
            stW1[n1 * T + e] = w2 ? w[e] - sw * w2[e] : w[e];
            stZ1[n1 * T + e] = z2 ? z[e] - sw * z2[e] : z[e];
        }
        if (tid == 0) stL1[n1] = L;
        n1 += 1;
    };
    auto push2 = [&](double L, const float* w, const float* z) {
        if (n2 >= MAXS) { retcode = RNDE_ERR_TAPE_FULL; return; }
        for (int e = tid; e < T; e += NT) { stW2[n2 * T + e] = w[e]; stZ2[n2 * T + e] = z[e]; }
        if (tid == 0) stL2[n2] = L;
        n2 += 1;
    };
    // increments for a step of size dt after an accepted step (RSwM3.setup of the oracle)
    auto noise_setup = [&](double dt) {
        n2 = 0;
        __syncthreads();
        if (n1 == 0) {
            fresh(dt, sdW, sdZ, false);
            __syncthreads();
            push2(dt, sdW, sdZ);
        } else {
            double dttmp = 0.0;
            bool first = true;
            while (n1 > 0) {
                n1 -= 1;
                const double L1 = stL1[n1];
                const float* L2 = stW1 + n1 * T; const float* L3 = stZ1 + n1 * T;
                const double qtmp = (dt - dttmp) / L1;
                if (qtmp > 1.0) {
                    dttmp += L1;
                    for (int e = tid; e < T; e += NT) { sdW[e] = first ? L2[e] : sdW[e] + L2[e]; sdZ[e] = first ? L3[e] : sdZ[e] + L3[e]; }
                    push2(L1, L2, L3);
                    first = false;
                    __syncthreads();
                } else {
                    bridge(qtmp, L1, L2, L3, sTmp, sE2);      // bW, bZ in scratch tiles
                    __syncthreads();
                    for (int e = tid; e < T; e += NT) { sdW[e] = first ? sTmp[e] : sdW[e] + sTmp[e]; sdZ[e] = first ? sE2[e] : sdZ[e] + sE2[e]; }
                    first = false;
                    const int slot = n1;
                    if ((1.0 - qtmp) * L1 > 1e-15) {       // remainder of the piece stays a future (in place: same slot)
                        for (int e = tid; e < T; e += NT) { stW1[slot * T + e] = L2[e] - sTmp[e]; stZ1[slot * T + e] = L3[e] - sE2[e]; }
                        if (tid == 0) stL1[slot] = (1.0 - qtmp) * L1;
                        n1 += 1;
                    }
                    if (qtmp * L1 > 1e-15) push2(qtmp * L1, sTmp, sE2);
                    dttmp = dt;
                    __syncthreads();
                    break;
                }
            }
            const double left = dt - dttmp;
            if (left > 0.0) {
                fresh(left, sTmp, sE2, false);
                __syncthreads();
                for (int e = tid; e < T; e += NT) { sdW[e] = first ? sTmp[e] : sdW[e] + sTmp[e]; sdZ[e] = first ? sE2[e] : sdZ[e] + sE2[e]; }
                push2(left, sTmp, sE2);
            }
        }
        cur_dt = dt;
        __syncthreads();
    };
    // the attempt of size cur_dt was rejected: shrink to dtnew on the same Brownian path (RSwM3.reject of the oracle)
    auto noise_reject = [&](double dtnew) {
        const double qq = dtnew / cur_dt;
        double dttmp = 0.0;
        bool any = false;
        for (int e = tid; e < T; e += NT) { sTmp[e] = 0.f; sE2[e] = 0.f; }       // dWtmp, dZtmp
        __syncthreads();
        while (n2 > 0) {
            const double L1 = stL2[n2 - 1];
            if (dttmp + L1 < (1.0 - qq) * cur_dt) {
                n2 -= 1;
                const float* L2 = stW2 + n2 * T; const float* L3 = stZ2 + n2 * T;
                dttmp += L1;
                for (int e = tid; e < T; e += NT) { sTmp[e] = any ? sTmp[e] + L2[e] : L2[e]; sE2[e] = any ? sE2[e] + L3[e] : L3[e]; }
                any = true;
                push1(L1, L2, L3, 0.f, nullptr, nullptr);
                __syncthreads();
            } else break;
        }
        const double dtK = cur_dt - dttmp;
        // K2 = dW - dWtmp, K3 = dZ - dZtmp (in place in sdW / sdZ)
        if (any) for (int e = tid; e < T; e += NT) { sdW[e] = sdW[e] - sTmp[e]; sdZ[e] = sdZ[e] - sE2[e]; }
        __syncthreads();
        const double qK = qq * cur_dt / dtK;
        bridge(qK, dtK, sdW, sdZ, sTmp, sE2);       // bW, bZ
        __syncthreads();
        const double cut = (1.0 - qK) * dtK;
        if (cut > 1e-15) push1(cut, sdW, sdZ, 1.f, sTmp, sE2);
        __syncthreads();
        for (int e = tid; e < T; e += NT) { sdW[e] = sTmp[e]; sdZ[e] = sE2[e]; }
        n2 = 0;
        __syncthreads();
        push2(dtnew, sdW, sdZ);
        cur_dt = dtnew;
        __syncthreads();
    };

    const double order = 1.5;
    const double beta2 = 2.0 / (5.0 * order), beta1 = 7.0 / (10.0 * order);
    const double gamma = 0.9, qmin = 0.2, qmax = 9.0 / 8.0, qoldinit = 1e-4, delta = 1.0 / 26.0;
    const double dtmax = (double)P.t1 - (double)P.t0;
    const float atol = P.abstol, rtol = P.reltol;
    double t = (double)P.t0;
    int n_saved = 0, naccept = 0, nreject = 0;
    if (P.reg_kind != RNDE_REG_NONE) {
        if (q == 0 && tid == 0 && P.saveval) P.saveval[0] = (P.reg_kind == RNDE_REG_STIFF_SCALED) ? (float)(1.0 / 10.6) : 0.f;
        n_saved = 1;
    }

    // ---- sde_determine_initdt ----
    double dt;
    {
        eval_fg(sU, sK[0], sU, sG[0]);                    // f0, g(u0)
        double d01[2];
        norms([&](int e, float* o) {
            const float sk = atol + fabsf(sU[e]) * rtol;
            const float g0 = 3.f * sG[0][e], f0 = sK[0][e];
            o[0] = sU[e] / sk;
            o[1] = fmaxf(fabsf(f0 + g0), fabsf(f0 - g0)) / sk;
        }, std::integral_constant<int, 2>{}, d01);
        const double d0 = d01[0], d1 = d01[1];
        double dt0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * (d0 / d1);
        dt0 = fmin(dt0, dtmax);
        const float dt0f = (float)dt0;
        for (int e = tid; e < T; e += NT) sH0[e] = sU[e] + dt0f * sK[0][e];
        __syncthreads();
        eval_fg(sH0, sK[1], sH0, sG[1]);                  // f1, g(u1)
        double d2v[1];
        norms([&](int e, float* o) {
            const float sk = atol + fabsf(sU[e]) * rtol;
            const float g0 = 3.f * sG[0][e], g1 = 3.f * sG[1][e];
            const float dg = fmaxf(fabsf(g0 - g1), fabsf(g0 + g1));
            const float df = sK[1][e] - sK[0][e];
            o[0] = fmaxf(fabsf(df + dg), fabsf(df - dg)) / sk;
        }, std::integral_constant<int, 1>{}, d2v);
        const double d2 = d2v[0] / dt0;
        const double md = fmax(d1, d2);
        const double dt1 = (md <= 1e-15) ? fmax(1e-6, dt0 * 1e-3) : pow(10.0, -(2.0 + log10(md)) / (order + 0.5));
        dt = fmin(fmin(100.0 * dt0, dt1), dtmax);
        dt = (double)(float)dt;
    }
    const float dt_init = (float)dt;
    double qold = qoldinit, q11 = 1.0;
    dt = fmin(dt, (double)P.t1 - t);
    noise_setup(dt);
    int it = 0;
    float dt_last = 0.f;
    while (t < (double)P.t1 && retcode == RNDE_OK) {
        if (it >= P.max_steps) { retcode = RNDE_ERR_MAXITERS; break; }
        it += 1;
        const float dtc = (float)dt, sqdt = (float)sqrt(dt);
        const float sqrt3 = (float)sqrt(3.0);
        // chi1..3 per element are recomputed where needed (cheap) instead of stored
        auto chi1 = [&](int e) { const float w = sdW[e]; return (w * w - dtc) / (2.f * sqdt); };
        auto chi2 = [&](int e) { return (sdW[e] + sdZ[e] / sqrt3) / 2.f; };
        auto chi3 = [&](int e) { const float w = sdW[e]; return (w * w * w - 3.f * w * dtc) / (6.f * dtc); };
        eval_fg(sU, sK[0], sU, sG[0]);
        for (int e = tid; e < T; e += NT) {
            const float k1 = sK[0][e], g1 = sG[0][e], u = sU[e];
            sH0[e] = u + dtc * tb.a021 * k1 + tb.b021 * chi2(e) * g1;
            sH1[e] = u + dtc * tb.a121 * k1 + sqdt * tb.b121 * g1;
        }
        __syncthreads();
        eval_fg(sH0, sK[1], sH1, sG[1]);
        for (int e = tid; e < T; e += NT) {
            const float k1 = sK[0][e], k2 = sK[1][e], g1 = sG[0][e], g2 = sG[1][e], u = sU[e];
            sH0[e] = u + dtc * (tb.a031 * k1 + tb.a032 * k2) + chi2(e) * (tb.b031 * g1 + tb.b032 * g2);
            sH1[e] = u + dtc * (tb.a131 * k1 + tb.a132 * k2) + sqdt * (tb.b131 * g1 + tb.b132 * g2);
        }
        __syncthreads();
        eval_fg(sH0, sK[2], sH1, sG[2]);
        if (P.alg == 1) for (int e = tid; e < T; e += NT) sTmp[e] = sH0[e];       // H02 for the stiffness estimate
        for (int e = tid; e < T; e += NT) {
            const float k1 = sK[0][e], k2 = sK[1][e], k3 = sK[2][e], g1 = sG[0][e], g2 = sG[1][e], g3 = sG[2][e], u = sU[e];
            const float h0 = u + dtc * (tb.a041 * k1 + tb.a042 * k2 + tb.a043 * k3) + chi2(e) * (tb.b041 * g1 + tb.b042 * g2 + tb.b043 * g3);
            const float h1 = u + dtc * (tb.a141 * k1 + tb.a142 * k2 + tb.a143 * k3) + sqdt * (tb.b141 * g1 + tb.b142 * g2 + tb.b143 * g3);
            sH0[e] = h0; sH1[e] = h1;
        }
        __syncthreads();
        eval_fg(sH0, sK[3], sH1, sG[3]);
        for (int e = tid; e < T; e += NT) {
            const float k1 = sK[0][e], k2 = sK[1][e], k3 = sK[2][e], k4 = sK[3][e];
            const float g1 = sG[0][e], g2 = sG[1][e], g3 = sG[2][e], g4 = sG[3][e], u = sU[e];
            const float E2 = chi2(e) * (tb.be31 * g1 + tb.be32 * g2 + tb.be33 * g3 + tb.be34 * g4) +
                             chi3(e) * (tb.be41 * g1 + tb.be42 * g2 + tb.be43 * g3 + tb.be44 * g4);
            sE2[e] = E2;
            sUn[e] = u + dtc * (tb.al1 * k1 + tb.al2 * k2 + tb.al3 * k3 + tb.al4 * k4) + E2 +
                     sdW[e] * (tb.be11 * g1 + tb.be12 * g2 + tb.be13 * g3 + tb.be14 * g4) +
                     chi1(e) * (tb.be21 * g1 + tb.be22 * g2 + tb.be23 * g3 + tb.be24 * g4);
        }
        __syncthreads();
        double nv[3];
        const float deltaf = (float)delta;
        if (P.alg == 1) {
            norms([&](int e, float* o) {
                const float E1 = dtc * (sK[0][e] + sK[1][e] + sK[2][e] + sK[3][e]);
                o[0] = (deltaf * E1 + sE2[e]) / (atol + fmaxf(fabsf(sU[e]), fabsf(sUn[e])) * rtol);
                o[1] = sK[3][e] - sK[2][e];
                o[2] = sH0[e] - sTmp[e];
            }, std::integral_constant<int, 3>{}, nv);
        } else {
            norms([&](int e, float* o) {
                const float E1 = dtc * (sK[0][e] + sK[1][e] + sK[2][e] + sK[3][e]);
                o[0] = (deltaf * E1 + sE2[e]) / (atol + fmaxf(fabsf(sU[e]), fabsf(sUn[e])) * rtol);
            }, std::integral_constant<int, 1>{}, nv);
        }
        const double EEst = (double)(float)nv[0];
        const double eig = (P.alg == 1) ? nv[1] / fmax(nv[2], 1e-300) : 1.0;
        if (EEst != EEst) { retcode = RNDE_ERR_NAN; break; }
        double qv;
        if (EEst == 0.0) qv = 1.0 / qmax;
        else {
            q11 = pow(EEst, beta1);
            qv = q11 / pow(qold, beta2);
            qv = fmax(1.0 / qmax, fmin(1.0 / qmin, qv / gamma));
        }
        const bool accept = EEst <= 1.0;
        if (q == 0 && tid == 0 && P.log && it <= P.log_cap) { P.log[(it - 1) * 3] = (float)dt; P.log[(it - 1) * 3 + 1] = (float)EEst; P.log[(it - 1) * 3 + 2] = accept ? 1.f : 0.f; }
        if (accept) {
            naccept += 1;
            qold = fmax(EEst, qoldinit);
            t = t + dt;
            if (fabs((double)P.t1 - t) < 10.0 * 1.1920928955078125e-07 * fmax(fabs((double)P.t1), 1.0)) t = (double)P.t1;
            if (P.tape) {
                if (naccept <= P.tape_cap) {
                    float* tp = P.tape + (((size_t)(naccept - 1) * P.Q + q) * 3) * T;
                    for (int e = tid; e < T; e += NT) { tp[e] = sU[e]; tp[T + e] = sdW[e]; tp[2 * T + e] = sdZ[e]; }
                    if (q == 0 && tid == 0) {
                        float* ts = P.tape_steps + 4 * (naccept - 1);
                        ts[0] = dtc; ts[1] = (float)EEst; ts[2] = (P.alg == 1) ? (float)nv[1] : 0.f; ts[3] = (P.alg == 1) ? (float)nv[2] : 0.f;
                    }
                } else retcode = RNDE_ERR_TAPE_FULL;
                __syncthreads();
            }
            for (int e = tid; e < T; e += NT) sU[e] = sUn[e];
            dt_last = (float)dt;
            if (P.reg_kind != RNDE_REG_NONE) {
                if (n_saved < P.max_saved) {
                    float sv;
                    if (P.reg_kind == RNDE_REG_STIFF_SCALED) { const double a = fabs(eig); sv = (float)(((a == 0.0 || a != a) ? 0.0 : a) / 10.6); }
                    else sv = (float)(EEst * dt);
                    if (q == 0 && tid == 0 && P.saveval) P.saveval[n_saved] = sv;
                } else retcode = RNDE_ERR_TAPE_FULL;
                n_saved += 1;
            }
            __syncthreads();
            if (!(t < (double)P.t1)) break;
            const double dtnew = dt / qv;
            dt = (double)(float)fmin(dtmax, dtnew);
            dt = fmin(dt, (double)P.t1 - t);
            noise_setup(dt);
        } else {
            nreject += 1;
            double dtnew = dt / fmin(1.0 / qmin, q11 / gamma);
            dtnew = (double)(float)dtnew;
            noise_reject(dtnew);
            dt = dtnew;
        }
    }
    __syncthreads();
    for (int e = tid; e < T; e += NT) {
        const int r = e / NP, n = e - r * NP;
        if (n < Nloc) P.u_out[(size_t)D * (c0 + n) + r] = sU[e];
    }
    if (q == 0 && tid == 0) {
        SdeStats s;
        s.nfe1 = nfe1; s.nfe2 = nfe2; s.naccept = naccept; s.nreject = nreject; s.n_saved = n_saved; s.draws = draw; s.retcode = retcode; s.pad = 0;
        s.t_final = (float)t; s.dt_init = dt_init; s.dt_last = dt_last; s.pad2 = 0.f;
        *P.stats = s;
    }
}

}  // namespace rnde
