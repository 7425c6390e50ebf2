/*
 * rnde_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * PARITY UNPINNED: the reference's tests hold no golden vectors
 * (test/test_node.jl:1-89 is @code_warntype only) and no Julia toolchain exists
 * in this image, so this oracle could not be checked against the reference
 * itself.  It is a restatement of:
 *   - src/models/neural_ode.jl:48-144      (problem set-up, return tuple, nfe)
 *   - src/models/basic.jl:16-28            (TDChain: vcat(x, t) before every layer)
 *   - experiments/mnist_node.jl:41-54      (MLPDynamics: tanh on both layers)
 *   - experiments/mnist_node.jl:62-103     (the three regulariser closures)
 *   - test/test_node.jl:28-89              (func = EEst*dt, abs(eigen_est*dt))
 *   - OrdinaryDiffEq 5.50.0 / DiffEqBase 6.53.4 / DiffEqCallbacks 2.16.0
 *     (Manifest.toml:964,242-248,250; un-vendored) as recalled in
 *     Tsit5 step, RMS norm, PI controller,
 *     Hairer-Wanner initial dt, FSAL, nf accounting, SavingCallback, eigen_est,
 *     AutoSwitch.
 * It is validated against oracle/torch_oracle.py (autograd through the same
 * algorithm) in FP64, against analytic ODEs and scipy (tests/test_oracle_*.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this file's shared objects.
 *
 * Build: oracle/Makefile  ->  oracle/_build/liborc_f32.so, liborc_f64.so
 * Arithmetic: include/regnde_canon.h ("canonical order"); the FP32 build is the
 * bit-level specification the CUDA kernels are tested against.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "regnde_canon.h"

#ifdef ORC_F64
typedef double REAL;
#define R_FMA(a, b, c) __builtin_fma((a), (b), (c))
#define R_SQRT(a) __builtin_sqrt((a))
#define R_TANH(a) tanh((a))
#define R_ABS(a) fabs((a))
#else
typedef float REAL;
#define R_FMA(a, b, c) __builtin_fmaf((a), (b), (c))
#define R_SQRT(a) __builtin_sqrtf((a))
#define R_TANH(a) canon_tanhf((a))
#define R_ABS(a) fabsf((a))
#endif

#define ORC_OK 0
#define ORC_ERR_MAXITERS 1
#define ORC_ERR_DTMIN 2
#define ORC_ERR_NAN 3
#define ORC_ERR_ARG 4

enum { ACT_ID = 0, ACT_TANH = 1 };
enum { ALG_TSIT5 = 0, ALG_AUTO_TSIT5 = 1 };
enum { REG_NONE = 0, REG_ERR_DT = 1, REG_STIFF_DT_ABS = 2, REG_STIFF_SCALED = 3, REG_ERR_PLUS_STIFF = 4 };

typedef struct {
    int D, H, B;
    int act1, act2;
    int time_dep;
    int kblock1, kblock2;  /* canonical K blocking of layer 1/2 (<=0: whole K) */
    int alg;
    int reg_kind;
    int max_steps;         /* maxiters */
    int nthreads;          /* <=0: all */
    double t0, t1, abstol, reltol;
    double dtmin;
    /* forced step sequence (replay): if n_forced>0 the controller is bypassed:
     * attempt i uses forced_dt[i] and is accepted iff forced_accept[i]. */
    int n_forced;
    const double* forced_dt;
    const int* forced_accept;
    /* saveat (neural_ode.jl:79-108,146-180; DiffEqBase saveat semantics): sorted times in
     * [t0,t1]; a time equal to a step end copies u, others use the Tsit5 free interpolant; t0 in saveat saves u0 */
    int n_saveat;
    const double* saveat;
    /* chain field (n_layers > 0): Flux Chain of Dense layers without time input, width[l-1] -> width[l] with
     * width[-1] = width[n_layers-1] = D, optional elementwise tanh on the input first -- the Latent-ODE generator
     * dynamics Chain(x -> tanh.(x), Dense(20,50,tanh), ..., Dense(50,20,tanh)) (experiments/latent_ode.jl:109-121),
     * called as re(p)(u) because time_dep = false (neural_ode.jl:57).  H must hold max(width). */
    int n_layers;
    int width[8];
    int act[8];
    int pre_act;
    /* arith = 1 (FIXED24): the layer products of the 2-layer field are exact truncated fixed-point products -- the
     * arithmetic of the tensor-core forward stepper (csrc/fwd4x_kernel.cuh); see fixed24_* below.
     * arith = 2 (SPLITK): the fma-chain order of csrc/fwd4s_kernel.cuh; see rhs_eval_splitk. */
    int arith;
    /* FFJORD field (csq_extra = 1 or 3; forward here, reverse sweep in rnde_oracle_bwd.inc): the state
     * holds D - csq_extra data rows z plus [delta_logp (; ||f||^2; ||e^T J||^2)] (src/models/ffjord.jl:53-66); the field is
     * MLPDynamics(D - csq_extra, H) of three ConcatSquash layers with softplus between (experiments/ffjord_tabular.jl:47-105),
     * evaluated together with e^T J for the fixed Hutchinson noise csq_noise ((D - csq_extra) x B, REAL, column-major). */
    int csq_extra;
    const void* csq_noise;
} orc_config;

typedef struct {
    int nf, naccept, nreject, n_saved, retcode;
    double t_final, dt_last, dt_init;
} orc_stats;

typedef struct {
    orc_config cfg;
    size_t np;
    /* tape: per accepted step */
    int cap, nsteps;
    REAL* tp_t; REAL* tp_dt; REAL* tp_eest; REAL* tp_eig;
    REAL** tp_uprev; REAL** tp_k1;
    REAL* u0; REAL* p;
    /* attempted-step log */
    int log_cap, log_n; double* log_dt; int* log_acc; double* log_eest;
    REAL* saveval; int n_saved;
    REAL* usave;            /* n_saveat x (D*B) saved states */
    int* save_step;         /* accepted-step index each save belongs to (-1: t0) */
    REAL* save_theta;       /* theta of the save inside its step (1 = copied end state) */
    int n_usaved;
    orc_stats st;
    /* Appendix A.6 with detach_dt = all_but_first: the initial-dt heuristic stays on the tape (set_detach) */
    int first_dt_tracked;   /* 0: every dt frozen (default); 1: dt_1 = initial_dt(theta, x) differentiated */
    int last_clamped;       /* the last attempt's dt was tf - t (tracked through t when first_dt_tracked) */
    REAL id_d0, id_d1, id_d2, id_dt0, id_dt1;   /* scalars of the initial-dt heuristic, kept for its adjoint */
} orc_handle;

static size_t n_params(const orc_config* c) {
    int td = c->time_dep ? 1 : 0;
    if (c->csq_extra > 0) {
        const size_t Dz = (size_t)(c->D - c->csq_extra), H = (size_t)c->H;
        return (H * Dz + 4 * H) + (H * H + 4 * H) + (Dz * H + 4 * Dz);
    }
    if (c->n_layers > 0) {
        size_t n = 0;
        for (int l = 0; l < c->n_layers; ++l) { const int K = l ? c->width[l - 1] : c->D; n += (size_t)c->width[l] * K + c->width[l]; }
        return n;
    }
    return (size_t)c->H * (c->D + td) + c->H + (size_t)c->D * (c->H + td) + c->D;
}

/* ------------------------------------------------------------------ */
/* canonical reductions                                                */
/* ------------------------------------------------------------------ */
/* Canonical combination of per-block partial sums p[0..n): adjacent blocks are paired
 * first, pairs are then accumulated left to right:  ((p0+p1) + (p2+p3)) + (p4+p5) ...
 * (one CTA of the cluster-4 kernel owns two adjacent blocks). */
static REAL combine_blocks(const REAL* p, int n) {
    REAL tot = 0;
    for (int b = 0; b < n; b += 2) {
        REAL pair = (b + 1 < n) ? p[b] + p[b + 1] : p[b];
        tot = (b == 0) ? pair : tot + pair;
    }
    return tot;
}
/* per-column sum of squares of v[0..D) with row blocking kb: inside a block, groups of 4
 * consecutive rows form one fma chain (acc = fma(v,v,acc) from 0), group sums are added in
 * order, block sums are combined by combine_blocks. */
static REAL col_sumsq(const REAL* v, int D, int kb) {
    if (kb < 0) {       /* arith = 2 (SPLITK): blocks of -kb rows (one CTA each); groups of 4 rows added in order inside a block, blocks in order */
        REAL tot = 0;
        for (int r0 = 0, b = 0; r0 < D; r0 += -kb, ++b) {
            const int r1 = r0 - kb < D ? r0 - kb : D;
            REAL s = 0;
            for (int g0 = r0; g0 < r1; g0 += 4) {
                const int g1 = g0 + 4 < r1 ? g0 + 4 : r1;
                REAL q = 0;
                for (int r = g0; r < g1; ++r) q = R_FMA(v[r], v[r], q);
                s = (g0 == r0) ? q : s + q;
            }
            tot = (b == 0) ? s : tot + s;
        }
        return tot;
    }
    REAL part[64]; int nb = 0;
    for (int r0 = 0; r0 < D; r0 += kb) {
        int r1 = r0 + kb < D ? r0 + kb : D;
        REAL s = 0;
        for (int g0 = r0; g0 < r1; g0 += 4) {
            int g1 = g0 + 4 < r1 ? g0 + 4 : r1;
            REAL q = 0;
            for (int r = g0; r < g1; ++r) q = R_FMA(v[r], v[r], q);
            s = (g0 == r0) ? q : s + q;
        }
        part[nb++] = s;
    }
    return combine_blocks(part, nb);
}
/* total over columns: 32-way interleaved chains then xor-butterfly 16,8,4,2,1 */
static REAL cols_total(const REAL* q, int B) {
    REAL s[32];
    for (int l = 0; l < 32; ++l) s[l] = 0;
    for (int j = 0; j < B; ++j) s[j & 31] = s[j & 31] + q[j];
    for (int off = 16; off >= 1; off >>= 1) {
        REAL n[32];
        for (int l = 0; l < 32; ++l) n[l] = s[l] + s[l ^ off];
        memcpy(s, n, sizeof(s));
    }
    return s[0];
}
static REAL rms_from_total(REAL tot, long long count) { return R_SQRT(tot / (REAL)count); }

/* ------------------------------------------------------------------ */
/* vector field: y = act2(W2*[act1(W1*[z;t]+b1);t]+b2), per column       */
/* ------------------------------------------------------------------ */
static REAL act_apply(int act, REAL s) { return act == ACT_TANH ? R_TANH(s) : s; }

/* z: D x B (col-major), out k: D x B, optional hout: H x B */
/* chain field, one column: a_0 = pre(z); a_l = act_l(W_l a_{l-1} + b_l).  Canonical order: the inputs are cut into 4
 * contiguous quarters of ceil(K/4); each quarter is an fma chain in ascending order starting from 0 (an empty quarter
 * is +0); the quarters are added as (q0 + q1) + (q2 + q3); then + bias, then the activation.  (The kernel gives the
 * four quarters of an output to four adjacent lanes: csrc/chain.cuh quad_dense.)  acts (may be NULL) receives
 * a_0 .. a_{L-1} back to back (the last entry is the output). */
static void chain_column(const orc_config* c, const REAL* p, const REAL* zj, REAL* out, REAL* acts) {
    REAL a[2][1024];
    int cur = 0;
    const int D = c->D;
    for (int i = 0; i < D; ++i) a[0][i] = c->pre_act == ACT_TANH ? act_apply(ACT_TANH, zj[i]) : zj[i];
    size_t ao = 0;
    if (acts) { memcpy(acts, a[0], sizeof(REAL) * D); ao = D; }
    const REAL* W = p;
    int K = D;
    for (int l = 0; l < c->n_layers; ++l) {
        const int M = c->width[l];
        const REAL* b = W + (size_t)M * K;
        const int kb = (K + 3) / 4;
        for (int o = 0; o < M; ++o) {
            REAL qv[4];
            for (int blk = 0; blk < 4; ++blk) {
                const int i0 = blk * kb, i1 = i0 + kb < K ? i0 + kb : K;
                REAL acc = 0;
                for (int i = i0; i < i1; ++i) acc = R_FMA(W[(size_t)M * i + o], a[cur][i], acc);
                qv[blk] = acc;
            }
            const REAL tot = (qv[0] + qv[1]) + (qv[2] + qv[3]);
            a[cur ^ 1][o] = act_apply(c->act[l], tot + b[o]);
        }
        cur ^= 1;
        if (acts) { memcpy(acts + ao, a[cur], sizeof(REAL) * M); ao += M; }
        W = b + M; K = M;
    }
    memcpy(out, a[cur], sizeof(REAL) * D);
}

/* ------------------------------------------------------------------ */
/* FIXED24 layer arithmetic (arith = 1)                                 */
/* A dot product over one block of K inputs is evaluated as an exact integer expression, so that its value does not  */
/* depend on the order of summation (tensor cores, any tiling):                                                      */
/*   weights, per output row and block:  e_w = exponent(max_k |w|) (2^(e_w-1) <= max < 2^e_w), q_w = rint(w * 2^(22-e_w))  */
/*   inputs, per column and block:       e_x likewise over the block's rows,                q_x = rint(x * 2^(22-e_x))  */
/*   balanced base-256 digits q = d0*65536 + d1*256 + d2 (d1, d2 in [-128,127], |d0| <= 64)                           */
/*   I_ac = sum_k dw_a[k] * dx_c[k] (exact integers),  T = I_00*2^24 + (I_01+I_10)*2^16 + (I_02+I_11+I_20)*2^8 + (I_12+I_21)  */
/*   value = float(T) * 2^(e_w + e_x - 36)      -- only the digit product d2*d2 (2^-30 of full scale) is dropped          */
/* max < 2^-102 (incl. 0 and subnormals) quantises to 0.  Layer 1 has 4 blocks of D/4 rows (= the CTAs of a cluster),  */
/* combined as ((p0+p1)+p2)+p3; layer 2 one block of H.  Float32 build only.                                          */
/* ------------------------------------------------------------------ */
typedef struct { int valid; int D, H, R; int8_t* w1[3]; int* ew1; int8_t* w2[3]; int* ew2; } fixed24_weights;
static fixed24_weights g_f24 = {0};

static int f24_exponent(float amax) {      /* e with 2^(e-1) <= amax < 2^e; returns INT32_MIN when the block quantises to zero */
    uint32_t u; memcpy(&u, &amax, 4);
    const int eb = (int)((u >> 23) & 0xFF);
    if (eb < 24) return -2147483647 - 1;
    return eb - 126;
}
static float f24_pow2(int e) {             /* 2^e, clamped to the normal range (e < -126 gives 0) */
    if (e < -126) return 0.0f;
    if (e > 127) e = 127;
    uint32_t u = (uint32_t)(e + 127) << 23; float f; memcpy(&f, &u, 4); return f;
}
static void f24_digits(float x, int e, int8_t* d0, int8_t* d1, int8_t* d2) {
    if (e == -2147483647 - 1) { *d0 = *d1 = *d2 = 0; return; }
    const float sc = fminf(fmaxf(x * f24_pow2(22 - e), -8388607.0f), 8388607.0f);   /* never binds for finite data (|x| < 2^e) */
    int q = (int)lrintf(sc);
    const int b2 = ((q + 128) & 255) - 128; q = (q - b2) >> 8;
    const int b1 = ((q + 128) & 255) - 128; q = (q - b1) >> 8;
    *d0 = (int8_t)q; *d1 = (int8_t)b1; *d2 = (int8_t)b2;
}
static void fixed24_prepare(const orc_config* c, const REAL* p) {
    const int D = c->D, H = c->H, td = c->time_dep ? 1 : 0, R = D / 4;
    fixed24_weights* q = &g_f24;
    for (int a = 0; a < 3; ++a) { free(q->w1[a]); free(q->w2[a]); q->w1[a] = (int8_t*)malloc((size_t)H * D); q->w2[a] = (int8_t*)malloc((size_t)D * H); }
    free(q->ew1); free(q->ew2);
    q->ew1 = (int*)malloc(sizeof(int) * H * 4); q->ew2 = (int*)malloc(sizeof(int) * D);
    const REAL* W1 = p;
    const REAL* W2 = p + (size_t)H * (D + td) + H;
    for (int o = 0; o < H; ++o)
        for (int b = 0; b < 4; ++b) {
            float amax = 0;
            for (int k = b * R; k < (b + 1) * R; ++k) { const float a = fabsf((float)W1[(size_t)H * k + o]); if (a > amax) amax = a; }
            const int e = f24_exponent(amax);
            q->ew1[o * 4 + b] = e;
            for (int k = b * R; k < (b + 1) * R; ++k)
                f24_digits((float)W1[(size_t)H * k + o], e, &q->w1[0][(size_t)o * D + k], &q->w1[1][(size_t)o * D + k], &q->w1[2][(size_t)o * D + k]);
        }
    for (int r = 0; r < D; ++r) {
        float amax = 0;
        for (int k = 0; k < H; ++k) { const float a = fabsf((float)W2[(size_t)D * k + r]); if (a > amax) amax = a; }
        const int e = f24_exponent(amax);
        q->ew2[r] = e;
        for (int k = 0; k < H; ++k) f24_digits((float)W2[(size_t)D * k + r], e, &q->w2[0][(size_t)r * H + k], &q->w2[1][(size_t)r * H + k], &q->w2[2][(size_t)r * H + k]);
    }
    q->D = D; q->H = H; q->R = R; q->valid = 1;
}
/* exact block product: dw[3] point at K digits of one weight row, dx[3] at K digits of one input column */
static float f24_block(const int8_t* const dw[3], int ew, const int8_t* const dx[3], int ex, int K) {
    if (ew == -2147483647 - 1 || ex == -2147483647 - 1) return 0.0f;
    int I00 = 0, I01 = 0, I02 = 0, I10 = 0, I11 = 0, I12 = 0, I20 = 0, I21 = 0;
    for (int k = 0; k < K; ++k) {
        const int w0 = dw[0][k], w1 = dw[1][k], w2 = dw[2][k], x0 = dx[0][k], x1 = dx[1][k], x2 = dx[2][k];
        I00 += w0 * x0; I01 += w0 * x1; I02 += w0 * x2; I10 += w1 * x0; I11 += w1 * x1; I12 += w1 * x2; I20 += w2 * x0; I21 += w2 * x1;
    }
    const long long T = ((long long)I00 << 24) + ((long long)(I01 + I10) << 16) + ((long long)(I02 + I11 + I20) << 8) + (long long)(I12 + I21);
    return (float)T * f24_pow2(ew + ex - 36);
}
static void rhs_eval_fixed24(const orc_config* c, const REAL* p, const REAL* z, REAL t, REAL* k, REAL* hout) {
    const int D = c->D, H = c->H, B = c->B, td = c->time_dep ? 1 : 0, R = D / 4;
    const fixed24_weights* q = &g_f24;
    const REAL* W1 = p;
    const REAL* b1 = W1 + (size_t)H * (D + td);
    const REAL* W2 = b1 + H;
    const REAL* b2 = W2 + (size_t)D * (H + td);
#pragma omp parallel
    {
        int8_t* dz[3]; int8_t* dh[3];
        for (int a = 0; a < 3; ++a) { dz[a] = (int8_t*)malloc((size_t)D); dh[a] = (int8_t*)malloc((size_t)H); }
        float* hh = (float*)malloc(sizeof(float) * H);
#pragma omp for schedule(static)
        for (int j = 0; j < B; ++j) {
            const REAL* zj = z + (size_t)D * j;
            int ex[4];
            for (int b = 0; b < 4; ++b) {
                float amax = 0;
                for (int i = b * R; i < (b + 1) * R; ++i) { const float a = fabsf((float)zj[i]); if (a > amax) amax = a; }
                ex[b] = f24_exponent(amax);
                for (int i = b * R; i < (b + 1) * R; ++i) f24_digits((float)zj[i], ex[b], &dz[0][i], &dz[1][i], &dz[2][i]);
            }
            for (int o = 0; o < H; ++o) {
                float pb[4];
                for (int b = 0; b < 4; ++b) {
                    const int8_t* dw[3] = {q->w1[0] + (size_t)o * D + b * R, q->w1[1] + (size_t)o * D + b * R, q->w1[2] + (size_t)o * D + b * R};
                    const int8_t* dx[3] = {dz[0] + b * R, dz[1] + b * R, dz[2] + b * R};
                    pb[b] = f24_block(dw, q->ew1[o * 4 + b], dx, ex[b], R);
                }
                float v = ((pb[0] + pb[1]) + pb[2]) + pb[3];      /* the reducer CTA adds the four CTAs' partials in rank order */
                if (td) v = __builtin_fmaf((float)W1[(size_t)H * D + o], (float)t, v);
                v = v + (float)b1[o];
                hh[o] = (float)act_apply(c->act1, (REAL)v);
            }
            if (hout) for (int o = 0; o < H; ++o) hout[(size_t)H * j + o] = (REAL)hh[o];
            float amax = 0;
            for (int o = 0; o < H; ++o) { const float a = fabsf(hh[o]); if (a > amax) amax = a; }
            const int eh = f24_exponent(amax);
            for (int o = 0; o < H; ++o) f24_digits(hh[o], eh, &dh[0][o], &dh[1][o], &dh[2][o]);
            REAL* kj = k + (size_t)D * j;
            for (int r = 0; r < D; ++r) {
                const int8_t* dw[3] = {q->w2[0] + (size_t)r * H, q->w2[1] + (size_t)r * H, q->w2[2] + (size_t)r * H};
                const int8_t* dx[3] = {dh[0], dh[1], dh[2]};
                float v = f24_block(dw, q->ew2[r], dx, eh, H);
                if (td) v = __builtin_fmaf((float)W2[(size_t)D * H + r], (float)t, v);
                v = v + (float)b2[r];
                kj[r] = act_apply(c->act2, (REAL)v);
            }
        }
        for (int a = 0; a < 3; ++a) { free(dz[a]); free(dh[a]); }
        free(hh);
    }
}

/* ------------------------------------------------------------------ */
/* FFJORD field: ConcatSquash MLP and its transposed-Jacobian product    */
/* Canonical order (one column): every product W x and W^T u cuts its contraction index into 4 contiguous quarters of */
/* ceil(K/4), each an fma chain in ascending order from 0, combined as (q0+q1)+(q2+q3) -- the order of quad_dense in    */
/* csrc/chain.cuh.  A layer is r = fma(W x + B, g, fma(bW, t, bB)) with the gate g = sigmoid(G*t); the transposed chain  */
/* multiplies by the gate first (u = g .* v, then W^T u: the same value as (W .* g)^T v of ffjord_tabular.jl:71 with M    */
/* instead of M*K roundings); the row sums (trace, kinetic terms) are ascending fma chains from 0.                       */
/* ------------------------------------------------------------------ */
#ifdef ORC_F64
static REAL csq_sigmoid(REAL x) { const REAL e = exp(-fabs(x)); return x >= 0 ? 1.0 / (1.0 + e) : e / (1.0 + e); }
static REAL csq_softplus(REAL x) { const REAL l = log1p(exp(-fabs(x))); return x > 0 ? x + l : l; }
#else
static REAL csq_sigmoid(REAL x) { return canon_sigmoidf(x); }
static REAL csq_softplus(REAL x) { return canon_softplusf(x); }
#endif
static REAL quad_sum(const REAL* w, size_t wstride, const REAL* x, int K) {
    const int kb = (K + 3) / 4;
    REAL qv[4];
    for (int blk = 0; blk < 4; ++blk) {
        const int i0 = blk * kb, i1 = i0 + kb < K ? i0 + kb : K;
        REAL acc = 0;
        for (int i = i0; i < i1; ++i) acc = R_FMA(w[wstride * (size_t)i], x[i], acc);
        qv[blk] = acc;
    }
    return (qv[0] + qv[1]) + (qv[2] + qv[3]);
}
typedef struct { const REAL *W, *B, *bW, *bB, *G; int M, K; } csq_layer;
static const REAL* csq_take(const REAL* p, int M, int K, csq_layer* L) {
    L->M = M; L->K = K; L->W = p; L->B = p + (size_t)M * K; L->bW = L->B + M; L->bB = L->bW + M; L->G = L->bB + M;
    return L->G + M;
}
/* r = layer(x), g = its gates (kept for the transposed chain) */
static void csq_forward(const csq_layer* L, const REAL* x, REAL t, REAL* r, REAL* g) {
    for (int o = 0; o < L->M; ++o) {
        g[o] = csq_sigmoid(L->G[o] * t);
        const REAL lin = quad_sum(L->W + o, (size_t)L->M, x, L->K) + L->B[o];        /* W column-major M x K: row o has stride M */
        r[o] = R_FMA(lin, g[o], R_FMA(L->bW[o], t, L->bB[o]));
    }
}
/* out = W^T (g .* v) */
static void csq_back(const csq_layer* L, const REAL* g, const REAL* v, REAL* out, REAL* u) {
    for (int o = 0; o < L->M; ++o) u[o] = g[o] * v[o];
    for (int k = 0; k < L->K; ++k) out[k] = quad_sum(L->W + (size_t)L->M * k, 1, u, L->M);
}
static void csq_column(const orc_config* c, const REAL* p, const REAL* zj, const REAL* ej, REAL t, REAL* out) {
    const int X = c->csq_extra, Dz = c->D - X, H = c->H;
    csq_layer L1, L2, L3;
    p = csq_take(p, H, Dz, &L1); p = csq_take(p, H, H, &L2); csq_take(p, Dz, H, &L3);
    REAL r1[1024], r2[1024], g1[1024], g2[1024], g3[1024], v[1024], u[1024], eJ[1024];
    REAL a[1024] = {0}, w[1024] = {0};
    csq_forward(&L1, zj, t, r1, g1);
    for (int i = 0; i < H; ++i) a[i] = csq_softplus(r1[i]);
    csq_forward(&L2, a, t, r2, g2);
    for (int i = 0; i < H; ++i) a[i] = csq_softplus(r2[i]);
    csq_forward(&L3, a, t, out, g3);                           /* rows 0 .. Dz-1: f(z, t) */
    csq_back(&L3, g3, ej, v, u);                               /* H */
    for (int i = 0; i < H; ++i) w[i] = csq_sigmoid(r2[i]) * v[i];
    csq_back(&L2, g2, w, v, u);                                /* H */
    for (int i = 0; i < H; ++i) w[i] = csq_sigmoid(r1[i]) * v[i];
    csq_back(&L1, g1, w, eJ, u);                               /* Dz: e^T J */
    REAL tr = 0, f2 = 0, j2 = 0;
    for (int i = 0; i < Dz; ++i) { tr = R_FMA(eJ[i], ej[i], tr); f2 = R_FMA(out[i], out[i], f2); j2 = R_FMA(eJ[i], eJ[i], j2); }
    out[Dz] = -tr;
    if (X == 3) { out[Dz + 1] = f2; out[Dz + 2] = j2; }
}

/* ------------------------------------------------------------------ */
/* SPLITK layer arithmetic (arith = 2): the order of csrc/fwd4s_kernel.cuh (8x8 register tiles, packed fmas, the        */
/* contraction index dealt out to 8 / 4 lanes).  Layer 1: the state rows form 4 blocks of R = D/4 (one CTA each); inside  */
/* a block, chain s (s = 0..7) runs over the rows 8i + s, i = 0 .. ceil(R/8)-1, in ascending order from 0 (an index past   */
/* the block is the padded fma(0, 0, acc)); the 8 chains are added as the xor-butterfly tree                              */
/* ((c0+c1)+(c2+c3)) + ((c4+c5)+(c6+c7)); the blocks as ((p0+p1)+p2)+p3 (the reducer CTA, rank order).  Layer 2: chain s   */
/* (s = 0..3) runs over the hidden units 4i + s; (c0+c1)+(c2+c3).  Then time column, bias, activation as everywhere.      */
/* ------------------------------------------------------------------ */
static void rhs_eval_splitk(const orc_config* c, const REAL* p, const REAL* z, REAL t, REAL* k, REAL* hout) {
    const int D = c->D, H = c->H, B = c->B, td = c->time_dep ? 1 : 0, R = D / 4;
    const int KS1 = (R + 7) / 8, KS2 = (H + 3) / 4;
    const REAL* W1 = p;
    const REAL* b1 = W1 + (size_t)H * (D + td);
    const REAL* W2 = b1 + H;
    const REAL* b2 = W2 + (size_t)D * (H + td);
    const REAL zero = 0;
#pragma omp parallel
    {
        REAL* ch = (REAL*)malloc(sizeof(REAL) * 8 * (size_t)(H > D ? H : D));
        REAL* pb = (REAL*)malloc(sizeof(REAL) * 4 * (size_t)H);
        REAL* hh = (REAL*)malloc(sizeof(REAL) * (size_t)H);
#pragma omp for schedule(static)
        for (int j = 0; j < B; ++j) {
            const REAL* zj = z + (size_t)D * j;
            for (int b = 0; b < 4; ++b) {
                for (int e = 0; e < 8 * H; ++e) ch[e] = 0;
                for (int i = 0; i < KS1; ++i)
                    for (int s = 0; s < 8; ++s) {
                        const int kk = 8 * i + s;
                        REAL* acc = ch + (size_t)s * H;
                        if (kk < R) {
                            const REAL xv = zj[b * R + kk];
                            const REAL* w = W1 + (size_t)H * (b * R + kk);
                            for (int o = 0; o < H; ++o) acc[o] = R_FMA(w[o], xv, acc[o]);
                        } else {
                            for (int o = 0; o < H; ++o) acc[o] = R_FMA(zero, zero, acc[o]);
                        }
                    }
                for (int o = 0; o < H; ++o)
                    pb[(size_t)b * H + o] = ((ch[o] + ch[H + o]) + (ch[2 * H + o] + ch[3 * H + o])) + ((ch[4 * H + o] + ch[5 * H + o]) + (ch[6 * H + o] + ch[7 * H + o]));
            }
            for (int o = 0; o < H; ++o) {
                REAL v = ((pb[o] + pb[H + o]) + pb[2 * H + o]) + pb[3 * H + o];
                if (td) v = R_FMA(W1[(size_t)H * D + o], t, v);
                v = v + b1[o];
                hh[o] = act_apply(c->act1, v);
            }
            if (hout) memcpy(hout + (size_t)H * j, hh, sizeof(REAL) * H);
            for (int e = 0; e < 4 * D; ++e) ch[e] = 0;
            for (int i = 0; i < KS2; ++i)
                for (int s = 0; s < 4; ++s) {
                    const int kk = 4 * i + s;
                    REAL* acc = ch + (size_t)s * D;
                    if (kk < H) {
                        const REAL xv = hh[kk];
                        const REAL* w = W2 + (size_t)D * kk;
                        for (int o = 0; o < D; ++o) acc[o] = R_FMA(w[o], xv, acc[o]);
                    } else {
                        for (int o = 0; o < D; ++o) acc[o] = R_FMA(zero, zero, acc[o]);
                    }
                }
            REAL* kj = k + (size_t)D * j;
            for (int o = 0; o < D; ++o) {
                REAL v = (ch[o] + ch[D + o]) + (ch[2 * D + o] + ch[3 * D + o]);
                if (td) v = R_FMA(W2[(size_t)D * H + o], t, v);
                v = v + b2[o];
                kj[o] = act_apply(c->act2, v);
            }
        }
        free(ch); free(pb); free(hh);
    }
}

static void rhs_eval(const orc_config* c, const REAL* p, const REAL* z, REAL t, REAL* k, REAL* hout) {
    const int D = c->D, H = c->H, B = c->B, td = c->time_dep ? 1 : 0;
    if (c->csq_extra > 0) {
        (void)hout;
        const REAL* e = (const REAL*)c->csq_noise;
        const int Dz = D - c->csq_extra;
#pragma omp parallel for schedule(static)
        for (int j = 0; j < B; ++j) csq_column(c, p, z + (size_t)D * j, e + (size_t)Dz * j, t, k + (size_t)D * j);
        return;
    }
    if (c->arith == 1 && c->n_layers == 0 && sizeof(REAL) == 4) { rhs_eval_fixed24(c, p, z, t, k, hout); return; }
    if (c->arith == 2 && c->n_layers == 0) { rhs_eval_splitk(c, p, z, t, k, hout); return; }
    if (c->n_layers > 0) {
        (void)t; (void)hout;
#pragma omp parallel for schedule(static)
        for (int j = 0; j < B; ++j) chain_column(c, p, z + (size_t)D * j, k + (size_t)D * j, NULL);
        return;
    }
    const REAL* W1 = p;
    const REAL* b1 = W1 + (size_t)H * (D + td);
    const REAL* W2 = b1 + H;
    const REAL* b2 = W2 + (size_t)D * (H + td);
    const int kb1 = c->kblock1 > 0 ? c->kblock1 : D;
    const int kb2 = c->kblock2 > 0 ? c->kblock2 : H;
#pragma omp parallel
    {
        REAL* pacc = (REAL*)malloc(sizeof(REAL) * (size_t)(H > D ? H : D));
        REAL* s = (REAL*)malloc(sizeof(REAL) * (size_t)(H > D ? H : D));
        REAL* pairb = (REAL*)malloc(sizeof(REAL) * (size_t)(H > D ? H : D));
        REAL* hh = (REAL*)malloc(sizeof(REAL) * (size_t)H);
#pragma omp for schedule(static)
        for (int j = 0; j < B; ++j) {
            const REAL* zj = z + (size_t)D * j;
            /* layer 1 */
            /* blocks of kb1 inputs: fma chain per block; blocks combined pairwise-then-sequentially */
            for (int i0 = 0, b = 0; i0 < D; i0 += kb1, ++b) {
                int i1 = i0 + kb1 < D ? i0 + kb1 : D;
                for (int o = 0; o < H; ++o) pacc[o] = 0;
                for (int i = i0; i < i1; ++i) {
                    const REAL xv = zj[i];
                    const REAL* w = W1 + (size_t)H * i;
                    for (int o = 0; o < H; ++o) pacc[o] = R_FMA(w[o], xv, pacc[o]);
                }
                if ((b & 1) == 0) for (int o = 0; o < H; ++o) pairb[o] = pacc[o];
                else for (int o = 0; o < H; ++o) pairb[o] = pairb[o] + pacc[o];
                if ((b & 1) == 1 || i1 == D) {
                    if (b < 2) for (int o = 0; o < H; ++o) s[o] = pairb[o];
                    else for (int o = 0; o < H; ++o) s[o] = s[o] + pairb[o];
                }
            }
            for (int o = 0; o < H; ++o) {
                REAL v = s[o];
                if (td) v = R_FMA(W1[(size_t)H * D + o], t, v);
                v = v + b1[o];
                hh[o] = act_apply(c->act1, v);
            }
            if (hout) memcpy(hout + (size_t)H * j, hh, sizeof(REAL) * H);
            /* layer 2 */
            for (int i0 = 0, b = 0; i0 < H; i0 += kb2, ++b) {
                int i1 = i0 + kb2 < H ? i0 + kb2 : H;
                for (int o = 0; o < D; ++o) pacc[o] = 0;
                for (int i = i0; i < i1; ++i) {
                    const REAL xv = hh[i];
                    const REAL* w = W2 + (size_t)D * i;
                    for (int o = 0; o < D; ++o) pacc[o] = R_FMA(w[o], xv, pacc[o]);
                }
                if ((b & 1) == 0) for (int o = 0; o < D; ++o) pairb[o] = pacc[o];
                else for (int o = 0; o < D; ++o) pairb[o] = pairb[o] + pacc[o];
                if ((b & 1) == 1 || i1 == H) {
                    if (b < 2) for (int o = 0; o < D; ++o) s[o] = pairb[o];
                    else for (int o = 0; o < D; ++o) s[o] = s[o] + pairb[o];
                }
            }
            REAL* kj = k + (size_t)D * j;
            for (int o = 0; o < D; ++o) {
                REAL v = s[o];
                if (td) v = R_FMA(W2[(size_t)D * H + o], t, v);
                v = v + b2[o];
                kj[o] = act_apply(c->act2, v);
            }
        }
        free(pacc); free(s); free(pairb); free(hh);
    }
}

/* ------------------------------------------------------------------ */
/* one Tsit5 attempt                                                   */
/* ------------------------------------------------------------------ */
typedef struct {
    REAL *k[8];   /* k[1..7] */
    REAL *z[8];   /* z[2..7]; z[7] = u_new */
    REAL *h[8];   /* h[2..7] hidden activations of the stage evals (h[1] given by caller when needed) */
    REAL *utilde, *atmp, *colq;
} step_ws;

static const double A_[8][7] = {
    {0}, {0},
    {0, TS_A21},
    {0, TS_A31, TS_A32},
    {0, TS_A41, TS_A42, TS_A43},
    {0, TS_A51, TS_A52, TS_A53, TS_A54},
    {0, TS_A61, TS_A62, TS_A63, TS_A64, TS_A65},
    {0, TS_A71, TS_A72, TS_A73, TS_A74, TS_A75, TS_A76}};
static const double BT_[8] = {0, TS_BT1, TS_BT2, TS_BT3, TS_BT4, TS_BT5, TS_BT6, TS_BT7};
static const double C_[8] = {0, 0, TS_C1, TS_C2, TS_C3, TS_C4, 1.0, 1.0};

static REAL stage_time(REAL t, REAL dt, int i) {
    if (i >= 6) return t + dt;
    return R_FMA((REAL)C_[i], dt, t);
}

/* z_i = uprev + dt*(sum_j a_ij k_j): c = a_i1*k1; c = fma(a_ij,k_j,c); z = fma(dt,c,uprev).
 * stage 2 follows upstream's  a = dt*a21; uprev + a*k1. */
static void stage_combo(const orc_config* c, int i, REAL dt, const REAL* uprev, REAL* const* k, REAL* z) {
    const size_t n = (size_t)c->D * c->B;
    if (i == 2) {
        const REAL a = dt * (REAL)TS_A21;
#pragma omp parallel for schedule(static)
        for (size_t e = 0; e < n; ++e) z[e] = R_FMA(a, k[1][e], uprev[e]);
        return;
    }
    REAL a[7];
    for (int j = 1; j < i; ++j) a[j] = (REAL)A_[i][j];
#pragma omp parallel for schedule(static)
    for (size_t e = 0; e < n; ++e) {
        REAL s = a[1] * k[1][e];
        for (int j = 2; j < i; ++j) s = R_FMA(a[j], k[j][e], s);
        z[e] = R_FMA(dt, s, uprev[e]);
    }
}

/* performs the attempt; k[1] must hold fsalfirst.  Returns EEst, eigen_est. */
static void tsit5_attempt(const orc_config* c, const REAL* p, const REAL* uprev, REAL t, REAL dt, step_ws* w,
                          REAL* EEst, REAL* eig, int want_h) {
    const int D = c->D, B = c->B;
    const size_t n = (size_t)D * B;
    const long long cnt = (long long)D * B;
    const int kb = (c->arith == 2 && c->n_layers == 0) ? -(D / 4) : (c->kblock1 > 0 ? c->kblock1 : D);
    for (int i = 2; i <= 7; ++i) {
        stage_combo(c, i, dt, uprev, w->k, w->z[i]);
        rhs_eval(c, p, w->z[i], stage_time(t, dt, i), w->k[i], want_h ? w->h[i] : NULL);
    }
    if (c->alg == ALG_AUTO_TSIT5) {
        /* eigen_est = norm(k7-k6)/norm(u-g6)  (Appendix A.2) */
        REAL tot1, tot2;
#pragma omp parallel for schedule(static)
        for (int j = 0; j < B; ++j) {
            REAL* d = w->atmp + (size_t)D * j;
            for (int i = 0; i < D; ++i) d[i] = w->k[7][(size_t)D * j + i] - w->k[6][(size_t)D * j + i];
            w->colq[j] = col_sumsq(d, D, kb);
        }
        tot1 = cols_total(w->colq, B);
#pragma omp parallel for schedule(static)
        for (int j = 0; j < B; ++j) {
            REAL* d = w->atmp + (size_t)D * j;
            for (int i = 0; i < D; ++i) d[i] = w->z[7][(size_t)D * j + i] - w->z[6][(size_t)D * j + i];
            w->colq[j] = col_sumsq(d, D, kb);
        }
        tot2 = cols_total(w->colq, B);
        *eig = rms_from_total(tot1, cnt) / rms_from_total(tot2, cnt);
    } else {
        *eig = 1;
    }
    REAL bt[8];
    for (int i = 1; i <= 7; ++i) bt[i] = (REAL)BT_[i];
    const REAL atol = (REAL)c->abstol, rtol = (REAL)c->reltol;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < B; ++j) {
        for (int i = 0; i < D; ++i) {
            size_t e = (size_t)D * j + i;
            REAL s = bt[1] * w->k[1][e];
            for (int q = 2; q <= 7; ++q) s = R_FMA(bt[q], w->k[q][e], s);
            REAL ut = dt * s;
            REAL a0 = R_ABS(uprev[e]), a1 = R_ABS(w->z[7][e]);
            REAL m = a0 > a1 ? a0 : a1;
            REAL den = R_FMA(m, rtol, atol);
            w->utilde[e] = ut;
            w->atmp[e] = ut / den;
        }
        w->colq[j] = col_sumsq(w->atmp + (size_t)D * j, D, kb);
    }
    (void)n;
    *EEst = rms_from_total(cols_total(w->colq, B), cnt);
}

/* Tsit5 free interpolant weights b_i(theta), i = 1..7 (Appendix A.9), Horner with fma */
static void interp_weights(REAL th, REAL* b) {
    const REAL th2 = th * th;
    b[1] = th * R_FMA(th, R_FMA(th, R_FMA(th, (REAL)TS_R14, (REAL)TS_R13), (REAL)TS_R12), (REAL)TS_R11);
    b[2] = th2 * R_FMA(th, R_FMA(th, (REAL)TS_R24, (REAL)TS_R23), (REAL)TS_R22);
    b[3] = th2 * R_FMA(th, R_FMA(th, (REAL)TS_R34, (REAL)TS_R33), (REAL)TS_R32);
    b[4] = th2 * R_FMA(th, R_FMA(th, (REAL)TS_R44, (REAL)TS_R43), (REAL)TS_R42);
    b[5] = th2 * R_FMA(th, R_FMA(th, (REAL)TS_R54, (REAL)TS_R53), (REAL)TS_R52);
    b[6] = th2 * R_FMA(th, R_FMA(th, (REAL)TS_R64, (REAL)TS_R63), (REAL)TS_R62);
    b[7] = th2 * R_FMA(th, R_FMA(th, (REAL)TS_R74, (REAL)TS_R73), (REAL)TS_R72);
}

/* ------------------------------------------------------------------ */
/* initial dt (Hairer-Wanner, Appendix A.5)                            */
/* ------------------------------------------------------------------ */
static REAL initial_dt(const orc_config* c, const REAL* p, const REAL* u0, const REAL* f0, REAL t0, REAL dtmax, REAL* scratch /*3*D*B*/,
                       REAL* colq, REAL* keep /* d0, d1, d2, dt0, dt1 */) {
    const int D = c->D, B = c->B;
    const long long cnt = (long long)D * B;
    const int kb = (c->arith == 2 && c->n_layers == 0) ? -(D / 4) : (c->kblock1 > 0 ? c->kblock1 : D);
    const REAL atol = (REAL)c->abstol, rtol = (REAL)c->reltol;
    REAL* tmp = scratch; REAL* u1 = scratch + (size_t)D * B; REAL* f1 = scratch + 2 * (size_t)D * B;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < B; ++j) {
        for (int i = 0; i < D; ++i) { size_t e = (size_t)D * j + i; tmp[e] = u0[e] / R_FMA(R_ABS(u0[e]), rtol, atol); }
        colq[j] = col_sumsq(tmp + (size_t)D * j, D, kb);
    }
    REAL d0 = rms_from_total(cols_total(colq, B), cnt);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < B; ++j) {
        for (int i = 0; i < D; ++i) { size_t e = (size_t)D * j + i; tmp[e] = f0[e] / R_FMA(R_ABS(u0[e]), rtol, atol); }
        colq[j] = col_sumsq(tmp + (size_t)D * j, D, kb);
    }
    REAL d1 = rms_from_total(cols_total(colq, B), cnt);
    REAL dt0;
    if (d0 < (REAL)1e-5 || d1 < (REAL)1e-5) dt0 = (REAL)1e-6;
    else dt0 = (d0 / d1) / (REAL)100;
    if (dt0 > dtmax) dt0 = dtmax;
    const size_t n = (size_t)D * B;
#pragma omp parallel for schedule(static)
    for (size_t e = 0; e < n; ++e) u1[e] = R_FMA(dt0, f0[e], u0[e]);
    rhs_eval(c, p, u1, t0 + dt0, f1, NULL);
#pragma omp parallel for schedule(static)
    for (int j = 0; j < B; ++j) {
        for (int i = 0; i < D; ++i) { size_t e = (size_t)D * j + i; tmp[e] = (f1[e] - f0[e]) / R_FMA(R_ABS(u0[e]), rtol, atol); }
        colq[j] = col_sumsq(tmp + (size_t)D * j, D, kb);
    }
    REAL d2 = rms_from_total(cols_total(colq, B), cnt) / dt0;
    REAL md = d1 > d2 ? d1 : d2;
    REAL dt1;
    if (md <= (REAL)1e-15) {
        REAL a = dt0 * (REAL)1e-3;
        dt1 = a > (REAL)1e-6 ? a : (REAL)1e-6;
    } else {
#ifdef ORC_F64
        double l10 = canon_log2(md) * 0.30102999566398120;
        dt1 = canon_exp10(-(2.0 + l10) / 5.0);
#else
        float l10 = canon_log10f(md);
        float ex = -(2.0f + l10) / 5.0f;
        dt1 = (float)canon_exp10((double)ex);
#endif
    }
    REAL dt = (REAL)100 * dt0;
    if (dt1 < dt) dt = dt1;
    if (dtmax < dt) dt = dtmax;
    if (dt < (REAL)c->dtmin) dt = (REAL)c->dtmin;
    if (keep) { keep[0] = d0; keep[1] = d1; keep[2] = d2; keep[3] = dt0; keep[4] = dt1; }
    return dt;
}

/* ------------------------------------------------------------------ */
/* saved value (the reference's func closures)                         */
/* ------------------------------------------------------------------ */
static REAL saved_value(int kind, REAL EEst, REAL eig, REAL dt) {
    const REAL stab = (REAL)1 / (REAL)(float)TS_STABILITY_SIZE;
    switch (kind) {
        case REG_ERR_DT: return EEst * dt;                        /* neural_ode.jl:116, mnist_node.jl:67 */
        case REG_STIFF_DT_ABS: return R_ABS(eig * dt);            /* test_node.jl:75 */
        case REG_STIFF_SCALED: {                                  /* mnist_node.jl:76-79 */
            REAL s = R_ABS(eig);
            return stab * ((s == 0 || s != s) ? (REAL)0 : s);
        }
        case REG_ERR_PLUS_STIFF: {                                /* mnist_node.jl:88-97 */
            REAL e = EEst * dt;
            REAL a = (e == 0 || e != e) ? (REAL)0 : e;
            REAL b = (eig == 0 || eig != eig) ? (REAL)0 : eig;
            return (a + ((REAL)0.1f * stab) * b) * (REAL)1;
        }
        default: return 0;
    }
}

/* ------------------------------------------------------------------ */
/* handle                                                              */
/* ------------------------------------------------------------------ */
#ifdef ORC_F64
#define FN(name) orc64_##name
#else
#define FN(name) orc32_##name
#endif

int FN(create)(const orc_config* cfg, void** out) {
    if (!cfg || cfg->D <= 0 || cfg->H <= 0 || cfg->B <= 0) return ORC_ERR_ARG;
    if (cfg->kblock1 > 0 && (cfg->D + cfg->kblock1 - 1) / cfg->kblock1 > 64) return ORC_ERR_ARG;   /* col_sumsq part[64] */
    if (cfg->arith == 2 && (cfg->D % 4 != 0 || cfg->n_layers > 0 || cfg->csq_extra != 0)) return ORC_ERR_ARG;
    if (cfg->csq_extra != 0 && ((cfg->csq_extra != 1 && cfg->csq_extra != 3) || cfg->D <= cfg->csq_extra || !cfg->csq_noise ||
                                cfg->n_layers > 0 || cfg->arith != 0 || cfg->D > 1024 || cfg->H > 1024)) return ORC_ERR_ARG;
    orc_handle* h = (orc_handle*)calloc(1, sizeof(orc_handle));
    h->cfg = *cfg;
    if (h->cfg.max_steps <= 0) h->cfg.max_steps = 1000000;
    if (h->cfg.dtmin <= 0) h->cfg.dtmin = 1e-10;
    h->np = n_params(cfg);
    h->cap = 0; h->nsteps = 0;
    *out = h;
    return ORC_OK;
}

static void tape_clear(orc_handle* h) {
    for (int i = 0; i < h->nsteps; ++i) { free(h->tp_uprev[i]); free(h->tp_k1[i]); }
    h->nsteps = 0;
}
static void tape_push(orc_handle* h, REAL t, REAL dt, REAL eest, REAL eig, const REAL* uprev, const REAL* k1) {
    if (h->nsteps == h->cap) {
        int nc = h->cap ? 2 * h->cap : 64;
        h->tp_t = (REAL*)realloc(h->tp_t, sizeof(REAL) * nc);
        h->tp_dt = (REAL*)realloc(h->tp_dt, sizeof(REAL) * nc);
        h->tp_eest = (REAL*)realloc(h->tp_eest, sizeof(REAL) * nc);
        h->tp_eig = (REAL*)realloc(h->tp_eig, sizeof(REAL) * nc);
        h->tp_uprev = (REAL**)realloc(h->tp_uprev, sizeof(REAL*) * nc);
        h->tp_k1 = (REAL**)realloc(h->tp_k1, sizeof(REAL*) * nc);
        h->saveval = (REAL*)realloc(h->saveval, sizeof(REAL) * (nc + 1));
        h->cap = nc;
    }
    size_t n = (size_t)h->cfg.D * h->cfg.B;
    int i = h->nsteps++;
    h->tp_t[i] = t; h->tp_dt[i] = dt; h->tp_eest[i] = eest; h->tp_eig[i] = eig;
    h->tp_uprev[i] = (REAL*)malloc(sizeof(REAL) * n); memcpy(h->tp_uprev[i], uprev, sizeof(REAL) * n);
    h->tp_k1[i] = (REAL*)malloc(sizeof(REAL) * n); memcpy(h->tp_k1[i], k1, sizeof(REAL) * n);
}
static void log_push(orc_handle* h, double dt, int acc, double eest) {
    if (h->log_n == h->log_cap) {
        int nc = h->log_cap ? 2 * h->log_cap : 128;
        h->log_dt = (double*)realloc(h->log_dt, sizeof(double) * nc);
        h->log_acc = (int*)realloc(h->log_acc, sizeof(int) * nc);
        h->log_eest = (double*)realloc(h->log_eest, sizeof(double) * nc);
        h->log_cap = nc;
    }
    h->log_dt[h->log_n] = dt; h->log_acc[h->log_n] = acc; h->log_eest[h->log_n] = eest; h->log_n++;
}

void FN(destroy)(void* hv) {
    orc_handle* h = (orc_handle*)hv;
    if (!h) return;
    tape_clear(h);
    free(h->tp_t); free(h->tp_dt); free(h->tp_eest); free(h->tp_eig); free(h->tp_uprev); free(h->tp_k1);
    free(h->u0); free(h->p); free(h->log_dt); free(h->log_acc); free(h->log_eest); free(h->saveval);
    free(h->usave); free(h->save_step); free(h->save_theta);
    free(h);
}

static step_ws* ws_alloc(const orc_config* c) {
    step_ws* w = (step_ws*)calloc(1, sizeof(step_ws));
    size_t n = (size_t)c->D * c->B, nh = (size_t)c->H * c->B;
    for (int i = 1; i <= 7; ++i) { w->k[i] = (REAL*)malloc(sizeof(REAL) * n); w->z[i] = (REAL*)malloc(sizeof(REAL) * n); w->h[i] = (REAL*)malloc(sizeof(REAL) * nh); }
    w->utilde = (REAL*)malloc(sizeof(REAL) * n); w->atmp = (REAL*)malloc(sizeof(REAL) * n);
    w->colq = (REAL*)malloc(sizeof(REAL) * c->B);
    return w;
}
static void ws_free(step_ws* w) {
    for (int i = 1; i <= 7; ++i) { free(w->k[i]); free(w->z[i]); free(w->h[i]); }
    free(w->utilde); free(w->atmp); free(w->colq); free(w);
}

/* forward solve.  u_out: D x B.  saveval_out (may be NULL) gets n_saved values. */
int FN(forward)(void* hv, const REAL* x, const REAL* p, REAL* u_out, orc_stats* st_out) {
    orc_handle* h = (orc_handle*)hv;
    const orc_config* c = &h->cfg;
#ifdef _OPENMP
    if (c->nthreads > 0) omp_set_num_threads(c->nthreads);
#endif
    const int D = c->D, B = c->B;
    const size_t n = (size_t)D * B;
    tape_clear(h);
    h->log_n = 0;
    free(h->u0); free(h->p);
    h->u0 = (REAL*)malloc(sizeof(REAL) * n); memcpy(h->u0, x, sizeof(REAL) * n);
    h->p = (REAL*)malloc(sizeof(REAL) * h->np); memcpy(h->p, p, sizeof(REAL) * h->np);
    if (!h->saveval) { h->saveval = (REAL*)malloc(sizeof(REAL) * 65); }
    if (c->arith == 1 && c->n_layers == 0 && sizeof(REAL) == 4) fixed24_prepare(c, h->p);
    step_ws* w = ws_alloc(c);
    REAL* u = (REAL*)malloc(sizeof(REAL) * n);
    REAL* scratch = (REAL*)malloc(sizeof(REAL) * 3 * n);
    memcpy(u, x, sizeof(REAL) * n);
    orc_stats st; memset(&st, 0, sizeof(st));

    const REAL t0 = (REAL)c->t0, tf = (REAL)c->t1;
    REAL t = t0;
    free(h->usave); free(h->save_step); free(h->save_theta);
    h->usave = NULL; h->save_step = NULL; h->save_theta = NULL; h->n_usaved = 0;
    int save_idx = 0;
    if (c->n_saveat > 0) {
        h->usave = (REAL*)calloc((size_t)c->n_saveat * n, sizeof(REAL));
        h->save_step = (int*)malloc(sizeof(int) * c->n_saveat);
        h->save_theta = (REAL*)malloc(sizeof(REAL) * c->n_saveat);
        while (save_idx < c->n_saveat && (REAL)c->saveat[save_idx] <= t0) {
            memcpy(h->usave + (size_t)save_idx * n, x, sizeof(REAL) * n);
            h->save_step[save_idx] = -1; h->save_theta[save_idx] = 0; save_idx++;
        }
    }
    const REAL dtmax = tf - t0;
    /* controller constants (Appendix A.4), converted once to REAL */
    const REAL gamma = (REAL)(9.0 / 10.0), qmin = (REAL)(1.0 / 5.0), qmax = (REAL)10;
    const REAL beta1 = (REAL)(7.0 / 50.0), beta2 = (REAL)(2.0 / 25.0), qoldinit = (REAL)1e-4;
    REAL qold = qoldinit, q11 = 1;
    /* SavingCallback initial entry: EEst=1, dt=0, eigen_est=1 (Appendix A.7) */
    h->n_saved = 0;
    if (c->reg_kind != REG_NONE) h->saveval[h->n_saved++] = saved_value(c->reg_kind, (REAL)1, (REAL)1, (REAL)0);
    /* initialize!: fsalfirst */
    rhs_eval(c, p, u, t, w->k[1], NULL); st.nf += 1;
    REAL dt;
    if (c->n_forced > 0) dt = (REAL)c->forced_dt[0];
    else { REAL keep[5]; dt = initial_dt(c, p, u, w->k[1], t, dtmax, scratch, w->colq, keep);
           h->id_d0 = keep[0]; h->id_d1 = keep[1]; h->id_d2 = keep[2]; h->id_dt0 = keep[3]; h->id_dt1 = keep[4]; }
    h->last_clamped = 0;
    st.nf += 2;
    st.dt_init = dt;
    /* AutoSwitch state (Appendix A.8) */
    int as_count = 0, as_stiff = 0;
    REAL eig_prev = 1;
    int iter = 0, accept_prev = 1;
    REAL dtpropose = dt;
    int rc = ORC_OK;
    while (t < tf) {
        if (iter >= c->max_steps) { rc = ORC_ERR_MAXITERS; break; }
        /* loopheader! */
        if (iter > 0) {
            if (accept_prev) dt = dtpropose;
            else if (c->n_forced == 0) {
                REAL f = q11 / gamma, lim = (REAL)1 / qmin;
                dt = dt / (lim < f ? lim : f);
            }
        }
        iter++;
        if (c->alg == ALG_AUTO_TSIT5 && c->n_forced == 0) {
            REAL stiffness = R_ABS(eig_prev * dt / (REAL)TS_STABILITY_SIZE);
            int stiff = stiffness > (REAL)(9.0 / 10.0);
            as_count = stiff ? (as_count < 0 ? 1 : as_count + 1) : (as_count > 0 ? -1 : as_count - 1);
            if (!as_stiff && as_count > 10) { dt = dt * (REAL)2; as_stiff = 1; st.nf += 1; }
            else if (as_stiff && as_count < -3) { dt = dt / (REAL)2; as_stiff = 0; st.nf += 1; }
        }
        if (c->n_forced > 0) {
            if (iter - 1 >= c->n_forced) { rc = ORC_ERR_ARG; break; }
            dt = (REAL)c->forced_dt[iter - 1];
        } else {
            /* fix_dt_at_bounds!, modify_dt_for_tstops! */
            if (dt > dtmax) dt = dtmax;
            if (dt < (REAL)c->dtmin) dt = (REAL)c->dtmin;
            REAL rem = tf - t;
            h->last_clamped = rem < dt;
            if (rem < dt) dt = rem;
        }
        REAL EEst, eig;
        tsit5_attempt(c, p, u, t, dt, w, &EEst, &eig, 0);
        st.nf += 6;
        if (EEst != EEst) { rc = ORC_ERR_NAN; log_push(h, dt, 0, EEst); break; }
        /* loopfooter!: stepsize_controller! */
        REAL q;
        if (EEst == 0) q = (REAL)1 / qmax;
        else {
#ifdef ORC_F64
            q11 = canon_pow(EEst, beta1);
            q = q11 / canon_pow(qold, beta2);
#else
            q11 = canon_powf(EEst, beta1);
            q = q11 / canon_powf(qold, beta2);
#endif
            REAL qq = q / gamma, hi = (REAL)1 / qmin, lo = (REAL)1 / qmax;
            qq = hi < qq ? hi : qq;
            q = lo > qq ? lo : qq;
        }
        int accept = c->n_forced > 0 ? c->forced_accept[iter - 1] : (EEst <= (REAL)1);
        log_push(h, dt, accept, EEst);
        if (c->alg == ALG_AUTO_TSIT5) eig_prev = eig;
        if (accept) {
            st.naccept++;
            tape_push(h, t, dt, EEst, eig, u, w->k[1]);
            if (q >= (REAL)1 && q <= (REAL)1) q = 1;   /* qsteady_min = qsteady_max = 1 */
            qold = EEst > qoldinit ? EEst : qoldinit;
            REAL dtnew = dt / q;
            const REAL tprev = t;
            t = t + dt;
            /* savevalues!: every pending saveat time <= t */
            while (save_idx < c->n_saveat && (REAL)c->saveat[save_idx] <= t) {
                const REAL tau = (REAL)c->saveat[save_idx];
                REAL* dst = h->usave + (size_t)save_idx * n;
                h->save_step[save_idx] = st.naccept - 1;
                if (tau == t) { memcpy(dst, w->z[7], sizeof(REAL) * n); h->save_theta[save_idx] = 1; }
                else {
                    const REAL th = (tau - tprev) / dt;
                    REAL b[8]; interp_weights(th, b);
                    h->save_theta[save_idx] = th;
#pragma omp parallel for schedule(static)
                    for (size_t e = 0; e < n; ++e) {
                        REAL sacc = b[1] * w->k[1][e];
                        for (int i = 2; i <= 7; ++i) sacc = R_FMA(b[i], w->k[i][e], sacc);
                        dst[e] = R_FMA(dt, sacc, u[e]);
                    }
                }
                save_idx++;
            }
            dtpropose = dtnew < dtmax ? dtnew : dtmax;
            if (dtpropose < (REAL)c->dtmin) dtpropose = (REAL)c->dtmin;
            memcpy(u, w->z[7], sizeof(REAL) * n);
            REAL* tmp = w->k[1]; w->k[1] = w->k[7]; w->k[7] = tmp;   /* FSAL */
            if (c->reg_kind != REG_NONE) h->saveval[h->n_saved++] = saved_value(c->reg_kind, EEst, eig, dt);
            st.dt_last = dt;
        } else {
            st.nreject++;
            if (dt <= (REAL)c->dtmin) { rc = ORC_ERR_DTMIN; break; }
        }
        accept_prev = accept;
    }
    memcpy(u_out, u, sizeof(REAL) * n);
    h->n_usaved = save_idx;
    st.t_final = t; st.n_saved = h->n_saved; st.retcode = rc;
    h->st = st;
    if (st_out) *st_out = st;
    ws_free(w); free(u); free(scratch);
    return rc;
}

int FN(get_saveval)(void* hv, REAL* out, int cap) {
    orc_handle* h = (orc_handle*)hv;
    int n = h->n_saved < cap ? h->n_saved : cap;
    memcpy(out, h->saveval, sizeof(REAL) * n);
    return h->n_saved;
}
/* saved states: out is n_saveat x (D*B); returns the number actually saved */
int FN(get_usave)(void* hv, REAL* out) {
    orc_handle* h = (orc_handle*)hv;
    if (h->n_usaved > 0) memcpy(out, h->usave, sizeof(REAL) * (size_t)h->n_usaved * h->cfg.D * h->cfg.B);
    return h->n_usaved;
}
int FN(get_log)(void* hv, double* dt, int* acc, double* eest, int cap) {
    orc_handle* h = (orc_handle*)hv;
    int n = h->log_n < cap ? h->log_n : cap;
    if (dt) memcpy(dt, h->log_dt, sizeof(double) * n);
    if (acc) memcpy(acc, h->log_acc, sizeof(int) * n);
    if (eest) memcpy(eest, h->log_eest, sizeof(double) * n);
    return h->log_n;
}
int FN(get_step)(void* hv, int j, double* t, double* dt, double* eest, double* eig) {
    orc_handle* h = (orc_handle*)hv;
    if (j < 0 || j >= h->nsteps) return ORC_ERR_ARG;
    *t = h->tp_t[j]; *dt = h->tp_dt[j]; *eest = h->tp_eest[j]; *eig = h->tp_eig[j];
    return ORC_OK;
}

/* field evaluation exported for unit tests */
int FN(rhs)(const orc_config* cfg, const REAL* p, const REAL* z, double t, REAL* k, REAL* hout) {
    if (cfg->arith == 1 && cfg->n_layers == 0 && sizeof(REAL) == 4) fixed24_prepare(cfg, p);
    rhs_eval(cfg, p, z, (REAL)t, k, hout);
    return ORC_OK;
}

/* The adjoint is compiled twice: with cotangents in REAL (the plain FP32/FP64 adjoint) and, for
 * the FP32 build, with cotangents in double over the SAME FP32 forward values ("truth" for
 * judging FP32 adjoints: the regulariser gradient cancels ~1e7-sized cotangents). */
#define ADJ REAL
#define VJP_FN rhs_vjp
#define BWD_FN FN(backward)
#include "rnde_oracle_bwd.inc"
#undef ADJ
#undef VJP_FN
#undef BWD_FN
#ifndef ORC_F64
#define ADJ double
#define VJP_FN rhs_vjp_hi
#define BWD_FN FN(backward_hi)
#include "rnde_oracle_bwd.inc"
#undef ADJ
#undef VJP_FN
#undef BWD_FN
#endif

/* Appendix A.6: 0 = every dt frozen, 1 = the first dt (initial-dt heuristic) stays on the tape (recalled upstream default),
 * 2 = diagnostic: the backward returns the first-dt term alone */
int FN(set_detach)(void* hv, int first_dt_tracked) { ((orc_handle*)hv)->first_dt_tracked = first_dt_tracked; return ORC_OK; }
int FN(sizeof_real)(void) { return (int)sizeof(REAL); }
