// fwd4s_kernel.cuh -- "cluster-4 / split-K" forward stepper: the north-star kernel of round 2 for the MNIST-shaped
// field (RNDE_ARITH_SPLITK; D = 4R with 128 < R <= 256, H <= 128).
//
// Same algorithm, cluster decomposition, hidden exchange (st.async + mbarrier), tape traffic (bulk stores) and on-device
// controller as fwd4_kernel.cuh; what changes is the inner product arithmetic, chosen from measurements on B200
// (tools/microbench3.cu, tools/microbench4.cu, profiles/r2a_microbench4.txt):
//   * fwd4_kernel's 4x4 register tile needs two LDS.128 per 16 FMA; the shared-memory return path (128 B/clk/SM) caps
//     that shape at ~50 % of the FFMA pipe and the kernel sat on the cap (layer phases at 61 / 57 lane-FMA/clk).
//   * here every thread owns an 8x8 tile (4 LDS.128 per 64 FMA = 1 B/FMA) and issues the FMAs as packed FFMA2
//     (fma.rn.f32x2: two IEEE fmas per instruction, the broadcast operand taken from a single register), which halves the
//     issue slots of the phase.  Only 3136 outputs x 16 columns exist per CTA and layer, so the contraction index is dealt
//     out to 8 lanes (layer 1, K = R) / 4 lanes (layer 2, K = H) of the same warp -- interleaved: lane s takes k = S*i + s --
//     and the partial tiles are combined by an xor-shuffle reduce-scatter that leaves every lane with the finished sums
//     of ITS state tile: 4 rows x 4 columns (layer 2) or one hidden row x 8 columns (layer 1, sent straight to the
//     reducer CTA).  Lane bits: b0 = s&1, b1 = column group, b2 = row-group bit 0, b3 (b4) = the other bits of s: the eight
//     lanes of a quarter warp read two rows x two row groups of the weights (pitch = 4 mod 32) and two rows x two column
//     groups of the inputs -- conflict free with dense [k][16] input tiles, which are also the tape tiles.
//   * the state (k1..k7) stays in registers, 16 values per thread as before; u_prev and the stage input live in shared
//     memory (the 64 accumulators need the room).
// The order of every rounding is fixed (oracle/rnde_oracle.c rhs_eval_splitk, arith = 2): results are bit-identical to
// the oracle.  Replaces the same reference code as fwd_kernel.cuh (solve(...) at src/models/neural_ode.jl:131-137).
#pragma once
#include "common.cuh"
#include "fwd_kernel.cuh"
#include "fwd4_kernel.cuh"

namespace rnde {

constexpr int V5_G = 4, V5_NP = 16, V5_NT = 256;

struct V5Layout {
    int R, KS1, KS2, NRG, NMG, HS, P1, P2, NG;
    int oW1, oW1t, ob1, oW2, oW2t, ob2, oZ, oH, oPart, oKt, oU, oRed, oCP, oTot, oCtl, oBar, total;
};

__host__ __device__ inline int v5_pitch(int n) {      // multiple of 4 with (pitch mod 32) not in {0, 8, 24}: conflict-free weight reads
    int p = (n + 3) / 4 * 4;
    while ((p & 31) == 0 || (p & 31) == 8 || (p & 31) == 24) p += 4;
    return p;
}

__host__ __device__ inline V5Layout make_v5_layout(int D, int H) {
    V5Layout L;
    L.R = D / 4;
    L.KS1 = (L.R + 7) / 8; L.KS2 = (H + 3) / 4;
    L.NRG = (L.R + 7) / 8; L.NMG = (H + 7) / 8;
    L.HS = (H + V5_G - 1) / V5_G;
    // pitches from the true dimensions: the last tile of a row reads up to 7 words of the following row (outputs discarded)
    L.P1 = v5_pitch(H); L.P2 = v5_pitch(L.R);
    L.NG = (L.R + 3) / 4;
    int o = 0;
    L.oW1 = o; o += 8 * L.KS1 * L.P1 + 8;       // [k = local row, padded to 8*KS1][m = hidden]
    L.oW1t = o; o += L.P1;
    L.ob1 = o; o += L.P1;
    L.oW2 = o; o += 4 * L.KS2 * L.P2 + 8;       // [k = hidden, padded to 4*KS2][m = local row]
    L.oW2t = o; o += L.P2 + 8;                  // rows up to 8*NRG - 1 are read (outputs past R discarded)
    L.ob2 = o; o += L.P2 + 8;
    L.oZ = o; o += 8 * L.KS1 * V5_NP;           // stage input (rows past R stay zero); rows [0, R) are the tape tile
    L.oH = o; o += 4 * L.KS2 * V5_NP;           // hidden activations (rows past H stay zero)
    L.oPart = o; o += V5_G * L.HS * V5_NP;
    L.oKt = o; o += L.R * V5_NP;                // layer-2 output staged for the bulk store to the tape
    L.oU = o; o += L.R * V5_NP;                 // u_prev
    L.oRed = o; o += 3 * 2 * L.NRG * V5_NP;
    L.oCP = o; o += 3 * V5_G * V5_NP;
    L.oTot = o; o += 4;
    L.oCtl = o; o += 32;
    L.oBar = o; o += 8;
    L.total = o;
    return L;
}

__host__ inline bool v5_shape_ok(int D, int H) {
    if (D % 4 != 0 || D < 64) return false;
    const int R = D / 4;
    if ((R + 7) / 8 > 28 || (H + 7) / 8 > 14 || H < 4) return false;      // 7 warps of tiles per layer phase
    return true;
}

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo2(u64 v) { return __uint_as_float((unsigned)(v & 0xffffffffull)); }
__device__ __forceinline__ float hi2(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
// two IEEE round-to-nearest fmas in one instruction (FFMA2); the same roundings as two rn_fmaf
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 shfl2(u64 v, int mask) {
    return pk2(__shfl_xor_sync(0xffffffffu, lo2(v), mask), __shfl_xor_sync(0xffffffffu, hi2(v), mask));
}
// one k-step of the 8x8 outer product: acc[a][j] (rows 2a, 2a+1; column j) += w[rows] * x[j]
__device__ __forceinline__ void tile_step(u64 (&acc)[4][8], const float4 w0, const float4 w1, const float4 x0, const float4 x1) {
    const u64 wv[4] = {pk2(w0.x, w0.y), pk2(w0.z, w0.w), pk2(w1.x, w1.y), pk2(w1.z, w1.w)};
    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[a][j] = fma2(wv[a], pk2(xv[j], xv[j]), acc[a][j]);
}

template <int HC, int DC>
__global__ void __launch_bounds__(V5_NT, 1) fwd4s_kernel(const KParams P) {
    constexpr int G = V5_G, NP = V5_NP, NT = V5_NT;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int q = blockIdx.x / G;
    const int D = (DC > 0) ? DC : P.D, td = P.td;
    const int H = (HC > 0) ? HC : P.H;
    const V5Layout L = make_v5_layout(D, H);
    const int R = L.R, KS1 = L.KS1, KS2 = L.KS2, NRG = L.NRG, NMG = L.NMG, HS = L.HS, P1 = L.P1, P2 = L.P2, NG = L.NG;
    const int r0 = rank * R;
    const int c0 = q * NP;
    const int Nloc = max(0, min(NP, P.B - c0));
    const int HSloc = max(0, min(HS, H - rank * HS));
    float* sW1 = smem + L.oW1; float* sW1t = smem + L.oW1t; float* sb1 = smem + L.ob1;
    float* sW2 = smem + L.oW2; float* sW2t = smem + L.oW2t; float* sb2 = smem + L.ob2;
    float* sZ = smem + L.oZ; float* sH = smem + L.oH; float* sPart = smem + L.oPart; float* sKt = smem + L.oKt; float* sU = smem + L.oU;
    float* sRed = smem + L.oRed; float* sCP = smem + L.oCP; float* sTot = smem + L.oTot;
    Ctl* ctl = reinterpret_cast<Ctl*>(smem + L.oCtl);
    const uint32_t barP = smem_u32(smem + L.oBar), barH = barP + 8;

    const float* gW1 = P.p;
    const float* gb1 = gW1 + (size_t)H * (D + td);
    const float* gW2 = gb1 + H;
    const float* gb2 = gW2 + (size_t)D * (H + td);

    // ---- thread roles (see the header) ------------------------------------------------------------
    const int b0 = lane & 1, cg = (lane >> 1) & 1, b2 = (lane >> 2) & 1, b3 = (lane >> 3) & 1, b4 = (lane >> 4) & 1;
    // layer 2 / state ownership: row group rg (8 rows), k chunk s2 = b0 + 2*b3; the lane ends on rows [row0, row0+4) x columns [col0, col0+4)
    const int rg = 4 * warp + 2 * b4 + b2;
    const int s2 = b0 + 2 * b3;
    const bool own = rg < NRG;
    const int rgc = own ? rg : 0;
    const int row0 = 8 * rgc + 4 * b0, col0 = 8 * cg + 4 * b3;
    const int cvalid = own ? max(0, min(4, R - row0)) : 0;
    // layer 1: hidden group mg (8 hidden units), k chunk s1 = b0 + 2*b3 + 4*b4; the lane ends on hidden unit mrow x columns [8cg, 8cg+8)
    const int mg = 2 * warp + b2;
    const int s1 = b0 + 2 * b3 + 4 * b4;
    const bool actA = mg < NMG;
    const int mgc = actA ? mg : 0;
    const int mrow = 8 * mgc + 4 * b0 + 2 * b3 + b4;

    // ---- stage weights (zero padding: k past R / H, rows past H / R) -----------------------------
    for (int e = tid; e < 8 * KS1 * P1 + 8; e += NT) {
        const int k = e / P1, m = e - k * P1;
        sW1[e] = (k < R && m < H) ? __ldg(gW1 + (size_t)(r0 + k) * H + m) : 0.f;
    }
    for (int e = tid; e < 4 * KS2 * P2 + 8; e += NT) {
        const int k = e / P2, m = e - k * P2;
        sW2[e] = (k < H && m < R) ? __ldg(gW2 + (size_t)D * k + r0 + m) : 0.f;
    }
    for (int m = tid; m < P1; m += NT) {
        sW1t[m] = (td && m < H) ? __ldg(gW1 + (size_t)H * D + m) : 0.f;
        sb1[m] = (m < H) ? __ldg(gb1 + m) : 0.f;
    }
    for (int m = tid; m < P2 + 8; m += NT) {
        sW2t[m] = (td && m < R) ? __ldg(gW2 + (size_t)D * H + r0 + m) : 0.f;
        sb2[m] = (m < R) ? __ldg(gb2 + r0 + m) : 0.f;
    }
    for (int e = tid; e < 8 * KS1 * NP; e += NT) sZ[e] = 0.f;
    for (int e = tid; e < 4 * KS2 * NP; e += NT) sH[e] = 0.f;
    for (int e = tid; e < R * NP; e += NT) {      // u_prev <- x
        const int m = e / NP, n = e - m * NP;
        sU[e] = (n < Nloc) ? __ldg(P.x + (size_t)D * (c0 + n) + r0 + m) : 0.f;
    }
    if (tid == 0) {
        Ctl c;
        c.t = P.t0; c.dt = 0.f; c.dtpropose = 0.f; c.qold = (float)1e-4; c.q11 = 1.f; c.eig_prev = 1.f; c.EEst = 1.f; c.eig = 1.f;
        c.qold_pow = canon_powf((float)1e-4, (float)(2.0 / 25.0)); c.qold_pow_next = c.qold_pow;
        c.dt_init = 0.f; c.dt_last = 0.f;
        c.accept = 0; c.accept_prev = 1; c.done = 0; c.iter = 0; c.nf = 0; c.naccept = 0; c.nreject = 0; c.n_saved = 0;
        c.retcode = RNDE_OK; c.as_count = 0; c.as_stiff = 0;
        if (P.reg_kind != RNDE_REG_NONE) {
            if (blockIdx.x == 0 && P.saveval) P.saveval[0] = saved_value(P.reg_kind, 1.f, 1.f, 0.f);
            c.n_saved = 1;
        }
        *ctl = c;
        mbar_init(barP, 1);
        mbar_init(barH, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float kk[7][16];      // kk[j-1] = k_j on the lane's 4x4 tile (element i*4+j = row row0+i, column col0+j)
#pragma unroll
    for (int a = 0; a < 7; ++a)
#pragma unroll
        for (int e = 0; e < 16; ++e) kk[a][e] = 0.f;
    __syncthreads();
    cluster_sync_all();     // weights staged, mbarriers initialised and visible cluster-wide

    unsigned norm_seq = 0, bar_gen = 0;
    unsigned* xseq_ptr = reinterpret_cast<unsigned*>(P.peers[P.rank]) + P.flag_off + 32;
    const unsigned xseq_base = (P.nranks > 1) ? *xseq_ptr : 0u;
    uint32_t ev_parity = 0;
    int dbg_n = 0;
    auto mark = [&](int id) {
#ifdef RNDE_TIMELINE
        if (P.dbg && blockIdx.x == 0 && tid == 0 && dbg_n < 4000) { P.dbg[dbg_n * 2] = id; P.dbg[dbg_n * 2 + 1] = clock64(); dbg_n++; }
#else
        (void)id; (void)dbg_n;
#endif
    };
    const uint32_t bytesP = (uint32_t)((G - 1) * HSloc * NP * 4);
    const uint32_t bytesH = (uint32_t)((H - HSloc) * NP * 4);

    auto load_tile = [&](const float* base, float (&v)[16]) {      // the lane's 4x4 tile of a [row][16] shared-memory array
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 t4 = (i < cvalid) ? *reinterpret_cast<const float4*>(base + (row0 + i) * NP + col0) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[i * 4] = t4.x; v[i * 4 + 1] = t4.y; v[i * 4 + 2] = t4.z; v[i * 4 + 3] = t4.w;
        }
    };
    auto store_tile = [&](float* base, const float (&v)[16]) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (i < cvalid) *reinterpret_cast<float4*>(base + (row0 + i) * NP + col0) = make_float4(v[i * 4], v[i * 4 + 1], v[i * 4 + 2], v[i * 4 + 3]);
    };

    int pend_rec = -1;      // record whose layer-2 output sits in sKt, not yet handed to the bulk-store engine
    // ---- one field evaluation: out = f(sZ, tstage); the caller has written the stage input to sZ ----
    auto rhs = [&](float (&out)[16], const float tstage, const int rec) {
        mark(0);
        if (tid == 0) { mbar_expect_tx(barP, bytesP); mbar_expect_tx(barH, bytesH); }
        if (rec >= 0 || pend_rec >= 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {     // tape: this evaluation's input tile (sZ) and the previous evaluation's output tile (sKt) as bulk stores
            if (rec >= 0) bulk_store(P.tapeZ + (((size_t)rec * P.Q + q) * D + r0) * NP, sZ, (uint32_t)(R * NP * 4));
            if (pend_rec >= 0) bulk_store(P.tapeK + (((size_t)pend_rec * P.Q + q) * D + r0) * NP, sKt, (uint32_t)(R * NP * 4));
            if (rec >= 0 || pend_rec >= 0) bulk_commit();
        }
        pend_rec = -1;
        mark(1);
        // ---- phase A: layer 1, 8 lanes per 8x8 tile of (hidden, column), k = 8i + s1 ----------------
        if (warp < (NMG + 1) / 2) {
            u64 acc[4][8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[a][j] = 0ull;
            const float* wp = sW1 + s1 * P1 + 8 * mgc;
            const float* xp = sZ + s1 * NP + 8 * cg;
            // (an explicit register double buffer for the next k-step's operands was measured SLOWER -- 1.55 vs 1.37 ms per
            // solve, profiles/r2d_*: the loop then no longer unrolls and ptxas' own hoisting of the unrolled loads does better)
#pragma unroll 5
            for (int i = 0; i < KS1; ++i) {
                const float4 w0 = *reinterpret_cast<const float4*>(wp + i * 8 * P1);
                const float4 w1 = *reinterpret_cast<const float4*>(wp + i * 8 * P1 + 4);
                const float4 x0 = *reinterpret_cast<const float4*>(xp + i * 8 * NP);
                const float4 x1 = *reinterpret_cast<const float4*>(xp + i * 8 * NP + 4);
                tile_step(acc, w0, w1, x0, x1);
            }
            // reduce-scatter over the 8 lanes of the tile: ((c0+c1)+(c2+c3)) + ((c4+c5)+(c6+c7))
            u64 h1[2][8];       // rows 4*b0 + {0,1 | 2,3}
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const u64 send = b0 ? acc[a][j] : acc[a + 2][j];
                    const u64 keep = b0 ? acc[a + 2][j] : acc[a][j];
                    h1[a][j] = add2(keep, shfl2(send, 1));
                }
            u64 h2[8];          // rows 4*b0 + 2*b3 + {0,1}
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const u64 send = b3 ? h1[0][j] : h1[1][j];
                const u64 keep = b3 ? h1[1][j] : h1[0][j];
                h2[j] = add2(keep, shfl2(send, 8));
            }
            float r8[8];        // row 4*b0 + 2*b3 + b4
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float send = b4 ? lo2(h2[j]) : hi2(h2[j]);
                const float keep = b4 ? hi2(h2[j]) : lo2(h2[j]);
                r8[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
            if (actA && mrow < H) {     // this CTA's partial of hidden unit mrow -> the CTA that reduces it
                const int d = mrow / HS, ml = mrow - d * HS;
                float* dst = sPart + (rank * HS + ml) * NP + 8 * cg;
                const float4 v0 = make_float4(r8[0], r8[1], r8[2], r8[3]), v1 = make_float4(r8[4], r8[5], r8[6], r8[7]);
                if (d == rank) { *reinterpret_cast<float4*>(dst) = v0; *reinterpret_cast<float4*>(dst + 4) = v1; }
                else {
                    const uint32_t ra = mapa_u32(smem_u32(dst), d), rb = mapa_u32(barP, d);
                    st_async_f4(ra, v0, rb); st_async_f4(ra + 16, v1, rb);
                }
            }
        }
        mark(4);
        if (tid == 0) bulk_wait_read();      // sZ, sKt and the hidden slice of the previous evaluation may be overwritten from here on
        __syncthreads();
        mark(5);
        mbar_wait(barP, ev_parity);
        mark(6);
        // ---- phase B: fixed-order sum over the 4 CTAs, time column, bias, activation, all-gather ----
        if (tid < HSloc * 4) {
            const int ml = tid >> 2, n4 = (tid & 3) * 4;
            const int m = rank * HS + ml;
            float4 s = *reinterpret_cast<const float4*>(sPart + ml * NP + n4);
#pragma unroll
            for (int c = 1; c < G; ++c) {
                const float4 pc = *reinterpret_cast<const float4*>(sPart + (c * HS + ml) * NP + n4);
                s.x = s.x + pc.x; s.y = s.y + pc.y; s.z = s.z + pc.z; s.w = s.w + pc.w;
            }
            float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float v = sv[j];
                if (td) v = rn_fmaf(sW1t[m], tstage, v);
                v = v + sb1[m];
                sv[j] = act_apply(P.act1, v);
            }
            const float4 h4 = make_float4(sv[0], sv[1], sv[2], sv[3]);
            float* dst = sH + m * NP + n4;
            *reinterpret_cast<float4*>(dst) = h4;
            const uint32_t da = smem_u32(dst);
#pragma unroll
            for (int d = 1; d < G; ++d) {
                const int peer = (rank + d) & (G - 1);
                st_async_f4(mapa_u32(da, peer), h4, mapa_u32(barH, peer));
            }
        }
        if (rec >= 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mark(7);
        __syncthreads();
        if (rec >= 0 && tid == 0 && HSloc > 0) {      // this CTA's hidden slice of the tape
            bulk_store(P.tapeH + (((size_t)rec * P.Q + q) * H + rank * HS) * NP, sH + rank * HS * NP, (uint32_t)(HSloc * NP * 4));
            bulk_commit();
        }
        mark(8);
        mbar_wait(barH, ev_parity);
        mark(9);
        ev_parity ^= 1u;
        // ---- phase C: layer 2, 4 lanes per 8x8 tile of (row, column), k = 4i + s2 -------------------
        if (warp < (NRG + 3) / 4) {
            u64 acc[4][8];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[a][j] = 0ull;
            const float* wp = sW2 + s2 * P2 + 8 * rgc;
            const float* xp = sH + s2 * NP + 8 * cg;
#pragma unroll 5
            for (int i = 0; i < KS2; ++i) {
                const float4 w0 = *reinterpret_cast<const float4*>(wp + i * 4 * P2);
                const float4 w1 = *reinterpret_cast<const float4*>(wp + i * 4 * P2 + 4);
                const float4 x0 = *reinterpret_cast<const float4*>(xp + i * 4 * NP);
                const float4 x1 = *reinterpret_cast<const float4*>(xp + i * 4 * NP + 4);
                tile_step(acc, w0, w1, x0, x1);
            }
            // reduce-scatter over the 4 lanes of the tile: (c0+c1) + (c2+c3); rows by b0, columns by b3
            u64 h1[2][8];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const u64 send = b0 ? acc[a][j] : acc[a + 2][j];
                    const u64 keep = b0 ? acc[a + 2][j] : acc[a][j];
                    h1[a][j] = add2(keep, shfl2(send, 1));
                }
            u64 h2[2][4];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const u64 send = b3 ? h1[a][j] : h1[a][j + 4];
                    const u64 keep = b3 ? h1[a][j + 4] : h1[a][j];
                    h2[a][j] = add2(keep, shfl2(send, 8));
                }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float wt = sW2t[row0 + i], bb = sb2[row0 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = (i & 1) ? hi2(h2[i >> 1][j]) : lo2(h2[i >> 1][j]);
                    if (td) v = rn_fmaf(wt, tstage, v);
                    v = v + bb;
                    out[i * 4 + j] = act_apply(P.act2, v);
                }
            }
            if (rec >= 0) store_tile(sKt, out);
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) out[e] = 0.f;
        }
        pend_rec = rec;      // its output tile is stored to the tape at the next block-wide barrier (next evaluation or flush_tape)
        mark(10);
    };
    auto flush_tape = [&]() {
        if (pend_rec >= 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) { bulk_store(P.tapeK + (((size_t)pend_rec * P.Q + q) * D + r0) * NP, sKt, (uint32_t)(R * NP * 4)); bulk_commit(); }
            pend_rec = -1;
        }
    };

    // ---- canonical norms from register tiles: val(e, out[NV]) for the lane's 16 elements ------------
    // per column: the 4 rows of a lane are one fma chain (row group g = row0/4), the groups of the CTA are added in order,
    // the CTAs in rank order (oracle col_sumsq, arith = 2), the columns by cols_total.
    auto norms = [&](auto val, auto nv_tag, float* result) {
        constexpr int NV = decltype(nv_tag)::value;
        if (own) {
            const int g = row0 >> 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float qv[NV];
#pragma unroll
                for (int v = 0; v < NV; ++v) qv[v] = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (i < cvalid) {
                        float vv[NV];
                        val(i * 4 + j, vv);
#pragma unroll
                        for (int v = 0; v < NV; ++v) qv[v] = rn_fmaf(vv[v], vv[v], qv[v]);
                    }
                }
#pragma unroll
                for (int v = 0; v < NV; ++v) sRed[(v * 2 * NRG + g) * NP + col0 + j] = qv[v];
            }
        }
        __syncthreads();
        mark(11);
        const unsigned slot = norm_seq & 1u;
        float* gcol = P.colsum + (size_t)slot * 3 * P.colsum_stride;
        if (tid < NP * NV) {
            const int n = tid % NP, v = tid / NP;
            const float* rp = sRed + (v * 2 * NRG) * NP + n;
            float s = rp[0];
            for (int g = 1; g < NG; ++g) s = s + rp[g * NP];
            st_cluster_f32(mapa_u32(smem_u32(sCP + (v * G + rank) * NP + n), 0), s);
        }
        cluster_sync_all();
        mark(12);
        if (rank == 0 && tid < NP * NV) {
            const int n = tid % NP, v = tid / NP;
            float tot = sCP[(v * G) * NP + n];
#pragma unroll
            for (int c = 1; c < G; ++c) tot = tot + sCP[(v * G + c) * NP + n];
            if (n < Nloc) publish_colsum(P, (size_t)slot * 3 * P.colsum_stride + (size_t)v * P.colsum_stride + P.col_offset + q * NP + n, tot);
        }
        mark(13);
        if (P.nranks > 1) xrank_arrive_wait(P, rank == 0, xseq_base + norm_seq + 1u, gridDim.x / G);      // all ranks' clusters, one NVLink round
        else grid_barrier(P.bar, gridDim.x, bar_gen);
        mark(14);
        if (warp < NV) {
            const float* g = gcol + (size_t)warp * P.colsum_stride;
            float s = 0.f;
            for (int j0 = lane; j0 < P.Bglobal; j0 += 32 * 8) {      // 8 loads in flight, added in the canonical order
                float v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = (j0 + 32 * u < P.Bglobal) ? __ldcg(g + j0 + 32 * u) : 0.f;
#pragma unroll
                for (int u = 0; u < 8; ++u) if (j0 + 32 * u < P.Bglobal) s = s + v[u];
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) sTot[warp] = rn_sqrtf(rn_divf(s, (float)P.norm_count));
        }
        __syncthreads();
#pragma unroll
        for (int v = 0; v < NV; ++v) result[v] = sTot[v];
        norm_seq += 1;
        __syncthreads();
        mark(15);
    };

    const float dtmax = P.t1 - P.t0;
    const float gamma = (float)(9.0 / 10.0), qmin = (float)(1.0 / 5.0), qmax = 10.f;
    const float beta1 = (float)(7.0 / 50.0), beta2 = (float)(2.0 / 25.0), qoldinit = (float)1e-4;
    const bool limited = P.need_tape || P.reg_kind != RNDE_REG_NONE;
    const bool forced = P.n_forced > 0;      // fixed-work replay (rnde_set_forced_steps): every attempt takes the recorded dt and is accepted

    // loopheader! (thread 0): choose dt for the next attempt or finish.  Returns via ctl.
    auto loopheader = [&]() {
        if (tid == 0) {
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
                if (forced) dt = __ldg(P.forced_dt + min(c.iter, P.n_forced - 1));
                c.iter += 1;
                if (P.alg == RNDE_ALG_AUTO_TSIT5 && !forced) {
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
    };

    // One field evaluation per trip; `stage`: 0 = fsalfirst = f(u0,t0); 1 = f(u0 + dt0*f0) of the initial-dt heuristic; 2..7 = Tsit5 stages
    int stage = 0;
    float t = P.t0, dt = 0.f, a2 = 0.f, dt0 = 0.f, d1_keep = 0.f, d0_keep = 0.f;
    int srec = -1;
    while (true) {
        float tstage;
        int rec = -1;
        {
            float up[16], zc[16];
            load_tile(sU, up);
            if (stage == 0) {
#pragma unroll
                for (int e = 0; e < 16; ++e) zc[e] = up[e];
                tstage = P.t0;
                rec = P.need_tape ? 0 : -1;
            } else if (stage == 1) {
#pragma unroll
                for (int e = 0; e < 16; ++e) zc[e] = rn_fmaf(dt0, kk[0][e], up[e]);
                tstage = P.t0 + dt0;
                rec = (P.need_tape && P.a6) ? P.rec_init : -1;      // Appendix A.6: this evaluation stays on the tape
            } else {
                switch (stage) {
                    case 2: combo_stage<2>(kk, up, dt, a2, zc); break;
                    case 3: combo_stage<3>(kk, up, dt, a2, zc); break;
                    case 4: combo_stage<4>(kk, up, dt, a2, zc); break;
                    case 5: combo_stage<5>(kk, up, dt, a2, zc); break;
                    case 6: combo_stage<6>(kk, up, dt, a2, zc); break;
                    default: combo_stage<7>(kk, up, dt, a2, zc); break;
                }
                tstage = stage_time(t, dt, stage);
                rec = srec >= 0 ? srec + (stage - 2) : -1;
            }
            store_tile(sZ, zc);
        }
        float out[16];
        rhs(out, tstage, rec);
        if (stage == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) kk[0][e] = out[e];
            float up[16];
            load_tile(sU, up);
            float d01[2];
            norms([&](int e, float* o) {
                const float sk = rn_fmaf(fabsf(up[e]), P.reltol, P.abstol);
                o[0] = rn_divf(up[e], sk);
                o[1] = rn_divf(kk[0][e], sk);
            }, std::integral_constant<int, 2>{}, d01);
            const float d0 = d01[0], d1 = d01[1];
            if (d0 < (float)1e-5 || d1 < (float)1e-5) dt0 = (float)1e-6;
            else dt0 = rn_divf(rn_divf(d0, d1), 100.f);
            if (dt0 > dtmax) dt0 = dtmax;
            d1_keep = d1; d0_keep = d0;
            stage = 1;
            continue;
        }
        if (stage == 1) {
            float up[16];
            load_tile(sU, up);
            float d2v[1];
            norms([&](int e, float* o) {
                const float sk = rn_fmaf(fabsf(up[e]), P.reltol, P.abstol);
                o[0] = rn_divf(out[e] - kk[0][e], sk);
            }, std::integral_constant<int, 1>{}, d2v);
            if (tid == 0) {
                const float d1 = d1_keep;
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
                float dti = 100.f * dt0;
                if (dt1 < dti) dti = dt1;
                if (dtmax < dti) dti = dtmax;
                if (dti < P.dtmin) dti = P.dtmin;
                ctl->dt = dti; ctl->dtpropose = dti; ctl->dt_init = dti; ctl->nf = 3;
                if (blockIdx.x == 0 && P.initdt) { P.initdt[0] = d0_keep; P.initdt[1] = d1; P.initdt[2] = d2; P.initdt[3] = dt0; P.initdt[4] = dt1; }
            }
            __syncthreads();
            loopheader();
            if (ctl->done) break;
            t = ctl->t; dt = ctl->dt; a2 = dt * (float)TS_A21;
            srec = P.need_tape ? 1 + 6 * ctl->naccept : -1;
            stage = 2;
            continue;
        }
        switch (stage) {
#define RNDE_KEEP(J) case J: _Pragma("unroll") for (int e = 0; e < 16; ++e) kk[J - 1][e] = out[e]; break;
            RNDE_KEEP(2) RNDE_KEEP(3) RNDE_KEEP(4) RNDE_KEEP(5) RNDE_KEEP(6)
            default: _Pragma("unroll") for (int e = 0; e < 16; ++e) kk[6][e] = out[e]; break;
#undef RNDE_KEEP
        }
        if (stage < 7) { stage += 1; continue; }

        // ---- all 7 stages done: embedded error estimate (+ eigen_est), controller ---------------
        float up[16], un[16];      // u_prev and u_new = the stage-7 input, still in sZ
        load_tile(sU, up);
        load_tile(sZ, un);
        auto atmp_val = [&](int e) -> float {
            float s = ts_bt(1) * kk[0][e];
#pragma unroll
            for (int j = 2; j <= 7; ++j) s = rn_fmaf(ts_bt(j), kk[j - 1][e], s);
            const float ut = dt * s;
            const float a0 = fabsf(up[e]), a1 = fabsf(un[e]);
            const float m = a0 > a1 ? a0 : a1;
            return rn_divf(ut, rn_fmaf(m, P.reltol, P.abstol));
        };
        float EEst, eig = 1.f, en1 = 0.f, en2 = 0.f;
        if (P.alg == RNDE_ALG_AUTO_TSIT5) {
            float o3[3];
            norms([&](int e, float* o) {
                float s = ts_a(6, 1) * kk[0][e];
#pragma unroll
                for (int j = 2; j <= 5; ++j) s = rn_fmaf(ts_a(6, j), kk[j - 1][e], s);
                const float g6 = rn_fmaf(dt, s, up[e]);
                o[0] = kk[6][e] - kk[5][e];
                o[1] = un[e] - g6;
                o[2] = atmp_val(e);
            }, std::integral_constant<int, 3>{}, o3);
            eig = rn_divf(o3[0], o3[1]); en1 = o3[0]; en2 = o3[1];
            EEst = o3[2];
        } else {
            float o1[1];
            norms([&](int e, float* o) { o[0] = atmp_val(e); }, std::integral_constant<int, 1>{}, o1);
            EEst = o1[0];
        }
        if (tid == 32) {
            const float qn = EEst > qoldinit ? EEst : qoldinit;
            ctl->qold_pow_next = canon_powf(qn, beta2);
        }
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
                    qv = rn_divf(c.q11, c.qold_pow);
                    float qq = rn_divf(qv, gamma);
                    const float hi = rn_divf(1.f, qmin), lo = rn_divf(1.f, qmax);
                    qq = hi < qq ? hi : qq;
                    qv = lo > qq ? lo : qq;
                }
                const int accept = forced ? 1 : (EEst <= 1.f);
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
        mark(16);
        const int accepted = ctl->accept, finished = ctl->done;
        if (tid == 0 && accepted) ctl->qold_pow = ctl->qold_pow_next;
        if (!accepted) {      // the retried attempt rewrites the same tape records: the pending tile first, then let every bulk store land
            flush_tape();
            if (tid == 0) bulk_wait_all();
        }
        if (accepted) {   // apply_step!: u <- u_new, fsalfirst <- fsallast
            store_tile(sU, un);
#pragma unroll
            for (int e = 0; e < 16; ++e) kk[0][e] = kk[6][e];
        }
        __syncthreads();
        if (finished) break;
        loopheader();
        if (ctl->done) break;
        t = ctl->t; dt = ctl->dt; a2 = dt * (float)TS_A21;
        srec = P.need_tape ? 1 + 6 * ctl->naccept : -1;
        stage = 2;
    }

    // ---- write back ------------------------------------------------------------------------------
    flush_tape();
    if (tid == 0) bulk_wait_all();
    __syncthreads();
    for (int e = tid; e < R * NP; e += NT) {
        const int n = e / R, m = e - n * R;
        if (n < Nloc) P.u_out[(size_t)D * (c0 + n) + r0 + m] = sU[m * NP + n];
    }
    if (blockIdx.x == 0 && tid == 0) {
        DevStats s;
        s.nf = ctl->nf; s.naccept = ctl->naccept; s.nreject = ctl->nreject; s.n_saved = ctl->n_saved; s.retcode = ctl->retcode;
        s.t_final = ctl->t; s.dt_last = ctl->dt_last; s.dt_init = ctl->dt_init;
        *P.stats = s;
        if (P.nranks > 1) *xseq_ptr = xseq_base + norm_seq;
    }
    cluster_sync_all();
}

}  // namespace rnde
