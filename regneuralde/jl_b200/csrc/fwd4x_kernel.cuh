// fwd4x_kernel.cuh -- the cluster-4 forward stepper with its two layer products on the tensor cores, EXACTLY:
// the FIXED24 arithmetic (oracle/rnde_oracle.c fixed24_*; RNDE_ARITH_FIXED24).
//
// A tensor core does not specify how it rounds a K-long floating-point sum, so a bit-reproducible forward solve cannot
// use floating-point MMAs.  Integer MMAs are exact: every operand of a layer product is scaled to 23-bit fixed point
// (weights per output row and block, inputs per column and block), cut into three signed base-256 digits, and the eight
// digit products with a + c <= 3 run as tcgen05.mma.kind::i8 (s8 x s8 -> s32 in TMEM).  Integer partial sums are
// order-independent, so the K range can be split over several issuing threads and accumulators freely; the eight
// integer tiles are combined as one 64-bit integer T and rounded to Float32 ONCE:  value = float(T) * 2^(e_w + e_x - 36).
// The CPU oracle computes the same integers with plain loops.  Everything else (stage combinations, norms, controller,
// tape, exchange of the hidden layer between the 4 CTAs of a cluster, canonical tanh) is fwd4_kernel.cuh unchanged.
//
// Per field evaluation and CTA (196 rows x 16 columns, hidden 100):
//   column max of z over the CTA's rows -> digits of z -> B1 = [X0 | X1 | X2] (48 x 224 int8, K-major core matrices)
//   GEMM 1: hidden partial = W1[:, rows] z        A1 digit d x B1[0 : 16*(3 - max(d-1,0))...]   3 MMAs per 32-row k-step, 2 issuers
//   TMEM -> T -> Float32 partial -> st.async scatter to the reducer CTA -> ((p0+p1)+p2)+p3, time column, bias, tanh, all-gather
//   column max of h -> digits of h -> B2;  GEMM 2: rows x hidden, 2 M tiles, one issuer each
//   TMEM -> T -> Float32 -> time column, bias, act2 (TMEM lane = state row) -> transposed through shared memory to the 4x4 tiles
#pragma once
#include "common.cuh"
#include "fwd_kernel.cuh"
#include "fwd4_kernel.cuh"
#include "wgrad_tc_kernel.cuh"      // tcgen05 wrappers, umma_desc

namespace rnde {

struct V4XLayout {          // byte offsets
    int R, HS, NGC, K1, K2, sbo1, sbo2, a1g, a2g, slice1, slice2;
    int oA1, oA2, oB1, oB2, oPart, oH, oZb, oEw1, oEw2, oW1t, ob1, oW2t, ob2, oCmax, oRed, oCP, oTot, oCtl, oBar, total;
};

__host__ __device__ inline V4XLayout make_v4x_layout(int D, int H) {
    V4XLayout L;
    const int KB = D / 8;
    L.R = 2 * KB; L.HS = (H + V2_G - 1) / V2_G; L.NGC = (KB + 3) / 4;
    L.K1 = round_up(L.R, 32); L.K2 = round_up(H, 32);
    L.sbo1 = (L.K1 / 16) * 128; L.sbo2 = (L.K2 / 16) * 128;
    L.a1g = (H + 7) / 8; L.a2g = (L.R + 7) / 8;
    L.slice1 = L.a1g * L.sbo1; L.slice2 = L.a2g * L.sbo2;
    const int HP = round_up(H, 4);
    int o = 0;
    L.oA1 = o; o += 3 * L.slice1;
    L.oA2 = o; o += 3 * L.slice2;
    L.oB1 = o; o += 6 * L.sbo1;
    L.oB2 = o; o += 6 * L.sbo2;
    L.oPart = o; o += V2_G * L.HS * V2_NP * 4;
    L.oH = o; o += HP * V2_NP * 4;
    L.oZb = o; o += L.R * V2_NP * 4;
    L.oEw1 = o; o += HP * 4;
    L.oEw2 = o; o += round_up(L.R, 4) * 4;
    L.oW1t = o; o += HP * 4;
    L.ob1 = o; o += HP * 4;
    L.oW2t = o; o += round_up(L.R, 4) * 4;
    L.ob2 = o; o += round_up(L.R, 4) * 4;
    L.oCmax = o; o += 4 * V2_NP * 4;           // column maxima: [z, h] x [2 parities] x 16
    L.oRed = o; o += 3 * 2 * L.NGC * V2_NP * 4;
    L.oCP = o; o += 3 * V2_G * V2_NP * 4;
    L.oTot = o; o += 16;
    L.oCtl = o; o += 160;
    L.oBar = o; o += 64;
    L.total = o;
    return L;
}

// one M = 128 tile of hidden units, at most two of state rows; the padded M tiles' overruns stay inside the allocation
__host__ inline bool v4x_shape_ok(int D, int H) {
    if (!v2_shape_ok(D, H) || H > 128) return false;
    const V4XLayout L = make_v4x_layout(D, H);
    if (L.R > 256 || L.R <= 128) return false;
    if (L.oA1 + 2 * L.slice1 + 16 * L.sbo1 > L.total || L.oA2 + 2 * L.slice2 + 32 * L.sbo2 > L.total) return false;
    return true;
}

constexpr int F24_NONE = -2147483647 - 1;
__device__ __forceinline__ int f24_exponent(const float amax) {      // 2^(e-1) <= amax < 2^e; F24_NONE: the block quantises to 0
    const int eb = (int)((__float_as_uint(amax) >> 23) & 0xFFu);
    return eb < 24 ? F24_NONE : eb - 126;
}
__device__ __forceinline__ float f24_pow2(int e) {      // 2^e, clamped to the normal range (e < -126 gives 0)
    if (e < -126) return 0.f;
    if (e > 127) e = 127;
    return __uint_as_float((uint32_t)(e + 127) << 23);
}
// q = rint(x * scale) as three balanced base-256 digits (branch-free; scale = 0 for a block that quantises to zero)
__device__ __forceinline__ float f24_scale(const int e) { return e == F24_NONE ? 0.f : f24_pow2(22 - e); }
__device__ __forceinline__ void f24_digits_s(const float x, const float scale, int& d0, int& d1, int& d2) {
    int q = __float2int_rn(fminf(fmaxf(x * scale, -8388607.f), 8388607.f));     // the clamp never binds for finite data
    d2 = (q << 24) >> 24; q = (q - d2) >> 8;
    d1 = (q << 24) >> 24;
    d0 = (q - d1) >> 8;
}
__device__ __forceinline__ void f24_digits(const float x, const int e, int& d0, int& d1, int& d2) { f24_digits_s(x, f24_scale(e), d0, d1, d2); }
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16i(uint32_t taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld8i_nowait(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// One issuer's 128 accumulator columns are 8 blocks of 16 columns: I00 I01 I02 | I10 I11 I12 | I20 I21.  The digit
// products of equal weight are added as 32-bit integers (exact: each is below 2^22): A = I00, B = I01 + I10,
// C = I02 + I11 + I20, E = I12 + I21, accumulated over issuers; T = A*2^24 + B*2^16 + C*2^8 + E is then formed exactly in
// Float64 (|T| < 2^45) and rounded to Float32 once.  Eight columns (half = 0 / 1) at a time keep the register footprint small.
__device__ __forceinline__ void f24_gather8(const uint32_t taddr, const int half, int (&A)[8], int (&B)[8], int (&Cc)[8], int (&E)[8]) {
    const uint32_t t = taddr + 8 * half;
    int v0[8], v1[8];
    tmem_ld8i_nowait(t, v0); tmem_ld8i_nowait(t + 16, v1); tmem_wait_ld();            // I00, I01
#pragma unroll
    for (int j = 0; j < 8; ++j) { A[j] += v0[j]; B[j] += v1[j]; }
    tmem_ld8i_nowait(t + 32, v0); tmem_ld8i_nowait(t + 48, v1); tmem_wait_ld();       // I02, I10
#pragma unroll
    for (int j = 0; j < 8; ++j) { Cc[j] += v0[j]; B[j] += v1[j]; }
    tmem_ld8i_nowait(t + 64, v0); tmem_ld8i_nowait(t + 80, v1); tmem_wait_ld();       // I11, I12
#pragma unroll
    for (int j = 0; j < 8; ++j) { Cc[j] += v0[j]; E[j] += v1[j]; }
    tmem_ld8i_nowait(t + 96, v0); tmem_ld8i_nowait(t + 112, v1); tmem_wait_ld();      // I20, I21
#pragma unroll
    for (int j = 0; j < 8; ++j) { Cc[j] += v0[j]; E[j] += v1[j]; }
}
__device__ __forceinline__ float f24_value(const int a, const int b, const int c, const int e, const float scale) {
    const double t = ((double)a * 16777216.0 + (double)b * 65536.0) + ((double)c * 256.0 + (double)e);      // exact
    return __double2float_rn(t) * scale;
}

__global__ void __launch_bounds__(V2_NT, 1) fwd4x_kernel(const KParams P) {
    constexpr int G = V2_G, NP = V2_NP, NT = V2_NT;
    extern __shared__ __align__(16) float smem[];
    unsigned char* sb = reinterpret_cast<unsigned char*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int q = blockIdx.x / G;
    const int D = P.D, td = P.td, H = P.H;
    const V4XLayout L = make_v4x_layout(D, H);
    const int KB = D / 8, R = L.R, HS = L.HS, NGC = L.NGC;
    const int r0 = rank * R;
    const int c0 = q * NP;
    const int Nloc = max(0, min(NP, P.B - c0));
    const int HSloc = max(0, min(HS, H - rank * HS));
    float* sPart = reinterpret_cast<float*>(sb + L.oPart); float* sH = reinterpret_cast<float*>(sb + L.oH);
    float* sZb = reinterpret_cast<float*>(sb + L.oZb);
    int* sEw1 = reinterpret_cast<int*>(sb + L.oEw1); int* sEw2 = reinterpret_cast<int*>(sb + L.oEw2);
    float* sW1t = reinterpret_cast<float*>(sb + L.oW1t); float* sb1 = reinterpret_cast<float*>(sb + L.ob1);
    float* sW2t = reinterpret_cast<float*>(sb + L.oW2t); float* sb2 = reinterpret_cast<float*>(sb + L.ob2);
    unsigned* sCmax = reinterpret_cast<unsigned*>(sb + L.oCmax);
    float* sRed = reinterpret_cast<float*>(sb + L.oRed); float* sCP = reinterpret_cast<float*>(sb + L.oCP); float* sTot = reinterpret_cast<float*>(sb + L.oTot);
    Ctl* ctl = reinterpret_cast<Ctl*>(sb + L.oCtl);
    const uint32_t sbase = smem_u32(sb);
    const uint32_t barP = sbase + L.oBar, barH = barP + 8, barM1 = barP + 16, barM2 = barP + 24;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sb + L.oBar + 32);

    const float* gW1 = P.p;
    const float* gb1 = gW1 + (size_t)H * (D + td);
    const float* gW2 = gb1 + H;
    const float* gb2 = gW2 + (size_t)D * (H + td);

    // state ownership: the 4x4 tiles of fwd4_kernel (two K-blocks of KB rows per CTA)
    const bool own = tid < 2 * NGC * 4;
    const int cblk = tid / (NGC * 4), ctile = tid % (NGC * 4);
    const int cmt = ctile >> 2, cn0 = (ctile & 3) * 4;
    const int crow0 = cblk * KB + cmt * 4;
    const int cvalid = own ? min(4, KB - cmt * 4) : 0;

    // ---- quantise this CTA's weight slices into the A operands (once per launch) -------------------------------
    for (int e = tid; e < L.total / 4; e += NT) smem[e] = 0.f;
    __syncthreads();
    unsigned* sMax1 = reinterpret_cast<unsigned*>(sZb);            // scratch: row maxima (H + R words <= R*16)
    unsigned* sMax2 = sMax1 + round_up(H, 4);
    for (int e = tid; e < R * H; e += NT) {          // W1[m, r0 + k]: m fastest (coalesced)
        const int k = e / H, m = e - k * H;
        atomicMax(sMax1 + m, __float_as_uint(fabsf(__ldg(gW1 + (size_t)(r0 + k) * H + m))));
    }
    for (int e = tid; e < H * R; e += NT) {          // W2[r0 + i, k]: i fastest (coalesced)
        const int k = e / R, i = e - k * R;
        atomicMax(sMax2 + i, __float_as_uint(fabsf(__ldg(gW2 + (size_t)D * k + r0 + i))));
    }
    __syncthreads();
    for (int m = tid; m < H; m += NT) sEw1[m] = f24_exponent(__uint_as_float(sMax1[m]));
    for (int i = tid; i < R; i += NT) sEw2[i] = f24_exponent(__uint_as_float(sMax2[i]));
    __syncthreads();
    for (int e = tid; e < R * H; e += NT) {
        const int k = e / H, m = e - k * H;
        int d0, d1, d2; f24_digits(__ldg(gW1 + (size_t)(r0 + k) * H + m), sEw1[m], d0, d1, d2);
        const int off = (m >> 3) * L.sbo1 + (k >> 4) * 128 + (m & 7) * 16 + (k & 15);
        sb[L.oA1 + off] = (unsigned char)d0; sb[L.oA1 + L.slice1 + off] = (unsigned char)d1; sb[L.oA1 + 2 * L.slice1 + off] = (unsigned char)d2;
    }
    for (int e = tid; e < H * R; e += NT) {
        const int k = e / R, i = e - k * R;
        int d0, d1, d2; f24_digits(__ldg(gW2 + (size_t)D * k + r0 + i), sEw2[i], d0, d1, d2);
        const int off = (i >> 3) * L.sbo2 + (k >> 4) * 128 + (i & 7) * 16 + (k & 15);
        sb[L.oA2 + off] = (unsigned char)d0; sb[L.oA2 + L.slice2 + off] = (unsigned char)d1; sb[L.oA2 + 2 * L.slice2 + off] = (unsigned char)d2;
    }
    __syncthreads();
    for (int e = tid; e < R * NP; e += NT) sZb[e] = 0.f;          // the scratch held the row maxima
    for (int m = tid; m < H; m += NT) {
        sW1t[m] = td ? __ldg(gW1 + (size_t)H * D + m) : 0.f;
        sb1[m] = __ldg(gb1 + m);
    }
    for (int i = tid; i < R; i += NT) {
        sW2t[i] = td ? __ldg(gW2 + (size_t)D * H + r0 + i) : 0.f;
        sb2[i] = __ldg(gb2 + r0 + i);
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
        mbar_init(barP, 1); mbar_init(barH, 1); mbar_init(barM1, 2); mbar_init(barM2, 2);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {        // TMEM: 512 columns = 2 GEMMs x 2 issuers x 128 int32 accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // ---- state in registers -------------------------------------------------------------------
    float uprev[16], zc[16], kk[7][16];   // kk[j-1] = k_j
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = cn0 + j;
            uprev[i * 4 + j] = (i < cvalid && n < Nloc) ? __ldg(P.x + (size_t)D * (c0 + n) + r0 + crow0 + i) : 0.f;
        }
#pragma unroll
    for (int a = 0; a < 7; ++a)
#pragma unroll
        for (int e = 0; e < 16; ++e) kk[a][e] = 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    cluster_sync_all();     // operands staged, mbarriers initialised and visible cluster-wide

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
    // instruction descriptors: D = S32, A = B = signed 8-bit, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    auto idesc_n = [](int n) { return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24); };
    const uint32_t idesc48 = idesc_n(48), idesc32 = idesc_n(32);
    const int issuer = (lane == 0 && warp < 2) ? warp : -1;
    const uint64_t dA1 = umma_desc(sbase + L.oA1, 128, L.sbo1), dB1 = umma_desc(sbase + L.oB1, 128, L.sbo1);
    const uint64_t dA2 = umma_desc(sbase + L.oA2, 128, L.sbo2), dB2 = umma_desc(sbase + L.oB2, 128, L.sbo2);
    const uint64_t sl1 = (uint64_t)(L.slice1 >> 4), sl2 = (uint64_t)(L.slice2 >> 4);
    const int nk1 = L.K1 / 32, nk2 = L.K2 / 32;
    const int quad = warp & 3;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;

    // ---- one field evaluation: out = f(zin, tstage) in the FIXED24 arithmetic --------------------------------------
    auto rhs = [&](const float (&zin)[16], float (&out)[16], const float tstage, const int rec) {
        mark(0);
        const uint32_t par = ev_parity;
        unsigned* cmz = sCmax + par * NP;               // column maxima of z (this evaluation)
        unsigned* cmh = sCmax + 2 * NP + par * NP;      // column maxima of h
        if (tid == 0) { mbar_expect_tx(barP, bytesP); mbar_expect_tx(barH, bytesH); }
        if (tid < NP) { sCmax[(par ^ 1u) * NP + tid] = 0u; sCmax[2 * NP + (par ^ 1u) * NP + tid] = 0u; }    // next evaluation's slots
        if (own) {      // the stage input goes to shared memory once (FP32, [row][column]); its column maxima by shared atomics
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float m = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) if (i < cvalid) m = fmaxf(m, fabsf(zin[i * 4 + j]));
                atomicMax(cmz + cn0 + j, __float_as_uint(m));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < cvalid) *reinterpret_cast<float4*>(sZb + (crow0 + i) * NP + cn0) = make_float4(zin[i * 4], zin[i * 4 + 1], zin[i * 4 + 2], zin[i * 4 + 3]);
        }
        __syncthreads();
        mark(1);
        // digits of z -> B1: one item = 16 consecutive rows of one column = one 16-byte chunk per digit row (few live registers,
        // wide stores; the per-thread 4x4 tiles would need 24 sub-word stores each)
        for (int ch = tid; ch < NP * (L.K1 / 16); ch += NT) {
            const int n = ch & (NP - 1), k0 = (ch >> 4) * 16;
            const float sc = f24_scale(f24_exponent(__uint_as_float(cmz[n])));
            uint32_t w[3][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                int d0 = 0, d1 = 0, d2 = 0;
                if (k0 + i < R) f24_digits_s(sZb[(k0 + i) * NP + n], sc, d0, d1, d2);
                w[0][i >> 2] |= (uint32_t)(d0 & 255) << (8 * (i & 3));
                w[1][i >> 2] |= (uint32_t)(d1 & 255) << (8 * (i & 3));
                w[2][i >> 2] |= (uint32_t)(d2 & 255) << (8 * (i & 3));
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int row = c * NP + n;
                *reinterpret_cast<uint4*>(sb + L.oB1 + (row >> 3) * L.sbo1 + (k0 >> 4) * 128 + (row & 7) * 16) = make_uint4(w[c][0], w[c][1], w[c][2], w[c][3]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        mark(2);
        if (issuer >= 0) {      // GEMM 1: k-steps split over the two issuers, three digit rows of A per k-step
            tc_fence_after();
            const uint32_t dcol = tmem_base + 128 * issuer;
            const int k0 = issuer ? (nk1 + 1) / 2 : 0, k1 = issuer ? nk1 : (nk1 + 1) / 2;
            for (int ks = k0; ks < k1; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 16);
                const uint32_t accf = ks == k0 ? 0u : 1u;
                tc_mma_i8(dcol, dA1 + adv, dB1 + adv, idesc48, accf);                  // d0 x [X0 X1 X2]
                tc_mma_i8(dcol + 48, dA1 + sl1 + adv, dB1 + adv, idesc48, accf);       // d1 x [X0 X1 X2]
                tc_mma_i8(dcol + 96, dA1 + 2 * sl1 + adv, dB1 + adv, idesc32, accf);   // d2 x [X0 X1]
            }
            tc_commit(barM1);
        }
        mbar_wait(barM1, par);
        tc_fence_after();
        mark(3);
        {   // TMEM lane = hidden unit (warps w and w+4 read the same lanes: they take 8 columns each): exact integer T, one rounding,
            // then to the reducer CTA of that hidden slice
            const int m = quad * 32 + lane;
            const int half = warp >> 2;
            const int ew = m < H ? sEw1[m] : F24_NONE;
            const int d = m / HS, ml = m - d * HS;
            float* dst = sPart + (rank * HS + ml) * NP;
            int A[8], B[8], Cc[8], E[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { A[j] = 0; B[j] = 0; Cc[j] = 0; E[j] = 0; }
            f24_gather8(tmem_base + tlane, half, A, B, Cc, E);
            if (nk1 > (nk1 + 1) / 2) f24_gather8(tmem_base + tlane + 128, half, A, B, Cc, E);
            if (m < H) {
#pragma unroll
                for (int c4 = 0; c4 < 2; ++c4) {
                    float pv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int ex = f24_exponent(__uint_as_float(cmz[half * 8 + c4 * 4 + j]));
                        const float sc = (ew == F24_NONE || ex == F24_NONE) ? 0.f : f24_pow2(ew + ex - 36);
                        pv[j] = f24_value(A[c4 * 4 + j], B[c4 * 4 + j], Cc[c4 * 4 + j], E[c4 * 4 + j], sc);
                    }
                    const float4 o = make_float4(pv[0], pv[1], pv[2], pv[3]);
                    float* dp = dst + half * 8 + c4 * 4;
                    if (d == rank) *reinterpret_cast<float4*>(dp) = o;
                    else st_async_f4(mapa_u32(smem_u32(dp), d), o, mapa_u32(barP, d));
                }
            }
        }
        tc_fence_before();
        mark(4);
        __syncthreads();
        mark(5);
        mbar_wait(barP, par);
        mark(6);
        // phase B: fixed-order reduction over the 4 CTAs, time column, bias, activation, all-gather
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
            if (rec >= 0) *reinterpret_cast<float4*>(P.tapeH + ((size_t)rec * P.Q + q) * H * NP + (size_t)m * NP + n4) = h4;
        }
        mark(7);
        __syncthreads();
        mark(8);
        mbar_wait(barH, par);
        mark(9);
        // column maxima of h over all hidden units, then its digits -> B2 (one 16-byte chunk = 16 hidden units of a column)
        {
            const int n = tid & (NP - 1), sl = tid >> 4;
            float m = 0.f;
            for (int r = sl; r < H; r += NT / NP) m = fmaxf(m, fabsf(sH[r * NP + n]));
            atomicMax(cmh + n, __float_as_uint(m));
        }
        __syncthreads();
        for (int ch = tid; ch < NP * (L.K2 / 8); ch += NT) {      // one item = 8 consecutive hidden units of one column
            const int n = ch & (NP - 1), k0 = (ch >> 4) * 8;
            const float sc = f24_scale(f24_exponent(__uint_as_float(cmh[n])));
            uint32_t w[3][2] = {{0, 0}, {0, 0}, {0, 0}};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int d0 = 0, d1 = 0, d2 = 0;
                if (k0 + i < H) f24_digits_s(sH[(k0 + i) * NP + n], sc, d0, d1, d2);
                w[0][i >> 2] |= (uint32_t)(d0 & 255) << (8 * (i & 3));
                w[1][i >> 2] |= (uint32_t)(d1 & 255) << (8 * (i & 3));
                w[2][i >> 2] |= (uint32_t)(d2 & 255) << (8 * (i & 3));
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const int row = c * NP + n;
                *reinterpret_cast<uint2*>(sb + L.oB2 + (row >> 3) * L.sbo2 + (k0 >> 4) * 128 + (row & 7) * 16 + (k0 & 15)) = make_uint2(w[c][0], w[c][1]);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        mark(20);
        if (issuer >= 0) {      // GEMM 2: issuer = M tile of this CTA's rows
            tc_fence_after();
            const uint32_t dcol = tmem_base + 256 + 128 * issuer;
            const uint64_t aoff = (uint64_t)((issuer * 16 * L.sbo2) >> 4);
            for (int ks = 0; ks < nk2; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 16);
                const uint32_t accf = ks == 0 ? 0u : 1u;
                tc_mma_i8(dcol, dA2 + aoff + adv, dB2 + adv, idesc48, accf);
                tc_mma_i8(dcol + 48, dA2 + sl2 + aoff + adv, dB2 + adv, idesc48, accf);
                tc_mma_i8(dcol + 96, dA2 + 2 * sl2 + aoff + adv, dB2 + adv, idesc32, accf);
            }
            tc_commit(barM2);
        }
        mbar_wait(barM2, par);
        tc_fence_after();
        mark(21);
        {   // TMEM lane = local row: T -> Float32 -> time column, bias, activation; transposed through shared memory
            const int row = (warp < 4 ? 0 : 128) + quad * 32 + lane;
            const int ew = row < R ? sEw2[row] : F24_NONE;
            const float wt = row < R ? sW2t[row] : 0.f, bb = row < R ? sb2[row] : 0.f;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                int A[8], B[8], Cc[8], E[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) { A[j] = 0; B[j] = 0; Cc[j] = 0; E[j] = 0; }
                f24_gather8(tmem_base + tlane + 256 + (warp < 4 ? 0u : 128u), half, A, B, Cc, E);
                if (row < R) {
#pragma unroll
                    for (int c4 = 0; c4 < 2; ++c4) {
                        float pv[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int ex = f24_exponent(__uint_as_float(cmh[half * 8 + c4 * 4 + j]));
                            const float sc = (ew == F24_NONE || ex == F24_NONE) ? 0.f : f24_pow2(ew + ex - 36);
                            float v = f24_value(A[c4 * 4 + j], B[c4 * 4 + j], Cc[c4 * 4 + j], E[c4 * 4 + j], sc);
                            if (td) v = rn_fmaf(wt, tstage, v);
                            v = v + bb;
                            pv[j] = act_apply(P.act2, v);
                        }
                        *reinterpret_cast<float4*>(sZb + row * NP + half * 8 + c4 * 4) = make_float4(pv[0], pv[1], pv[2], pv[3]);
                    }
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        mark(22);
        ev_parity ^= 1u;
        if (own) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < cvalid) o4 = *reinterpret_cast<const float4*>(sZb + (crow0 + i) * NP + cn0);
                out[i * 4] = o4.x; out[i * 4 + 1] = o4.y; out[i * 4 + 2] = o4.z; out[i * 4 + 3] = o4.w;
                if (rec >= 0 && i < cvalid) {
                    const size_t off = ((size_t)rec * P.Q + q) * D * NP + (size_t)(r0 + crow0 + i) * NP + cn0;
                    *reinterpret_cast<float4*>(P.tapeK + off) = o4;
                    *reinterpret_cast<float4*>(P.tapeZ + off) = make_float4(zin[i * 4], zin[i * 4 + 1], zin[i * 4 + 2], zin[i * 4 + 3]);
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) out[e] = 0.f;
        }
        mark(10);
    };

    // ---- canonical norms from register tiles: val(e, out[NV]) for the thread's 16 elements ---------
    auto norms = [&](auto val, auto nv_tag, float* result) {
        constexpr int NV = decltype(nv_tag)::value;
        if (own) {
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
                for (int v = 0; v < NV; ++v) sRed[((v * 2 + cblk) * NGC + cmt) * NP + cn0 + j] = qv[v];
            }
        }
        __syncthreads();
        mark(11);
        const unsigned slot = norm_seq & 1u;
        float* gcol = P.colsum + (size_t)slot * 3 * P.colsum_stride;
        if (tid < NP * NV) {
            const int n = tid % NP, v = tid / NP;
            float bs[2];
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const float* rp = sRed + ((v * 2 + b) * NGC) * NP + n;
                float s = rp[0];
                for (int g = 1; g < NGC; ++g) s = s + rp[g * NP];
                bs[b] = s;
            }
            st_cluster_f32(mapa_u32(smem_u32(sCP + (v * G + rank) * NP + n), 0), bs[0] + bs[1]);
        }
        cluster_sync_all();
        mark(12);
        if (rank == 0 && tid < NP * NV) {
            const int n = tid % NP, v = tid / NP;
            float tot = sCP[(v * G) * NP + n];
#pragma unroll
            for (int c = 1; c < G; ++c) tot = tot + sCP[(v * G + c) * NP + n];
            if (n < Nloc) publish_colsum(P, (size_t)slot * 3 * P.colsum_stride + (size_t)v * P.colsum_stride + P.col_offset + q * NP + n, tot);
            if (P.nranks > 1) __threadfence_system();
        }
        mark(13);
        grid_barrier(P.bar, gridDim.x, bar_gen);
        mark(14);
        xrank_barrier(P, xseq_base + norm_seq + 1u);
        const int warp = tid >> 5, lane = tid & 31;
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
                c.iter += 1;
                if (P.alg == RNDE_ALG_AUTO_TSIT5) {
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

    // One field evaluation per trip; `stage` says what it is for:
    //   0: fsalfirst = f(u0,t0) (initialize!)   1: f(u0 + dt0*f0) of the initial-dt heuristic
    //   2..7: Tsit5 stages of the current attempt.
    int stage = 0;
    float t = P.t0, dt = 0.f, a2 = 0.f, dt0 = 0.f, d1_keep = 0.f, d0_keep = 0.f;
    int srec = -1;
    while (true) {
        float tstage;
        int rec = -1;
        if (stage == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) zc[e] = uprev[e];
            tstage = P.t0;
            rec = P.need_tape ? 0 : -1;
        } else if (stage == 1) {
#pragma unroll
            for (int e = 0; e < 16; ++e) zc[e] = rn_fmaf(dt0, kk[0][e], uprev[e]);
            tstage = P.t0 + dt0;
            rec = (P.need_tape && P.a6) ? P.rec_init : -1;      // Appendix A.6: this evaluation stays on the tape
        } else {
            const int i = stage;
            switch (i) {
                case 2: combo_stage<2>(kk, uprev, dt, a2, zc); break;
                case 3: combo_stage<3>(kk, uprev, dt, a2, zc); break;
                case 4: combo_stage<4>(kk, uprev, dt, a2, zc); break;
                case 5: combo_stage<5>(kk, uprev, dt, a2, zc); break;
                case 6: combo_stage<6>(kk, uprev, dt, a2, zc); break;
                default: combo_stage<7>(kk, uprev, dt, a2, zc); break;
            }
            tstage = stage_time(t, dt, i);
            rec = srec >= 0 ? srec + (i - 2) : -1;
        }
        float out[16];
        rhs(zc, out, tstage, rec);
        if (stage == 0) {
#pragma unroll
            for (int e = 0; e < 16; ++e) kk[0][e] = out[e];
            // initial dt (Hairer-Wanner, Appendix A.5), first half
            float d01[2];
            norms([&](int e, float* o) {
                const float sk = rn_fmaf(fabsf(uprev[e]), P.reltol, P.abstol);
                o[0] = rn_divf(uprev[e], sk);
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
            float d2v[1];
            norms([&](int e, float* o) {
                const float sk = rn_fmaf(fabsf(uprev[e]), P.reltol, P.abstol);
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
        // Tsit5 stage: keep k_stage
        switch (stage) {
#define RNDE_KEEP(J) case J: _Pragma("unroll") for (int e = 0; e < 16; ++e) kk[J - 1][e] = out[e]; break;
            RNDE_KEEP(2) RNDE_KEEP(3) RNDE_KEEP(4) RNDE_KEEP(5) RNDE_KEEP(6)
            default: _Pragma("unroll") for (int e = 0; e < 16; ++e) kk[6][e] = out[e]; break;
#undef RNDE_KEEP
        }
        if (stage < 7) { stage += 1; continue; }

        // ---- all 7 stages done: embedded error estimate (+ eigen_est), controller ---------------
        auto atmp_val = [&](int e) -> float {
            float s = ts_bt(1) * kk[0][e];
#pragma unroll
            for (int j = 2; j <= 7; ++j) s = rn_fmaf(ts_bt(j), kk[j - 1][e], s);
            const float ut = dt * s;
            const float a0 = fabsf(uprev[e]), a1 = fabsf(zc[e]);
            const float m = a0 > a1 ? a0 : a1;
            return rn_divf(ut, rn_fmaf(m, P.reltol, P.abstol));
        };
        float EEst, eig = 1.f, en1 = 0.f, en2 = 0.f;
        if (P.alg == RNDE_ALG_AUTO_TSIT5) {
            float o3[3];
            norms([&](int e, float* o) {
                // g6 (stage-6 input) is recomputed with the stage-6 combination, bit-identically
                float s = ts_a(6, 1) * kk[0][e];
#pragma unroll
                for (int j = 2; j <= 5; ++j) s = rn_fmaf(ts_a(6, j), kk[j - 1][e], s);
                const float g6 = rn_fmaf(dt, s, uprev[e]);
                o[0] = kk[6][e] - kk[5][e];
                o[1] = zc[e] - g6;
                o[2] = atmp_val(e);
            }, std::integral_constant<int, 3>{}, o3);
            eig = rn_divf(o3[0], o3[1]); en1 = o3[0]; en2 = o3[1];
            EEst = o3[2];
        } else {
            float o1[1];
            norms([&](int e, float* o) { o[0] = atmp_val(e); }, std::integral_constant<int, 1>{}, o1);
            EEst = o1[0];
        }
        if (tid == 32) {  // next step's qold^beta2, side by side with thread 0's EEst^beta1 (both are ~1 k-cycle double-precision pows)
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
        mark(16);
        const int accepted = ctl->accept, finished = ctl->done;
        if (tid == 0 && accepted) ctl->qold_pow = ctl->qold_pow_next;     // qold was updated: its power follows
        __syncthreads();
        if (accepted) {   // apply_step!: u <- u_new, fsalfirst <- fsallast
#pragma unroll
            for (int e = 0; e < 16; ++e) { uprev[e] = zc[e]; kk[0][e] = kk[6][e]; }
        }
        if (finished) break;
        loopheader();
        if (ctl->done) break;
        t = ctl->t; dt = ctl->dt; a2 = dt * (float)TS_A21;
        srec = P.need_tape ? 1 + 6 * ctl->naccept : -1;
        stage = 2;
    }

    // ---- write back ------------------------------------------------------------------------------
    if (own) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = cn0 + j;
                if (i < cvalid && n < Nloc) P.u_out[(size_t)D * (c0 + n) + r0 + crow0 + i] = uprev[i * 4 + j];
            }
    }
    if (blockIdx.x == 0 && tid == 0) {
        DevStats s;
        s.nf = ctl->nf; s.naccept = ctl->naccept; s.nreject = ctl->nreject; s.n_saved = ctl->n_saved; s.retcode = ctl->retcode;
        s.t_final = ctl->t; s.dt_last = ctl->dt_last; s.dt_init = ctl->dt_init;
        *P.stats = s;
        if (P.nranks > 1) *xseq_ptr = xseq_base + norm_seq;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
    cluster_sync_all();
}

}  // namespace rnde
