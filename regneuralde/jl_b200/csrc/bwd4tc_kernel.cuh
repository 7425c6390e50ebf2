// bwd4tc_kernel.cuh -- the reverse sweep of the cluster-4 decomposition with its two products per field evaluation
// on the 5th-generation tensor cores (tcgen05, accumulators in TMEM).
//
// The backward pass is held to a tolerance (gradient <= 1e-4 relative), not to bit-identity like the forward solve,
// so its contractions may leave the canonical FFMA order.  Per evaluation on the tape, per CTA (196 state rows of 16
// columns, hidden dimension 100):
//   GEMM 1  hbar_partial[100 x 16] = W2[rows, :]^T delta2[rows x 16]      M=128 N=16 K=196(208)   13 k-steps
//   GEMM 2  zbar[196 x 16]         = W1[:, rows]^T delta1[100 x 16]       2 x (M=128 N=16) K=100(112)  7 k-steps
// as tcgen05.mma.kind::f16 on BF16 operands with FP32 accumulation.  The two operands are split differently, because
// their truncation errors act differently on the regulariser gradient (the error-estimate
// cotangents are O(10) on the seven stages of a step and cancel to O(1e-2); measured with oracle/ emulation):
//   * the COTANGENT (B operand) is split into three BF16 terms hi + mid + lo = the Float32 value exactly: its truncation
//     error differs from stage to stage, so any of it survives the cancellation (two terms = 16 bits made the sweep 5x
//     noisier than a CPU Float32 adjoint);
//   * the WEIGHTS (A operand) keep two terms (16 bits): the same perturbed weights act on every stage, which perturbs the
//     small result relatively, not the large terms absolutely (truncating weights to 12 bits changes nothing measurable).
// Products per k-step: A_hi x [B_hi | B_mid | B_lo] (N = 48) and A_lo x [B_hi | B_mid] (N = 32); the dropped A_lo x B_lo is
// 2^-26 relative.  hi/mid/lo products accumulate in separate TMEM columns and are added small-to-large in the epilogue.
// (A third weight term or the 3xTF32 split of the weight-gradient kernel would not fit: 39 200 weights per CTA at 4 bytes.)  Weights are split once per launch into the UMMA K-major core-matrix layout (A operands); the
// cotangents are split and written as B operands by the threads that own them; one elected thread issues the MMAs;
// results come back with tcgen05.ld (TMEM lane = output row).  The exchange of the hidden cotangent between the four
// CTAs of the cluster (st.async + mbarrier), the tape traffic and all per-step cotangent arithmetic are those of
// bwd4_kernel.cuh.  Replaces the same reference code: the Tracker tape through solve(...), experiments/mnist_node.jl:229-232.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"
#include "a6.cuh"
#include "fwd4_kernel.cuh"
#include "bwd4_kernel.cuh"
#include "wgrad_tc_kernel.cuh"      // tcgen05 wrappers, umma_desc

namespace rnde {

constexpr int B4T_LBO_B = 144;

struct B4TLayout {          // byte offsets
    int K1, K2, R, H, HS;   // K1 = padded rows (GEMM-1 contraction), K2 = padded hidden (GEMM-2 contraction)
    int sbo1, sbo2;         // 8-row group strides of K1-long and K2-long A operands (bytes)
    int sbb1, sbb2;         // the same for the B operands, whose K-adjacent core matrices are B4T_LBO_B = 144 bytes apart
                            // (16 bytes of padding: the owners' 4-byte stores would otherwise be 8-way bank conflicted)
    int a2_groups, a1_groups;
    int oA2hi, oA2lo, oA1hi, oA1lo, oB1, oB2, oPart, oD1, oZb, oBar, oA6, total;   // oB*: [hi | mid | lo], 2 * sbb* bytes each
};

__host__ __device__ inline B4TLayout make_b4t_layout(int D, int H) {
    B4TLayout L;
    L.R = D / 4; L.H = H; L.HS = (H + V2_G - 1) / V2_G;
    L.K1 = round_up(L.R, 16); L.K2 = round_up(H, 16);
    L.sbo1 = L.K1 * 16; L.sbo2 = L.K2 * 16;
    L.sbb1 = (L.K1 / 8) * B4T_LBO_B; L.sbb2 = (L.K2 / 8) * B4T_LBO_B;
    L.a2_groups = (L.R + 8) / 8; L.a1_groups = (H + 8) / 8;      // + one row: the time column of W1 / W2 (a6.cuh)
    int o = 0;
    L.oA2hi = o; o += L.a2_groups * L.sbo2;
    L.oA2lo = o; o += L.a2_groups * L.sbo2;
    L.oA1hi = o; o += L.a1_groups * L.sbo1;
    L.oA1lo = o; o += L.a1_groups * L.sbo1;
    L.oB1 = o; o += 6 * L.sbb1;
    L.oPart = o; o += V2_G * L.HS * V2_NP * 4;
    L.oD1 = o; o += round_up(H, 4) * V2_NP * 4;
    // the B operand of GEMM 2 lives in the delta2 / transposition tile: written after the tile's bulk store has been read,
    // consumed by GEMM 2 before the transposition overwrites it
    L.oZb = o; L.oB2 = o; o += (L.R * V2_NP * 4 > 6 * L.sbb2) ? L.R * V2_NP * 4 : 6 * L.sbb2;
    L.oBar = o; o += 64;
    L.oA6 = o; o += 64;      // scalars of the initial-dt adjoint between its two halves (a6.cuh)
    L.total = o;
    return L;
}

// rows <= 256 (two M=128 tiles), hidden <= 128, and the operand overruns of the padded M tiles stay inside the allocation
__host__ inline bool b4t_shape_ok(int D, int H) {
    if (!v2_shape_ok(D, H) || D % 4 != 0) return false;
    const B4TLayout L = make_b4t_layout(D, H);
    if (L.R >= 256 || L.R <= 128 || H >= 128) return false;      // one accumulator row past R and past H carries the time cotangent
    if (L.oA2lo + 32 * L.sbo2 > L.total || L.oA1lo + 16 * L.sbo1 > L.total) return false;
    return true;
}

__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void bf16_split(const float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// x = hi + mid + lo exactly (8 + 8 + 8 significant bits; both remainders are exact Float32 subtractions)
__device__ __forceinline__ void bf16_split3(const float x, __nv_bfloat16& hi, __nv_bfloat16& mid, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    const float r1 = x - __bfloat162float(hi);
    mid = __float2bfloat16_rn(r1);
    lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}

#ifdef RNDE_A6_COMPILED_OUT      // developer switch (tools/sweep_variants.py): what the first-dt additions cost the hot loop
#define A6ON false
#else
#define A6ON A6C      // the <true> instantiation is only launched with P.a6 != 0
#endif
#ifdef RNDE_A6_NO_TAU
#define A6TAU false
#else
#define A6TAU A6ON
#endif
#ifdef RNDE_A6_NO_ZB
#define A6ZB false
#else
#define A6ZB A6ON
#endif
#ifdef RNDE_A6_NO_TASK
#define A6TASK false
#else
#define A6TASK A6ON
#endif
// A6C: compile the first-dt additions in (a6.cuh); the <false> instantiation serves detach = all and forced-step replays
template <bool A6C>
__global__ void __launch_bounds__(V2_NT, 1) bwd4tc_kernel(const KParams P) {
    constexpr int G = V2_G, NP = V2_NP, NT = V2_NT;
    extern __shared__ __align__(16) float smem[];
    unsigned char* sb = reinterpret_cast<unsigned char*>(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_ctarank();
    const int q = blockIdx.x / G;
    const int D = P.D, td = P.td, H = P.H;
    const B4TLayout L = make_b4t_layout(D, H);
    const int KB = D / 8, R = L.R, HS = L.HS, NGC = (KB + 3) / 4;
    const int r0 = rank * R;
    const int c0 = q * NP;
    const int Nloc = max(0, min(NP, P.B - c0));
    const int HSloc = max(0, min(HS, H - rank * HS));
    float* sPart = reinterpret_cast<float*>(sb + L.oPart);
    float* sD1 = reinterpret_cast<float*>(sb + L.oD1);
    float* sZb = reinterpret_cast<float*>(sb + L.oZb);
    const uint32_t sbase = smem_u32(sb);
    const uint32_t barP = sbase + L.oBar, barH = barP + 8, barM1 = barP + 16, barM2 = barP + 24;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sb + L.oBar + 32);
    const float* gW1 = P.p;
    const float* gW2 = gW1 + (size_t)H * (D + td) + H;

    // the same 4x4 state tiles as fwd4_kernel / bwd4_kernel (two K-blocks of KB rows per CTA)
    const bool own = tid < 2 * NGC * 4;
    const int cblk = tid / (NGC * 4), ctile = tid % (NGC * 4);
    const int cmt = ctile >> 2, cn0 = (ctile & 3) * 4;
    const int crow0 = cblk * KB + cmt * 4;
    const int cvalid = own ? min(4, KB - cmt * 4) : 0;

    // ---- stage the split weights as A operands (K-major, no swizzle: 8-row x 16-byte core matrices) -------------
    for (int e = tid; e < L.total / 4; e += NT) smem[e] = 0.f;
    __syncthreads();
    auto put = [&](int ohi, int olo, int sbo, int row, int k, float w) {
        __nv_bfloat16 hi, lo; bf16_split(w, hi, lo);
        const int off = (row >> 3) * sbo + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(sb + ohi + off) = hi;
        *reinterpret_cast<__nv_bfloat16*>(sb + olo + off) = lo;
    };
    for (int e = tid; e < H * R; e += NT) {          // A1[m][k] = W2[r0 + k, m]   (k fastest: coalesced)
        const int m = e / R, k = e - m * R;
        put(L.oA1hi, L.oA1lo, L.sbo1, m, k, __ldg(gW2 + (size_t)D * m + r0 + k));
    }
    for (int e = tid; e < R * H; e += NT) {          // A2[i][k] = W1[k, r0 + i]   (k fastest: coalesced)
        const int i = e / H, k = e - i * H;
        put(L.oA2hi, L.oA2lo, L.sbo2, i, k, __ldg(gW1 + (size_t)H * (r0 + i) + k));
    }
    if (P.a6 && td) {      // a6.cuh: the time columns as one more operand row each -- the MMAs then deliver sum_r W2[r, t] delta2[r, :] in
        // accumulator row H of GEMM 1 (this CTA's rows) and sum_h W1[h, t] delta1[h, :] in accumulator row R of GEMM 2
        for (int k = tid; k < R; k += NT) put(L.oA1hi, L.oA1lo, L.sbo1, H, k, __ldg(gW2 + (size_t)D * H + r0 + k));
        for (int k = tid; k < H; k += NT) put(L.oA2hi, L.oA2lo, L.sbo2, R, k, __ldg(gW1 + (size_t)H * D + k));
    }
    if (tid == 0) {
        mbar_init(barP, 1); mbar_init(barH, 1); mbar_init(barM1, 4); mbar_init(barM2, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {        // TMEM: 2 GEMMs x 4 issuers x (16 A*B_hi | 16 A*B_mid | 16 A_hi*B_lo) FP32 accumulator columns = 384 of 512
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // instruction descriptor: D = F32, A = B = BF16, both K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc32 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(32 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t idesc48 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(48 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    // Four threads (lane 0 of warps 0-3) issue the MMAs of a GEMM side by side, each into its own accumulator: the tiles
    // are tiny (M128 N32/48 K16) and a single thread issues one only every ~70 cycles.  Per k-step: A_hi x [B_hi | B_mid | B_lo]
    // (N = 48: the three terms of a B operand are stored back to back) and A_lo x [B_hi | B_mid] (N = 32, the same first 32 columns).
    const int issuer = ((tid & 31) == 0 && warp < 4) ? warp : -1;
    const uint64_t dA1hi = umma_desc(sbase + L.oA1hi, 128, L.sbo1), dA1lo = umma_desc(sbase + L.oA1lo, 128, L.sbo1);
    const uint64_t dB1 = umma_desc(sbase + L.oB1, B4T_LBO_B, L.sbb1);
    const uint64_t dA2hi = umma_desc(sbase + L.oA2hi, 128, L.sbo2), dA2lo = umma_desc(sbase + L.oA2lo, 128, L.sbo2);
    const uint64_t dB2 = umma_desc(sbase + L.oB2, B4T_LBO_B, L.sbb2);
    const int nk1 = L.K1 / 16, nk2 = L.K2 / 16;
    // GEMM 1: issuer w takes k-steps [k1lo, k1hi); GEMM 2: issuer w takes M tile w >> 1 and half (w & 1) of the k-steps
    const int k1lo = issuer >= 0 ? nk1 * issuer / 4 : 0, k1hi = issuer >= 0 ? nk1 * (issuer + 1) / 4 : 0;
    const int k2lo = issuer >= 0 ? ((issuer & 1) ? (nk2 + 1) / 2 : 0) : 0, k2hi = issuer >= 0 ? ((issuer & 1) ? nk2 : (nk2 + 1) / 2) : 0;

    // cotangents in registers
    float ubar[16], upb[16], kb[7][16];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = cn0 + j;
            ubar[i * 4 + j] = (i < cvalid && n < Nloc) ? __ldg(P.du + (size_t)D * (c0 + n) + r0 + crow0 + i) : 0.f;
            upb[i * 4 + j] = 0.f;
        }
#pragma unroll
    for (int a = 0; a < 7; ++a)
#pragma unroll
        for (int e = 0; e < 16; ++e) kb[a][e] = 0.f;
    __syncthreads();
    cluster_sync_all();

    uint32_t ev_parity = 0;
    const uint32_t bytesP = (uint32_t)((G - 1) * HSloc * NP * 4);
    const uint32_t bytesH = (uint32_t)((H - HSloc) * NP * 4);
    const size_t tileD = (size_t)D * NP, tileH = (size_t)H * NP;
    auto offD = [&](int rec, int i) -> size_t { return ((size_t)rec * P.Q + q) * tileD + (size_t)(r0 + crow0 + i) * NP + cn0; };
    const int quad = warp & 3;
    const uint32_t tlane = (uint32_t)(quad * 32) << 16;

#ifdef RNDE_TIMELINE      // phase accumulators (tools/bwd4tc_timeline.py); compiled out of the product build
    long long tl_acc[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tl_t = clock64();
#define TLB(k) do { const long long _n = clock64(); tl_acc[k] += _n - tl_t; tl_t = _n; } while (0)
#else
#define TLB(k) do { } while (0)
#endif
    // k of the next record (for act2'), fetched one evaluation ahead so that its L2/HBM latency hides behind GEMM 2
    float4 kvn[4];
    int kvn_rec = -1;
    // VJP of record `rec`: cur = kbar of that evaluation (in), zb = W1^T delta1 for the tile (out)
    auto vjp = [&](const float (&cur)[16], float (&zb)[16], const int rec, const int rec_next) {
        TLB(0);
        if (tid == 0) { mbar_expect_tx(barP, bytesP); mbar_expect_tx(barH, bytesH); }
        if (rec_next >= 0) {        // pull the next record's tiles towards L2
            if (own) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (i < cvalid) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tapeK + offD(rec_next, i)));
            }
            if (tid < HSloc * 4)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tapeH + ((size_t)rec_next * P.Q + q) * tileH + (size_t)(rank * HS + (tid >> 2)) * NP + (tid & 3) * 4));
        }
        // delta2 = kbar * act2'(k): to the tape (wgrad operand) and, split, into the B operand of GEMM 1
        if (own) {
            float d2v[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float4 d2 = make_float4(cur[i * 4], cur[i * 4 + 1], cur[i * 4 + 2], cur[i * 4 + 3]);
                if (i < cvalid) {
                    float* tp = P.tapeK + offD(rec, i);
                    if (P.act2 == RNDE_ACT_TANH) {
#ifdef RNDE_EXP_NOKV
                        const float4 kv = make_float4(0.1f, 0.2f, 0.3f, 0.4f);
#else
                        const float4 kv = (kvn_rec == rec) ? kvn[i] : __ldcg(reinterpret_cast<const float4*>(tp));
#endif
                        d2.x = d2.x * (1.f - kv.x * kv.x); d2.y = d2.y * (1.f - kv.y * kv.y);
                        d2.z = d2.z * (1.f - kv.z * kv.z); d2.w = d2.w * (1.f - kv.w * kv.w);
                    }
                    *reinterpret_cast<float4*>(sZb + (crow0 + i) * NP + cn0) = d2;      // delta2 replaces k on the tape: bulk store below
                } else d2 = make_float4(0.f, 0.f, 0.f, 0.f);
                d2v[i][0] = d2.x; d2v[i][1] = d2.y; d2v[i][2] = d2.z; d2v[i][3] = d2.w;
            }
            // element (n, k = local row): crow0 is even, so rows (i, i+1) are an aligned bf16 pair inside one 8-group
#ifndef RNDE_EXP_NOSPLIT
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = cn0 + j;
#pragma unroll
                for (int i = 0; i < 4; i += 2) {
                    const int k = crow0 + i;
                    const int off = (n >> 3) * L.sbb1 + (k >> 3) * B4T_LBO_B + (n & 7) * 16 + (k & 7) * 2;
                    __nv_bfloat16 h2[2], m2[2], l2[2];
                    bf16_split3(d2v[i][j], h2[0], m2[0], l2[0]);
                    bf16_split3(d2v[i + 1][j], h2[1], m2[1], l2[1]);
                    if (i + 1 < cvalid) {       // rows past cvalid belong to the next K-block's tiles: never touch them
                        *reinterpret_cast<uint32_t*>(sb + L.oB1 + off) = *reinterpret_cast<const uint32_t*>(h2);
                        *reinterpret_cast<uint32_t*>(sb + L.oB1 + 2 * L.sbb1 + off) = *reinterpret_cast<const uint32_t*>(m2);
                        *reinterpret_cast<uint32_t*>(sb + L.oB1 + 4 * L.sbb1 + off) = *reinterpret_cast<const uint32_t*>(l2);
                    } else if (i < cvalid) {
                        *reinterpret_cast<__nv_bfloat16*>(sb + L.oB1 + off) = h2[0];
                        *reinterpret_cast<__nv_bfloat16*>(sb + L.oB1 + 2 * L.sbb1 + off) = m2[0];
                        *reinterpret_cast<__nv_bfloat16*>(sb + L.oB1 + 4 * L.sbb1 + off) = l2[0];
                    }
                }
            }
#else
            if (d2v[0][0] == 1234.5f) sb[L.oB1] = 1;
#endif
        }
#ifndef RNDE_EXP_NOFENCE
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
        __syncthreads();
        if (tid == 0) {     // the delta2 tile of this CTA's rows -> tape (wgrad operand), by the bulk-store engine
            bulk_store(P.tapeK + (((size_t)rec * P.Q + q) * D + r0) * NP, sZb, (uint32_t)(R * NP * 4));
            bulk_commit();
        }
        TLB(1);
        if (issuer >= 0) {     // GEMM 1 (a K quarter per issuer; descriptors advance by 2 core matrices = 256 bytes per k-step)
            tc_fence_after();
            const uint32_t dcol = tmem_base + 48 * issuer;
            for (int ks = k1lo; ks < k1hi; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 16), advb = (uint64_t)(ks * (2 * B4T_LBO_B / 16));
                tc_mma_bf16(dcol, dA1hi + adv, dB1 + advb, idesc48, ks == k1lo ? 0u : 1u);
                tc_mma_bf16(dcol, dA1lo + adv, dB1 + advb, idesc32, 1u);
            }
            tc_commit(barM1);
        }
        TLB(2);
        mbar_wait(barM1, ev_parity);
        tc_fence_after();
        TLB(3);
        if (warp < 4) {     // hbar partial rows (TMEM lane = hidden unit) -> the reducer CTA of that hidden slice
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
            for (int term = 2; term >= 0; --term)       // lo, mid, hi: the small terms are added first
                for (int is = 0; is < 4; ++is) {        // 4 issuers (k-step quarters)
                    float t[16];
                    tmem_ld16(tmem_base + tlane + 48 * is + 16 * term, t);
                    if (nk1 * is / 4 < nk1 * (is + 1) / 4) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += t[j];
                    }
                }
            const int m = quad * 32 + lane;
            if (m == H && A6TAU) {      // time cotangent of layer 2, this CTA's rows (a6.cuh): kept per record and column, summed in the a6 task
                float4* tp = reinterpret_cast<float4*>(P.a6_tau + (((size_t)rec * P.Q + q) * G + rank) * 32);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) tp[c4] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
            }
            if (m < H) {
                const int d = m / HS, ml = m - d * HS;
                float* dst = sPart + (rank * HS + ml) * NP;
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const float4 o = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
                    if (d == rank) *reinterpret_cast<float4*>(dst + c4 * 4) = o;
                    else st_async_f4(mapa_u32(smem_u32(dst + c4 * 4), d), o, mapa_u32(barP, d));
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        mbar_wait(barP, ev_parity);
        TLB(4);
        if (tid < HSloc * 4) {      // sum over the 4 CTAs, delta1 = hbar * act1'(h), all-gather, tape
            const int ml = tid >> 2, n4 = (tid & 3) * 4;
            const int m = rank * HS + ml;
            float4 s = *reinterpret_cast<const float4*>(sPart + ml * NP + n4);
#pragma unroll
            for (int c = 1; c < G; ++c) {
                const float4 pc = *reinterpret_cast<const float4*>(sPart + (c * HS + ml) * NP + n4);
                s.x += pc.x; s.y += pc.y; s.z += pc.z; s.w += pc.w;
            }
            const size_t oh = ((size_t)rec * P.Q + q) * tileH + (size_t)m * NP + n4;
            if (P.act1 == RNDE_ACT_TANH) {
                const float4 hv = __ldcg(reinterpret_cast<const float4*>(P.tapeH + oh));
                s.x *= (1.f - hv.x * hv.x); s.y *= (1.f - hv.y * hv.y); s.z *= (1.f - hv.z * hv.z); s.w *= (1.f - hv.w * hv.w);
            }
            float* dst = sD1 + m * NP + n4;
            *reinterpret_cast<float4*>(dst) = s;
            const uint32_t da = smem_u32(dst);
#pragma unroll
            for (int d = 1; d < G; ++d) {
                const int peer = (rank + d) & (G - 1);
                st_async_f4(mapa_u32(da, peer), s, mapa_u32(barH, peer));
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (tid == 0) bulk_wait_read();      // the delta2 tile in sZb has left: sZb becomes GEMM 2's B operand, then the transposition tile
        __syncthreads();
        if (tid == 0 && HSloc > 0) {      // this CTA's slice of delta1 -> tape
            bulk_store(P.tapeD1 + (((size_t)rec * P.Q + q) * H + rank * HS) * NP, sD1 + rank * HS * NP, (uint32_t)(HSloc * NP * 4));
            bulk_commit();
        }
        mbar_wait(barH, ev_parity);
        TLB(5);
        // delta1 (H x 16, FP32) -> split B operand of GEMM 2: one 16-byte chunk = 8 consecutive hidden units of one column.
        // The operand overlays sZb, whose delta2 tile was handed to the bulk-store engine before GEMM 1: thread 0 has waited
        // for that read (below, before the barrier that precedes this loop).
        for (int ch = tid; ch < NP * (L.K2 / 8); ch += NT) {
            const int n = ch % NP, k0 = (ch / NP) * 8;
            __nv_bfloat16 h8[8], m8[8], l8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float x = (k0 + i < H) ? sD1[(k0 + i) * NP + n] : 0.f;
                bf16_split3(x, h8[i], m8[i], l8[i]);
            }
            const int off = (n >> 3) * L.sbb2 + (k0 >> 3) * B4T_LBO_B + (n & 7) * 16;
            *reinterpret_cast<uint4*>(sb + L.oB2 + off) = *reinterpret_cast<const uint4*>(h8);
            *reinterpret_cast<uint4*>(sb + L.oB2 + 2 * L.sbb2 + off) = *reinterpret_cast<const uint4*>(m8);
            *reinterpret_cast<uint4*>(sb + L.oB2 + 4 * L.sbb2 + off) = *reinterpret_cast<const uint4*>(l8);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        TLB(6);
        if (issuer >= 0) {     // GEMM 2: M tile issuer >> 1, half of the k-steps each
            tc_fence_after();
            const uint32_t dcol = tmem_base + 192 + 48 * issuer;
            const uint64_t aoff = (uint64_t)(((issuer >> 1) * 16 * L.sbo2) >> 4);
            for (int ks = k2lo; ks < k2hi; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 16), advb = (uint64_t)(ks * (2 * B4T_LBO_B / 16));
                tc_mma_bf16(dcol, dA2hi + aoff + adv, dB2 + advb, idesc48, ks == k2lo ? 0u : 1u);
                tc_mma_bf16(dcol, dA2lo + aoff + adv, dB2 + advb, idesc32, 1u);
            }
            tc_commit(barM2);
        }
        if (own && rec_next >= 0 && P.act2 == RNDE_ACT_TANH) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < cvalid) kvn[i] = __ldcg(reinterpret_cast<const float4*>(P.tapeK + offD(rec_next, i)));
        }
        kvn_rec = rec_next;
        mbar_wait(barM2, ev_parity);
        tc_fence_after();
        TLB(7);
        {   // TMEM lane = local row; transpose through shared memory into the 4x4 register tiles
            const int row = (warp < 4 ? 0 : 128) + quad * 32 + lane;
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
            for (int term = 2; term >= 0; --term)       // lo, mid, hi
                for (int is = 0; is < 2; ++is) {        // the 2 issuers of this M tile (k-step halves)
                    const bool used = is ? (nk2 > (nk2 + 1) / 2) : true;
                    float t[16];
                    tmem_ld16(tmem_base + tlane + 192 + (warp < 4 ? 0u : 96u) + 48 * is + 16 * term, t);
                    if (used) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] += t[j];
                    }
                }
            if (row == R && A6TAU) {      // time cotangent of layer 1 (the same in every CTA of the cluster)
                float4* tp = reinterpret_cast<float4*>(P.a6_tau + (((size_t)rec * P.Q + q) * G + rank) * 32 + 16);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) tp[c4] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
            }
            if (row < R) {
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
                    *reinterpret_cast<float4*>(sZb + row * NP + c4 * 4) = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
            }
        }
        tc_fence_before();
        __syncthreads();
        TLB(8);
        ev_parity ^= 1u;
#pragma unroll
        for (int e = 0; e < 16; ++e) zb[e] = 0.f;
        if (own) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (i < cvalid) {
                    const float4 z4 = *reinterpret_cast<const float4*>(sZb + (crow0 + i) * NP + cn0);
                    zb[i * 4] = z4.x; zb[i * 4 + 1] = z4.y; zb[i * 4 + 2] = z4.z; zb[i * 4 + 3] = z4.w;
                }
            }
        }
    };

    const float atol = P.abstol, rtol = P.reltol;
    const float cntf = (float)P.norm_count;
    const float stab = rn_divf(1.0f, (float)TS_STABILITY_SIZE);

    // task list, newest evaluation first: for s = nsteps-1..0: stages 7..2; then the initial record 0
    // (+ with P.a6 >= 2 one task before the last: the evaluation of the initial-dt heuristic, a6.cuh)
    const int na6 = (A6TASK && P.a6 >= 2) ? 1 : 0;
    const int ntask = 6 * P.nsteps + 1 + na6;
    float* sA6 = reinterpret_cast<float*>(sb + L.oA6);
    float dt = 0.f, gB = 0.f;
    bool use_eig = false;
    int recU1 = 0, recG6 = 0;
    // With the first-dt additions compiled in, the loop runs the steps only and the tail tasks (the a6 task, record 0) follow it with
    // a second inlined copy of the VJP: branches on them inside the loop cost every record (profiles/r2p_a6_sweep_variants.txt).
    const int nloop = A6C ? 6 * P.nsteps : ntask;
    for (int task = 0; task < nloop; ++task) {
        const bool last = !A6C && (task == ntask - 1);
        const bool a6task = false;
        const bool tail = last || a6task;
        const int s = tail ? 0 : P.nsteps - 1 - task / 6;
        const int i = tail ? 7 : 7 - task % 6;
        const int rec = last ? 0 : (a6task ? P.rec_x : 6 * s + i - 1);
        TLB(10);
        if (!tail && i == 7) {
            // ---- entering step s: reset per-step cotangents, add the saved-value cotangents --------
            const StepRec sr = P.steps[s];
            dt = sr.dt;
            const float EEst = sr.eest, eig = sr.eig, n1 = sr.n1, n2 = sr.n2;
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
            use_eig = (eigbar != 0.f) && (n1 > 0.f) && (n2 > 0.f);
            const float gE = use_eest ? eestbar / (cntf * EEst) : 0.f;
            const float n1b = use_eig ? eigbar / n2 : 0.f;
            const float n2b = use_eig ? -eigbar * n1 / (n2 * n2) : 0.f;
            const float gA = use_eig ? n1b / (cntf * n1) : 0.f;
            gB = use_eig ? n2b / (cntf * n2) : 0.f;
            recU1 = 6 * s + 6; recG6 = 6 * s + 5;
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int e = 0; e < 16; ++e) kb[a][e] = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) upb[e] = 0.f;
            TLB(11);
            if ((use_eest || use_eig) && own) {
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    if (ii < cvalid) {
                        const float4 up4 = __ldcg(reinterpret_cast<const float4*>(P.tapeZ + offD(6 * s, ii)));
                        const float4 un4 = __ldcg(reinterpret_cast<const float4*>(P.tapeZ + offD(recU1, ii)));
                        float kv[7][4];
#pragma unroll
                        for (int j = 0; j < 7; ++j) {
                            const float4 k4 = __ldcg(reinterpret_cast<const float4*>(P.tapeK + offD(6 * s + j, ii)));
                            kv[j][0] = k4.x; kv[j][1] = k4.y; kv[j][2] = k4.z; kv[j][3] = k4.w;
                        }
                        float4 g64 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (use_eig) g64 = __ldcg(reinterpret_cast<const float4*>(P.tapeZ + offD(recG6, ii)));
                        const float upv[4] = {up4.x, up4.y, up4.z, up4.w}, unv[4] = {un4.x, un4.y, un4.z, un4.w};
                        const float g6v[4] = {g64.x, g64.y, g64.z, g64.w};
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) {
                            const int e = ii * 4 + jj;
                            const float up = upv[jj], un = unv[jj];
                            const float live = (cn0 + jj < Nloc) ? 1.f : 0.f;      // columns past the batch contribute nothing
                            if (use_eest) {
                                float ssum = c_BT[1] * kv[0][jj];
#pragma unroll
                                for (int j = 2; j <= 7; ++j) ssum = rn_fmaf(c_BT[j], kv[j - 1][jj], ssum);
                                const float ut = dt * ssum;
                                const float a0 = fabsf(up), a1 = fabsf(un);
                                const float mx = a0 > a1 ? a0 : a1;
                                const float rden = __frcp_rn(rn_fmaf(mx, rtol, atol));     // one reciprocal instead of three IEEE divisions
                                const float at = ut * rden;
                                const float ab = live * gE * at;
                                const float utb = ab * rden;
                                const float mb = (-ab * at * rden) * rtol;
                                // max(|up|, |un|): the larger branch gets the cotangent, a tie splits it (branch-free selects;
                                // the strictly larger magnitude is non-zero, so the three-way sign equals the two-way one there)
                                const float wu = a0 > a1 ? 1.f : (a1 > a0 ? 0.f : 0.5f);
                                const float su = up > 0.f ? 1.f : (up < 0.f ? -1.f : 0.f), sn = un > 0.f ? 1.f : (un < 0.f ? -1.f : 0.f);
                                upb[e] += wu * mb * su;
                                ubar[e] += (1.f - wu) * mb * sn;
                                const float dtu = dt * utb;
#pragma unroll
                                for (int j = 1; j <= 7; ++j) kb[j - 1][e] += c_BT[j] * dtu;
                            }
                            if (use_eig) {
                                const float ga = live * gA * (kv[6][jj] - kv[5][jj]);
                                const float gb = live * gB * (un - g6v[jj]);
                                kb[6][e] += ga; kb[5][e] -= ga;
                                ubar[e] += gb;       // the matching -gb on g6 is applied after stage 6's VJP
                            }
                        }
                    }
                }
            }
        }
        TLB(12);
#ifdef RNDE_TIMELINE
        __syncthreads();      // separates this thread's own time from waiting for the other warps
#endif
        TLB(9);
        // kbar of this evaluation
        float cur[16], zb[16];
        switch (i) {
#define RNDE_CUR(J) case J: _Pragma("unroll") for (int e = 0; e < 16; ++e) cur[e] = kb[J - 1][e]; break;
            RNDE_CUR(2) RNDE_CUR(3) RNDE_CUR(4) RNDE_CUR(5) RNDE_CUR(6)
            default: _Pragma("unroll") for (int e = 0; e < 16; ++e) cur[e] = kb[6][e]; break;
#undef RNDE_CUR
        }
        int rec_next = -1;
        if (!last) {
            const int t2 = task + 1;
            if (t2 == ntask - 1) rec_next = 0;
            else if (na6 && t2 == ntask - 2) rec_next = P.rec_x;
            else { const int s2 = P.nsteps - 1 - t2 / 6, i2 = 7 - t2 % 6; rec_next = 6 * s2 + i2 - 1; }
        }
        if (!last && i == 2 && s > 0 && own) {   // one evaluation ahead of step s-1's entry: start its records towards L2
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                if (ii < cvalid) {
#pragma unroll
                    for (int j = 0; j < 7; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tapeK + offD(6 * (s - 1) + j, ii)));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tapeZ + offD(6 * (s - 1), ii)));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tapeZ + offD(6 * (s - 1) + 6, ii)));
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(P.tapeZ + offD(6 * (s - 1) + 5, ii)));
                }
            }
        }
        vjp(cur, zb, rec, rec_next);
        if (last) {
            // initial fsalfirst = f(u0, t0): dx = ubar + zbar
            if (P.dx && own) {
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int n = cn0 + jj;
                        if (ii < cvalid && n < Nloc) P.dx[(size_t)D * (c0 + n) + r0 + crow0 + ii] = ubar[ii * 4 + jj] + zb[ii * 4 + jj];
                    }
            }
            break;
        }
        // cotangent of the stage input z_i
        if (i == 7) {
#pragma unroll
            for (int e = 0; e < 16; ++e) zb[e] += ubar[e];
        }
        if (i == 6 && use_eig && own) {
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                if (ii < cvalid) {
                    const float4 un4 = __ldcg(reinterpret_cast<const float4*>(P.tapeZ + offD(recU1, ii)));
                    const float4 g64 = __ldcg(reinterpret_cast<const float4*>(P.tapeZ + offD(recG6, ii)));
                    const float unv[4] = {un4.x, un4.y, un4.z, un4.w}, g6v[4] = {g64.x, g64.y, g64.z, g64.w};
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (cn0 + jj < Nloc) zb[ii * 4 + jj] -= gB * (unv[jj] - g6v[jj]);
                }
            }
        }
        if (A6ZB && own && (s == 0 || s == P.nsteps - 1)) {      // a6.cuh: the explicit dt of z_i needs this cotangent later (first and last step)
            const int slot = (s == 0 ? 0 : 6) + i - 2;
#pragma unroll
            for (int ii = 0; ii < 4; ++ii)
                if (ii < cvalid) *reinterpret_cast<float4*>(P.a6_zb + offD(slot, ii)) = make_float4(zb[ii * 4], zb[ii * 4 + 1], zb[ii * 4 + 2], zb[ii * 4 + 3]);
        }
        switch (i) {
            case 2: bwd_distribute<2>(kb, zb, dt); break;
            case 3: bwd_distribute<3>(kb, zb, dt); break;
            case 4: bwd_distribute<4>(kb, zb, dt); break;
            case 5: bwd_distribute<5>(kb, zb, dt); break;
            case 6: bwd_distribute<6>(kb, zb, dt); break;
            default: bwd_distribute<7>(kb, zb, dt); break;
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) upb[e] += zb[e];
        if (i == 2) {
            // hand over to the previous step: u_new(prev) = uprev, k7(prev) = k1
#pragma unroll
            for (int e = 0; e < 16; ++e) { ubar[e] = upb[e]; kb[6][e] = kb[0][e]; }
        }
    }
    if constexpr (A6C) {
        for (int tt = 0; tt < 1 + na6; ++tt) {
            const bool a6task = na6 && tt == 0;
            const int rec = a6task ? P.rec_x : 0;
            float cur[16], zb[16];
            if (a6task) RNDE_A6_TASK_PRE(sA6, cur, sPart, a6_sweep_partial(P, own, cvalid, offD(0, 0), (size_t)P.Q * tileD, NP, q, rank, G));
            else {
#pragma unroll
                for (int e = 0; e < 16; ++e) cur[e] = kb[6][e];
            }
            vjp(cur, zb, rec, a6task ? 0 : -1);
            if (a6task) {      // its time cotangent: the two sums the tensor cores left for this CTA
                const float* tp = P.a6_tau + (((size_t)rec * P.Q + q) * G + rank) * 32;
                RNDE_A6_TASK_POST(sA6, zb, sPart, (tid < 16 && td) ? (double)(__ldcg(tp + tid) + (rank == 0 ? __ldcg(tp + 16 + tid) : 0.f)) : 0.0);
                continue;
            }
            // initial fsalfirst = f(u0, t0): dx = ubar + zbar
            if (P.dx && own) {
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int n = cn0 + jj;
                        if (ii < cvalid && n < Nloc) P.dx[(size_t)D * (c0 + n) + r0 + crow0 + ii] = ubar[ii * 4 + jj] + zb[ii * 4 + jj];
                    }
            }
        }
    }
#ifdef RNDE_TIMELINE
    if (P.dbg && blockIdx.x == 0 && tid == 0) for (int k = 0; k < 14; ++k) P.dbg[k] = tl_acc[k];
#endif
    if (A6ON && P.a6 == 1) {      // the two extra VJPs run as separate launches (reference-exact data parallel): leave this CTA's partial sum
        const double part = a6_sweep_partial(P, own, cvalid, offD(0, 0), (size_t)P.Q * tileD, NP, q, rank, G);
        a6_block_sum<NT>(part, sPart, P.a6_part);
    }
    if (tid == 0) bulk_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    }
    cluster_sync_all();
}

}  // namespace rnde
