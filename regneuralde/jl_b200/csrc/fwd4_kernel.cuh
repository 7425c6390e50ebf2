// fwd4_kernel.cuh -- "cluster-4 / register-resident" forward stepper: the north-star kernel for
// the MNIST-shaped field (D = 8*kblock, e.g. 784 = 8*98; H <= 128).
//
// Same algorithm and the same canonical arithmetic as fwd_kernel.cuh (bit-identical results);
// the decomposition is chosen from measurements on B200 (profiles/r1_microbench*.txt):
//   * at most 15 clusters of 8 CTAs with >160 KB shared memory are co-resident, but 33 clusters
//     of 4 are -> 32 clusters x 16 columns cover the reference batch of 512 on 128 SMs;
//   * DSMEM pushes + barrier.cluster cost ~2400 cycles per exchange at cluster 8 and the barrier's
//     release/acquire compiles to MEMBAR.ALL.GPU + CCTL.IVALL; st.async with mbarrier complete_tx
//     moves the same bytes in ~700 cycles with no cluster barrier at all.
// Layout per CTA (rank r of 4): state rows [r*2*KB, (r+1)*2*KB) = two canonical K-blocks, NP = 16
// columns.  Every thread owns one 4x4 (rows x columns) tile of the state: uprev, k1..k7 and the
// current stage input live in REGISTERS for the whole solve (the thread that produces a layer-2
// output tile is the one that consumes it in the stage combinations and the error estimate), so
// shared memory only holds the weights (W1 slice 2KB x H, W2 slice H x 2KB), the stage input Z
// (operand of layer 1) and the hidden exchange buffers.
// Per field evaluation:  Z -> [A] partial pre-activations, one thread tile per (K-block, 4 hidden,
// 4 cols) -> block pair-sum -> st.async scatter to the CTA reducing that hidden slice -> [B] fixed
// order sum over the 4 CTAs, time column, bias, tanh, st.async all-gather -> [C] layer 2 for the
// thread's own tile (full K = H chain), time column, bias, tanh -> registers.
#pragma once
#include "common.cuh"
#include "fwd_kernel.cuh"

namespace rnde {

constexpr int V2_G = 4;
constexpr int V2_NP = 16;
constexpr int V2_NT = 256;

struct V2Layout {
    int KB, KBP, R, RPAD, HP, HS, NGC, NGH;
    int oW1, oW1t, ob1, oW2, oW2t, ob2, oZ, oP1, oPart, oH, oRed, oCP, oTot, oCtl, oBar, oKt, total;
};

__host__ __device__ inline V2Layout make_v2_layout(int D, int H) {
    V2Layout L;
    L.KB = D / 8;
    L.KBP = round_up(L.KB, 4);
    L.R = 2 * L.KB;
    L.RPAD = 2 * L.KBP;
    L.HP = round_up(H, 4);
    L.HS = (H + V2_G - 1) / V2_G;
    L.NGC = (L.KB + 3) / 4;
    L.NGH = (H + 3) / 4;
    int o = 0;
    L.oW1 = o; o += L.R * L.HP;            // [k = local row][m = hidden]
    L.oW1t = o; o += L.HP;
    L.ob1 = o; o += L.HP;
    L.oW2 = o; o += H * L.RPAD;            // [k = hidden][m = padded local row]
    L.oW2t = o; o += L.RPAD;
    L.ob2 = o; o += L.RPAD;
    L.oZ = o; o += L.R * V2_NP;
    L.oP1 = o; o += L.HP * V2_NP;
    L.oPart = o; o += V2_G * L.HS * V2_NP;
    L.oH = o; o += L.HP * V2_NP;
    L.oRed = o; o += 3 * 2 * L.NGC * V2_NP;
    L.oCP = o; o += 3 * V2_G * V2_NP;
    L.oTot = o; o += 4;
    L.oCtl = o; o += 32;
    L.oBar = o; o += 8;                    // two 8-byte mbarriers (+pad), 16B aligned
    L.oKt = o; o += L.R * V2_NP;           // layer-2 output staged for the bulk store to the tape
    L.total = o;
    return L;
}

__host__ inline bool v2_shape_ok(int D, int H) {
    if (D % 8 != 0 || D < 64) return false;
    const int KB = D / 8;
    if (2 * ((KB + 3) / 4) * 4 > V2_NT) return false;     // phase-C tiles
    if (2 * ((H + 3) / 4) * 4 > V2_NT) return false;      // phase-A tiles
    if (((H + V2_G - 1) / V2_G) * 4 > V2_NT) return false;
    return true;
}

// ---- mbarrier / st.async wrappers ---------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void st_async_f4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(remote_addr), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w), "r"(remote_bar) : "memory");
}

// ---- bulk (TMA engine) stores shared -> global: the tape tiles leave the SM without occupying the threads' store path ----
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// z_I = uprev + dt * sum_j a_Ij k_j with the canonical association, specialised per stage so the
// 16 elements of the thread's tile are straight-line independent FMA chains.
template <int I>
__device__ __forceinline__ void combo_stage(const float (&kk)[7][16], const float (&uprev)[16], const float dt, const float a2, float (&zc)[16]) {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        if constexpr (I == 2) {
            zc[e] = rn_fmaf(a2, kk[0][e], uprev[e]);
        } else {
            float s = c_A[I][1] * kk[0][e];
#pragma unroll
            for (int j = 2; j < I; ++j) s = rn_fmaf(c_A[I][j], kk[j - 1][e], s);
            zc[e] = rn_fmaf(dt, s, uprev[e]);
        }
    }
}

// HC / KBC: compile-time hidden size and K-block (0 = take them from the launch parameters).  The
// <100, 98> instantiation is the MNIST field: constant strides and trip counts let ptxas use
// immediate offsets and hoist the shared-memory operand loads several iterations ahead.
template <int HC, int KBC>
__global__ void __launch_bounds__(V2_NT, 1) fwd4_kernel(const KParams P) {
    constexpr int G = V2_G, NP = V2_NP, NT = V2_NT;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int rank = (int)cluster_ctarank();
    const int q = blockIdx.x / G;
    const int D = P.D, td = P.td;
    const int H = (HC > 0) ? HC : P.H;
    const V2Layout L = make_v2_layout(D, H);
    const int KB = (KBC > 0) ? KBC : L.KB;
    const int KBP = (KBC > 0) ? ((KBC + 3) / 4 * 4) : L.KBP;
    const int R = 2 * KB, RPAD = 2 * KBP;
    const int HP = (HC > 0) ? ((HC + 3) / 4 * 4) : L.HP;
    const int HS = L.HS, NGC = (KB + 3) / 4, NGH = (H + 3) / 4;
    const int r0 = rank * R;
    const int c0 = q * NP;
    const int Nloc = max(0, min(NP, P.B - c0));
    const int HSloc = max(0, min(HS, H - rank * HS));
    float* sW1 = smem + L.oW1; float* sW1t = smem + L.oW1t; float* sb1 = smem + L.ob1;
    float* sW2 = smem + L.oW2; float* sW2t = smem + L.oW2t; float* sb2 = smem + L.ob2;
    float* sZ = smem + L.oZ; float* sP1 = smem + L.oP1; float* sPart = smem + L.oPart; float* sH = smem + L.oH;
    float* sRed = smem + L.oRed; float* sCP = smem + L.oCP; float* sTot = smem + L.oTot;
    float* sKt = smem + L.oKt;
    Ctl* ctl = reinterpret_cast<Ctl*>(smem + L.oCtl);
    const uint32_t barP = smem_u32(smem + L.oBar), barH = barP + 8;

    const float* gW1 = P.p;
    const float* gb1 = gW1 + (size_t)H * (D + td);
    const float* gW2 = gb1 + H;
    const float* gb2 = gW2 + (size_t)D * (H + td);

    // ---- thread roles -------------------------------------------------------------------------
    // phase C / state ownership: tile (block cblk, row group cmt, column group cnt)
    const bool own = tid < 2 * NGC * 4;
    const int cblk = tid / (NGC * 4), ctile = tid % (NGC * 4);
    const int cmt = ctile >> 2, cn0 = (ctile & 3) * 4;
    const int crow0 = cblk * KB + cmt * 4;                 // dense local row of the tile
    const int cprow0 = cblk * KBP + cmt * 4;               // padded row (W2 slice, biases)
    const int cvalid = own ? min(4, KB - cmt * 4) : 0;     // rows of the tile inside the block
    // phase A: tile (K-half akh, hidden group amt, column group)
    const bool actA = tid < 2 * NGH * 4;
    const int akh = tid / (NGH * 4), atile = tid % (NGH * 4);
    const int am0 = (atile >> 2) * 4, an0 = (atile & 3) * 4;

    // ---- stage weights ------------------------------------------------------------------------
    for (int e = tid; e < R * HP; e += NT) {
        const int k = e / HP, m = e - k * HP;
        sW1[e] = (m < H) ? __ldg(gW1 + (size_t)(r0 + k) * H + m) : 0.f;
    }
    for (int e = tid; e < H * RPAD; e += NT) {
        const int k = e / RPAD, mp = e - k * RPAD;
        const int b = mp / KBP, i = mp - b * KBP;
        sW2[e] = (i < KB) ? __ldg(gW2 + (size_t)D * k + r0 + b * KB + i) : 0.f;
    }
    for (int m = tid; m < HP; m += NT) {
        sW1t[m] = (td && m < H) ? __ldg(gW1 + (size_t)H * D + m) : 0.f;
        sb1[m] = (m < H) ? __ldg(gb1 + m) : 0.f;
    }
    for (int mp = tid; mp < RPAD; mp += NT) {
        const int b = mp / KBP, i = mp - b * KBP;
        sW2t[mp] = (td && i < KB) ? __ldg(gW2 + (size_t)D * H + r0 + b * KB + i) : 0.f;
        sb2[mp] = (i < KB) ? __ldg(gb2 + r0 + b * KB + i) : 0.f;
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
    __syncthreads();
    cluster_sync_all();     // weights staged, mbarriers initialised and visible cluster-wide

    unsigned norm_seq = 0, bar_gen = 0;
    // exact mode: norms are numbered consecutively across launches (same count on every rank)
    unsigned* xseq_ptr = reinterpret_cast<unsigned*>(P.peers[P.rank]) + P.flag_off + 32;
    const unsigned xseq_base = (P.nranks > 1) ? *xseq_ptr : 0u;
    uint32_t ev_parity = 0;
    int dbg_n = 0;
    auto mark = [&](int id) {
#ifdef RNDE_TIMELINE      // phase timeline for tools/gpu_check.py timeline; compiled out of the product build
        if (P.dbg && blockIdx.x == 0 && tid == 0 && dbg_n < 4000) { P.dbg[dbg_n * 2] = id; P.dbg[dbg_n * 2 + 1] = clock64(); dbg_n++; }
#else
        (void)id; (void)dbg_n;
#endif
    };
    const uint32_t bytesP = (uint32_t)((G - 1) * HSloc * NP * 4);
    const uint32_t bytesH = (uint32_t)((H - HSloc) * NP * 4);

    int pend_rec = -1;      // record whose layer-2 output sits in sKt, not yet handed to the bulk-store engine
    // ---- one field evaluation: out = f(zin, tstage); zin (registers) is also staged to sZ ---------
    auto rhs = [&](const float (&zin)[16], float (&out)[16], const float tstage, const int rec) {
        mark(0);
        if (tid == 0) { mbar_expect_tx(barP, bytesP); mbar_expect_tx(barH, bytesH); }
        if (own) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < cvalid) *reinterpret_cast<float4*>(sZ + (crow0 + i) * NP + cn0) = make_float4(zin[i * 4], zin[i * 4 + 1], zin[i * 4 + 2], zin[i * 4 + 3]);
        }
        if (rec >= 0 || pend_rec >= 0) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (tid == 0) {     // tape: this evaluation's input tile (sZ) and the previous evaluation's output tile (sKt) as bulk stores
            if (rec >= 0) bulk_store(P.tapeZ + (((size_t)rec * P.Q + q) * D + r0) * NP, sZ, (uint32_t)(R * NP * 4));
            if (pend_rec >= 0) bulk_store(P.tapeK + (((size_t)pend_rec * P.Q + q) * D + r0) * NP, sKt, (uint32_t)(R * NP * 4));
            if (rec >= 0 || pend_rec >= 0) bulk_commit();
        }
        pend_rec = -1;
        mark(1);
        // phase A: one canonical K-block (KB rows) per thread tile, software pipelined operand loads
        float acc[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0.f;
        if (actA) {
            const float* wp = sW1 + (akh * KB) * HP + am0;
            const float* xp = sZ + (akh * KB) * NP + an0;
            float4 w = *reinterpret_cast<const float4*>(wp), x = *reinterpret_cast<const float4*>(xp);
            // operands of iteration k+1 are fetched before the 16 FMAs of iteration k (the last
            // iteration re-fetches row KB-1, harmlessly); 98 = 14 x 7 for the MNIST instantiation
#pragma unroll (KBC > 0 ? 14 : 4)
            for (int k = 0; k < KB; ++k) {
                const int kn = (k + 1 < KB) ? k + 1 : k;
                const float4 wn = *reinterpret_cast<const float4*>(wp + kn * HP);
                const float4 xn = *reinterpret_cast<const float4*>(xp + kn * NP);
                const float wv[4] = {w.x, w.y, w.z, w.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i * 4 + j] = rn_fmaf(wv[i], xv[j], acc[i * 4 + j]);
                w = wn; x = xn;
            }
            if (akh == 1) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(sP1 + (am0 + i) * NP + an0) = make_float4(acc[i * 4], acc[i * 4 + 1], acc[i * 4 + 2], acc[i * 4 + 3]);
            }
        }
        mark(2);
        __syncthreads();
        mark(3);
        if (actA && akh == 0) {
            // canonical pair sum of this CTA's two blocks, then scatter rows to their reducer CTA
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int m = am0 + i;
                if (m < H) {
                    const float4 o = *reinterpret_cast<const float4*>(sP1 + m * NP + an0);
                    const float4 v = make_float4(acc[i * 4] + o.x, acc[i * 4 + 1] + o.y, acc[i * 4 + 2] + o.z, acc[i * 4 + 3] + o.w);
                    const int d = m / HS, ml = m - d * HS;
                    float* dst = sPart + (rank * HS + ml) * NP + an0;
                    if (d == rank) *reinterpret_cast<float4*>(dst) = v;
                    else st_async_f4(mapa_u32(smem_u32(dst), d), v, mapa_u32(barP, d));
                }
            }
        }
        mark(4);
        if (tid == 0) bulk_wait_read();      // sZ, sKt and the hidden slice of the previous evaluation may be overwritten from here on
        __syncthreads();
        mark(5);
        mbar_wait(barP, ev_parity);
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
        // phase C: layer 2 for the thread's own tile (full K = H chain)
#pragma unroll
        for (int e = 0; e < 16; ++e) acc[e] = 0.f;
        if (own) {
            const float* wp = sW2 + cprow0;
            const float* xp = sH + cn0;
            float4 w = *reinterpret_cast<const float4*>(wp), x = *reinterpret_cast<const float4*>(xp);
#pragma unroll (HC > 0 ? 10 : 4)
            for (int k = 0; k < H; ++k) {
                const int kn = (k + 1 < H) ? k + 1 : k;
                const float4 wn = *reinterpret_cast<const float4*>(wp + kn * RPAD);
                const float4 xn = *reinterpret_cast<const float4*>(xp + kn * NP);
                const float wv[4] = {w.x, w.y, w.z, w.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i * 4 + j] = rn_fmaf(wv[i], xv[j], acc[i * 4 + j]);
                w = wn; x = xn;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float wt = sW2t[cprow0 + i], bb = sb2[cprow0 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float v = acc[i * 4 + j];
                    if (td) v = rn_fmaf(wt, tstage, v);
                    v = v + bb;
                    out[i * 4 + j] = act_apply(P.act2, v);
                }
                if (rec >= 0 && i < cvalid)
                    *reinterpret_cast<float4*>(sKt + (crow0 + i) * NP + cn0) = make_float4(out[i * 4], out[i * 4 + 1], out[i * 4 + 2], out[i * 4 + 3]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) out[e] = 0.f;
        }
        pend_rec = rec;      // its output tile is stored to the tape at the next block-wide barrier (next evaluation or flush_tape)
        mark(10);
    };
    // the last staged output tile must reach the tape before anything reuses sKt / before the kernel ends
    auto flush_tape = [&]() {
        if (pend_rec >= 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid == 0) { bulk_store(P.tapeK + (((size_t)pend_rec * P.Q + q) * D + r0) * NP, sKt, (uint32_t)(R * NP * 4)); bulk_commit(); }
            pend_rec = -1;
        }
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
        if (!accepted) {      // the retried attempt rewrites the same tape records: the pending tile first, then let every bulk store land
            flush_tape();
            if (tid == 0) bulk_wait_all();
        }
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
    flush_tape();
    if (tid == 0) bulk_wait_all();
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
    cluster_sync_all();
}

}  // namespace rnde
