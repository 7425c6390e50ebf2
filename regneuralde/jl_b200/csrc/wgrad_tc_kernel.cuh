// wgrad_tc_kernel.cuh -- the weight-gradient contractions on the 5th-gen tensor cores (tcgen05),
// Float32 accuracy preserved by the 3xTF32 split  a*b ~= ah*bh + ah*bl + al*bh  (north_star item 2:
// "tensor cores only where batch x hidden makes a real dense contraction": here the contraction
// length is nrec*B ~ 1e5).  Same operands and FP64 partial sums as wgrad_kernel.cuh.
//
// Contraction order (round 2).  The regulariser puts +-O(10) cotangents on the stages of one step that cancel to
// O(1e-2), and the tensor core ACCUMULATES WITH TRUNCATION: a running sum that swings to O(10) and
// back picks up a one-sided error at every add, coherent over all columns and steps -- measured 2.3x (error estimate) and
// 5.4x (error + stiffness estimate) the error of a CPU Float32 adjoint, reproduced on the CPU by truncating adds
// (profiles/r2b_grad_err.txt; with the FFMA contraction the same sweep is 4x BETTER than the CPU adjoint).  So the cancelling
// terms must meet INSIDE one MMA, whose 8-term dot product is summed before it touches the accumulator: the K index of
// an MMA K-block is now the 6 records of one step (+ record 0 in the first step's spare slot + a zero) for ONE batch
// column, instead of 8 columns of one record.  A stage = (step, column tile, column quad) = 4 K-blocks; the accumulator
// only ever sees the small per-step sums.
//
// One CTA = one 128 x 128 output tile x one contiguous range of stages (split-K).
// Warp roles (416 threads):
//   warps 0-7  load + split: thread = (operand, row); per stage 6-7 LDG.128 (4 columns of each record of the group; the
//              loads of stages g+1 and g+2 are in flight while stage g is converted), a register transposition to record-major,
//              hi = tf32_rn(x), lo = tf32_rn(x - hi), 16 STS.128 into the UMMA K-major no-swizzle core-matrix layout
//              (8-row groups 1152 bytes apart: the 32 rows of a warp store conflict free), fence.proxy.async,
//              arrive on ready[stage].
//   warp 8     one elected lane issues 4 K-blocks x 3 tcgen05.mma.kind::tf32 (M128 N128 K8) per stage
//              into a TMEM accumulator, tcgen05.commit -> empty[stage]; every `flush` stages the
//              accumulator buffer is committed to the epilogue and the other TMEM buffer is used.
//   warps 9-12 epilogue: tcgen05.ld the 128 x 128 FP32 accumulator (lane = output row); every chunk is stored to its
//              own FP32 slot; the FP64 summation over (split, chunk) happens in the fixed-order reduce kernel.
// Shared memory: 3 stages x (A_hi, A_lo, B_hi, B_lo) x 18 KB = 216 KB.  TMEM: 2 x 128 columns.
#pragma once
#include "common.cuh"
#include "fwd4_kernel.cuh"      // mbarrier wrappers
#include "wgrad_kernel.cuh"     // tile_of, rec_time, cp_async helpers

namespace rnde {

constexpr int WGT_M = 128, WGT_N = 128;
constexpr int WGT_KC = 32;                 // contraction entries per stage (4 UMMA K-blocks of 8)
constexpr int WGT_STAGES = 3;
constexpr int WGT_FLUSH_MIN = 16;          // stages per FP32 accumulation chunk (>= 512 entries)
constexpr int WGT_MAXCHUNK = 16;           // chunks per split (bounds the partial workspace)
constexpr int WGT_LOADERS = 256;
constexpr int WGT_THREADS = 32 * (8 + 1 + 4);
constexpr int WGT_SBO = 1152;                                     // stride of 8-row groups: 1024 + 128 (conflict-free row-parallel stores)
constexpr int WGT_PART_BYTES = 16 * WGT_SBO;                      // one operand part of a stage: 128 rows x 32 entries, 18 KB
constexpr int WGT_STAGE_BYTES = 4 * WGT_PART_BYTES;               // A_hi, A_lo, B_hi, B_lo
constexpr size_t WGT_SMEM = (size_t)WGT_STAGES * WGT_STAGE_BYTES + 256;
constexpr int WGT_SPLITS = 21;             // 7 tiles x 21 splits = 147 CTAs: one wave on 148 SMs

// ---- tcgen05 wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, SWIZZLE_NONE canonical layout: 8-row x 16-byte core matrices; K-adjacent core matrices
// LBO bytes apart, 8-row groups SBO bytes apart (cute/arch/mma_sm100_desc.hpp SmemDescriptor).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;        // descriptor version 1 (Blackwell)
    return d;                       // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }

// A: [tiles][M][16], Bm: [tiles][Nrows][16] (physical tile = rec*Q + q); partial: [split][chunk][Naug][M] floats.
// Stage sg = ((s*Q + q) << 2) | cq : step s, column tile q, columns [4cq, 4cq+4) of the tile; K slot j of a column = record
// 1+6s+j (j < 6); slot 6 holds record 0 in the first step's group and, when rec_extra >= 0, record rec_extra in the second
// step's (the evaluation of the initial-dt heuristic, a6.cuh; its time is t + dt of the pseudo-step steps[nsteps]); slot 7 and
// every other slot 6 are zero.
__global__ void __launch_bounds__(WGT_THREADS, 1) wgrad_tc_kernel(const float* __restrict__ A, int M, const float* __restrict__ Bm, int Nrows,
                                                                 int nsteps, int Q, const StepRec* __restrict__ steps, float t0, int td,
                                                                 int flush, float* __restrict__ partial, int rec_extra) {
    constexpr int NP = 16;
    extern __shared__ __align__(1024) unsigned char tsm[];
    const uint32_t sbase = smem_u32(tsm);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + (size_t)WGT_STAGES * WGT_STAGE_BYTES);
    const uint32_t bar0 = sbase + WGT_STAGES * WGT_STAGE_BYTES;
    auto bar_ready = [&](int s) { return bar0 + 8 * s; };
    auto bar_empty = [&](int s) { return bar0 + 8 * (WGT_STAGES + s); };
    auto bar_accfull = [&](int a) { return bar0 + 8 * (2 * WGT_STAGES + a); };
    auto bar_accempty = [&](int a) { return bar0 + 8 * (2 * WGT_STAGES + 2 + a); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WGT_STAGES + 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m_base = blockIdx.x * WGT_M, n_base = blockIdx.y * WGT_N;
    const int Naug = Nrows + 2;
    const int NS = max(nsteps, 1) * Q * 4;                 // stages in all
    const int per = (NS + gridDim.z - 1) / gridDim.z;
    const int stage0 = blockIdx.z * per, stage1 = min(NS, stage0 + per);
    const int nstage = max(stage1 - stage0, 0);
    const int nchunk = (nstage + flush - 1) / flush;

    if (tid == 0) {
        for (int s = 0; s < WGT_STAGES; ++s) { mbar_init(bar_ready(s), WGT_LOADERS); mbar_init(bar_empty(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_accfull(a), 1); mbar_init(bar_accempty(a), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 8) {        // TMEM allocation: 256 columns (two 128-column FP32 accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ================= loaders / splitters: thread = (operand, row) =================
        const int opnd = tid >> 7, row = tid & 127;
        const float* gsrc = opnd ? Bm : A;
        const int nrows_src = opnd ? Nrows : M;                  // rows that exist in the tape operand
        const int grow = (opnd ? n_base : m_base) + row;
        const bool real_row = grow < nrows_src;
        const bool aug_t = opnd && grow == Nrows, aug_1 = opnd && grow == Nrows + 1;
        const uint32_t dst_off = (uint32_t)((opnd ? 2 * WGT_PART_BYTES : 0) + (row >> 3) * WGT_SBO + (row & 7) * 16);
        auto tf32_rn = [](float v) -> float { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return __uint_as_float(r); };
        auto issue = [&](int g, float4 (&buf)[7]) {
            const int sg = stage0 + g;
            const int grp = sg >> 2, cq = sg & 3;
            const int s = grp / Q, q = grp - s * Q;
#pragma unroll
            for (int j = 0; j < 7; ++j) {
                const int rec = (j < 6) ? 1 + 6 * s + j : (s == 0 ? 0 : rec_extra);
                const bool have = (j < 6) ? (s < nsteps) : (s == 0 || (s == 1 && rec_extra >= 0));
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (have) {
                    if (real_row) v = __ldg(reinterpret_cast<const float4*>(gsrc + (((size_t)rec * Q + q) * nrows_src + grow) * NP + cq * 4));
                    else if (aug_1) v = make_float4(1.f, 1.f, 1.f, 1.f);
                }
                buf[j] = v;
            }
            // the time row: fetch the step record (t, dt, ...) like any other operand -- asynchronously; the stage times are
            // formed from it in convert_store (a dependent load here would stall this thread, and with it the stage, once per stage)
            if (aug_t && td) {
                buf[0] = (s < nsteps) ? __ldg(reinterpret_cast<const float4*>(steps + s)) : make_float4(0.f, 0.f, 0.f, 0.f);
                float tx = 0.f;      // time of slot 6: t0 for record 0, t + dt of the pseudo-step for the extra record
                if (s == 0) tx = t0;
                else if (s == 1 && rec_extra >= 0) { const float4 ps = __ldg(reinterpret_cast<const float4*>(steps + nsteps)); tx = ps.x + ps.y; }
                buf[1] = make_float4((s < nsteps) ? 1.f : 0.f, tx, 0.f, 0.f);      // slots 0-5 live; time of slot 6 (0 when empty)
            }
        };
        auto time_row = [&](float4 (&buf)[7]) {      // expand (t, dt) of the step into the 7 slot times (same arithmetic as rec_time)
            const float st = buf[0].x, sdt = buf[0].y;
            const bool live6 = buf[1].x != 0.f;
            const float tz = buf[1].y;
#pragma unroll
            for (int j = 0; j < 6; ++j) { const float tv = live6 ? stage_time(st, sdt, j + 2) : 0.f; buf[j] = make_float4(tv, tv, tv, tv); }
            buf[6] = make_float4(tz, tz, tz, tz);
        };
        // x = hi + lo (+ <= 2^-24 |x|), both rounded to nearest TF32; column c of the quad -> K-block c, slots 0-3 | 4-7
        auto convert_store = [&](int g, const float4 (&buf)[7]) {
            unsigned char* st = tsm + (size_t)(g % WGT_STAGES) * WGT_STAGE_BYTES + dst_off;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v[8], hi[8], lo[8];
#pragma unroll
                for (int j = 0; j < 7; ++j) v[j] = (c == 0) ? buf[j].x : (c == 1) ? buf[j].y : (c == 2) ? buf[j].z : buf[j].w;
                v[7] = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) { hi[j] = tf32_rn(v[j]); lo[j] = tf32_rn(v[j] - hi[j]); }
                *reinterpret_cast<float4*>(st + c * 256) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(st + c * 256 + 128) = make_float4(hi[4], hi[5], hi[6], hi[7]);
                *reinterpret_cast<float4*>(st + WGT_PART_BYTES + c * 256) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(st + WGT_PART_BYTES + c * 256 + 128) = make_float4(lo[4], lo[5], lo[6], lo[7]);
            }
        };
        auto publish = [&](int g, float4 (&buf)[7]) {
            const int b = g % WGT_STAGES;
            if (aug_t && td) time_row(buf);
            if (g >= WGT_STAGES) mbar_wait(bar_empty(b), (uint32_t)((g / WGT_STAGES - 1) & 1));
            convert_store(g, buf);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive_local(bar_ready(b));
        };
        // software pipeline over three register buffers: the loads of stages g+1 and g+2 are in flight while stage g is converted
        float4 bufA[7], bufB[7], bufC[7];
        if (nstage > 0) issue(0, bufA);
        if (nstage > 1) issue(1, bufB);
        for (int g = 0; g < nstage; g += 3) {
            if (g + 2 < nstage) issue(g + 2, bufC);
            publish(g, bufA);
            if (g + 1 < nstage) {
                if (g + 3 < nstage) issue(g + 3, bufA);
                publish(g + 1, bufB);
            }
            if (g + 2 < nstage) {
                if (g + 4 < nstage) issue(g + 4, bufB);
                publish(g + 2, bufC);
            }
        }
    } else if (warp == 8) {
        // ================= MMA issuer =================
        // instruction descriptor: D=F32, A=B=TF32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(WGT_N >> 3) << 17) | ((uint32_t)(WGT_M >> 4) << 24);
        for (int g = 0; g < nstage; ++g) {
            const int b = g % WGT_STAGES;
            const int chunk = g / flush, acc = chunk & 1;
            const bool first_in_chunk = (g % flush) == 0;
            if (first_in_chunk && chunk >= 2) mbar_wait(bar_accempty(acc), (uint32_t)((chunk / 2 - 1) & 1));
            mbar_wait(bar_ready(b), (uint32_t)((g / WGT_STAGES) & 1));
            tc_fence_after();
            if (lane == 0) {
                const uint32_t st = sbase + b * WGT_STAGE_BYTES;
                const uint32_t d_tmem = tmem_base + acc * 128;
#pragma unroll
                for (int kb = 0; kb < WGT_KC / 8; ++kb) {
                    const uint64_t a_hi = umma_desc(st + kb * 256, 128, WGT_SBO);
                    const uint64_t a_lo = umma_desc(st + WGT_PART_BYTES + kb * 256, 128, WGT_SBO);
                    const uint64_t b_hi = umma_desc(st + 2 * WGT_PART_BYTES + kb * 256, 128, WGT_SBO);
                    const uint64_t b_lo = umma_desc(st + 3 * WGT_PART_BYTES + kb * 256, 128, WGT_SBO);
                    // the small products first: the accumulator receives hi*hi last
                    tc_mma_tf32(d_tmem, a_hi, b_lo, idesc, (first_in_chunk && kb == 0) ? 0u : 1u);
                    tc_mma_tf32(d_tmem, a_lo, b_hi, idesc, 1u);
                    tc_mma_tf32(d_tmem, a_hi, b_hi, idesc, 1u);
                }
                tc_commit(bar_empty(b));                                  // stage buffer free when these MMAs retire
                if ((g % flush) == flush - 1 || g == nstage - 1) tc_commit(bar_accfull(acc));
            }
            __syncwarp();
        }
    } else {
        // ================= epilogue: TMEM -> FP32 chunk slots in global memory =================
        const int quad = warp & 3;                    // TMEM lanes [32*quad, 32*quad+32) are this warp's
        const int row = quad * 32 + lane;
        const int m = m_base + row;
        for (int chunk = 0; chunk < WGT_MAXCHUNK; ++chunk) {
            float* prow = partial + ((size_t)blockIdx.z * WGT_MAXCHUNK + chunk) * Naug * M;     // [n][m], m fastest
            if (chunk >= nchunk) break;      // unused slots are neither written nor read (the reduce recomputes every split's chunk count)
            const int acc = chunk & 1;
            mbar_wait(bar_accfull(acc), (uint32_t)((chunk / 2) & 1));
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < WGT_N / 16; ++c) {
                uint32_t v[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * 128 + c * 16);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                    : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (m < M) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int n = n_base + c * 16 + j;
                        if (n < Naug) prow[(size_t)n * M + m] = __uint_as_float(v[j]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive_local(bar_accempty(acc));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
    }
}

// fixed-order FP64 sum over splits of the [split][n][m] partials; scatter into Flux.destructure layout
__global__ void wgrad_tc_reduce_kernel(const float* __restrict__ partial, int nsplit, int NS, int flush, int M, int Nrows, int td,
                                       float* __restrict__ outW, float* __restrict__ outb) {
    const int Naug = Nrows + 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * Naug) return;
    const int n = idx / M, m = idx - n * M;
    const int per = (NS + nsplit - 1) / nsplit;
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) {       // the chunks each split really produced (same arithmetic as wgrad_tc_kernel), in order
        const int stage0 = sp * per, stage1 = min(NS, stage0 + per);
        const int nstage = max(stage1 - stage0, 0);
        const int nchunk = (nstage + flush - 1) / flush;
        // all chunk loads of a split in flight at once; the additions keep their order.  (Two splits at once -- 32 values -- spill
        // and triple the time: measured.)
        const float* pp = partial + (((size_t)sp * WGT_MAXCHUNK) * Naug + n) * M + m;
        float v[WGT_MAXCHUNK];
#pragma unroll
        for (int ch = 0; ch < WGT_MAXCHUNK; ++ch) v[ch] = (ch < nchunk) ? __ldcs(pp + (size_t)ch * Naug * M) : 0.f;
#pragma unroll
        for (int ch = 0; ch < WGT_MAXCHUNK; ++ch) if (ch < nchunk) s += (double)v[ch];
    }
    if (n < Nrows) outW[(size_t)M * n + m] = (float)s;
    else if (n == Nrows) { if (td) outW[(size_t)M * Nrows + m] = (float)s; }
    else outb[m] = (float)s;
}

static int launch_wgrad_tc(int D, int H, int td, int nrec, int Q, const float* tapeZ, const float* tapeD2, const float* tapeH,
                           const float* tapeD1, const StepRec* steps, float t0, float* ws, float* dp, cudaStream_t st, int64_t* launches, int rec_extra = -1) {
    cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WGT_SMEM);
    const int nsteps = (nrec - 1) / 6;
    const int NS = (nsteps > 0 ? nsteps : 1) * Q * 4;       // stages: (step, column tile, column quad)
    int nsplit = WGT_SPLITS;
    if (NS < nsplit) nsplit = NS;
    if (nsplit < 1) nsplit = 1;
    const int nstage = (NS + nsplit - 1) / nsplit;
    int flush = (nstage + WGT_MAXCHUNK - 1) / WGT_MAXCHUNK;
    if (flush < WGT_FLUSH_MIN) flush = WGT_FLUSH_MIN;
    float* wsd = ws;
    float* dW1 = dp;
    float* db1 = dW1 + (size_t)H * (D + td);
    float* dW2 = db1 + H;
    float* db2 = dW2 + (size_t)D * (H + td);
    {   // dW1aug = delta1 . [Z; t; 1]^T      (M = H, N = D + 2)
        dim3 grid((H + WGT_M - 1) / WGT_M, (D + 2 + WGT_N - 1) / WGT_N, nsplit);
        wgrad_tc_kernel<<<grid, WGT_THREADS, WGT_SMEM, st>>>(tapeD1, H, tapeZ, D, nsteps, Q, steps, t0, td, flush, wsd, rec_extra);
        const int tot = H * (D + 2);
        wgrad_tc_reduce_kernel<<<(tot + 255) / 256, 256, 0, st>>>(wsd, nsplit, NS, flush, H, D, td, dW1, db1);
    }
    {   // dW2aug = delta2 . [Hact; t; 1]^T   (M = D, N = H + 2)
        dim3 grid((D + WGT_M - 1) / WGT_M, (H + 2 + WGT_N - 1) / WGT_N, nsplit);
        wgrad_tc_kernel<<<grid, WGT_THREADS, WGT_SMEM, st>>>(tapeD2, D, tapeH, H, nsteps, Q, steps, t0, td, flush, wsd, rec_extra);
        const int tot = D * (H + 2);
        wgrad_tc_reduce_kernel<<<(tot + 255) / 256, 256, 0, st>>>(wsd, nsplit, NS, flush, D, H, td, dW2, db2);
    }
    if (launches) *launches += 4;
    return (int)cudaGetLastError();
}

}  // namespace rnde
