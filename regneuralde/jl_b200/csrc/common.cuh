// common.cuh -- shared device-side helpers for the regnde sm_100a kernels.
// Arithmetic follows include/regnde_canon.h (canonical order): every fused
// multiply-add is an explicit rn_fmaf; the translation unit is built with
// -fmad=false so nothing else is contracted.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "regnde.h"
#include "regnde_canon.h"

namespace rnde {

struct StepRec { float t, dt, eest, eig, n1, n2, pad0, pad1; };   // n1/n2: rms(k7-k6), rms(u-g6)

struct DevStats {
    int nf, naccept, nreject, n_saved, retcode;
    float t_final, dt_last, dt_init;
};

// Kernel parameters shared by the forward and backward steppers.
struct KParams {
    int D, H, B;        // state dim, hidden dim, local columns
    int R;              // rows owned by one CTA (G*R >= D)
    int kblock;         // canonical K-block of layer 1 and of the norms
    int HS;             // hidden rows reduced by one CTA of the cluster
    int Q;              // number of column tiles (= clusters)
    int act1, act2, td;
    int alg, reg_kind, max_steps, tape_cap, need_tape;
    float t0, t1, abstol, reltol, dtmin;
    long long norm_count;   // D * global batch
    int Bglobal;            // columns entering the norm (== B unless distributed)
    int col_offset;         // first global column of this rank
    const float* x; const float* p; float* u_out; float* saveval;
    float* colsum;          // [2 slots][3][Bpad_global]
    int colsum_stride;      // Bpad_global
    unsigned int* bar;      // grid barrier {count, generation}
    StepRec* steps; DevStats* stats;
    float* tapeZ; float* tapeK; float* tapeH; float* tapeD1;   // [rec][tile][rows][NP]
    // backward only
    const float* du; const float* dsaveval; float* dx;
    float* scal;            // per-step scalar adjoints (dtbar, tbar) [2*tape_cap]
    int nsteps;
    long long* dbg;         // optional phase timeline (clock64 stamps), developer diagnostics only
    // reference-exact data parallel mode: every rank's exchange buffer (IPC-mapped peer memory over NVLink)
    int nranks, rank;
    unsigned long long peers[8];    // base of rank r's exchange buffer: [colsum 2x3xstride floats][flags 64 u32][seq u32]
    unsigned int flag_off;          // offset (in 4-byte words) of the flag array inside an exchange buffer
    // saveat (multi-save functors, neural_ode.jl:79-108,146-180): sorted times, states written as feat x nsave x batch
    const float* saveat; int n_saveat;
    float* usave; const float* dusave;
    // fixed-work replay (rnde_set_forced_steps): dt of attempt i, every attempt accepted
    const float* forced_dt; int n_forced;
    // Appendix A.6 with the first dt on the tape (rnde_set_detach, a6.cuh): dt_1 = initial_dt(theta, x) is differentiated
    int a6;                 // forward: tape the evaluation of the initial-dt heuristic; sweep: accumulate dL/d(dt_1)
    int a6_scalar;          // this rank adds the terms that exist once per solve (saved-value cotangent x explicit dt): rank 0 in the exact mode
    int a6_mode;            // bwd_kernel: 0 = sweep, 1 = VJP of f(u0 + dt0 f0, t0 + dt0), 2 = VJP of f0 added to record 0
    int rec_init;           // tape record the forward writes the initial-dt evaluation to (copied behind the last step for wgrad)
    int rec_x;              // record the adjoint of that evaluation is written to: 6 * nsteps + 6 (stage 7 of a pseudo-step)
    float* initdt;          // [8] d0, d1, d2, dt0, dt1, last attempt cut to land on t1 (0/1)
    double* a6_part;        // [gridDim.x] per-CTA partial sums of the running kernel
    const float* a6_sum;    // [2] reduced over CTAs (and ranks): dL/d(dt_1) of the sweep; <u1bar, f0> + tbar of phase 1
    float* a6_u1bar;        // [tile][row][NP] cotangent of u1 = u0 + dt0 f0 (phase 1 -> phase 2)
    const float* a6_f0;     // [tile][row][NP] copy of record 0's k = f0, taken before the sweep replaces it by delta2
    float* a6_zb;           // tensor-core sweep: cotangents of the stage inputs of the first / last step, [12][tile][row][NP]
    const float* a6_kc;     // tensor-core sweep: copies of k_1..k_6 of the first / last step, [12][tile][row][NP]
    float* a6_tau;          // tensor-core sweep: per record and CTA the two time cotangents, [rec][Q][G][2][16]
    // chain field (chain.cuh): layer widths / activations, pre-activation, tape rows per column, shared-memory offsets (floats)
    int n_layers; int lw[8]; int la[8]; int pre_act; int hrows; int chain_np; int oCW, oCA, oCB, oCH;
    // FFJORD field (csq.cuh): Hutchinson noise ((D - csq_extra) x B, column-major), augmented rows, shared-memory offset of its region
    const float* noise; int csq_extra; int oCS;
    int oCSP;               // > 0: the FFJORD parameters are staged in shared memory at this offset (floats)
    int csq_reverse;        // 1: the field is -f(z, t0 + t1 - t): the flow integrated backwards (rnde_set_reverse_time, `sample`)
};

// Tsit5 free interpolant weights b_1..b_7(theta); same Horner/fma order as the oracle's interp_weights
__device__ __forceinline__ void interp_weights(const float th, float* b) {
    const float th2 = th * th;
    b[1] = th * rn_fmaf(th, rn_fmaf(th, rn_fmaf(th, (float)TS_R14, (float)TS_R13), (float)TS_R12), (float)TS_R11);
    b[2] = th2 * rn_fmaf(th, rn_fmaf(th, (float)TS_R24, (float)TS_R23), (float)TS_R22);
    b[3] = th2 * rn_fmaf(th, rn_fmaf(th, (float)TS_R34, (float)TS_R33), (float)TS_R32);
    b[4] = th2 * rn_fmaf(th, rn_fmaf(th, (float)TS_R44, (float)TS_R43), (float)TS_R42);
    b[5] = th2 * rn_fmaf(th, rn_fmaf(th, (float)TS_R54, (float)TS_R53), (float)TS_R52);
    b[6] = th2 * rn_fmaf(th, rn_fmaf(th, (float)TS_R64, (float)TS_R63), (float)TS_R62);
    b[7] = th2 * rn_fmaf(th, rn_fmaf(th, (float)TS_R74, (float)TS_R73), (float)TS_R72);
}

// ---- small PTX wrappers ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_f4(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

template <int G>
__device__ __forceinline__ void group_sync() {
    if constexpr (G > 1) cluster_sync_all(); else __syncthreads();
}

// Grid barrier over all CTAs of a co-resident persistent grid: one release-add on a monotonically
// increasing arrival counter, then acquire-polling until it reaches nblocks*(generation+1).
// (bar[0] is zeroed by the host before every launch.)
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned nblocks, unsigned& gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned target = nblocks * (gen + 1u);
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
        while ((int)(ld_acquire_gpu(bar) - target) < 0) { }
    }
    gen += 1;
    __syncthreads();
}

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// Cross-GPU barrier of the exact data-parallel mode.  Called after the local grid barrier: block 0 publishes
// "rank r reached norm `seq`" into every peer's flag array (peer memory, st.release.sys makes this rank's column
// sums -- already written into the peers' buffers -- visible first); every CTA then waits for all ranks' flags.
__device__ __forceinline__ void xrank_barrier(const KParams& P, unsigned seq) {
    if (P.nranks <= 1) return;
    if (threadIdx.x == 0) {
        if (blockIdx.x == 0) {
            __threadfence_system();
            for (int r = 0; r < P.nranks; ++r)
                st_release_sys(reinterpret_cast<unsigned*>(P.peers[r]) + P.flag_off + P.rank, seq);
        }
        const unsigned* mine = reinterpret_cast<const unsigned*>(P.peers[P.rank]) + P.flag_off;
        // bounded like ar_sum_kernel's wait: a rank that never launched (host exception, handle mismatch) must not hang
        // the persistent kernels of all the others for ever -- tens of seconds of polling, then the launch fails with a trap
        for (int r = 0; r < P.nranks; ++r) {
            long long spins = 0;
            while ((int)(ld_acquire_sys(mine + r) - seq) < 0) { if (++spins > (1ll << 25)) __trap(); }
        }
    }
    __syncthreads();
}
// Merged barrier of the exact data-parallel mode (round 2): the cluster's publishing CTA has written its column sums into every
// rank's exchange buffer; one release-add per rank on that rank's arrival counter then says so, and every CTA of every rank
// waits until its own counter has seen all Q x nranks clusters of this norm.  One NVLink round instead of a local grid
// barrier + eight serialised st.release.sys flags + a poll of eight flags.  Call with the whole CTA; `publisher` = this CTA
// published column sums (cluster rank 0); `seq` counts norms from 1 across launches (same on every rank).
__device__ __forceinline__ void xrank_arrive_wait(const KParams& P, bool publisher, unsigned seq, unsigned nclusters) {
    constexpr unsigned ARRIVE = 48;      // word inside the flag block
    __syncthreads();                     // the publishing threads' peer stores are ordered before thread 0's release
    if (threadIdx.x == 0) {
        if (publisher) {
            // ONE system-scope fence orders the column sums before the arrivals; the arrivals themselves are relaxed
            // fire-and-forget reductions (a release on each of them would wait for the NVLink round trip once per rank)
            __threadfence_system();
            for (int r = 0; r < P.nranks; ++r)
                asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(reinterpret_cast<unsigned*>(P.peers[r]) + P.flag_off + ARRIVE) : "memory");
        }
        const unsigned target = seq * nclusters * (unsigned)P.nranks;
        const unsigned* mine = reinterpret_cast<const unsigned*>(P.peers[P.rank]) + P.flag_off + ARRIVE;
        long long spins = 0;
        unsigned cur;
        do {
            asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(mine) : "memory");
            if (++spins > (1ll << 26)) __trap();
        } while ((int)(cur - target) < 0);
        __threadfence_system();      // acquire side: the peers' column sums are visible to the fold below
    }
    __syncthreads();
}
// write one per-column sum into every rank's exchange buffer (own buffer included)
__device__ __forceinline__ void publish_colsum(const KParams& P, size_t word_off, float v) {
    if (P.nranks <= 1) { P.colsum[word_off] = v; return; }
    for (int r = 0; r < P.nranks; ++r) reinterpret_cast<float*>(P.peers[r])[word_off] = v;
}

__device__ __forceinline__ float act_apply(int act, float s) { return act == RNDE_ACT_TANH ? canon_tanhf(s) : s; }

// 4 consecutive floats from global memory; vector load when 16B aligned and in range.
__device__ __forceinline__ float4 ldg4(const float* __restrict__ ptr, int remaining) {
    if (remaining >= 4 && ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0)) return __ldg(reinterpret_cast<const float4*>(ptr));
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (remaining > 0) v.x = __ldg(ptr);
    if (remaining > 1) v.y = __ldg(ptr + 1);
    if (remaining > 2) v.z = __ldg(ptr + 2);
    if (remaining > 3) v.w = __ldg(ptr + 3);
    return v;
}

__host__ __device__ inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Tsit5 coefficients in Float32 (each converted once from its Float64 literal on the host,
// exactly like the oracle's (REAL)TS_xx casts); filled by init_constants() in regnde.cu.
__constant__ float c_A[8][8];
__constant__ float c_BT[8];
__constant__ float c_C[8];
__device__ __forceinline__ float ts_a(int i, int j) { return c_A[i][j]; }
__device__ __forceinline__ float ts_bt(int i) { return c_BT[i]; }
__device__ __forceinline__ float ts_c(int i) { return c_C[i]; }
__device__ __forceinline__ float stage_time(float t, float dt, int i) {
    if (i >= 6) return t + dt;
    return rn_fmaf(ts_c(i), dt, t);
}

}  // namespace rnde
