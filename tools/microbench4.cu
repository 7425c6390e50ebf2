// microbench4.cu -- round-2 design numbers for the forward stepper's layer phases (run on the GPU box):
//   (1) FFMA2 (fma.rn.f32x2, two IEEE fmas per instruction) pipe rate against scalar FFMA, 4/7/8 warps per SM
//   (2) the proposed inner-loop shape: 8x8 register tile, contraction index split over 4 (or 8) lanes, per k-step
//       4 x LDS.128 + 32 FFMA2 (scalar-broadcast operand form), with the stepper's real shared-memory layouts
//       (weights [k][pitch 196 | 100], inputs [k][16]) and lane mapping, followed by the xor-shuffle reduce-scatter that
//       leaves every lane with 16 (or 8) finished sums
//   (3) the same tile with scalar FFMA, for comparison
//   (4) grid barrier variants over 128 CTAs
// Output: lane-FMA per clock per SM (FFMA peak 128).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench4 tools/microbench4.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)(v & 0xffffffffu)); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }

__global__ void ffma2_ilp(float* out, long long* cyc, int iters, float a, float b) {
    u64 x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = pk(threadIdx.x * 0.001f + i, 1.f + i);
    const u64 aa = pk(a, a), bb = pk(b, b);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma2(x[i], aa, bb);
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += lo(x[i]) + hi(x[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Layer-2-like phase: out[m][n] = sum_k W[k][m] * X[k][n], m in [0, 8*MG), n in [0,16), K = 4*KS (S = 4 interleaved chunks)
// or K = 8*KS (S = 8).  lane bits: b0 = s&1, b1 = cg, b2 = mg&1, then S=4: b3 = s>>1, b4 = (mg>>1)&1, warp = mg>>2
//                                                                  S=8: b3,b4 = s>>1,          warp = mg>>1
template <int S, bool PACKED, bool SCATTER>
__global__ void tile88(float* out, long long* cyc, int KS, int WP, int MG, int reps) {
    extern __shared__ __align__(16) float sm[];
    const int K = S * KS;
    float* sW = sm;                     // [K][WP]
    float* sX = sm + K * WP + 8;        // [K][16]
    for (int e = threadIdx.x; e < K * WP + 8 + K * 16; e += blockDim.x) sm[e] = ((e * 7) % 13) * 0.01f - 0.05f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = lane & 1, cg = (lane >> 1) & 1;
    int s, mg;
    if (S == 4) { s = b0 | (((lane >> 3) & 1) << 1); mg = ((lane >> 2) & 1) | (((lane >> 4) & 1) << 1) | (warp << 2); }
    else        { s = b0 | (((lane >> 3) & 3) << 1); mg = ((lane >> 2) & 1) | (warp << 1); }
    const bool act = (S == 4 ? (warp << 2) : (warp << 1)) < MG;      // warp-uniform: the shuffles below need every lane of an active warp
    if (mg >= MG) mg = MG - 1;
    const float* wp = sW + s * WP + mg * 8;
    const float* xp = sX + s * 16 + cg * 8;
    float res = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (act) {
            if (PACKED) {
                u64 acc[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0ull;
#pragma unroll 5
                for (int i = 0; i < KS; ++i) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wp + i * S * WP);
                    const float4 w1 = *reinterpret_cast<const float4*>(wp + i * S * WP + 4);
                    const float4 x0 = *reinterpret_cast<const float4*>(xp + i * S * 16);
                    const float4 x1 = *reinterpret_cast<const float4*>(xp + i * S * 16 + 4);
                    const u64 wv[4] = {pk(w0.x, w0.y), pk(w0.z, w0.w), pk(w1.x, w1.y), pk(w1.z, w1.w)};
                    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                    for (int a = 0; a < 4; ++a)
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[a][j] = fma2(wv[a], pk(xv[j], xv[j]), acc[a][j]);
                }
                float a[8][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) { a[2 * i][j] = lo(acc[i][j]); a[2 * i + 1][j] = hi(acc[i][j]); }
                if (SCATTER) {
                    // round 1 (xor 1): rows 0-3 stay with b0 = 0, rows 4-7 with b0 = 1
                    float h1[4][8];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float send = b0 ? a[i][j] : a[i + 4][j];
                            const float keep = b0 ? a[i + 4][j] : a[i][j];
                            h1[i][j] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                        }
                    const int b3 = (lane >> 3) & 1;
                    if (S == 4) {
                        float h2[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float send = b3 ? h1[i][j] : h1[i][j + 4];
                                const float keep = b3 ? h1[i][j + 4] : h1[i][j];
                                h2[i][j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                            }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) res += h2[i][j];
                    } else {
                        float h2[2][8];
#pragma unroll
                        for (int i = 0; i < 2; ++i)
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float send = b3 ? h1[i][j] : h1[i + 2][j];
                                const float keep = b3 ? h1[i + 2][j] : h1[i][j];
                                h2[i][j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                            }
                        const int b4 = (lane >> 4) & 1;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float send = b4 ? h2[0][j] : h2[1][j];
                            const float keep = b4 ? h2[1][j] : h2[0][j];
                            res += keep + __shfl_xor_sync(0xffffffffu, send, 16);
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) res += a[i][j];
                }
            } else {
                float a[8][8];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[i][j] = 0.f;
#pragma unroll 5
                for (int i = 0; i < KS; ++i) {
                    const float4 w0 = *reinterpret_cast<const float4*>(wp + i * S * WP);
                    const float4 w1 = *reinterpret_cast<const float4*>(wp + i * S * WP + 4);
                    const float4 x0 = *reinterpret_cast<const float4*>(xp + i * S * 16);
                    const float4 x1 = *reinterpret_cast<const float4*>(xp + i * S * 16 + 4);
                    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
                    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                    for (int ii = 0; ii < 8; ++ii)
#pragma unroll
                        for (int j = 0; j < 8; ++j) a[ii][j] = __fmaf_rn(wv[ii], xv[j], a[ii][j]);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) res += a[i][j];
            }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- grid barriers over a co-resident grid ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
    unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
// (a) the stepper's barrier: one counter, everyone polls it with ld.acquire
__global__ void gbar_a(unsigned* bar, long long* out, int iters) {
    unsigned gen = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned target = gridDim.x * (gen + 1u);
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            while ((int)(ld_acquire_gpu(bar) - target) < 0) { }
        }
        gen += 1;
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}
// (b) arrive on one counter, poll with relaxed loads and one fence at the end
__global__ void gbar_b(unsigned* bar, long long* out, int iters) {
    unsigned gen = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned target = gridDim.x * (gen + 1u);
            asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
            while ((int)(ld_relaxed_gpu(bar) - target) < 0) { }
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        gen += 1;
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}
// (c) data-carrying all-gather without a counter: every CTA writes a (value, generation) pair into its own slot; every CTA
// polls all slots with 128 threads (one slot per thread) -- the barrier and the exchange of the per-CTA norm partials in one round
__global__ void gbar_c(unsigned long long* slots, long long* out, int iters) {
    __shared__ int ok;
    __syncthreads();
    const long long t0 = clock64();
    float acc = 0.f;
    for (int i = 0; i < iters; ++i) {
        const unsigned gen = i + 1;
        if (threadIdx.x == 0) {
            const unsigned long long v = ((unsigned long long)gen << 32) | __float_as_uint(1.0f + blockIdx.x);
            asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(slots + blockIdx.x), "l"(v) : "memory");
        }
        if (threadIdx.x < gridDim.x) {
            unsigned long long v;
            do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(slots + threadIdx.x) : "memory"); } while ((unsigned)(v >> 32) < gen);
            acc += __uint_as_float((unsigned)v);
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
    if (acc == -1.f) ok = 1;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    long long h[148];
    for (int warps : {4, 7, 8, 16}) {
        const int iters = 20000;
        ffma2_ilp<<<148, warps * 32>>>(out, cyc, iters, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("pure FFMA2 ILP16  warps/SM %2d : %.1f lane-FMA/clk/SM\n", warps, (double)warps * 32 * 32 * iters / (double)h[0]);
        fflush(stdout);
    }
    const int reps = 200;
    // layer 2: K = 100 (S = 4, KS = 25), weights pitch 196, 25 row groups -> 7 warps (224 threads, 200 active)
    // layer 1: K = 200 (S = 8, KS = 25), weights pitch 100, 13 hidden groups -> 7 warps (224 threads, 208 active)
    auto run = [&](const char* name, auto kern, int S, int KS, int WP, int MG, int threads) {
        const size_t smem = sizeof(float) * ((size_t)S * KS * WP + 8 + (size_t)S * KS * 16);
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<148, threads, smem>>>(out, cyc, KS, WP, MG, reps);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        const double fma = (double)MG * 2 * 64 * S * KS * reps;      // useful lane-FMAs per CTA
        printf("%-58s thr %3d : %7.0f cyc/phase, %.1f lane-FMA/clk/SM (%s)\n", name, threads, (double)h[0] / reps, fma / (double)h[0], cudaGetErrorString(e));
        fflush(stdout);
    };
    run("L2 8x8 S=4 FFMA2 + reduce-scatter, 25 row groups", tile88<4, true, true>, 4, 25, 196, 25, 224);
    run("L2 8x8 S=4 FFMA2, no scatter,      25 row groups", tile88<4, true, false>, 4, 25, 196, 25, 224);
    run("L2 8x8 S=4 FFMA  (scalar),         25 row groups", tile88<4, false, false>, 4, 25, 196, 25, 224);
    run("L2 8x8 S=4 FFMA2 + reduce-scatter, 24 row groups", tile88<4, true, true>, 4, 25, 196, 24, 192);
    run("L2 8x8 S=4 FFMA2 + reduce-scatter, 32 row groups (8 full warps)", tile88<4, true, true>, 4, 25, 260, 32, 256);
    run("L2 8x8 S=4 FFMA2 + reduce-scatter, 16 row groups (4 full warps)", tile88<4, true, true>, 4, 25, 196, 16, 128);
    run("L1 8x8 S=8 FFMA2 + reduce-scatter, 13 hidden groups", tile88<8, true, true>, 8, 25, 100, 13, 224);
    run("L1 8x8 S=8 FFMA2, no scatter,      13 hidden groups", tile88<8, true, false>, 8, 25, 100, 13, 224);
    run("L1 8x8 S=8 FFMA  (scalar),         13 hidden groups", tile88<8, false, false>, 8, 25, 100, 13, 224);
    run("L1 8x8 S=8 FFMA2 + reduce-scatter, 16 hidden groups (8 full warps)", tile88<8, true, true>, 8, 25, 132, 16, 256);

    unsigned* bar; cudaMalloc(&bar, 256 * 8);
    for (int grid : {128, 148}) {
        cudaMemset(bar, 0, 256 * 8);
        gbar_a<<<grid, 256>>>(bar, cyc, 2000); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("grid barrier (a) counter + ld.acquire poll   grid %3d : %lld cyc\n", grid, h[0]);
        cudaMemset(bar, 0, 256 * 8);
        gbar_b<<<grid, 256>>>(bar, cyc, 2000); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("grid barrier (b) counter + relaxed poll      grid %3d : %lld cyc\n", grid, h[0]);
        cudaMemset(bar, 0, 256 * 8);
        gbar_c<<<grid, 256>>>((unsigned long long*)bar, cyc, 2000); cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("grid all-gather (c) slot per CTA, 1 round    grid %3d : %lld cyc\n", grid, h[0]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
