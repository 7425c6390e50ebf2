// microbench6.cu -- FFMA2 operand patterns for the 8x8 outer-product step, NO shared-memory loads (operands in registers):
// which accumulate pattern does the register file / operand collector sustain at 2 cycles per FFMA2?
//   P0: acc[a][j] (row pair a, column j)  = fma2(w pair a, x_j broadcast, acc)        j inner   (the kernel's pattern)
//   P1: same, a inner (x_j reused over the 4 row pairs)
//   P2: acc[i][jp] (row i, column pair jp) = fma2(w_i broadcast, x pair jp, acc)       jp inner
//   P3: as P0 but x_j materialised as a register pair (x_j, x_j) first (8 extra MOV-pairs per k-step)
//   P4: scalar FFMA, j inner;  P5: scalar FFMA, i inner
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench6 tools/microbench6.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 pkv(float a, float b) { u64 r; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)(v & 0xffffffffu)); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
template <int P>
__global__ void __launch_bounds__(256, 1) pat(float* out, long long* cyc, int KS, int reps, float seed) {
    float w[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = seed + 0.01f * i + threadIdx.x * 1e-4f; x[i] = seed * 0.5f - 0.02f * i; }
    float res = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        u64 acc[32];
        float a[64];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0ull;
#pragma unroll
        for (int i = 0; i < 64; ++i) a[i] = 0.f;
#pragma unroll 5
        for (int k = 0; k < KS; ++k) {
            if (P == 0) {
#pragma unroll
                for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[aa * 8 + j] = fma2(pk(w[2 * aa], w[2 * aa + 1]), pk(x[j], x[j]), acc[aa * 8 + j]);
            } else if (P == 1) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int aa = 0; aa < 4; ++aa) acc[aa * 8 + j] = fma2(pk(w[2 * aa], w[2 * aa + 1]), pk(x[j], x[j]), acc[aa * 8 + j]);
            } else if (P == 2) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int jp = 0; jp < 4; ++jp) acc[i * 4 + jp] = fma2(pk(w[i], w[i]), pk(x[2 * jp], x[2 * jp + 1]), acc[i * 4 + jp]);
            } else if (P == 3) {
                u64 xx[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) xx[j] = pkv(x[j], x[j]);
#pragma unroll
                for (int aa = 0; aa < 4; ++aa)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[aa * 8 + j] = fma2(pk(w[2 * aa], w[2 * aa + 1]), xx[j], acc[aa * 8 + j]);
            } else if (P == 4) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[i * 8 + j] = __fmaf_rn(w[i], x[j], a[i * 8 + j]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j)
#pragma unroll
                    for (int i = 0; i < 8; ++i) a[i * 8 + j] = __fmaf_rn(w[i], x[j], a[i * 8 + j]);
            }
            // new operands every k-step (keeps the compiler from hoisting anything), no memory traffic
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("" : "+f"(w[i]), "+f"(x[i]));
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) res += lo(acc[i]) + hi(acc[i]);
#pragma unroll
        for (int i = 0; i < 64; ++i) res += a[i];
        __syncthreads();
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int P>
void run(const char* name, float* out, long long* cyc) {
    long long h[148];
    for (int threads : {32, 128, 256}) {
        pat<P><<<148, threads>>>(out, cyc, 25, 200, 0.3f);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-58s warps %d : %7.0f cyc/phase = %.2f cyc per 2 lane-FMA-instr (%s)\n", name, threads / 32, (double)h[0] / 200, (double)h[0] / 200 / 800 / (threads > 128 ? 2 : 1), cudaGetErrorString(e));
        fflush(stdout);
    }
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 1024); cudaMalloc(&cyc, sizeof(long long) * 148);
    run<0>("P0 fma2(w pair, x_j bcast), j inner (kernel)", out, cyc);
    run<1>("P1 fma2(w pair, x_j bcast), a inner", out, cyc);
    run<2>("P2 fma2(w_i bcast, x pair), jp inner", out, cyc);
    run<3>("P3 fma2(w pair, (x_j,x_j) pair), j inner", out, cyc);
    run<4>("P4 scalar FFMA, j inner (1600 instr)", out, cyc);
    run<5>("P5 scalar FFMA, i inner (1600 instr)", out, cyc);
    return 0;
}
