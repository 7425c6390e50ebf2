// microbench5.cu -- what limits the 8x8 / FFMA2 layer loop of fwd4s_kernel?  Variants of one phase (25 k-steps, layer-2 shape):
//   MODE 0: as in the kernel (4 LDS.128 + 32 FFMA2 per k-step, unroll 5)
//   MODE 1: no shared-memory loads at all (operands stay in registers): the FFMA2 issue rate of this accumulate pattern
//   MODE 2: loads only (results folded into one add): the LDS rate
//   MODE 3: explicit software pipelining, prefetch distance 1, unroll 1
//   MODE 4: explicit software pipelining, prefetch distance 1, unroll 5 (register rotation by the compiler)
//   MODE 5: scalar FFMA instead of FFMA2, unroll 5
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench5 tools/microbench5.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float lo(u64 v) { return __uint_as_float((unsigned)(v & 0xffffffffu)); }
__device__ __forceinline__ float hi(u64 v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ void step(u64 (&acc)[4][8], const float4 w0, const float4 w1, const float4 x0, const float4 x1) {
    const u64 wv[4] = {pk(w0.x, w0.y), pk(w0.z, w0.w), pk(w1.x, w1.y), pk(w1.z, w1.w)};
    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[a][j] = fma2(wv[a], pk(xv[j], xv[j]), acc[a][j]);
}
template <int MODE>
__global__ void __launch_bounds__(256, 1) phase(float* out, long long* cyc, int KS, int WP, int reps) {
    extern __shared__ __align__(16) float sm[];
    const int K = 4 * KS;
    float* sW = sm; float* sX = sm + K * WP + 8;
    for (int e = threadIdx.x; e < K * WP + 8 + K * 16; e += blockDim.x) sm[e] = ((e * 7) % 13) * 0.01f - 0.05f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b0 = lane & 1, cg = (lane >> 1) & 1;
    const int s = b0 | (((lane >> 3) & 1) << 1);
    int mg = ((lane >> 2) & 1) | (((lane >> 4) & 1) << 1) | (warp << 2);
    if (mg > 24) mg = 24;
    const float* wp = sW + s * WP + mg * 8;
    const float* xp = sX + s * 16 + cg * 8;
    float res = 0.f;
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        u64 acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0ull;
        if (MODE == 0) {
#pragma unroll 5
            for (int i = 0; i < KS; ++i) {
                const float4 w0 = *(const float4*)(wp + i * 4 * WP), w1 = *(const float4*)(wp + i * 4 * WP + 4);
                const float4 x0 = *(const float4*)(xp + i * 64), x1 = *(const float4*)(xp + i * 64 + 4);
                step(acc, w0, w1, x0, x1);
            }
        } else if (MODE == 1) {
            float4 w0 = *(const float4*)(wp), w1 = *(const float4*)(wp + 4), x0 = *(const float4*)(xp), x1 = *(const float4*)(xp + 4);
#pragma unroll 5
            for (int i = 0; i < KS; ++i) { step(acc, w0, w1, x0, x1); asm volatile("" : "+f"(w0.x), "+f"(x0.x)); }
        } else if (MODE == 2) {
            float4 a4 = make_float4(0, 0, 0, 0);
#pragma unroll 5
            for (int i = 0; i < KS; ++i) {
                const float4 w0 = *(const float4*)(wp + i * 4 * WP), w1 = *(const float4*)(wp + i * 4 * WP + 4);
                const float4 x0 = *(const float4*)(xp + i * 64), x1 = *(const float4*)(xp + i * 64 + 4);
                a4.x += w0.x + w1.x + x0.x + x1.x; a4.y += w0.y + w1.y + x0.y + x1.y; a4.z += w0.z + w1.z + x0.z + x1.z; a4.w += w0.w + w1.w + x0.w + x1.w;
            }
            acc[0][0] = pk(a4.x + a4.y, a4.z + a4.w);
        } else if (MODE == 3 || MODE == 4) {
            float4 w0 = *(const float4*)(wp), w1 = *(const float4*)(wp + 4), x0 = *(const float4*)(xp), x1 = *(const float4*)(xp + 4);
            if (MODE == 3) {
#pragma unroll 1
                for (int i = 0; i < KS; ++i) {
                    const int in = (i + 1 < KS) ? i + 1 : i;
                    const float4 nw0 = *(const float4*)(wp + in * 4 * WP), nw1 = *(const float4*)(wp + in * 4 * WP + 4);
                    const float4 nx0 = *(const float4*)(xp + in * 64), nx1 = *(const float4*)(xp + in * 64 + 4);
                    step(acc, w0, w1, x0, x1);
                    w0 = nw0; w1 = nw1; x0 = nx0; x1 = nx1;
                }
            } else {
#pragma unroll 5
                for (int i = 0; i < KS; ++i) {
                    const int in = (i + 1 < KS) ? i + 1 : i;
                    const float4 nw0 = *(const float4*)(wp + in * 4 * WP), nw1 = *(const float4*)(wp + in * 4 * WP + 4);
                    const float4 nx0 = *(const float4*)(xp + in * 64), nx1 = *(const float4*)(xp + in * 64 + 4);
                    step(acc, w0, w1, x0, x1);
                    w0 = nw0; w1 = nw1; x0 = nx0; x1 = nx1;
                }
            }
        } else {
            float a[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i][j] = 0.f;
#pragma unroll 5
            for (int i = 0; i < KS; ++i) {
                const float4 w0 = *(const float4*)(wp + i * 4 * WP), w1 = *(const float4*)(wp + i * 4 * WP + 4);
                const float4 x0 = *(const float4*)(xp + i * 64), x1 = *(const float4*)(xp + i * 64 + 4);
                const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                for (int ii = 0; ii < 8; ++ii)
#pragma unroll
                    for (int j = 0; j < 8; ++j) a[ii][j] = __fmaf_rn(wv[ii], xv[j], a[ii][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = pk(a[2 * i][j], a[2 * i + 1][j]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) res += lo(acc[i][j]) + hi(acc[i][j]);
        __syncthreads();
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = res;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, float* out, long long* cyc) {
    const int KS = 25, WP = 196, reps = 200;
    const size_t smem = sizeof(float) * ((size_t)4 * KS * WP + 8 + (size_t)4 * KS * 16);
    cudaFuncSetAttribute(phase<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    long long h[148];
    for (int threads : {32, 128, 224, 256}) {
        phase<MODE><<<148, threads, smem>>>(out, cyc, KS, WP, reps);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%-62s warps %d : %7.0f cyc/phase (%s)\n", name, threads / 32, (double)h[0] / reps, cudaGetErrorString(e));
        fflush(stdout);
    }
}
int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 1024); cudaMalloc(&cyc, sizeof(long long) * 148);
    run<0>("0: kernel loop (4 LDS.128 + 32 FFMA2 / k-step, unroll 5)", out, cyc);
    run<1>("1: no loads (operands in registers)", out, cyc);
    run<2>("2: loads only", out, cyc);
    run<3>("3: explicit prefetch distance 1, unroll 1", out, cyc);
    run<4>("4: explicit prefetch distance 1, unroll 5", out, cyc);
    run<5>("5: scalar FFMA, unroll 5", out, cyc);
    return 0;
}
