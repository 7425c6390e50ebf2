// microbench3.cu -- how many warps per scheduler does the FP32 pipe of one B200 SM need?
//   (1) pure FFMA, 16 independent accumulators per thread, 1 CTA per SM with 4/8/16/32 warps
//   (2) the stepper's inner-loop shape: per k-step 2 x LDS.128 (4 weights broadcast over 4 lanes, 4 inputs broadcast
//       over 8 lanes) + 16 FFMA on a 4x4 register tile, 8 or 16 warps
// Output: lane-FMA per clock per SM (peak 128).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench3 tools/microbench3.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void ffma_ilp(float* out, long long* cyc, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], a, b);
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// W: [K][128] floats, X: [K][16] floats in shared memory; thread tile 4 rows x 4 cols; rows by lane/4 + 8*warp...
template <int UNROLL>
__global__ void gemm_shape(float* out, long long* cyc, int K, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm;               // K x 128
    float* sX = sm + K * 128;     // K x 16
    for (int e = threadIdx.x; e < K * 144; e += blockDim.x) sm[e] = (e % 7) * 0.01f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = ((warp & 3) * 8 + lane / 4) * 4, n0 = (lane & 3) * 4;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll UNROLL
        for (int k = 0; k < K; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(sW + k * 128 + m0);
            const float4 x = *reinterpret_cast<const float4*>(sX + k * 16 + n0);
            const float wv[4] = {w.x, w.y, w.z, w.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i * 4 + j] = __fmaf_rn(wv[i], xv[j], acc[i * 4 + j]);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// 8x4 register tile: 3 LDS.128 (8 weights + 4 inputs) + 32 FFMA per k-step
template <int UNROLL>
__global__ void gemm_shape84(float* out, long long* cyc, int K, int reps) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm;               // K x 128
    float* sX = sm + K * 128;     // K x 16
    for (int e = threadIdx.x; e < K * 144; e += blockDim.x) sm[e] = (e % 7) * 0.01f;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = ((warp & 1) * 8 + lane / 4) * 8, n0 = (lane & 3) * 4;
    float acc[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = 0.f;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll UNROLL
        for (int k = 0; k < K; ++k) {
            const float4 w0 = *reinterpret_cast<const float4*>(sW + k * 128 + m0);
            const float4 w1 = *reinterpret_cast<const float4*>(sW + k * 128 + m0 + 4);
            const float4 x = *reinterpret_cast<const float4*>(sX + k * 16 + n0);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i * 4 + j] = __fmaf_rn(wv[i], xv[j], acc[i * 4 + j]);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, sizeof(float) * 148 * 1024);
    cudaMalloc(&cyc, sizeof(long long) * 148);
    long long h[148];
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 20000;
        ffma_ilp<<<148, warps * 32>>>(out, cyc, iters, 1.0001f, 0.5f);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("pure FFMA ILP16  warps/SM %2d : %.1f lane-FMA/clk/SM\n", warps, (double)warps * 32 * 16 * iters / (double)h[0]);
    }
    const int K = 100;
    const size_t smem = sizeof(float) * K * 144;
    cudaFuncSetAttribute(gemm_shape<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(gemm_shape<10>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int warps : {4, 8, 16}) {
        const int reps = 200;
        gemm_shape<4><<<148, warps * 32, smem>>>(out, cyc, K, reps);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("4x4 tile, 2 LDS.128 + 16 FFMA per k, unroll 4   warps/SM %2d : %.1f lane-FMA/clk/SM\n", warps, (double)warps * 32 * 16 * K * reps / (double)h[0]);
        gemm_shape<10><<<148, warps * 32, smem>>>(out, cyc, K, reps);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("4x4 tile, 2 LDS.128 + 16 FFMA per k, unroll 10  warps/SM %2d : %.1f lane-FMA/clk/SM\n", warps, (double)warps * 32 * 16 * K * reps / (double)h[0]);
    }
    cudaFuncSetAttribute(gemm_shape84<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int warps : {4, 8}) {
        const int reps = 200;
        gemm_shape84<4><<<148, warps * 32, smem>>>(out, cyc, K, reps);
        cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        printf("8x4 tile, 3 LDS.128 + 32 FFMA per k, unroll 4   warps/SM %2d : %.1f lane-FMA/clk/SM\n", warps, (double)warps * 32 * 32 * K * reps / (double)h[0]);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
