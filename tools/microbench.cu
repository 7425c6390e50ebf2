// microbench.cu -- B200 numbers the stepper design depends on (run on the GPU box):
//   1. FP32 FFMA peak (the roofline denominator for the stepper; not in MEASURED_PEAKS.json)
//   2. cluster barrier latency (cluster size 8), 3. DSMEM store bandwidth, 4. grid barrier latency,
//   5. max co-resident clusters of 8 CTAs x 226 KB shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__global__ void ffma_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = __fmaf_rn(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }

__global__ void cluster_barrier_kernel(long long* out, int iters) {
    cluster_sync_all();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) cluster_sync_all();
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}

// every CTA pushes `floats` floats to each of `fan` peers with 16B stores, then a cluster barrier
__global__ void dsmem_kernel(long long* out, int iters, int floats, int fan) {
    extern __shared__ __align__(16) float sm[];
    const uint32_t rank = cluster_ctarank();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    cluster_sync_all();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int d = 1; d <= fan; ++d) {
            const uint32_t peer = (rank + d) & 7;
            const uint32_t pa = mapa_u32(base + rank * floats * 4, peer);
            for (int e = threadIdx.x * 4; e < floats; e += blockDim.x * 4) {
                asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pa + e * 4), "f"(1.f), "f"(2.f), "f"(3.f), "f"((float)it) : "memory");
            }
        }
        cluster_sync_all();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__global__ void grid_barrier_kernel(unsigned* bar, long long* out, int iters) {
    unsigned gen = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            unsigned prev = atomicAdd(&bar[0], 1u);
            if (prev == gridDim.x - 1) { atomicExch(&bar[0], 0u); __threadfence(); atomicAdd(&bar[1], 1u); }
            else { while (ld_acquire_gpu(&bar[1]) == gen) { } }
            __threadfence();
        }
        gen++;
        __syncthreads();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = (t1 - t0) / iters;
}

__global__ void big_smem_kernel(float* o) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; __syncthreads(); if (o) o[0] = s[0]; }

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s SMs %d smem/block optin %zu clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.sharedMemPerBlockOptin, clk_khz);
    float* out; CK(cudaMalloc(&out, sizeof(float) * 148 * 16 * 256));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    // 1. FFMA peak
    {
        const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 20000;
        ffma_kernel<<<blocks, threads>>>(out, 1000, 1.0001f, 0.5f);
        CK(cudaDeviceSynchronize());
        float best = 1e9;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            ffma_kernel<<<blocks, threads>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double flops = 2.0 * 16 * (double)iters * blocks * threads;
        printf("FFMA_PEAK_TFLOPS %.2f (best of 5, %.3f ms)\n", flops / best / 1e9, best);
        // sustained ~2 s
        cudaEventRecord(e0);
        int n = 0; for (; n < 40; ++n) ffma_kernel<<<blocks, threads>>>(out, iters * 4, 1.0001f, 0.5f);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA_SUSTAINED_TFLOPS %.2f (%.0f ms)\n", flops * 4 * n / ms / 1e9, ms);
    }
    long long* lout; CK(cudaMalloc(&lout, sizeof(long long) * 1024));
    long long h[1024];
    // 5. occupancy of 8-CTA clusters with big smem
    for (int smem : {226 * 1024, 200 * 1024, 160 * 1024, 100 * 1024}) {
        CK(cudaFuncSetAttribute(big_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        for (int cs : {8, 4, 2}) {
            cudaLaunchConfig_t lc{}; lc.gridDim = dim3(cs * 32); lc.blockDim = dim3(256); lc.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            int ncl = 0; cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, big_smem_kernel, &lc);
            printf("MAX_ACTIVE_CLUSTERS smem=%dKB cluster=%d -> %d (%s)\n", smem / 1024, cs, ncl, cudaGetErrorString(e));
        }
    }
    // 2. cluster barrier
    {
        cudaLaunchConfig_t lc{}; lc.gridDim = dim3(128); lc.blockDim = dim3(256);
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        CK(cudaLaunchKernelEx(&lc, cluster_barrier_kernel, lout, 1000));
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, lout, sizeof(long long) * 128, cudaMemcpyDeviceToHost));
        printf("CLUSTER8_BARRIER_CYCLES %lld (256 threads/CTA, 16 clusters)\n", h[0]);
    }
    // 3. DSMEM push bandwidth
    for (int fan : {1, 7}) {
        for (int floats : {416, 1600, 3200}) {
            int smem = 8 * floats * 4;
            CK(cudaFuncSetAttribute(dsmem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            cudaLaunchConfig_t lc{}; lc.gridDim = dim3(128); lc.blockDim = dim3(256); lc.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            lc.attrs = at; lc.numAttrs = 1;
            CK(cudaLaunchKernelEx(&lc, dsmem_kernel, lout, 200, floats, fan));
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, lout, sizeof(long long) * 128, cudaMemcpyDeviceToHost));
            printf("DSMEM_PUSH fan=%d bytes_per_peer=%d -> %lld cycles per round incl. barrier (%.1f B/cycle/CTA sent)\n", fan, floats * 4, h[0],
                   (double)fan * floats * 4 / (double)h[0]);
        }
    }
    // 4. grid barrier
    for (int nb : {16, 128, 148}) {
        unsigned* bar; CK(cudaMalloc(&bar, 16)); CK(cudaMemset(bar, 0, 16));
        grid_barrier_kernel<<<nb, 256>>>(bar, lout, 500);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h, lout, sizeof(long long) * nb, cudaMemcpyDeviceToHost));
        printf("GRID_BARRIER_CYCLES nblocks=%d -> %lld\n", nb, h[0]);
        cudaFree(bar);
    }
    printf("done\n");
    return 0;
}
