// microbench2.cu -- how to move a few KB between the CTAs of a cluster on B200 (sm_100a).
// Variants (cluster of CS CTAs, each CTA sends `bytes` to each of `fan` peers per round):
//   0  st.shared::cluster.v4 + barrier.cluster (release/acquire)           [baseline, microbench.cu]
//   1  ld.shared::cluster.v4 pull after barrier.cluster
//   2  st.async.v4 ... mbarrier::complete_tx  (receiver waits on its local mbarrier, no cluster barrier)
//   3  cp.async.bulk.shared::cluster.shared::cta ... mbarrier::complete_tx (bulk DMA smem->peer smem)
//   4  global memory round trip: st.global + barrier.cluster + ld.global.cg
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench2 tools/microbench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_relaxed() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

template <int MODE>
__global__ void xchg_kernel(long long* out, int iters, int floats, int fan, float* gbuf, float* sink) {
    extern __shared__ __align__(128) float sm[];
    // layout: [recv: CS * floats][send: floats][mbar (8B)]
    const uint32_t rank = cluster_ctarank(), cs = cluster_nctarank();
    float* recv = sm;
    float* send = sm + cs * floats;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(send + floats);
    const uint32_t recv_a = (uint32_t)__cvta_generic_to_shared(recv);
    const uint32_t send_a = (uint32_t)__cvta_generic_to_shared(send);
    const uint32_t mbar_a = (uint32_t)__cvta_generic_to_shared(mbar);
    const int tid = threadIdx.x;
    for (int e = tid; e < floats; e += blockDim.x) send[e] = (float)(e + rank);
    if (tid == 0) { mbar_init(mbar_a, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    cluster_sync_all();
    const int cluster_id = blockIdx.x / cs;
    float acc = 0.f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if constexpr (MODE == 0) {
            for (int d = 1; d <= fan; ++d) {
                const uint32_t peer = (rank + d) % cs;
                const uint32_t pa = mapa_u32(recv_a + rank * floats * 4, peer);
                for (int e = tid * 4; e < floats; e += blockDim.x * 4) {
                    const float4 v = *reinterpret_cast<const float4*>(send + e);
                    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(pa + e * 4), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
                }
            }
            cluster_sync_all();
            acc += recv[((rank + 1) % cs) * floats + (tid % floats)];
        } else if constexpr (MODE == 1) {
            cluster_sync_all();     // peers' send buffers are ready
            for (int d = 1; d <= fan; ++d) {
                const uint32_t peer = (rank + d) % cs;
                const uint32_t pa = mapa_u32(send_a, peer);
                for (int e = tid * 4; e < floats; e += blockDim.x * 4) {
                    float4 v;
                    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(pa + e * 4) : "memory");
                    acc += v.x + v.y + v.z + v.w;
                }
            }
            cluster_sync_relaxed();  // everyone done reading before send buffers change
        } else if constexpr (MODE == 2) {
            if (tid == 0) mbar_expect_tx(mbar_a, (uint32_t)(fan * floats * 4));
            for (int d = 1; d <= fan; ++d) {
                const uint32_t peer = (rank + d) % cs;
                const uint32_t pa = mapa_u32(recv_a + rank * floats * 4, peer);
                const uint32_t pm = mapa_u32(mbar_a, peer);
                for (int e = tid * 4; e < floats; e += blockDim.x * 4) {
                    const float4 v = *reinterpret_cast<const float4*>(send + e);
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(pa + e * 4),
                                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(pm) : "memory");
                }
            }
            mbar_wait(mbar_a, it & 1);
            acc += recv[((rank + cs - 1) % cs) * floats + (tid % floats)];
            cluster_sync_relaxed();  // WAR protection for the benchmark loop (a real pipeline double-buffers instead)
        } else if constexpr (MODE == 3) {
            if (tid == 0) {
                mbar_expect_tx(mbar_a, (uint32_t)(fan * floats * 4));
                for (int d = 1; d <= fan; ++d) {
                    const uint32_t peer = (rank + d) % cs;
                    const uint32_t pa = mapa_u32(recv_a + rank * floats * 4, peer);
                    const uint32_t pm = mapa_u32(mbar_a, peer);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(pa), "r"(send_a),
                                 "r"((uint32_t)(floats * 4)), "r"(pm) : "memory");
                }
            }
            mbar_wait(mbar_a, it & 1);
            acc += recv[((rank + cs - 1) % cs) * floats + (tid % floats)];
            cluster_sync_relaxed();
        } else if constexpr (MODE == 4) {
            float* mine = gbuf + ((size_t)(cluster_id * cs + rank)) * floats;
            for (int e = tid * 4; e < floats; e += blockDim.x * 4) *reinterpret_cast<float4*>(mine + e) = *reinterpret_cast<const float4*>(send + e);
            cluster_sync_all();
            for (int d = 1; d <= fan; ++d) {
                const uint32_t peer = (rank + d) % cs;
                const float* src = gbuf + ((size_t)(cluster_id * cs + peer)) * floats;
                for (int e = tid * 4; e < floats; e += blockDim.x * 4) {
                    const float4 v = __ldcg(reinterpret_cast<const float4*>(src + e));
                    acc += v.x + v.y + v.z + v.w;
                }
            }
            cluster_sync_relaxed();
        }
    }
    long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = (t1 - t0) / iters;
    if (acc == 123.456f) sink[0] = acc;
    cluster_sync_all();
}

template <int MODE>
int run(const char* name, int cs, int floats, int fan, long long* lout, float* gbuf, float* sink) {
    const int nclusters = (cs == 8) ? 15 : 32;
    int smem = (cs * floats + floats) * 4 + 64;
    CK(cudaFuncSetAttribute(xchg_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaLaunchConfig_t lc{}; lc.gridDim = dim3(nclusters * cs); lc.blockDim = dim3(256); lc.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = 1;
    CK(cudaLaunchKernelEx(&lc, xchg_kernel<MODE>, lout, 200, floats, fan, gbuf, sink));
    CK(cudaDeviceSynchronize());
    long long h[256];
    CK(cudaMemcpy(h, lout, sizeof(long long) * nclusters * cs, cudaMemcpyDeviceToHost));
    long long mx = 0; for (int i = 0; i < nclusters * cs; ++i) if (h[i] > mx) mx = h[i];
    printf("%-28s cs=%d fan=%d bytes/peer=%5d -> %6lld cycles/round (max over CTAs %lld)  %.1f B/cycle/CTA\n", name, cs, fan, floats * 4, h[0], mx,
           (double)fan * floats * 4 / (double)h[0]);
    return 0;
}

int main() {
    long long* lout; CK(cudaMalloc(&lout, sizeof(long long) * 1024));
    float* gbuf; CK(cudaMalloc(&gbuf, sizeof(float) * 256 * 8192));
    float* sink; CK(cudaMalloc(&sink, 16));
    for (int cs : {4, 8}) {
        for (int fan : {1, cs - 1}) {
            for (int floats : {400, 1600, 3200}) {
                if (run<0>("st.shared::cluster + bar", cs, floats, fan, lout, gbuf, sink)) return 1;
                if (run<1>("bar + ld.shared::cluster", cs, floats, fan, lout, gbuf, sink)) return 1;
                if (run<2>("st.async + mbarrier", cs, floats, fan, lout, gbuf, sink)) return 1;
                if (run<3>("cp.async.bulk smem->smem", cs, floats, fan, lout, gbuf, sink)) return 1;
                if (run<4>("global st + bar + ld.cg", cs, floats, fan, lout, gbuf, sink)) return 1;
            }
        }
    }
    printf("done\n");
    return 0;
}
