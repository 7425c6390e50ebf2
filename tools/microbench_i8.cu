// microbench_i8.cu -- is an integer tcgen05.mma (kind::i8, s8 x s8 -> s32 in TMEM) exactly reproducible on the CPU?
// Groundwork for the exact tensor-core forward stepper: one CTA, A = 128 x K, B = N x K signed 8-bit digits
// in the K-major no-swizzle core-matrix layout (8 rows x 16 bytes), K = 32 per instruction, accumulators read back with
// tcgen05.ld and compared with a 64-bit integer reference.  Prints the number of mismatches for N = 16, 32, 48.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_i8 tools/microbench_i8.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// A: [128][K] int8 row-major in global, B: [N][K]; out: [128][N] int32
template <int N>
__global__ void __launch_bounds__(128, 1) i8_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int K, int32_t* __restrict__ out,
                                                   long long* cyc, int reps) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sbo = (K / 16) * 128;                    // bytes between 8-row groups; K-adjacent core matrices 128 bytes apart
    unsigned char* sA = sm;                            // 16 groups
    unsigned char* sB = sm + 16 * sbo;                 // N/8 groups
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 16 * sbo + (N / 8) * sbo);
    uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 1);
    for (int e = tid; e < 128 * K; e += 128) { const int r = e / K, k = e % K; sA[(r >> 3) * sbo + (k >> 4) * 128 + (r & 7) * 16 + (k & 15)] = (unsigned char)A[e]; }
    for (int e = tid; e < N * K; e += 128) { const int r = e / K, k = e % K; sB[(r >> 3) * sbo + (k >> 4) * 128 + (r & 7) * 16 + (k & 15)] = (unsigned char)B[e]; }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "n"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tslot;
    // instruction descriptor: D = S32 (c_format 2), A = B = signed 8-bit (format 1), K-major, N >> 3 at [17,23), M >> 4 at [24,29)
    const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const long long t0 = clock64();
    uint32_t parity = 0;
    for (int r = 0; r < reps; ++r) {
        if (tid == 0) {
            const uint64_t da = umma_desc(smem_u32(sA), 128, sbo), db = umma_desc(smem_u32(sB), 128, sbo);
            for (int ks = 0; ks < K / 32; ++ks) {
                const uint64_t adv = (uint64_t)(ks * 16);          // 2 core matrices = 256 bytes per k-step
                const uint32_t accf = ks == 0 ? 0u : 1u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(da + adv), "l"(db + adv), "r"(idesc), "r"(accf) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        }
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        } while (!done);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[0] = (t1 - t0) / reps;
    for (int c = 0; c < N / 16; ++c) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(c * 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c * 16 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(64) : "memory");
}

template <int N>
static void run(int K) {
    std::vector<int8_t> A(128 * K), B(N * K);
    srand(7 + N);
    for (auto& v : A) v = (int8_t)(rand() % 256 - 128);
    for (auto& v : B) v = (int8_t)(rand() % 256 - 128);
    int8_t *dA, *dB; int32_t* dO; long long* dC;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dO, 128 * N * 4); cudaMalloc(&dC, 8);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice); cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    const int sbo = (K / 16) * 128;
    const size_t smem = (size_t)(16 + N / 8) * sbo + 64;
    cudaFuncSetAttribute(i8_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    i8_kernel<N><<<1, 128, smem>>>(dA, dB, K, dO, dC, 50);
    cudaError_t e = cudaDeviceSynchronize();
    std::vector<int32_t> O(128 * N); long long cyc = 0;
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost); cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    long long bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            long long s = 0;
            for (int k = 0; k < K; ++k) s += (long long)A[m * K + k] * (long long)B[n * K + k];
            if (s != O[m * N + n]) { if (bad < 3) printf("   mismatch m=%d n=%d ref=%lld got=%d\n", m, n, s, O[m * N + n]); ++bad; }
        }
    printf("kind::i8 M=128 N=%d K=%d (%d MMAs): %lld mismatches of %d, %lld cycles per GEMM [%s]\n", N, K, K / 32, bad, 128 * N, cyc, cudaGetErrorString(e));
    cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
}

int main() {
    run<16>(224); run<32>(224); run<48>(224); run<48>(128);
    return 0;
}
