// head_kernel.cuh -- classifier head of ClassifierNODE and its loss, forward + gradient.
//   logits = W3*u + b3               /root/reference/src/models/supervised_classification.jl:44-45
//   loss   = mean_j logitcrossentropy(logits[:,j], y[:,j])   experiments/mnist_node.jl:135
// Outputs du = dloss/du (cotangent fed to rnde_backward), dp3 = [dW3; db3], the loss.
#pragma once
#include "common.cuh"

namespace rnde {

constexpr int HW_ROWS = 32, HW_SLICES = 8, HW_MAXC = 32;

// one warp per batch column
__global__ void __launch_bounds__(256) head_col_kernel(int D, int B, int C, const float* __restrict__ u, const float* __restrict__ p3,
                                                      const float* __restrict__ y, float scale, float* __restrict__ logits_out,
                                                      float* __restrict__ du, float* __restrict__ g_ws, float* __restrict__ loss_ws) {
    const int lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= B) return;
    const float* W3 = p3;
    const float* b3 = p3 + (size_t)C * D;
    const float* uj = u + (size_t)D * j;
    float logit = 0.f;      // lane c holds logits[c] (C <= 32)
    for (int c = 0; c < C; ++c) {
        float s = 0.f;
        for (int d = lane; d < D; d += 32) s = fmaf(__ldg(W3 + (size_t)C * d + c), __ldg(uj + d), s);
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (lane == c) logit = s + __ldg(b3 + c);
    }
    float mx = lane < C ? logit : -INFINITY;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float ex = lane < C ? expf(logit - mx) : 0.f;
    float se = ex;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) se += __shfl_xor_sync(0xffffffffu, se, off);
    const float lse = mx + logf(se);
    const float yv = lane < C ? __ldg(y + (size_t)C * j + lane) : 0.f;
    float lj = lane < C ? -yv * (logit - lse) : 0.f;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) lj += __shfl_xor_sync(0xffffffffu, lj, off);
    const float g = lane < C ? (ex / se - yv) * (scale / (float)B) : 0.f;   // d(mean CE)/dlogit
    if (lane < C) {
        g_ws[(size_t)C * j + lane] = g;
        if (logits_out) logits_out[(size_t)C * j + lane] = logit;
    }
    if (lane == 0) loss_ws[j] = lj;
    // du[:, j] = W3^T g
    // warp-uniform trip count: every lane takes part in the shuffles of the tail iteration too
    for (int d0 = 0; d0 < D; d0 += 32) {
        const int d = d0 + lane;
        float s = 0.f;
        for (int c = 0; c < C; ++c) {
            const float gc = __shfl_sync(0xffffffffu, g, c);
            if (d < D) s = fmaf(__ldg(W3 + (size_t)C * d + c), gc, s);
        }
        if (d < D) du[(size_t)D * j + d] = s;
    }
}

// dW3[c][d] = sum_j g[c][j] u[d][j]; db3[c] = sum_j g[c][j]; loss = scale * mean_j loss_j
// A block owns 32 consecutive state rows d (one coalesced 128-byte segment of u per batch column) and splits the batch
// over its 8 warps; lane = row, every lane keeps C accumulators; g is read as warp-uniform broadcasts.  The 8 partial sums
// are combined in a fixed order (deterministic).  The last block also reduces db3 and the loss.
__global__ void __launch_bounds__(HW_ROWS * HW_SLICES) head_wgrad_kernel(int D, int B, int C, const float* __restrict__ u, const float* __restrict__ g_ws,
                                                                    const float* __restrict__ loss_ws, float scale, float* __restrict__ dp3,
                                                                    float* __restrict__ loss_out) {
    __shared__ float part[HW_SLICES][HW_MAXC][HW_ROWS + 1];
    const int lane = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int nrb = (D + HW_ROWS - 1) / HW_ROWS;
    if ((int)blockIdx.x < nrb) {
        const int d = blockIdx.x * HW_ROWS + lane;
        float acc[HW_MAXC];
#pragma unroll
        for (int c = 0; c < HW_MAXC; ++c) acc[c] = 0.f;
        const int per = (B + HW_SLICES - 1) / HW_SLICES;
        const int j0 = sl * per, j1 = min(B, j0 + per);
        for (int jb = j0; jb < j1; jb += 8) {      // 8 columns of u in flight (one dependent load per column made this kernel 50 us)
            float uv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) uv[k] = (d < D && jb + k < j1) ? __ldg(u + (size_t)D * (jb + k) + d) : 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (jb + k < j1) {
                    const float* gj = g_ws + (size_t)C * (jb + k);
#pragma unroll
                    for (int c = 0; c < HW_MAXC; ++c) if (c < C) acc[c] = fmaf(__ldg(gj + c), uv[k], acc[c]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < HW_MAXC; ++c) if (c < C) part[sl][c][lane] = acc[c];
        __syncthreads();
        for (int e = threadIdx.x; e < C * HW_ROWS; e += blockDim.x) {
            const int c = e / HW_ROWS, l = e - c * HW_ROWS;
            float s = part[0][c][l];
#pragma unroll
            for (int k = 1; k < HW_SLICES; ++k) s += part[k][c][l];
            const int dd = blockIdx.x * HW_ROWS + l;
            if (dd < D) dp3[(size_t)C * dd + c] = s;
        }
    } else {        // the extra block: bias gradient and the loss
        for (int c = sl; c < C; c += HW_SLICES) {
            float s = 0.f;
            for (int j = lane; j < B; j += 32) s += __ldg(g_ws + (size_t)C * j + c);
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) dp3[(size_t)C * D + c] = s;
        }
        if (sl == 0) {
            float s = 0.f;
            for (int j = lane; j < B; j += 32) s += loss_ws[j];
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) loss_out[0] = scale * s / (float)B;
        }
    }
}

static int launch_head(int D, int B, int C, const float* u, const float* p3, const float* y, float scale, float* loss, float* logits,
                       float* du, float* dp3, float* ws, cudaStream_t st, int64_t* launches) {
    float* g_ws = ws;
    float* loss_ws = ws + (size_t)C * B;
    const int warps_per_block = 8;
    head_col_kernel<<<(B + warps_per_block - 1) / warps_per_block, 256, 0, st>>>(D, B, C, u, p3, y, scale, logits, du, g_ws, loss_ws);
    head_wgrad_kernel<<<(D + HW_ROWS - 1) / HW_ROWS + 1, HW_ROWS * HW_SLICES, 0, st>>>(D, B, C, u, g_ws, loss_ws, scale, dp3, loss);
    if (launches) *launches += 2;
    return (int)cudaGetLastError();
}

// lambda * agg(sv.saveval) and its cotangents, with the number of saved values read on the device (no host round trip
// between the forward solve and the backward sweep).  agg: 0 mean (mnist_node.jl:69,98), 1 maximum (:80), 2 sum
// (test/test_node.jl:55).  One block.
__global__ void reg_agg_kernel(const DevStats* __restrict__ stats, const int agg, const float lam, const float cot_scale,
                               const float* __restrict__ sv, float* __restrict__ dsv, const int cap, float* __restrict__ reg_out) {
    __shared__ float sval[32];
    __shared__ int sidx[32];
    const int n = min(stats->n_saved, cap), tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float acc = (agg == 1) ? -INFINITY : 0.f;
    int best = 0;
    for (int i = tid; i < n; i += blockDim.x) {
        const float v = sv[i];
        if (agg == 1) { if (v > acc) { acc = v; best = i; } }
        else acc += v;
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, acc, off);
        const int ob = __shfl_xor_sync(0xffffffffu, best, off);
        if (agg == 1) { if (o > acc || (o == acc && ob < best)) { acc = o; best = ob; } }
        else acc += o;
    }
    if (lane == 0) { sval[warp] = acc; sidx[warp] = best; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            if (agg == 1) { if (sval[w] > acc || (sval[w] == acc && sidx[w] < best)) { acc = sval[w]; best = sidx[w]; } }
            else acc += sval[w];
        }
        sval[0] = acc; sidx[0] = best;
    }
    __syncthreads();
    acc = sval[0]; best = sidx[0];
    const float cot = (agg == 0) ? (n > 0 ? cot_scale * lam / (float)n : 0.f) : cot_scale * lam;
    for (int i = tid; i <= cap; i += blockDim.x) dsv[i] = (i < n && (agg != 1 || i == best)) ? cot : 0.f;
    if (tid == 0) reg_out[0] = (n == 0) ? 0.f : (agg == 0 ? lam * (acc / (float)n) : lam * acc);
}

// ---- gradient all-reduce over NVLink peer memory (reference-exact data-parallel mode) -----------------------------------
// One-shot push all-reduce on the CUDA-IPC mapped exchange buffers (no NCCL call): every rank writes its vector into slot
// [parity][rank] of EVERY rank's buffer, publishes "call seq done" flags with st.release.sys, then each rank adds the slots
// it received in rank order -- the same order on all ranks, so the result is bitwise identical everywhere.  Slots are
// double buffered by call parity: a rank can only reach call k+2 after every rank finished pushing call k+1, i.e. after
// every rank finished reading call k.  Waiting is bounded (~seconds); a missing peer traps instead of hanging the GPU.
struct ArParams {
    unsigned long long peers[8];      // base of every rank's exchange buffer
    unsigned long long goff;          // byte offset of the gradient area inside a buffer: [flags 32 u32][done 32 u32][2][nranks][gcap] floats
    int nranks, rank; unsigned seq; long long n, gcap;
};
__global__ void ar_push_kernel(const ArParams A, const float* __restrict__ src) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    const size_t slot = ((size_t)(A.seq & 1u) * A.nranks + A.rank) * (size_t)A.gcap;
    for (int r = 0; r < A.nranks; ++r) {
        float* dst = reinterpret_cast<float*>(A.peers[r] + A.goff) + 64 + slot;
        for (long long i = tid; i < A.n; i += nth) dst[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned* done = reinterpret_cast<unsigned*>(A.peers[A.rank] + A.goff) + 32;
        const unsigned prev = atomicAdd(done, 1u);
        if (prev == gridDim.x - 1) {           // last block: everything this rank pushed is visible before the flags
            *done = 0u;
            __threadfence_system();
            for (int r = 0; r < A.nranks; ++r) st_release_sys(reinterpret_cast<unsigned*>(A.peers[r] + A.goff) + A.rank, A.seq);
        }
    }
}
__global__ void ar_sum_kernel(const ArParams A, float* __restrict__ dst) {
    if (threadIdx.x == 0) {
        const unsigned* mine = reinterpret_cast<const unsigned*>(A.peers[A.rank] + A.goff);
        for (int r = 0; r < A.nranks; ++r) {
            long long spins = 0;
            // bounded (~10 s): a peer that never arrives is an error, not a hang
            while ((int)(ld_acquire_sys(mine + r) - A.seq) < 0) { if (++spins > (1ll << 24)) __trap(); }
        }
    }
    __syncthreads();
    const float* base = reinterpret_cast<const float*>(A.peers[A.rank] + A.goff) + 64 + (size_t)(A.seq & 1u) * A.nranks * (size_t)A.gcap;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
    for (long long i = tid; i < A.n; i += nth) {
        float s = __ldcg(base + i);
        for (int r = 1; r < A.nranks; ++r) s += __ldcg(base + (size_t)r * A.gcap + i);
        dst[i] = s;
    }
}

__global__ void opt_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ v, long long n, float scale,
                                  float eta, float rho) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float delta = g[i] * scale;
    const float vn = rho * v[i] - eta * delta;
    v[i] = vn;
    p[i] = p[i] + vn;
}

// Flux.Optimise.Optimiser(WeightDecay(wd), ADAM(eta, beta)) (experiments/ffjord_tabular.jl:128; Flux 0.11.6 apply!)
__global__ void adam_update_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
                                   float eta, float b1, float b2, float b1p, float b2p, float eps, float wd) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float pi = p[i];
    const float d = g[i] + wd * pi;
    const float mn = b1 * m[i] + (1.f - b1) * d;
    const float vn = b2 * v[i] + (1.f - b2) * d * d;
    m[i] = mn; v[i] = vn;
    p[i] = pi - mn / (1.f - b1p) / (sqrtf(vn / (1.f - b2p)) + eps) * eta;
}

}  // namespace rnde
