// head_kernel.cuh -- classifier head of ClassifierNODE and its loss, forward + gradient.
//   logits = W3*u + b3               /root/reference/src/models/supervised_classification.jl:44-45
//   loss   = mean_j logitcrossentropy(logits[:,j], y[:,j])   experiments/mnist_node.jl:135
// Outputs du = dloss/du (cotangent fed to rnde_backward), dp3 = [dW3; db3], the loss.
#pragma once
#include "common.cuh"

namespace rnde {

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
    for (int d = lane; d < D; d += 32) {
        float s = 0.f;
        for (int c = 0; c < C; ++c) s = fmaf(__ldg(W3 + (size_t)C * d + c), __shfl_sync(0xffffffffu, g, c), s);
        du[(size_t)D * j + d] = s;
    }
}

// dW3[c][d] = sum_j g[c][j] u[d][j]; db3[c] = sum_j g[c][j]; loss = scale * mean_j loss_j
__global__ void __launch_bounds__(256) head_wgrad_kernel(int D, int B, int C, const float* __restrict__ u, const float* __restrict__ g_ws,
                                                        const float* __restrict__ loss_ws, float scale, float* __restrict__ dp3,
                                                        float* __restrict__ loss_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = C * D + C;
    if (idx < C * D) {
        const int d = idx / C, c = idx - d * C;
        float s = 0.f;
        for (int j = 0; j < B; ++j) s = fmaf(__ldg(g_ws + (size_t)C * j + c), __ldg(u + (size_t)D * j + d), s);
        dp3[idx] = s;
    } else if (idx < total) {
        const int c = idx - C * D;
        float s = 0.f;
        for (int j = 0; j < B; ++j) s += __ldg(g_ws + (size_t)C * j + c);
        dp3[idx] = s;
    }
    if (blockIdx.x == 0 && threadIdx.x < 32) {
        float s = 0.f;
        for (int j = threadIdx.x; j < B; j += 32) s += loss_ws[j];
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        if (threadIdx.x == 0) loss_out[0] = scale * s / (float)B;
    }
}

static int launch_head(int D, int B, int C, const float* u, const float* p3, const float* y, float scale, float* loss, float* logits,
                       float* du, float* dp3, float* ws, cudaStream_t st, int64_t* launches) {
    float* g_ws = ws;
    float* loss_ws = ws + (size_t)C * B;
    const int warps_per_block = 8;
    head_col_kernel<<<(B + warps_per_block - 1) / warps_per_block, 256, 0, st>>>(D, B, C, u, p3, y, scale, logits, du, g_ws, loss_ws);
    const int total = C * D + C;
    head_wgrad_kernel<<<(total + 255) / 256, 256, 0, st>>>(D, B, C, u, g_ws, loss_ws, scale, dp3, loss);
    if (launches) *launches += 2;
    return (int)cudaGetLastError();
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

}  // namespace rnde
