// a6.cuh -- the first step size on the tape (detach_dt = all_but_first).
//
// The reference makes tspan tracked (/root/reference/src/utils.jl:21-23), so t and dt are tracked scalars.  Every dt the
// controller PROPOSES is detached (DiffEqBase.value in loopfooter!), but the first one, dt_1 = initial_dt(theta, x) of
// the Hairer-Wanner heuristic (Appendix A.5), is not.  To first order the reference gradient is therefore the frozen-step
// discrete adjoint plus
//     dL/d(dt_1) * d(dt_1)/d(theta, x),
// where dt_1 enters three ways: it is the size of the first accepted step (times the constant factors kappa of rejected
// first attempts), it shifts the start time of every later step, and the last step, cut to land on t1, shrinks by it:
//     dL/d(dt_1) = dtbar_1 - [last step cut] dtbar_N + sum_{s >= 2} tbar_s
// (dtbar_s: explicit dt of step s with its start time fixed, tbar_s: start time with dt fixed).  The reverse sweeps
// accumulate this sum while they run (per-thread `dacc`, per-CTA partials in a6_part, a6_reduce_kernel); then two more
// field VJPs differentiate the heuristic itself (bwd_kernel.cuh, a6_mode 1 and 2):
//     sk = atol + |u0| rtol;  d0 = rms(u0/sk);  d1 = rms(f0/sk);  dt0 = min(0.01 d0/d1, dtmax)
//     u1 = u0 + dt0 f0;  f1 = f(u1, t0 + dt0);  d2 = rms((f1 - f0)/sk)/dt0
//     dt_1 = max(dtmin, min(100 dt0, (100 max(d1, d2))^(-1/5), dtmax))
// The evaluation f1 is taped by the forward kernels (record rec_init) and copied behind the last step (rec_x) so that
// the weight-gradient contraction picks its deltas up as stage 7 of a pseudo-step (t0, dt0); f0 is the same value as
// fsalfirst = record 0, whose k the sweep replaces by delta2 -- a copy is taken first, and the second VJP adds its deltas
// to record 0's.  Checked against oracle/rnde_oracle_bwd.inc (first_dt_tracked) which is checked against autograd through
// the same algorithm (tests/test_oracle_agreement.py).
#pragma once
#include "common.cuh"

namespace rnde {

struct A6Scal {
    float Dt;          // kappa * dL/d(dt_1): cotangent of the heuristic's result
    float d2bar;       // cotangent of d2
    float d1bar;       // cotangent of d1 through max(d1, d2) only
    float dt0bar;      // cotangent of dt0 through min(100 dt0, .), the md <= 1e-15 branch and d2 = rms(.)/dt0
    bool dt0_free;     // dt0 = 0.01 d0/d1 (neither the 1e-6 fallback nor the dtmax clamp)
};

__device__ inline A6Scal a6_scalars(const KParams& P) {
    A6Scal s;
    const float d0 = P.initdt[0], d1 = P.initdt[1], d2 = P.initdt[2], dt0 = P.initdt[3], dt1 = P.initdt[4];
    const float dti = P.stats->dt_init, dtmax = P.t1 - P.t0;
    s.Dt = P.a6_sum[0] * (P.steps[0].dt / dti);
    s.d2bar = 0.f; s.d1bar = 0.f; s.dt0bar = 0.f;
    float dt1b = 0.f;
    if (dti >= dtmax || dti <= P.dtmin) { }                    // clamped: a constant
    else if (100.f * dt0 < dt1) s.dt0bar = 100.f * s.Dt;
    else dt1b = s.Dt;
    if (dt1b != 0.f) {
        const float md = d1 > d2 ? d1 : d2;
        if (md <= (float)1e-15) { if (dt0 * (float)1e-3 > (float)1e-6) s.dt0bar += (float)1e-3 * dt1b; }
        else {                                                 // dt1 = (100 md)^(-1/5)
            const float mdbar = -dt1b * dt1 / (5.f * md);
            if (d1 > d2) s.d1bar = mdbar; else s.d2bar = mdbar;
        }
    }
    s.dt0bar -= s.d2bar * d2 / dt0;
    s.dt0_free = !(d0 < (float)1e-5 || d1 < (float)1e-5) && dt0 < dtmax;
    return s;
}

// sum of `v` over the CTA -> out[blockIdx.x]; scratch: NT/32 doubles of shared memory that nobody else uses any more
// (Float64: the terms are the regulariser's cancelling cotangents times O(1) factors -- the CPU adjoint sums them in double too)
template <int NT>
__device__ __forceinline__ void a6_block_sum(double v, float* scratch_f, double* out) {
    double* scratch = reinterpret_cast<double*>(scratch_f);      // 8-byte aligned: every caller passes a 16-byte aligned region
    __syncthreads();
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < NT / 32; ++w) s += scratch[w];
        out[blockIdx.x] = s;
    }
    __syncthreads();
}

// fixed-order sum of the per-CTA partials (in Float64) -> dst[slot]; one block
__global__ void a6_reduce_kernel(const double* __restrict__ part, int n, float* __restrict__ dst, int slot) {
    __shared__ double sh[256];
    double s = 0;
    for (int i = threadIdx.x; i < n; i += 256) s += part[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int w = 128; w >= 1; w >>= 1) { if ((int)threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w]; __syncthreads(); }
    if (threadIdx.x == 0) dst[slot] = (float)sh[0];
}

}  // namespace rnde
