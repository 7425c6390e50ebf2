/*
 * regnde.h -- C ABI of libregnde.so: the B200-native replacement for the body of
 * the reference's neural-ODE layer call and everything below it.
 *
 * The reference has no FFI; its hot path sits behind Julia callable structs.
 * Each entry point below names the reference interface it replaces:
 *
 *   rnde_create        TrackedNeuralODE(model, tspan, time_dep, regularize, solver...; kwargs...)
 *                      src/models/neural_ode.jl:10-32   (constructor: shapes, solver args, kwargs)
 *   rnde_forward       (n::TrackedNeuralODE{R,false})(x, p; func, tspan)  ->  (res, nfe, sv)
 *                      src/models/neural_ode.jl:48-77 (R=false), :110-144 (R=true);
 *                      solve(prob, Tsit5()|AutoTsit5(Tsit5()); callback=SavingCallback(func, sv), ...) :131-137
 *   rnde_backward      Tracker.gradient(...) through that call
 *                      experiments/mnist_node.jl:229-232, test/test_node.jl:25,47-57,79-89
 *   rnde_head_*        ClassifierNODE post-net Dense(784,10) + logitcrossentropy
 *                      src/models/supervised_classification.jl:44-45, experiments/mnist_node.jl:135
 *   rnde_stats         sol.destats.nf (neural_ode.jl:142) and length(sv.saveval)
 *
 * Conventions
 *   - All arrays are Float32, COLUMN-MAJOR features x batch (a Julia CuArray /
 *     Array passes its pointer unchanged); `p` is the Flux.destructure vector
 *     of the 2-layer time-concatenated field: W1 (H x (D+1)), b1 (H),
 *     W2 (D x (H+1)), b2 (D)   [src/models/basic.jl:16-28, experiments/mnist_node.jl:41-54].
 *   - *_dev pointers are device pointers owned by the caller; `stream` is a
 *     cudaStream_t passed as void*.  The *_host entry points take host pointers
 *     and do the H2D/D2H copies themselves (the end-to-end path).
 *   - Every function returns an rnde_status; nothing throws across the ABI.
 *     Solver failures (maxiters, dt<=dtmin, NaN) are reported in the return
 *     code AND in rnde_stats.retcode, mirroring upstream retcodes.
 *   - A handle is not thread-safe; distinct handles are independent.  A handle
 *     belongs to the device that was current in rnde_create / rnde_gru_create;
 *     its entry points run there whatever the caller's current device is and
 *     restore the caller's device before returning.
 *   - There is NO CPU fallback: with no CUDA device every call returns
 *     RNDE_ERR_CUDA.
 */
#ifndef REGNDE_H
#define REGNDE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RNDE_VERSION 100

typedef enum {
    RNDE_OK = 0,
    RNDE_ERR_MAXITERS = 1,     /* retcode :MaxIters   */
    RNDE_ERR_DTMIN = 2,        /* retcode :DtLessThanMin */
    RNDE_ERR_NAN = 3,          /* retcode :DtNaN / unstable */
    RNDE_ERR_ARG = 4,
    RNDE_ERR_UNSUPPORTED = 5,  /* shape does not fit any kernel variant */
    RNDE_ERR_CUDA = 6,
    RNDE_ERR_TAPE_FULL = 7,    /* more accepted steps than tape_capacity */
    RNDE_ERR_STATE = 8         /* backward without a taped forward */
} rnde_status;

enum { RNDE_ACT_IDENTITY = 0, RNDE_ACT_TANH = 1 };
/* solver_args... of the constructor: Tsit5() or AutoTsit5(Tsit5()) */
enum { RNDE_ALG_TSIT5 = 0, RNDE_ALG_AUTO_TSIT5 = 1 };
/* the `func` closures the reference passes to SavingCallback:
 *   ERR_DT         integrator.EEst * integrator.dt          neural_ode.jl:116, mnist_node.jl:67
 *   STIFF_DT_ABS   abs(integrator.eigen_est*integrator.dt)  test/test_node.jl:75
 *   STIFF_SCALED   stability_size*|eigen_est| (0/NaN guard)  mnist_node.jl:76-79
 *   ERR_PLUS_STIFF EEst*dt + 0.1*stability_size*eigen_est   mnist_node.jl:88-97 */
enum { RNDE_REG_NONE = 0, RNDE_REG_ERR_DT = 1, RNDE_REG_STIFF_DT_ABS = 2, RNDE_REG_STIFF_SCALED = 3, RNDE_REG_ERR_PLUS_STIFF = 4 };
/* CTA: weights + state of a column tile in one CTA's shared memory (small fields);
 * STREAM: weights streamed from L2 (any size, slow fallback); CLUSTER: 8-CTA clusters, state in
 * distributed shared memory; CLUSTER4: 4-CTA clusters, state in registers (MNIST-shaped fields). */
enum { RNDE_KERNEL_AUTO = 0, RNDE_KERNEL_CTA = 1, RNDE_KERNEL_STREAM = 2, RNDE_KERNEL_CLUSTER = 3, RNDE_KERNEL_CLUSTER4 = 4,
       RNDE_KERNEL_CHAIN = 5 /* CTA variant with 4-column tiles: chain fields, one CTA per SM at batch 512 */,
       RNDE_KERNEL_CHAIN8 = 6 /* the same with 8-column tiles: FFJORD handles whose batch needs more than one 4-column CTA per SM */ };
enum { RNDE_DIST_SINGLE = 0, RNDE_DIST_EXACT = 1, RNDE_DIST_INDEPENDENT = 2 };
/* FMA_CHAIN: blocked fma chains in a fixed order (every kernel variant).
 * FIXED24: exact truncated fixed-point products, order-independent, run as integer tensor-core MMAs by the cluster-4
 * variant (csrc/fwd4x_kernel.cuh); bit-identical to the oracle's arith = 1.  The two modes are two
 * different roundings of the same field (relative difference ~1e-7 per evaluation). */
enum { RNDE_ARITH_FMA_CHAIN = 0, RNDE_ARITH_FIXED24 = 1,
       /* SPLITK: fma chains like FMA_CHAIN, but the contraction index of each layer is dealt out to 8 (layer 1, per CTA) / 4
        * (layer 2) interleaved chains combined by a balanced tree -- the order of the 8x8-tile FFMA2 stepper of the cluster-4
        * variant (csrc/fwd4s_kernel.cuh); bit-identical to the oracle's arith = 2.  The fastest forward for MNIST-shaped fields. */
       RNDE_ARITH_SPLITK = 2 };

typedef struct rnde_config {
    int32_t struct_bytes;     /* sizeof(rnde_config), for versioning */
    int32_t state_dim;        /* D */
    int32_t hidden_dim;       /* H */
    int32_t batch;            /* B: columns held by this handle (this rank) */
    int32_t act_hidden;       /* RNDE_ACT_* of layer 1 */
    int32_t act_out;          /* RNDE_ACT_* of layer 2 (identity: TDChain test; tanh: MLPDynamics) */
    int32_t time_dep;         /* time_dep::Bool of the constructor (neural_ode.jl:8) */
    int32_t kblock;           /* canonical K-blocking of layer 1; 0 = library default */
    int32_t alg;              /* RNDE_ALG_* */
    int32_t reg_kind;         /* RNDE_REG_* ; NONE == regularize=false */
    int32_t max_steps;        /* maxiters; 0 = 1000000 */
    int32_t tape_capacity;    /* accepted steps the backward tape can hold; 0 = 256 */
    int32_t need_backward;    /* record the tape during forward */
    int32_t kernel_variant;   /* RNDE_KERNEL_* */
    int32_t dist_mode;        /* RNDE_DIST_* */
    int32_t rank, nranks;     /* data-parallel position (columns sharded by rank) */
    float t0, t1;             /* tspan */
    float abstol, reltol;     /* solver kwargs */
    float dtmin;              /* 0 = 1e-10 */
    int32_t max_saveat;       /* > 0: the handle serves the multi-save functors (saveat keyword); sizes the saveat buffer */
    int32_t n_layers;         /* 0: the 2-layer time-concatenated field above; 1..8: chain field, see layer_width */
    int64_t global_batch;     /* columns over all ranks (EXACT mode); 0 = batch */
    /* chain field (n_layers > 0): Flux Chain of Dense layers evaluated as re(p)(u), time_dep = 0 -- the Latent-ODE
     * generator dynamics (experiments/latent_ode.jl:109-121).  Layer l maps layer_width[l-1] -> layer_width[l]
     * (layer_width[-1] = layer_width[n_layers-1] = state_dim), activation layer_act[l] (RNDE_ACT_*); pre_act = RNDE_ACT_TANH
     * applies the leading `x -> tanh.(x)`.  p in Flux.destructure order: W_l (out x in, column-major), b_l.  hidden_dim is ignored. */
    int32_t pre_act;
    int32_t layer_width[8];
    int32_t layer_act[8];
    int32_t arith;            /* RNDE_ARITH_*: canonical arithmetic of the layer products (2-layer fields) */
    /* FFJORD field (src/models/ffjord.jl:53-66 over experiments/ffjord_tabular.jl:47-105): csq_extra = 1 or 3 makes the state
     * [z; delta_logp (; ||f||^2; ||e^T J||^2)] with state_dim COUNTING the extra rows, the field MLPDynamics(state_dim -
     * csq_extra, hidden_dim) of three ConcatSquash layers, evaluated with e^T J for the noise given by rnde_set_noise.
     * Forward solves only so far (need_backward must be 0). */
    int32_t csq_extra;
    int32_t reserved0;
} rnde_config;

typedef struct rnde_stats {
    int32_t nf;        /* sol.destats.nf */
    int32_t naccept;
    int32_t nreject;
    int32_t n_saved;   /* length(sv.saveval) = naccept + 1 when regularize */
    int32_t retcode;   /* rnde_status of the solve */
    float t_final;
    float dt_last;
    float dt_init;
} rnde_stats;

typedef struct rnde_handle rnde_handle;

int rnde_version(void);
const char* rnde_status_string(int status);
/* device query used by the host layer to fail loudly: number of visible CUDA devices */
int rnde_device_count(void);

int rnde_create(const rnde_config* cfg, rnde_handle** out);
void rnde_destroy(rnde_handle* h);
const char* rnde_last_error(const rnde_handle* h);
/* number of parameters of the field (length of p) and default kblock for (D, variant) */
int64_t rnde_num_params(const rnde_config* cfg);
int rnde_default_kblock(const rnde_config* cfg);
/* which kernel variant the handle resolved to (RNDE_KERNEL_*) and how many of
 * this library's kernels it has launched so far (bench.py's gpu_launches) */
int rnde_kernel_variant(const rnde_handle* h);
int64_t rnde_launch_count(const rnde_handle* h);

/* tspan override per call (the functor's `tspan` keyword, neural_ode.jl:53,58) */
int rnde_set_tspan(rnde_handle* h, float t0, float t1);

/* Fixed-work replay (SURVEY.md 8d "controller forced to a recorded dt list"): every attempt i of the following forward
 * solves takes dt_host[i] (the last entry repeats; the final step is still clamped to tspan) and is accepted whatever its
 * error estimate, so the number of field evaluations no longer depends on the weights or the data -- a measurement aid
 * (bench.py --fixed-work), not a reference code path.  n = 0 restores the adaptive controller.  RNDE_ARITH_SPLITK handles only. */
int rnde_set_forced_steps(rnde_handle* h, const float* dt_host, int32_t n);

/* What the backward pass differentiates.  `_convert_tspan` (/root/reference/src/utils.jl:21-23)
 * makes tspan tracked, so t and dt are tracked scalars in the reference; every step size the controller proposes is
 * detached (DiffEqBase.value in loopfooter!), the first one -- the Hairer-Wanner initial-dt heuristic -- is not.
 *   RNDE_DETACH_ALL_BUT_FIRST (default, the recalled upstream behaviour): the gradient includes
 *       dL/d(dt_1) * d(dt_1)/d(theta, x), where dt_1 is the first accepted step size, the shift of every later step's start
 *       time and the shortening of the last (clamped) step; costs two more field VJPs per backward pass.
 *   RNDE_DETACH_ALL: the discrete adjoint of the frozen step sequence only.
 * Takes effect from the next rnde_forward (the initial-dt evaluation has to be on the tape).  Handles that replay forced
 * steps always behave as RNDE_DETACH_ALL (their dt_1 does not come from the heuristic).
 *   RNDE_DETACH_FIRST_TERM_ONLY: diagnostic -- the backward returns that extra term alone (dp, dx), so that a test can
 *       check it against the oracle without the Float32 noise of the frozen-step gradient it is 1e-3 ... 1e-7 of. */
enum { RNDE_DETACH_ALL = 0, RNDE_DETACH_ALL_BUT_FIRST = 1, RNDE_DETACH_FIRST_TERM_ONLY = 2 };
int rnde_set_detach(rnde_handle* h, int32_t mode);

/* Forward solve.  x_dev (D x B), p_dev (num_params), u_out_dev (D x B),
 * saveval_dev (>= tape_capacity+1 floats, may be NULL when reg_kind==NONE).
 * stats_host may be NULL (fully asynchronous); otherwise the stream is
 * synchronised and the stats are filled in. */
int rnde_forward(rnde_handle* h, const float* x_dev, const float* p_dev, float* u_out_dev, float* saveval_dev,
                 rnde_stats* stats_host, void* stream);

/* Backward (discrete adjoint of the recorded accepted steps).
 * du_dev: dL/d res (D x B); dsaveval_dev: dL/d sv.saveval[i] (n_saved floats, may be NULL);
 * dp_dev (num_params) and dx_dev (D x B, may be NULL) are OVERWRITTEN. */
int rnde_backward(rnde_handle* h, const float* du_dev, const float* dsaveval_dev, float* dp_dev, float* dx_dev, void* stream);

/* Multi-save functors  (n::TrackedNeuralODE{R,true})(x, p; func)  (neural_ode.jl:79-108 unregularised, :146-180
 * regularised; latent-ODE call site time_series.jl:51).  rnde_set_saveat installs the sorted save times (host
 * array, n <= max_saveat, all inside tspan; n = 0 switches back to the single-save functors) -- the counterpart of
 * update_saveat! (neural_ode.jl:41-45).  Semantics follow solve(...; saveat) of OrdinaryDiffEq 5.50 (SURVEY.md
 * Appendix A.9): no tstops are added; after every accepted step each pending time <= t is produced, by copying u when
 * it equals the step end and by the Tsit5 free interpolant otherwise; a time equal to tspan[1] saves the input.
 * usave_dev is the reference's `res`: feat x nsave x batch, column-major (D*n*B floats); u_out_dev (final state,
 * D x B) may be NULL.  Backward: dusave_dev has the shape of usave_dev, du_dev (cotangent of the final state) may
 * be NULL. */
int rnde_set_saveat(rnde_handle* h, const float* saveat_host, int32_t n);
/* FFJORD handles (rnde_config.csq_extra > 0): the Hutchinson noise e of the next solves, (state_dim - csq_extra) x batch,
 * column-major, caller-owned device memory that must stay valid until the solve has run (the `e` argument of the
 * TrackedFFJORD functors, src/models/ffjord.jl:68-72). */
int rnde_set_noise(rnde_handle* h, const float* e_dev);
/* FFJORD handles: integrate the flow backwards, tspan[1] -> tspan[0] -- `sample` of the reference solves its ODEProblem over
 * [n.tspan[2], n.tspan[1]] (src/models/ffjord.jl:160-167).  With reverse != 0 the following forward solves integrate
 * dz/ds = -f(z, t0 + t1 - s) over s in [t0, t1], i.e. they return z(t0) for the state given at t1.  Forward solves only. */
int rnde_set_reverse_time(rnde_handle* h, int32_t reverse);
int rnde_forward_saveat(rnde_handle* h, const float* x_dev, const float* p_dev, float* u_out_dev, float* usave_dev, float* saveval_dev,
                        rnde_stats* stats_host, void* stream);
int rnde_backward_saveat(rnde_handle* h, const float* du_dev, const float* dusave_dev, const float* dsaveval_dev, float* dp_dev,
                         float* dx_dev, void* stream);

/* ---- Latent-ODE recognition RNN (BASELINE.json north_star (4)) ------------------------------
 * (p::LatentGRU)(x)  experiments/latent_ode.jl:39-99: update / reset / new-state gate networks
 * Chain(Dense(2L+2I+1, H, tanh), Dense(H, L | L | 2L, sigmoid | sigmoid | identity)), sequence consumed backwards
 * in time, observation-mask gating of the state update.  One persistent kernel per direction, weights in shared
 * memory.  x_dev: (2I+1) x T x B column-major (data; mask; delta-t row), p_dev: Flux.destructure order
 * (update_gate, reset_gate, new_state; each W1,b1,W2,b2), out_dev: 2L x B = vcat(y_mean, y_std).
 * rnde_gru_backward: dout_dev 2L x B -> dp_dev (overwritten).  x carries data, so no dx is produced. */
typedef struct rnde_gru_config {
    int32_t struct_bytes;
    int32_t in_dim;        /* I (37 in latent_ode.jl:105; BASELINE.json says 41: a parameter) */
    int32_t hidden_dim;    /* H (40) */
    int32_t latent_dim;    /* L (50) */
    int32_t batch;         /* B */
    int32_t seq_len;       /* T (49 observation times) */
    int32_t need_backward;
    int32_t reserved;
} rnde_gru_config;
typedef struct rnde_gru rnde_gru;
int64_t rnde_gru_num_params(const rnde_gru_config* cfg);
int rnde_gru_create(const rnde_gru_config* cfg, rnde_gru** out);
void rnde_gru_destroy(rnde_gru* g);
const char* rnde_gru_last_error(const rnde_gru* g);
int rnde_gru_forward(rnde_gru* g, const float* x_dev, const float* p_dev, float* out_dev, void* stream);
int rnde_gru_backward(rnde_gru* g, const float* dout_dev, float* dp_dev, void* stream);
int64_t rnde_gru_launch_count(const rnde_gru* g);

/* ---- Neural SDE (BASELINE.json north_star (4)) -------------------------------------------------
 * (n::TrackedNeuralDSDE{R,false})(x, p; func) -> (res, nfe1, nfe2, sv)   src/models/neural_sde.jl:84-146:
 * SDEProblem{false}(drift, diffusion, x, tspan, p) with drift Chain(Dense(D,H,tanh), Dense(H,D)) and DIAGONAL diffusion
 * Dense(D,D) (experiments/mnist_nsde.jl:73-74), solved by SOSRI() or AutoSOSRI2(SOSRI2()) at reltol = abstol = 1.4f-1
 * (:79-80), SavingCallback(func, sv) with func = EEst*dt (:48) or the scaled stiffness estimate (:52-56).
 * p = vcat(p_drift, p_diffusion) in Flux.destructure order (neural_sde.jl:16-18).  One persistent kernel: the adaptive
 * Roessler-SRI stepper with the RSwM3 noise bookkeeping on the device.  The reference's random stream (randn! of Julia's
 * MersenneTwister) cannot be reproduced: the caller SUPPLIES the standard normals, normals_dev[draw][row][column]
 * (n_draws x D x B floats); every request of the solver (dW then dZ of a fresh step, the two bridges of a rejected one)
 * consumes the next draws; rnde_sde_stats.draws reports how many were used, RNDE_ERR_ARG that they ran out.
 * Forward solves only: Tracker.gradient through the SDE solve (mnist_nsde.jl:201-204) is not built. */
enum { RNDE_SDE_SOSRI = 0, RNDE_SDE_AUTO_SOSRI2 = 1 };
typedef struct rnde_sde_config {
    int32_t struct_bytes;
    int32_t state_dim;     /* D (32) */
    int32_t hidden_dim;    /* H (64) */
    int32_t batch;         /* B: columns incl. the trajectory replication of ClassifierNSDE */
    int32_t alg;           /* RNDE_SDE_* */
    int32_t reg_kind;      /* RNDE_REG_NONE | RNDE_REG_ERR_DT | RNDE_REG_STIFF_SCALED */
    int32_t max_steps;     /* maxiters; 0 = 100000 */
    int32_t max_saved;     /* capacity of saveval; 0 = 1024 */
    float t0, t1, abstol, reltol;
} rnde_sde_config;
typedef struct rnde_sde_stats {
    int32_t nfe1, nfe2;    /* drift / diffusion evaluations (neural_sde.jl:46,50) */
    int32_t naccept, nreject, n_saved;
    int32_t draws;         /* normals consumed (in units of one D x B array) */
    int32_t retcode, reserved;
    float t_final, dt_init, dt_last, reserved2;
} rnde_sde_stats;
typedef struct rnde_sde rnde_sde;
int64_t rnde_sde_num_params(const rnde_sde_config* cfg);
int rnde_sde_create(const rnde_sde_config* cfg, rnde_sde** out);
void rnde_sde_destroy(rnde_sde* s);
const char* rnde_sde_last_error(const rnde_sde* s);
/* x_dev, u_out_dev: D x B column-major; saveval_dev: max_saved floats or NULL; stats_host filled after a stream sync */
int rnde_sde_forward(rnde_sde* s, const float* x_dev, const float* p_dev, const float* normals_dev, int32_t n_draws, float* u_out_dev,
                     float* saveval_dev, rnde_sde_stats* stats_host, void* stream);
/* introspection for tests: (dt, EEst, accepted) of the first `cap` attempts of the last solve, host array of 3*cap floats */
/* Reverse sweep of the solve (Tracker.gradient through solve(SDEProblem, SOSRI(); sensealg = SensitivityADPassThrough()),
 * src/models/neural_sde.jl:84-146, experiments/mnist_nsde.jl:201-204): the discrete adjoint of the accepted steps with the step sizes
 * and the Wiener increments frozen.  rnde_sde_enable_tape makes the following forward solves record their accepted steps (state,
 * dW, dZ, dt, EEst; at most tape_capacity of them, RNDE_ERR_TAPE_FULL beyond); rnde_sde_backward then takes the cotangents of the
 * final state (D x B, may be NULL) and of the saved values (rnde_sde_stats.n_saved entries, error-estimate regulariser only; may be
 * NULL) and writes dp (num_params) and dx (D x B, may be NULL). */
int rnde_sde_enable_tape(rnde_sde* s, int32_t tape_capacity);
int rnde_sde_backward(rnde_sde* s, const float* du_dev, const float* dsaveval_dev, float* dp_dev, float* dx_dev, void* stream);
int rnde_sde_get_log(rnde_sde* s, float* log_host, int32_t cap);
int64_t rnde_sde_launch_count(const rnde_sde* s);

/* Host-buffer variants (end-to-end path: copies inside the call). */
int rnde_forward_host(rnde_handle* h, const float* x_host, const float* p_host, float* u_out_host, float* saveval_host,
                      rnde_stats* stats_host);
int rnde_backward_host(rnde_handle* h, const float* du_host, const float* dsaveval_host, float* dp_host, float* dx_host);

/* Classifier head of ClassifierNODE (supervised_classification.jl:44-45) with the
 * experiment's loss (mnist_node.jl:135): logits = W3*u + b3 (C x B),
 * loss = mean_j logitcrossentropy(logits[:,j], y[:,j]).  p3 = [W3 (C x D) col-major; b3 (C)].
 * Writes loss (1 float), logits (C x B, may be NULL), du = dloss/du (D x B), dp3 (C*D + C). */
int rnde_head_loss_grad(rnde_handle* h, const float* u_dev, const float* p3_dev, const float* y_onehot_dev, int32_t n_classes,
                        float loss_scale, float* loss_dev, float* logits_dev, float* du_dev, float* dp3_dev, void* stream);

/* Sum of `buf_dev` (n floats, n <= num_params + 16384) over all ranks of a RNDE_DIST_EXACT group, in place, by a one-shot
 * push all-reduce over the CUDA-IPC mapped peer buffers (NVLink; no NCCL call): ranks are added in rank order on every
 * rank, so the result is bitwise identical everywhere.  Collective: every rank must call it the same number of times.
 * (SURVEY.md 8b rnde_allreduce_grads; with RNDE_DIST_SINGLE / INDEPENDENT handles use the host layer's NCCL all-reduce.) */
int rnde_allreduce_grads(rnde_handle* h, float* buf_dev, int64_t n, void* stream);

/* Statistics of the last rnde_forward on `h` (waits for that forward only, not for work queued behind it). */
int rnde_last_stats(rnde_handle* h, rnde_stats* out);

/* Regulariser term of the experiment's loss, lambda * agg(sv.saveval) (mnist_node.jl:69,80,98,146), and its
 * cotangents for rnde_backward, computed on the device from the saved values of the last forward on `h` -- the
 * number of saved values is read on the device, so the training step needs no host round trip between the forward
 * solve and the backward sweep.  agg: 0 mean, 1 maximum, 2 sum.  dsaveval_dev (tape_capacity+1 floats) is
 * overwritten with cot_scale * d(reg)/d(saveval[i]); reg_dev receives the value (1 float). */
int rnde_reg_agg(rnde_handle* h, int32_t agg, float lam, float cot_scale, const float* saveval_dev, float* dsaveval_dev, float* reg_dev,
                 void* stream);

/* update_parameters!(ps, gs, opt) with opt = Optimiser(InvDecay(gamma), Momentum(eta, rho))
 * (src/utils.jl:149-156, experiments/mnist_node.jl:130), in place on raw arrays:
 *   delta = g * inv_decay_scale   [InvDecay: 1/(1 + gamma*n), n = 1-based update count]
 *   v = rho*v - eta*delta ; p = p + v                        [Flux 0.11.6 Momentum]
 * `h` may be NULL (no handle state is used). */
int rnde_opt_update(rnde_handle* h, float* p_dev, const float* g_dev, float* v_dev, int64_t n, float inv_decay_scale, float eta, float rho,
                    void* stream);
/* Flux.Optimise.Optimiser(WeightDecay(wd), ADAM(eta, (beta1, beta2))) on raw arrays (experiments/ffjord_tabular.jl:128):
 *   delta = g + wd * p ;  m = beta1 m + (1 - beta1) delta ;  v = beta2 v + (1 - beta2) delta^2
 *   p = p - eta * m / (1 - beta1_pow) / (sqrt(v / (1 - beta2_pow)) + eps)        [Flux 0.11.6 ADAM, eps = 1e-8]
 * beta*_pow: the running powers beta^t of this update (t = 1-based), kept by the caller like Flux keeps them in its state.
 * `h` may be NULL. */
int rnde_adam_update(rnde_handle* h, float* p_dev, const float* g_dev, float* m_dev, float* v_dev, int64_t n, float eta, float beta1, float beta2,
                     float beta1_pow, float beta2_pow, float eps, float weight_decay, void* stream);

/* Reference-exact data-parallel mode (RNDE_DIST_EXACT): all ranks take the step sequence of the single batched
 * solve over the global batch.  The per-column sums of squares of every norm are written by each rank straight into
 * every peer's exchange buffer (CUDA IPC peer memory over NVLink) from inside the persistent kernel, followed by a
 * flag barrier; the fold over columns uses the GLOBAL column order, so results are bit-identical to one GPU.
 * Protocol: every rank calls rnde_dist_export, the 64-byte handles are all-gathered by the host layer
 * (torch.distributed / MPI / files), then every rank calls rnde_dist_import with all of them (rank order). */
#define RNDE_IPC_HANDLE_BYTES 64
int rnde_dist_export(rnde_handle* h, void* ipc_handle_out);
int rnde_dist_import(rnde_handle* h, const void* ipc_handles, int32_t nranks);

/* introspection for tests: per accepted step (t, dt, EEst, eigen_est), host arrays of length naccept */
int rnde_get_steps(rnde_handle* h, float* t, float* dt, float* eest, float* eig, int32_t cap);

/* test hooks: canonical device math (include/regnde_canon.h) evaluated on the GPU, for bit-level
 * comparison with the CPU build of the same header */
int rnde_test_tanh(const float* x_dev, float* y_dev, int64_t n, void* stream);
int rnde_test_tanh_bits(uint32_t first_bits, int64_t n, float* y_dev, void* stream);
int rnde_test_pow(const float* x_dev, float e, float* y_dev, float* l10_dev, int64_t n, void* stream);
/* y[i] = fn(float with bit pattern first_bits + i); fn: 0 canon_tanhf, 1 canon_sigmoidf, 2 canon_softplusf, 3 canon_expnegf */
int rnde_test_unary_bits(int32_t fn, uint32_t first_bits, int64_t n, float* y_dev, void* stream);
/* One evaluation of the FFJORD field (csrc/csq.cuh; src/models/ffjord.jl:53-66 over experiments/ffjord_tabular.jl:47-105):
 * k ((data_dim + extra) x batch) = [f(z, t); -sum(eJ .* e) (; ||f||^2; ||eJ||^2)] for z's first data_dim rows, the noise e
 * (data_dim x batch) and the MLPDynamics(data_dim, hidden) parameters p; extra = 1 or 3.  The field is not wired into a
 * stepper yet: this hook exists so the device evaluation can be compared bit for bit with oracle/rnde_oracle.c. */
int rnde_test_csq_rhs(int32_t data_dim, int32_t hidden, int32_t extra, int32_t batch, const float* p_dev, const float* z_dev,
                      const float* e_dev, float t, float* k_dev, void* stream);
int rnde_debug_timeline(rnde_handle* h, long long* out, int n);
/* developer diagnostic of the first-dt adjoint (rnde_set_detach): the two reduced sums of its separate-launch path */
int rnde_debug_a6(rnde_handle* h, float* out2);

#ifdef __cplusplus
}
#endif
#endif /* REGNDE_H */
