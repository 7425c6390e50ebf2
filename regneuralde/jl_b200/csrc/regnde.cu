// regnde.cu -- C ABI of libregnde.so (include/regnde.h): handle management, kernel
// variant selection, workspace / tape allocation in HBM, launches.
// No torch, no CPU fallback: every entry point needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <mutex>
#include <algorithm>
#include <cuda_runtime.h>
#include "regnde.h"
#include "common.cuh"
#include "fwd_kernel.cuh"
#include "fwd4_kernel.cuh"
#include "bwd_kernel.cuh"
#include "bwd4_kernel.cuh"
#include "bwd4tc_kernel.cuh"
#include "fwd4x_kernel.cuh"
#include "fwd4s_kernel.cuh"
#include "wgrad_kernel.cuh"
#include "wgrad_tc_kernel.cuh"
#include "head_kernel.cuh"
#include "gru_kernel.cuh"
#include "csq.cuh"
#include "sde_kernel.cuh"
#include "sde_bwd.cuh"

using namespace rnde;

struct rnde_handle {
    rnde_config cfg;
    int variant = 0;
    int G = 1, NP = 32, R = 0, HS = 0, Q = 0, kblock = 0;
    int num_sms = 0, device = 0;
    size_t smem_fwd = 0, smem_bwd = 0;
    int64_t np = 0;
    // device workspace
    float* colsum = nullptr; int colsum_stride = 0;     // exchange buffer: [colsum 2x3xstride][flags 64][seq...]
    unsigned long long peers[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool peers_open[8] = {false, false, false, false, false, false, false, false};
    bool dist_ready = false;
    unsigned int* bar = nullptr;
    StepRec* steps = nullptr;
    DevStats* stats = nullptr;
    float* tapeZ = nullptr; float* tapeK = nullptr; float* tapeH = nullptr; float* tapeD1 = nullptr;
    float* wg_ws = nullptr; size_t wg_ws_floats = 0;
    float* scal = nullptr;
    float* saveval_int = nullptr;       // used when the caller passes no saveval buffer
    float* dtile = nullptr;             // dx staging in tile layout
    float* head_ws = nullptr;
    float* saveat_dev = nullptr; int n_saveat = 0;
    float* forced_dev = nullptr; int n_forced = 0;      // fixed-work replay (rnde_set_forced_steps)
    // Appendix A.6 with the first dt on the tape (rnde_set_detach, a6.cuh)
    int detach = RNDE_DETACH_ALL_BUT_FIRST;
    float* initdt = nullptr; double* a6_part = nullptr; float* a6_sum = nullptr; float* a6_buf = nullptr; float* a6_f0 = nullptr;
    float* a6_zb = nullptr; float* a6_tau = nullptr; float* a6_kc = nullptr;
    size_t smem_a6 = 0;
    const float* noise = nullptr;       // FFJORD: caller-owned Hutchinson noise (rnde_set_noise)
    int csq_reverse = 0;                // FFJORD: integrate the flow backwards (rnde_set_reverse_time)
    int csq_stage = 0;                  // FFJORD: the parameters are staged in shared memory (one CTA per SM suffices and they fit)
    long long* dbg = nullptr;
    // host-path staging
    float *hx = nullptr, *hp = nullptr, *hu = nullptr, *hsv = nullptr, *hdu = nullptr, *hdsv = nullptr, *hdp = nullptr, *hdx = nullptr;
    DevStats* stats_pinned = nullptr;
    size_t goff_bytes = 0; long long gcap = 0; unsigned ar_seq = 0;     // gradient all-reduce area of the exchange buffer (EXACT mode)
    cudaEvent_t ev_stats = nullptr;      // recorded after the forward's stats copy: the backward waits on it, not on the stream
    const float* last_p = nullptr;
    rnde_stats last_stats{};
    bool have_tape = false;
    int64_t launches = 0;
    std::string err;
};

// The opt-in dynamic shared-memory limit is a property of the KERNEL, shared by all handles that launch it: only ever raise it
// (a later, smaller handle must not pull it below what an earlier handle still launches with).
static std::mutex g_smem_mutex;
static std::map<const void*, size_t> g_smem_attr;
static cudaError_t raise_smem_limit(const void* kern, size_t bytes) {
    std::lock_guard<std::mutex> lk(g_smem_mutex);
    // per device: the attribute is per (function, device); key on both
    int dev = 0; cudaGetDevice(&dev);
    const void* key = reinterpret_cast<const void*>(reinterpret_cast<uintptr_t>(kern) ^ ((uintptr_t)(dev + 1) << 52));
    size_t& cur = g_smem_attr[key];
    if (bytes <= cur) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) cur = bytes;
    return e;
}

static bool g_const_init[64] = {false};
static std::mutex g_const_mutex;      // handles may be created from several host threads

static int set_err(rnde_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg;
    return code;
}
#define CUDA_TRY(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess) {                                                                  \
            return set_err((h), RNDE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
        }                                                                                         \
    } while (0)

// A handle's workspace, constants and kernel attributes belong to the device that was current in rnde_create.  Entry
// points that touch the device run on it whatever the caller's current device is and put the caller's back on return
// (the library never changes the current device behind the caller's back).
struct DeviceScope {
    int prev = -1; bool switched = false; cudaError_t err = cudaSuccess;
    explicit DeviceScope(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) { err = cudaSetDevice(dev); switched = (err == cudaSuccess); }
    }
    ~DeviceScope() { if (switched) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};
#define ON_HANDLE_DEVICE(h)                                                                                          \
    DeviceScope _dev_scope((h)->device);                                                                             \
    if (_dev_scope.err != cudaSuccess) return set_err((h), RNDE_ERR_CUDA, std::string("selecting the handle's device: ") + cudaGetErrorString(_dev_scope.err))

static int init_constants(rnde_handle* h) {
    int dev = 0;
    CUDA_TRY(h, cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (dev < 64 && g_const_init[dev]) return RNDE_OK;
    float A[8][8]; float BT[8]; float C[8];
    memset(A, 0, sizeof(A)); memset(BT, 0, sizeof(BT)); memset(C, 0, sizeof(C));
    A[2][1] = (float)TS_A21;
    A[3][1] = (float)TS_A31; A[3][2] = (float)TS_A32;
    A[4][1] = (float)TS_A41; A[4][2] = (float)TS_A42; A[4][3] = (float)TS_A43;
    A[5][1] = (float)TS_A51; A[5][2] = (float)TS_A52; A[5][3] = (float)TS_A53; A[5][4] = (float)TS_A54;
    A[6][1] = (float)TS_A61; A[6][2] = (float)TS_A62; A[6][3] = (float)TS_A63; A[6][4] = (float)TS_A64; A[6][5] = (float)TS_A65;
    A[7][1] = (float)TS_A71; A[7][2] = (float)TS_A72; A[7][3] = (float)TS_A73; A[7][4] = (float)TS_A74; A[7][5] = (float)TS_A75;
    A[7][6] = (float)TS_A76;
    BT[1] = (float)TS_BT1; BT[2] = (float)TS_BT2; BT[3] = (float)TS_BT3; BT[4] = (float)TS_BT4; BT[5] = (float)TS_BT5;
    BT[6] = (float)TS_BT6; BT[7] = (float)TS_BT7;
    C[2] = (float)TS_C1; C[3] = (float)TS_C2; C[4] = (float)TS_C3; C[5] = (float)TS_C4; C[6] = 1.f; C[7] = 1.f;
    CUDA_TRY(h, cudaMemcpyToSymbol(c_A, A, sizeof(A)));
    CUDA_TRY(h, cudaMemcpyToSymbol(c_BT, BT, sizeof(BT)));
    CUDA_TRY(h, cudaMemcpyToSymbol(c_C, C, sizeof(C)));
    if (dev < 64) g_const_init[dev] = true;
    return RNDE_OK;
}

// ---- kernel variants --------------------------------------------------------
constexpr int NT_FWD = 256;
typedef void (*kern_t)(const KParams);

static kern_t fwd_kernel_for(int variant, int D = 0, int H = 0, int arith = 0, int csq = 0) {
    if (csq > 0) return variant == RNDE_KERNEL_CHAIN ? fwd_kernel<1, 4, 1, true, NT_FWD, 1> : (variant == RNDE_KERNEL_CHAIN8 ? fwd_kernel<1, 8, 1, true, NT_FWD, 1> : nullptr);
    switch (variant) {
        case RNDE_KERNEL_CTA: return fwd_kernel<1, 32, 4, true, NT_FWD>;
        case RNDE_KERNEL_STREAM: return fwd_kernel<1, 4, 1, false, NT_FWD>;
        case RNDE_KERNEL_CHAIN: return fwd_kernel<1, 4, 1, true, NT_FWD>;
        case RNDE_KERNEL_CLUSTER: return fwd_kernel<8, 32, 4, true, NT_FWD>;
        case RNDE_KERNEL_CLUSTER4:
            if (arith == RNDE_ARITH_FIXED24) return fwd4x_kernel;
            if (arith == RNDE_ARITH_SPLITK) return (H == 100 && D == 784) ? fwd4s_kernel<100, 784> : fwd4s_kernel<0, 0>;
            return (H == 100 && D == 784) ? fwd4_kernel<100, 98> : fwd4_kernel<0, 0>;
        default: return nullptr;
    }
}
// the reverse sweep of the cluster-4 decomposition runs its products on the tensor cores (bwd4tc_kernel.cuh) whenever the
// shape fits; RNDE_BWD_FFMA=1 selects the FFMA sweep (bwd4_kernel.cuh)
static bool bwd4_use_tc(int D, int H) {
    static const bool force_ffma = getenv("RNDE_BWD_FFMA") != nullptr;
    return !force_ffma && D > 0 && H > 0 && b4t_shape_ok(D, H);
}
// a6: the instantiation that also serves the first-dt term (a6.cuh); the plain one keeps the hot loop free of it
static kern_t bwd_kernel_for(int variant, int D = 0, int H = 0, bool a6 = true, int csq = 0) {
    if (csq > 0) return variant == RNDE_KERNEL_CHAIN ? bwd_kernel<1, 4, 1, true, NT_FWD, 1> : (variant == RNDE_KERNEL_CHAIN8 ? bwd_kernel<1, 8, 1, true, NT_FWD, 1> : nullptr);
    switch (variant) {
        case RNDE_KERNEL_CTA: return bwd_kernel<1, 32, 4, true, NT_FWD>;
        case RNDE_KERNEL_STREAM: return bwd_kernel<1, 4, 1, false, NT_FWD>;
        case RNDE_KERNEL_CHAIN: return bwd_kernel<1, 4, 1, true, NT_FWD>;
        case RNDE_KERNEL_CLUSTER: return bwd_kernel<8, 32, 4, true, NT_FWD>;
        case RNDE_KERNEL_CLUSTER4:
            if (bwd4_use_tc(D, H)) return a6 ? bwd4tc_kernel<true> : bwd4tc_kernel<false>;
            return (H == 100 && D == 784) ? bwd4_kernel<100, 98> : bwd4_kernel<0, 0>;
        default: return nullptr;
    }
}
// the two extra VJPs of Appendix A.6 (bwd_kernel a6_mode 1 / 2): the variant's own generic sweep, or for the cluster-4
// variant the generic kernel on the same cluster shape and tape layout (4 CTAs x 16 columns, weights streamed from L2)
static kern_t a6_kernel_for(int variant) {
    if (variant == RNDE_KERNEL_CLUSTER4) return bwd_kernel<4, 16, 4, false, NT_FWD>;
    return bwd_kernel_for(variant);
}
static void variant_shape(int variant, int* G, int* NP, bool* WS) {
    switch (variant) {
        case RNDE_KERNEL_CTA: *G = 1; *NP = 32; *WS = true; break;
        case RNDE_KERNEL_STREAM: *G = 1; *NP = 4; *WS = false; break;
        case RNDE_KERNEL_CHAIN: *G = 1; *NP = 4; *WS = true; break;
        case RNDE_KERNEL_CHAIN8: *G = 1; *NP = 8; *WS = true; break;
        case RNDE_KERNEL_CLUSTER4: *G = V2_G; *NP = V2_NP; *WS = true; break;
        default: *G = 8; *NP = 32; *WS = true; break;
    }
}

extern "C" int rnde_version(void) { return RNDE_VERSION; }

extern "C" const char* rnde_status_string(int s) {
    switch (s) {
        case RNDE_OK: return "ok";
        case RNDE_ERR_MAXITERS: return "maxiters reached";
        case RNDE_ERR_DTMIN: return "dt <= dtmin";
        case RNDE_ERR_NAN: return "NaN in error estimate";
        case RNDE_ERR_ARG: return "invalid argument";
        case RNDE_ERR_UNSUPPORTED: return "shape not supported by any kernel variant";
        case RNDE_ERR_CUDA: return "CUDA error";
        case RNDE_ERR_TAPE_FULL: return "tape capacity exceeded";
        case RNDE_ERR_STATE: return "backward requires a taped forward";
        default: return "unknown";
    }
}

extern "C" int rnde_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// chain fields: widest activation (incl. the state) and tape rows per column (a_0 .. a_{L-2})
static int chain_maxw(const rnde_config& c) {
    int m = c.state_dim;
    for (int l = 0; l < c.n_layers; ++l) m = std::max(m, (int)c.layer_width[l]);
    return m;
}
static int chain_hrows(const rnde_config& c) {
    int r = c.state_dim;
    for (int l = 0; l + 1 < c.n_layers; ++l) r += c.layer_width[l];
    return r;
}

extern "C" int64_t rnde_num_params(const rnde_config* c) {
    if (!c) return 0;
    if (c->csq_extra > 0) return csq_num_params(c->state_dim - c->csq_extra, c->hidden_dim);
    const int td = c->time_dep ? 1 : 0;
    if (c->n_layers > 0) {
        int64_t n = 0;
        int K = c->state_dim;
        for (int l = 0; l < c->n_layers; ++l) { n += (int64_t)c->layer_width[l] * K + c->layer_width[l]; K = c->layer_width[l]; }
        return n;
    }
    return (int64_t)c->hidden_dim * (c->state_dim + td) + c->hidden_dim + (int64_t)c->state_dim * (c->hidden_dim + td) + c->state_dim;
}

extern "C" int rnde_default_kblock(const rnde_config* c) {
    // canonical K-blocking of layer 1 (DESIGN.md "canonical arithmetic"): large states are
    // summed in 8 blocks (one per CTA of the cluster kernel), small ones in a single chain.
    if (!c) return 0;
    const int D = c->state_dim;
    return D >= 128 ? (D + 7) / 8 : D;
}

extern "C" const char* rnde_last_error(const rnde_handle* h) { return h ? h->err.c_str() : "null handle"; }
extern "C" int rnde_kernel_variant(const rnde_handle* h) { return h ? h->variant : 0; }
extern "C" int64_t rnde_launch_count(const rnde_handle* h) { return h ? h->launches : 0; }

static size_t smem_bytes_fwd(int variant, int D, int H, int R, int HS, int kblock, int arith = 0) {
    if (variant == RNDE_KERNEL_CLUSTER4 && arith == RNDE_ARITH_FIXED24) return (size_t)make_v4x_layout(D, H).total;
    if (variant == RNDE_KERNEL_CLUSTER4 && arith == RNDE_ARITH_SPLITK) return (size_t)make_v5_layout(D, H).total * sizeof(float);
    if (variant == RNDE_KERNEL_CLUSTER4) return (size_t)make_v2_layout(D, H).total * sizeof(float);
    int G, NP; bool WS;
    variant_shape(variant, &G, &NP, &WS);
    return (size_t)make_layout(G, NP, WS, D, H, R, HS, kblock).total * sizeof(float);
}
static size_t smem_bytes_bwd(int variant, int D, int H, int R, int HS, int kblock) {
    if (variant == RNDE_KERNEL_CLUSTER4) return bwd4_use_tc(D, H) ? (size_t)make_b4t_layout(D, H).total : (size_t)make_b4_layout(D, H).total * sizeof(float);
    int G, NP; bool WS;
    variant_shape(variant, &G, &NP, &WS);
    return (size_t)make_bwd_layout(G, NP, WS, D, H, R, HS).total * sizeof(float);
}

// extra shared memory of a chain field: all parameters + two ping-pong activation tiles
static size_t chain_smem_floats(const rnde_config& c, int NP, bool backward) {
    if (c.n_layers <= 0) return 0;
    return (size_t)round_up((int)rnde_num_params(&c), 4) + 2 * (size_t)round_up(chain_maxw(c), 4) * NP + (backward ? (size_t)chain_hrows(c) * NP : 0);
}

// the same for the reverse pass: noise tile, stage-input tile, the tiles of csq_bwd.cuh
static size_t csq_bwd_smem_floats(const rnde_config& c, int NP) {
    if (c.csq_extra <= 0) return 0;
    const int Dz = c.state_dim - c.csq_extra;
    return (size_t)Dz * NP + (size_t)c.state_dim * NP + (size_t)csq_bwd_tile_floats(Dz, c.hidden_dim, c.csq_extra, NP) + 4;
}
// extra shared memory of the FFJORD field: the noise tile and the tiles of one evaluation (csq.cuh)
static size_t csq_smem_floats(const rnde_config& c, int NP) {
    if (c.csq_extra <= 0) return 0;
    const int Dz = c.state_dim - c.csq_extra;
    return (size_t)Dz * NP + (size_t)csq_tile_floats(Dz, c.hidden_dim, NP) + 4;      // + alignment of the region's start
}

static void free_all(rnde_handle* h) {
    for (int i = 0; i < 8; ++i) if (h->peers_open[i]) cudaIpcCloseMemHandle((void*)h->peers[i]);
    cudaFree(h->colsum); cudaFree(h->bar); cudaFree(h->steps); cudaFree(h->stats);
    cudaFree(h->tapeZ); cudaFree(h->tapeK); cudaFree(h->tapeH); cudaFree(h->tapeD1); cudaFree(h->wg_ws); cudaFree(h->scal); cudaFree(h->saveval_int);
    cudaFree(h->a6_zb); cudaFree(h->a6_tau); cudaFree(h->a6_kc); cudaFree(h->initdt); cudaFree(h->a6_part); cudaFree(h->a6_sum); cudaFree(h->a6_buf); cudaFree(h->a6_f0);
    cudaFree(h->dtile); cudaFree(h->head_ws); cudaFree(h->dbg); cudaFree(h->saveat_dev); cudaFree(h->forced_dev);
    if (h->ev_stats) cudaEventDestroy(h->ev_stats);
    cudaFree(h->hx); cudaFree(h->hp); cudaFree(h->hu); cudaFree(h->hsv); cudaFree(h->hdu); cudaFree(h->hdsv); cudaFree(h->hdp); cudaFree(h->hdx);
    if (h->stats_pinned) cudaFreeHost(h->stats_pinned);
}

extern "C" void rnde_destroy(rnde_handle* h) {
    if (!h) return;
    DeviceScope scope(h->device);
    free_all(h);
    delete h;
}

static int try_variant(rnde_handle* h, int variant, size_t smem_limit, std::string* why) {
    const rnde_config& c = h->cfg;
    const int D = c.state_dim, H = c.hidden_dim, B = c.batch;
    int G, NP; bool WS;
    variant_shape(variant, &G, &NP, &WS);
    const int R = (D + G - 1) / G;
    const int HS = (H + G - 1) / G;
    const int Q = (B + NP - 1) / NP;
    if (variant == RNDE_KERNEL_CLUSTER4) {
        if (c.max_saveat > 0) { *why = "cluster-4 variant has no saveat path"; return 0; }
        if (c.arith == RNDE_ARITH_SPLITK) {
            if (!v5_shape_ok(D, H) || c.n_layers > 0) { *why = "RNDE_ARITH_SPLITK needs D % 4 == 0, 64 <= D <= 896, 4 <= H <= 112"; return 0; }
            if (c.need_backward && !v2_shape_ok(D, H)) { *why = "cluster-4 reverse sweep needs D % 8 == 0"; return 0; }
        } else if (!v2_shape_ok(D, H) || h->kblock != D / 8) { *why = "cluster-4 variant needs D % 8 == 0, kblock == D/8, H <= 128, D <= 1024"; return 0; }
        if (c.need_backward && !bwd_kernel_for(variant)) { *why = "cluster-4 backward not available"; return 0; }
    } else if (G > 1 && h->kblock != R) { *why = "cluster variant needs kblock == ceil(D/8)"; return 0; }
    if (G > 1 && (D < 64 || (G - 1) * R >= D)) { *why = "state too small for the cluster variant"; return 0; }
    const int nbl = (D + h->kblock - 1) / h->kblock;
    if (nbl > 64) { *why = "more than 64 canonical K-blocks"; return 0; }
    if (c.n_layers > 0 && variant != RNDE_KERNEL_CHAIN && variant != RNDE_KERNEL_CTA) { *why = "chain fields run on the CHAIN / CTA variants"; return 0; }
    if (c.csq_extra > 0 && variant != RNDE_KERNEL_CHAIN && variant != RNDE_KERNEL_CHAIN8) { *why = "the FFJORD field runs on the CHAIN variants (4- or 8-column tiles)"; return 0; }
    if (c.csq_extra == 0 && variant == RNDE_KERNEL_CHAIN8) { *why = "the 8-column CHAIN variant serves FFJORD handles"; return 0; }
    size_t sf = smem_bytes_fwd(variant, D, H, R, HS, h->kblock, c.arith) + sizeof(float) * (chain_smem_floats(c, NP, false) + csq_smem_floats(c, NP));
    size_t sb = c.need_backward ? smem_bytes_bwd(variant, D, H, R, HS, h->kblock) + sizeof(float) * (chain_smem_floats(c, NP, true) + csq_bwd_smem_floats(c, NP)) : 0;
    int csq_stage = 0;
    if (c.csq_extra > 0 && Q <= h->num_sms) {      // FFJORD: keep the parameters in shared memory when one CTA per SM is enough and they fit
        const size_t pbytes = sizeof(float) * (size_t)round_up(csq_num_params(D - c.csq_extra, H), 4);
        if (sf + pbytes <= smem_limit && (sb == 0 || sb + pbytes <= smem_limit)) { sf += pbytes; if (sb) sb += pbytes; csq_stage = 1; }
    }
    if (sf > smem_limit || sb > smem_limit) { *why = "shared memory: need " + std::to_string(std::max(sf, sb)) + " B"; return 0; }
    // all CTAs must be co-resident (persistent grid with a grid barrier)
    if (c.arith == RNDE_ARITH_SPLITK && variant != RNDE_KERNEL_CLUSTER4) { *why = "RNDE_ARITH_SPLITK is implemented by the cluster-4 variant"; return 0; }
    if (c.arith == RNDE_ARITH_FIXED24 && (variant != RNDE_KERNEL_CLUSTER4 || c.n_layers > 0 || !v4x_shape_ok(D, H))) {
        *why = "RNDE_ARITH_FIXED24 is implemented by the cluster-4 variant for 128 < D/4 <= 256, H <= 128"; return 0;
    }
    kern_t kf = fwd_kernel_for(variant, D, H, c.arith, c.csq_extra);
    if (raise_smem_limit((const void*)kf, sf) != cudaSuccess) { cudaGetLastError(); *why = "cudaFuncSetAttribute(fwd) failed"; return 0; }
    size_t sa = sb;
    if (c.need_backward) {
        for (int a = 0; a < 2; ++a) {      // both instantiations (with / without the first-dt additions)
            kern_t kb = bwd_kernel_for(variant, D, H, a != 0, c.csq_extra);
            if (raise_smem_limit((const void*)kb, sb) != cudaSuccess) { cudaGetLastError(); *why = "cudaFuncSetAttribute(bwd) failed"; return 0; }
        }
        if (variant == RNDE_KERNEL_CLUSTER4) {
            sa = (size_t)make_bwd_layout(4, 16, false, D, H, R, HS).total * sizeof(float);
            if (sa > smem_limit) { *why = "shared memory (initial-dt adjoint): need " + std::to_string(sa) + " B"; return 0; }
            if (raise_smem_limit((const void*)a6_kernel_for(variant), sa) != cudaSuccess) { cudaGetLastError(); *why = "cudaFuncSetAttribute(a6) failed"; return 0; }
        }
    }
    int max_cta = 0;
    if (G == 1) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kf, NT_FWD, sf) != cudaSuccess) { cudaGetLastError(); *why = "occupancy query failed"; return 0; }
        max_cta = per_sm * h->num_sms;
        if (c.need_backward) {
            int per_sm_b = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_b, bwd_kernel_for(variant, D, H, true, c.csq_extra), NT_FWD, sb);
            max_cta = std::min(max_cta, per_sm_b * h->num_sms);
        }
    } else {
        cudaLaunchConfig_t lc{};
        lc.gridDim = dim3(Q * G); lc.blockDim = dim3(NT_FWD); lc.dynamicSmemBytes = sf;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = G; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        lc.attrs = at; lc.numAttrs = 1;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, kf, &lc) != cudaSuccess) { cudaGetLastError(); *why = "cluster occupancy query failed"; return 0; }
        max_cta = ncl * G;
        if (c.need_backward) {
            lc.dynamicSmemBytes = sb;
            int nclb = 0;
            if (cudaOccupancyMaxActiveClusters(&nclb, bwd_kernel_for(variant, D, H), &lc) == cudaSuccess) max_cta = std::min(max_cta, nclb * G);
            else cudaGetLastError();
        }
    }
    if (Q * G > max_cta) { *why = "grid of " + std::to_string(Q * G) + " CTAs exceeds co-resident capacity " + std::to_string(max_cta); return 0; }
    h->variant = variant; h->G = G; h->NP = NP; h->R = R; h->HS = HS; h->Q = Q; h->smem_fwd = sf; h->smem_bwd = sb; h->smem_a6 = sa; h->csq_stage = csq_stage;
    return 1;
}

extern "C" int rnde_create(const rnde_config* cfg, rnde_handle** out) {
    if (!cfg || !out) return RNDE_ERR_ARG;
    *out = nullptr;
    if (cfg->struct_bytes != (int32_t)sizeof(rnde_config)) return RNDE_ERR_ARG;
    if (cfg->state_dim <= 0 || cfg->batch <= 0 || cfg->n_layers < 0 || cfg->n_layers > 8) return RNDE_ERR_ARG;
    if (cfg->n_layers == 0 && cfg->hidden_dim <= 0) return RNDE_ERR_ARG;
    if (cfg->n_layers > 0) {
        if (cfg->time_dep || cfg->layer_width[cfg->n_layers - 1] != cfg->state_dim || cfg->pre_act < 0 || cfg->pre_act > 1) return RNDE_ERR_ARG;
        int K = cfg->state_dim;
        for (int l = 0; l < cfg->n_layers; ++l) {
            const int M = cfg->layer_width[l];
            if (M <= 0 || M > 1024 || cfg->layer_act[l] < 0 || cfg->layer_act[l] > 1) return RNDE_ERR_ARG;
            // dense_wgrad_kernel: at most CW_MAXTILES 4x4 tiles of a layer's (out x (in + bias)) gradient, and its stage in shared memory
            if (cfg->need_backward && (((M + 3) / 4) * ((K + 1 + 3) / 4) > CW_MAXTILES || dense_wgrad_smem(M, K) > CW_SMEM_MAX)) return RNDE_ERR_UNSUPPORTED;
            K = M;
        }
    }
    if (cfg->act_hidden < 0 || cfg->act_hidden > 1 || cfg->act_out < 0 || cfg->act_out > 1) return RNDE_ERR_ARG;
    if (cfg->csq_extra != 0) {
        if ((cfg->csq_extra != 1 && cfg->csq_extra != 3) || cfg->state_dim <= cfg->csq_extra || cfg->n_layers != 0 || cfg->hidden_dim <= 0 ||
            cfg->arith != RNDE_ARITH_FMA_CHAIN || cfg->max_saveat > 0) return RNDE_ERR_ARG;
        if (cfg->dist_mode == RNDE_DIST_EXACT) return RNDE_ERR_UNSUPPORTED;
    }
    if (cfg->reg_kind < 0 || cfg->reg_kind > RNDE_REG_ERR_PLUS_STIFF || cfg->alg < 0 || cfg->alg > 1) return RNDE_ERR_ARG;
    if (!(cfg->t1 > cfg->t0) || !(cfg->abstol > 0.f) || !(cfg->reltol > 0.f)) return RNDE_ERR_ARG;
    if (cfg->dist_mode < RNDE_DIST_SINGLE || cfg->dist_mode > RNDE_DIST_INDEPENDENT) return RNDE_ERR_ARG;
    if (cfg->arith < RNDE_ARITH_FMA_CHAIN || cfg->arith > RNDE_ARITH_SPLITK) return RNDE_ERR_ARG;
    if (cfg->dist_mode == RNDE_DIST_EXACT && (cfg->nranks < 1 || cfg->nranks > 8 || cfg->rank < 0 || cfg->rank >= cfg->nranks)) return RNDE_ERR_ARG;
    if (rnde_device_count() <= 0) return RNDE_ERR_CUDA;
    rnde_handle* h = new rnde_handle();
    h->cfg = *cfg;
    if (h->cfg.max_steps <= 0) h->cfg.max_steps = 1000000;
    if (h->cfg.tape_capacity <= 0) h->cfg.tape_capacity = 256;
    if (h->cfg.dtmin <= 0.f) h->cfg.dtmin = 1e-10f;
    if (h->cfg.dist_mode != RNDE_DIST_EXACT) { h->cfg.global_batch = h->cfg.batch; h->cfg.nranks = 1; h->cfg.rank = 0; }
    else h->cfg.global_batch = (int64_t)h->cfg.batch * h->cfg.nranks;     // equal shards
    h->kblock = cfg->kblock > 0 ? std::min(cfg->kblock, cfg->state_dim) : rnde_default_kblock(cfg);
    h->np = rnde_num_params(cfg);
    if (cfg->n_layers > 0) h->cfg.hidden_dim = chain_maxw(*cfg);      // sizes the shared scratch of the generic kernels
    cudaDeviceProp prop;
    if (cudaGetDevice(&h->device) != cudaSuccess || cudaGetDeviceProperties(&prop, h->device) != cudaSuccess) { delete h; return RNDE_ERR_CUDA; }
    h->num_sms = prop.multiProcessorCount;
    const size_t smem_limit = prop.sharedMemPerBlockOptin;
    int rc = init_constants(h);
    if (rc != RNDE_OK) { delete h; return rc; }
    std::string why, all;
    int ok = 0;
    if (cfg->kernel_variant != RNDE_KERNEL_AUTO) {
        ok = try_variant(h, cfg->kernel_variant, smem_limit, &why);
        all = why;
    } else {
        // chain fields: 4-column tiles while they give every SM at most one CTA, else 32-column tiles
        const bool chain4 = (cfg->n_layers > 0 && (cfg->batch + 3) / 4 <= h->num_sms) || cfg->csq_extra > 0;
        const int order[5] = {chain4 ? RNDE_KERNEL_CHAIN : RNDE_KERNEL_CTA, chain4 ? RNDE_KERNEL_CTA : RNDE_KERNEL_CLUSTER4, RNDE_KERNEL_CLUSTER, RNDE_KERNEL_STREAM,
                              RNDE_KERNEL_CHAIN8};
        for (int i = 0; i < 5 && !ok; ++i) {
            ok = try_variant(h, order[i], smem_limit, &why);
            if (!ok) all += "[variant " + std::to_string(order[i]) + ": " + why + "] ";
        }
    }
    if (!ok) {
        fprintf(stderr, "regnde: no kernel variant fits D=%d H=%d B=%d: %s\n", cfg->state_dim, cfg->hidden_dim, cfg->batch, all.c_str());
        delete h;
        return RNDE_ERR_UNSUPPORTED;
    }
    // ---- workspace in HBM ----
    const rnde_config& c = h->cfg;
    const int D = c.state_dim, H = c.hidden_dim;
    h->colsum_stride = round_up((int)c.global_batch, 32);
    auto fail = [&](const char* what) { h->err = what; free_all(h); delete h; return RNDE_ERR_CUDA; };
    size_t xwords = (size_t)2 * 3 * h->colsum_stride + 128;
    if (h->cfg.dist_mode == RNDE_DIST_EXACT && h->cfg.nranks > 1) {      // + gradient all-reduce area: flags, done counter, 2 x nranks slots
        xwords = (xwords + 3) / 4 * 4;
        h->goff_bytes = xwords * sizeof(float);
        h->gcap = ((long long)h->np + 16384 + 3) / 4 * 4;
        xwords += 64 + (size_t)2 * h->cfg.nranks * (size_t)h->gcap;
    }
    if (cudaMalloc(&h->colsum, sizeof(float) * xwords) != cudaSuccess) return fail("cudaMalloc colsum");
    if (cudaMemset(h->colsum, 0, sizeof(float) * xwords) != cudaSuccess) return fail("cudaMemset colsum");
    h->peers[h->cfg.rank] = (unsigned long long)h->colsum;
    h->dist_ready = h->cfg.dist_mode != RNDE_DIST_EXACT || h->cfg.nranks == 1;
    if (cudaMalloc(&h->bar, sizeof(unsigned) * 4) != cudaSuccess) return fail("cudaMalloc bar");
    if (cudaMalloc(&h->steps, sizeof(StepRec) * (c.tape_capacity + 1)) != cudaSuccess) return fail("cudaMalloc steps");      // + the pseudo-step of a6.cuh
    if (cudaMalloc(&h->initdt, sizeof(float) * 8) != cudaSuccess || cudaMemset(h->initdt, 0, sizeof(float) * 8) != cudaSuccess) return fail("cudaMalloc initdt");
    if (cudaMalloc(&h->stats, sizeof(DevStats)) != cudaSuccess) return fail("cudaMalloc stats");
    if (cudaMalloc(&h->saveval_int, sizeof(float) * (c.tape_capacity + 1)) != cudaSuccess) return fail("cudaMalloc saveval");
    if (cudaMallocHost(&h->stats_pinned, sizeof(DevStats)) != cudaSuccess) return fail("cudaMallocHost stats");
    if (cudaEventCreateWithFlags(&h->ev_stats, cudaEventDisableTiming) != cudaSuccess) return fail("cudaEventCreate");
    if (c.need_backward) {
        // records: fsalfirst + 6 per step, + one pseudo-step and the initial-dt evaluation of Appendix A.6 (a6.cuh)
        const size_t nrec = 1 + (size_t)6 * (c.tape_capacity + 1) + 1;
        const size_t tile = (size_t)h->Q * h->NP;
        if (cudaMalloc(&h->a6_part, sizeof(double) * 2 * (size_t)h->Q * h->G) != cudaSuccess) return fail("cudaMalloc a6_part");
        if (cudaMalloc(&h->a6_sum, sizeof(float) * 4) != cudaSuccess) return fail("cudaMalloc a6_sum");
        if (cudaMalloc(&h->a6_buf, sizeof(float) * 3 * tile * D) != cudaSuccess) return fail("cudaMalloc a6_buf");
        if (cudaMalloc(&h->a6_f0, sizeof(float) * tile * D) != cudaSuccess) return fail("cudaMalloc a6_f0");
        if (h->variant == RNDE_KERNEL_CLUSTER4) {      // what the tensor-core sweep leaves for the first-dt term (a6.cuh)
            if (cudaMalloc(&h->a6_zb, sizeof(float) * 12 * tile * D) != cudaSuccess) return fail("cudaMalloc a6_zb");
            if (cudaMalloc(&h->a6_kc, sizeof(float) * 12 * tile * D) != cudaSuccess) return fail("cudaMalloc a6_kc");
            if (cudaMalloc(&h->a6_tau, sizeof(float) * nrec * h->Q * h->G * 32) != cudaSuccess) return fail("cudaMalloc a6_tau");
        }
        if (cudaMalloc(&h->tapeZ, sizeof(float) * nrec * tile * D) != cudaSuccess) return fail("cudaMalloc tapeZ (lower tape_capacity?)");
        if (cudaMalloc(&h->tapeK, sizeof(float) * nrec * tile * D) != cudaSuccess) return fail("cudaMalloc tapeK (lower tape_capacity?)");
        const size_t hrows = c.csq_extra > 0 ? (size_t)csq_tape_rows(D - c.csq_extra, H).total : (c.n_layers > 0 ? (size_t)chain_hrows(c) : (size_t)H);
        if (cudaMalloc(&h->tapeH, sizeof(float) * nrec * tile * hrows) != cudaSuccess) return fail("cudaMalloc tapeH");
        if (cudaMalloc(&h->tapeD1, sizeof(float) * nrec * tile * (c.csq_extra > 0 ? 1 : hrows)) != cudaSuccess) return fail("cudaMalloc tapeD1");      // FFJORD: unused
        h->wg_ws_floats = std::max(wgrad_workspace_floats(D, H), (size_t)2 * h->np + 8);
        if (cudaMalloc(&h->wg_ws, sizeof(float) * h->wg_ws_floats) != cudaSuccess) return fail("cudaMalloc wgrad workspace");
        if (cudaMalloc(&h->scal, sizeof(float) * 2 * c.tape_capacity) != cudaSuccess) return fail("cudaMalloc scal");
    }
    if (c.max_saveat > 0 && cudaMalloc(&h->saveat_dev, sizeof(float) * c.max_saveat) != cudaSuccess) return fail("cudaMalloc saveat");
    if (getenv("RNDE_DEBUG_TIMELINE")) { cudaMalloc(&h->dbg, sizeof(long long) * 8000); cudaMemset(h->dbg, 0, sizeof(long long) * 8000); }
    *out = h;
    return RNDE_OK;
}

// developer diagnostic: the two reduced sums of the first-dt adjoint (dL/d(dt_1); <u1bar, f0> + tbar), separate-launch path only
extern "C" int rnde_debug_a6(rnde_handle* h, float* out2) {
    if (!h || !out2 || !h->a6_sum) return RNDE_ERR_ARG;
    ON_HANDLE_DEVICE(h);
    CUDA_TRY(h, cudaDeviceSynchronize());
    CUDA_TRY(h, cudaMemcpy(out2, h->a6_sum, sizeof(float) * 2, cudaMemcpyDeviceToHost));
    return RNDE_OK;
}

extern "C" int rnde_debug_timeline(rnde_handle* h, long long* out, int n) {
    if (!h || !h->dbg) return RNDE_ERR_STATE;
    if (!out || n < 0 || n > 8000) return RNDE_ERR_ARG;
    ON_HANDLE_DEVICE(h);
    CUDA_TRY(h, cudaDeviceSynchronize());
    CUDA_TRY(h, cudaMemcpy(out, h->dbg, sizeof(long long) * n, cudaMemcpyDeviceToHost));
    return RNDE_OK;
}

extern "C" int rnde_dist_export(rnde_handle* h, void* ipc_handle_out) {
    if (!h || !ipc_handle_out) return RNDE_ERR_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == RNDE_IPC_HANDLE_BYTES, "IPC handle size");
    ON_HANDLE_DEVICE(h);
    cudaIpcMemHandle_t mh;
    CUDA_TRY(h, cudaIpcGetMemHandle(&mh, h->colsum));
    memcpy(ipc_handle_out, &mh, sizeof(mh));
    return RNDE_OK;
}

extern "C" int rnde_dist_import(rnde_handle* h, const void* ipc_handles, int32_t nranks) {
    if (!h || !ipc_handles || nranks != h->cfg.nranks) return RNDE_ERR_ARG;
    ON_HANDLE_DEVICE(h);
    const unsigned char* p = (const unsigned char*)ipc_handles;
    for (int r = 0; r < nranks; ++r) {
        if (r == h->cfg.rank || h->peers_open[r]) continue;      // a repeated import keeps the mappings it already has
        cudaIpcMemHandle_t mh;
        memcpy(&mh, p + (size_t)r * RNDE_IPC_HANDLE_BYTES, sizeof(mh));
        void* ptr = nullptr;
        CUDA_TRY(h, cudaIpcOpenMemHandle(&ptr, mh, cudaIpcMemLazyEnablePeerAccess));
        h->peers[r] = (unsigned long long)ptr;
        h->peers_open[r] = true;
    }
    h->dist_ready = true;
    return RNDE_OK;
}

extern "C" int rnde_set_tspan(rnde_handle* h, float t0, float t1) {
    if (!h || !(t1 > t0)) return RNDE_ERR_ARG;
    h->cfg.t0 = t0; h->cfg.t1 = t1;
    return RNDE_OK;
}

extern "C" int rnde_set_saveat(rnde_handle* h, const float* saveat_host, int32_t n) {
    if (!h || n < 0 || (n > 0 && !saveat_host)) return RNDE_ERR_ARG;
    if (n > h->cfg.max_saveat) return set_err(h, RNDE_ERR_ARG, "more saveat times than rnde_config.max_saveat");
    for (int i = 0; i < n; ++i) {
        if (!(saveat_host[i] >= h->cfg.t0 && saveat_host[i] <= h->cfg.t1)) return set_err(h, RNDE_ERR_ARG, "saveat time outside tspan");
        if (i > 0 && !(saveat_host[i] > saveat_host[i - 1])) return set_err(h, RNDE_ERR_ARG, "saveat times must be strictly increasing");
    }
    ON_HANDLE_DEVICE(h);
    if (n > 0) CUDA_TRY(h, cudaMemcpy(h->saveat_dev, saveat_host, sizeof(float) * n, cudaMemcpyHostToDevice));
    h->n_saveat = n;
    return RNDE_OK;
}

extern "C" int rnde_set_forced_steps(rnde_handle* h, const float* dt_host, int32_t n) {
    if (!h || n < 0 || (n > 0 && !dt_host)) return RNDE_ERR_ARG;
    if (n > 0 && (h->cfg.arith != RNDE_ARITH_SPLITK || h->variant != RNDE_KERNEL_CLUSTER4))
        return set_err(h, RNDE_ERR_UNSUPPORTED, "rnde_set_forced_steps: only the RNDE_ARITH_SPLITK stepper replays a step list");
    for (int i = 0; i < n; ++i) if (!(dt_host[i] > 0.f)) return set_err(h, RNDE_ERR_ARG, "forced dt must be positive");
    ON_HANDLE_DEVICE(h);
    if (n > h->n_forced || !h->forced_dev) {
        cudaFree(h->forced_dev); h->forced_dev = nullptr;
        if (n > 0) CUDA_TRY(h, cudaMalloc(&h->forced_dev, sizeof(float) * n));
    }
    if (n > 0) CUDA_TRY(h, cudaMemcpy(h->forced_dev, dt_host, sizeof(float) * n, cudaMemcpyHostToDevice));
    h->n_forced = n;
    return RNDE_OK;
}

extern "C" int rnde_set_reverse_time(rnde_handle* h, int32_t reverse) {
    if (!h) return RNDE_ERR_ARG;
    if (h->cfg.csq_extra <= 0) return set_err(h, RNDE_ERR_STATE, "rnde_set_reverse_time: FFJORD handles only");
    if (reverse && h->cfg.need_backward) return set_err(h, RNDE_ERR_UNSUPPORTED, "reverse-time solves are forward only");
    h->csq_reverse = reverse ? 1 : 0;
    return RNDE_OK;
}

extern "C" int rnde_set_noise(rnde_handle* h, const float* e_dev) {
    if (!h || !e_dev) return RNDE_ERR_ARG;
    if (h->cfg.csq_extra <= 0) return set_err(h, RNDE_ERR_STATE, "rnde_set_noise: the handle was not created with csq_extra");
    h->noise = e_dev;
    return RNDE_OK;
}

static void fill_params(const rnde_handle* h, KParams& P) {
    const rnde_config& c = h->cfg;
    memset(&P, 0, sizeof(P));
    P.D = c.state_dim; P.H = c.hidden_dim; P.B = c.batch;
    P.R = h->R; P.kblock = h->kblock; P.HS = h->HS; P.Q = h->Q;
    P.act1 = c.act_hidden; P.act2 = c.act_out; P.td = c.time_dep ? 1 : 0;
    P.alg = c.alg; P.reg_kind = c.reg_kind; P.max_steps = c.max_steps; P.tape_cap = c.tape_capacity; P.need_tape = c.need_backward ? 1 : 0;
    P.t0 = c.t0; P.t1 = c.t1; P.abstol = c.abstol; P.reltol = c.reltol; P.dtmin = c.dtmin;
    P.norm_count = (long long)c.state_dim * (long long)c.global_batch;   // SINGLE / INDEPENDENT: global_batch == batch
    P.Bglobal = (int)c.global_batch; P.col_offset = c.rank * c.batch;
    P.nranks = c.nranks; P.rank = c.rank; P.flag_off = (unsigned)(2 * 3 * h->colsum_stride);
    for (int i = 0; i < 8; ++i) P.peers[i] = h->peers[i];
    P.colsum = h->colsum; P.colsum_stride = h->colsum_stride; P.bar = h->bar; P.steps = h->steps; P.stats = h->stats;
    P.dbg = h->dbg;
    P.n_layers = c.n_layers; P.pre_act = c.pre_act; P.hrows = c.csq_extra > 0 ? csq_tape_rows(c.state_dim - c.csq_extra, c.hidden_dim).total : (c.n_layers > 0 ? chain_hrows(c) : 0); P.chain_np = c.n_layers > 0 ? (int)h->np : 0;
    for (int l = 0; l < 8; ++l) { P.lw[l] = c.layer_width[l]; P.la[l] = c.layer_act[l]; }
    P.tapeZ = h->tapeZ; P.tapeK = h->tapeK; P.tapeH = h->tapeH; P.tapeD1 = h->tapeD1; P.scal = h->scal;
    P.forced_dt = h->forced_dev; P.n_forced = h->n_forced;
    P.a6 = (c.need_backward && h->detach != RNDE_DETACH_ALL && h->n_forced == 0 && c.csq_extra == 0) ? 1 : 0;
    P.rec_init = 1 + 6 * (c.tape_capacity + 1);
    P.a6_scalar = (c.dist_mode == RNDE_DIST_EXACT && c.nranks > 1 && c.rank != 0) ? 0 : 1;
    P.initdt = h->initdt; P.a6_part = h->a6_part; P.a6_sum = h->a6_sum; P.a6_u1bar = h->a6_buf; P.a6_f0 = h->a6_f0; P.a6_zb = h->a6_zb; P.a6_tau = h->a6_tau; P.a6_kc = h->a6_kc;
}

// shared-memory offsets of the chain region: right after the generic kernel's own layout
static void set_chain_offsets(const rnde_handle* h, KParams& P, int base_floats) {
    if (h->cfg.n_layers <= 0) return;
    const int mw = round_up(chain_maxw(h->cfg), 4);
    P.oCW = round_up(base_floats, 4);
    P.oCA = P.oCW + round_up((int)h->np, 4);
    P.oCB = P.oCA + mw * h->NP;
    P.oCH = P.oCB + mw * h->NP;
}

static int launch(rnde_handle* h, kern_t k, const KParams& P, size_t smem, cudaStream_t st) {
    cudaLaunchConfig_t lc{};
    lc.gridDim = dim3(h->Q * h->G); lc.blockDim = dim3(NT_FWD); lc.dynamicSmemBytes = smem; lc.stream = st;
    cudaLaunchAttribute at[1];
    int na = 0;
    if (h->G > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = h->G; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        na++;
    }
    lc.attrs = at; lc.numAttrs = na;
    CUDA_TRY(h, cudaLaunchKernelEx(&lc, k, P));
    h->launches += 1;
    return RNDE_OK;
}

static int forward_impl(rnde_handle* h, const float* x_dev, const float* p_dev, float* u_out_dev, float* usave_dev, float* saveval_dev,
                        rnde_stats* stats_host, void* stream) {
    if (!h || !x_dev || !p_dev || (!u_out_dev && !usave_dev)) return RNDE_ERR_ARG;
    if (!h->dist_ready) return set_err(h, RNDE_ERR_STATE, "RNDE_DIST_EXACT: call rnde_dist_export / rnde_dist_import on every rank first");
    ON_HANDLE_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    KParams P;
    fill_params(h, P);
    P.x = x_dev; P.p = p_dev; P.u_out = u_out_dev;
    P.saveval = saveval_dev ? saveval_dev : h->saveval_int;
    if (usave_dev) {
        if (h->n_saveat <= 0) return set_err(h, RNDE_ERR_STATE, "rnde_forward_saveat needs rnde_set_saveat first");
        P.saveat = h->saveat_dev; P.n_saveat = h->n_saveat; P.usave = usave_dev;
    }
    if (h->cfg.n_layers > 0) {
        bool WS; int G, NP; variant_shape(h->variant, &G, &NP, &WS);
        set_chain_offsets(h, P, make_layout(G, NP, WS, h->cfg.state_dim, h->cfg.hidden_dim, h->R, h->HS, h->kblock).total);
    }
    if (h->cfg.csq_extra > 0) {
        if (!h->noise) return set_err(h, RNDE_ERR_STATE, "FFJORD handle: call rnde_set_noise first");
        bool WS; int G, NP; variant_shape(h->variant, &G, &NP, &WS);
        P.noise = h->noise; P.csq_extra = h->cfg.csq_extra; P.csq_reverse = h->csq_reverse;
        P.oCS = round_up(make_layout(G, NP, WS, h->cfg.state_dim, h->cfg.hidden_dim, h->R, h->HS, h->kblock).total, 4);
        if (h->csq_stage) P.oCSP = round_up(P.oCS + (int)csq_smem_floats(h->cfg, NP), 4);
    }
    CUDA_TRY(h, cudaMemsetAsync(h->bar, 0, sizeof(unsigned) * 4, st));
    int rc = launch(h, fwd_kernel_for(h->variant, h->cfg.state_dim, h->cfg.hidden_dim, h->cfg.arith, h->cfg.csq_extra), P, h->smem_fwd, st);
    if (rc != RNDE_OK) return rc;
    h->last_p = p_dev;
    h->have_tape = h->cfg.need_backward != 0;
    CUDA_TRY(h, cudaMemcpyAsync(h->stats_pinned, h->stats, sizeof(DevStats), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaEventRecord(h->ev_stats, st));
    if (stats_host) {
        CUDA_TRY(h, cudaStreamSynchronize(st));
        const DevStats& s = *h->stats_pinned;
        stats_host->nf = s.nf; stats_host->naccept = s.naccept; stats_host->nreject = s.nreject; stats_host->n_saved = s.n_saved;
        stats_host->retcode = s.retcode; stats_host->t_final = s.t_final; stats_host->dt_last = s.dt_last; stats_host->dt_init = s.dt_init;
        h->last_stats = *stats_host;
        if (s.retcode != RNDE_OK) return set_err(h, s.retcode, rnde_status_string(s.retcode));
    }
    return RNDE_OK;
}

extern "C" int rnde_forward(rnde_handle* h, const float* x_dev, const float* p_dev, float* u_out_dev, float* saveval_dev,
                            rnde_stats* stats_host, void* stream) {
    if (!u_out_dev) return RNDE_ERR_ARG;
    return forward_impl(h, x_dev, p_dev, u_out_dev, nullptr, saveval_dev, stats_host, stream);
}

extern "C" int rnde_forward_saveat(rnde_handle* h, const float* x_dev, const float* p_dev, float* u_out_dev, float* usave_dev, float* saveval_dev,
                                   rnde_stats* stats_host, void* stream) {
    if (!usave_dev) return RNDE_ERR_ARG;
    return forward_impl(h, x_dev, p_dev, u_out_dev, usave_dev, saveval_dev, stats_host, stream);
}

extern "C" int rnde_get_steps(rnde_handle* h, float* t, float* dt, float* eest, float* eig, int32_t cap) {
    if (!h) return RNDE_ERR_ARG;
    ON_HANDLE_DEVICE(h);
    CUDA_TRY(h, cudaDeviceSynchronize());
    DevStats s;
    CUDA_TRY(h, cudaMemcpy(&s, h->stats, sizeof(s), cudaMemcpyDeviceToHost));
    const int n = std::min(std::min(s.naccept, (int)cap), h->cfg.tape_capacity);
    std::vector<StepRec> r(std::max(n, 1));
    if (n > 0) CUDA_TRY(h, cudaMemcpy(r.data(), h->steps, sizeof(StepRec) * n, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; ++i) {
        if (t) t[i] = r[i].t;
        if (dt) dt[i] = r[i].dt;
        if (eest) eest[i] = r[i].eest;
        if (eig) eig[i] = r[i].eig;
    }
    return RNDE_OK;
}

// ---- backward -----------------------------------------------------------------
static int backward_impl(rnde_handle* h, const float* du_dev, const float* dusave_dev, const float* dsaveval_dev, float* dp_dev, float* dx_dev,
                         void* stream) {
    if (!h || (!du_dev && !dusave_dev) || !dp_dev) return RNDE_ERR_ARG;
    if (!h->have_tape || !h->cfg.need_backward) return set_err(h, RNDE_ERR_STATE, "rnde_backward needs a preceding rnde_forward on a handle created with need_backward=1");
    ON_HANDLE_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    // number of accepted steps of the forward on this handle: wait for its stats copy only, so that work queued
    // behind the forward (classifier head, regulariser aggregation) keeps the GPU busy while the host gets here
    CUDA_TRY(h, cudaEventSynchronize(h->ev_stats));
    const DevStats s = *h->stats_pinned;
    h->last_stats.nf = s.nf; h->last_stats.naccept = s.naccept; h->last_stats.nreject = s.nreject; h->last_stats.n_saved = s.n_saved;
    h->last_stats.retcode = s.retcode; h->last_stats.t_final = s.t_final; h->last_stats.dt_last = s.dt_last; h->last_stats.dt_init = s.dt_init;
    if (s.retcode != RNDE_OK) return set_err(h, RNDE_ERR_STATE, "forward solve failed; nothing to differentiate");
    KParams P;
    fill_params(h, P);
    P.p = h->last_p; P.du = du_dev; P.dsaveval = (h->cfg.reg_kind != RNDE_REG_NONE) ? dsaveval_dev : nullptr; P.dx = dx_dev;
    P.nsteps = s.naccept;
    if (dusave_dev) {
        if (h->n_saveat <= 0) return set_err(h, RNDE_ERR_STATE, "rnde_backward_saveat needs rnde_set_saveat first");
        P.saveat = h->saveat_dev; P.n_saveat = h->n_saveat; P.dusave = dusave_dev;
    }
    if (h->cfg.n_layers > 0) {
        bool WS; int G, NP; variant_shape(h->variant, &G, &NP, &WS);
        set_chain_offsets(h, P, make_bwd_layout(G, NP, WS, h->cfg.state_dim, h->cfg.hidden_dim, h->R, h->HS).total);
    }
    if (h->cfg.csq_extra > 0) {
        if (!h->noise) return set_err(h, RNDE_ERR_STATE, "FFJORD handle: call rnde_set_noise first");
        if (dusave_dev) return set_err(h, RNDE_ERR_UNSUPPORTED, "FFJORD handles have no saveat path");
        bool WS; int G, NP; variant_shape(h->variant, &G, &NP, &WS);
        P.noise = h->noise; P.csq_extra = h->cfg.csq_extra;
        P.oCS = round_up(make_bwd_layout(G, NP, WS, h->cfg.state_dim, h->cfg.hidden_dim, h->R, h->HS).total, 4);
        if (h->csq_stage) P.oCSP = round_up(P.oCS + (int)csq_bwd_smem_floats(h->cfg, NP), 4);
    }
    // Appendix A.6 (a6.cuh): the sweep also accumulates dL/d(dt_1); two more VJPs then differentiate the initial-dt heuristic.
    // Their records sit behind the last step: 6N+1..6N+5 empty, 6N+6 the evaluation f(u0 + dt0 f0, t0 + dt0).
    const bool a6 = P.a6 && s.naccept > 0;
    P.a6 = a6 ? 1 : 0;
    // the tensor-core contraction takes the extra record in a spare K slot (of the second step's group), where it lies;
    // the other contractions see it as stage 7 of a pseudo-step behind the last one (records 6N+1..6N+5 empty)
    static const bool force_ffma = getenv("RNDE_WGRAD_FFMA") != nullptr;
    const bool wg_tc = h->NP == 16 && !force_ffma && h->cfg.n_layers == 0;
    const bool wg_slot = wg_tc && s.naccept >= 2;      // ... in the spare K slot of the SECOND step's group
    // the cluster-4 sweeps differentiate the heuristic themselves (one more VJP between two grid-wide sums -- over all ranks in the
    // reference-exact mode --, a6.cuh); the other variants launch the generic kernel twice after the sweep
    static const bool a6_external = getenv("RNDE_A6_EXTERNAL") != nullptr;      // developer switch: always the separate launches
    const bool a6_inkernel = a6 && !a6_external && h->variant == RNDE_KERNEL_CLUSTER4;
    if (a6_inkernel) { P.a6 = (h->detach == RNDE_DETACH_FIRST_TERM_ONLY) ? 3 : 2; P.rec_x = wg_slot ? P.rec_init : 6 * s.naccept + 6; }
    const size_t tileN = (size_t)h->Q * h->NP;
    const size_t recD = tileN * h->cfg.state_dim;
    const size_t recH = tileN * (h->cfg.n_layers > 0 ? (size_t)chain_hrows(h->cfg) : (size_t)h->cfg.hidden_dim);
    if (a6 && !a6_inkernel) CUDA_TRY(h, cudaMemcpyAsync(h->a6_f0, h->tapeK, sizeof(float) * recD, cudaMemcpyDeviceToDevice, st));      // f0, before delta2 replaces it
    if (a6 && !wg_slot) {
        const size_t rx = (size_t)6 * s.naccept + 6, ri = (size_t)P.rec_init;
        CUDA_TRY(h, cudaMemcpyAsync(h->tapeZ + rx * recD, h->tapeZ + ri * recD, sizeof(float) * recD, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(h, cudaMemcpyAsync(h->tapeK + rx * recD, h->tapeK + ri * recD, sizeof(float) * recD, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(h, cudaMemcpyAsync(h->tapeH + rx * recH, h->tapeH + ri * recH, sizeof(float) * recH, cudaMemcpyDeviceToDevice, st));
        CUDA_TRY(h, cudaMemsetAsync(h->tapeK + (rx - 5) * recD, 0, sizeof(float) * 5 * recD, st));
        CUDA_TRY(h, cudaMemsetAsync(h->tapeD1 + (rx - 5) * recH, 0, sizeof(float) * 5 * recH, st));
        // ... and their inputs too: the contraction multiplies them by the zero deltas, and stale memory may hold NaN
        CUDA_TRY(h, cudaMemsetAsync(h->tapeZ + (rx - 5) * recD, 0, sizeof(float) * 5 * recD, st));
        CUDA_TRY(h, cudaMemsetAsync(h->tapeH + (rx - 5) * recH, 0, sizeof(float) * 5 * recH, st));
    }
    if (a6 && h->a6_kc) {      // k_1..k_6 of the first and of the last step, before the sweep replaces them by delta2 (a6.cuh)
        CUDA_TRY(h, cudaMemcpyAsync(h->a6_kc, h->tapeK, sizeof(float) * 6 * recD, cudaMemcpyDeviceToDevice, st));
        if (s.naccept > 1)
            CUDA_TRY(h, cudaMemcpyAsync(h->a6_kc + 6 * recD, h->tapeK + (size_t)6 * (s.naccept - 1) * recD, sizeof(float) * 6 * recD, cudaMemcpyDeviceToDevice, st));
    }
    CUDA_TRY(h, cudaMemsetAsync(h->bar, 0, sizeof(unsigned) * 4, st));
    int rc = launch(h, bwd_kernel_for(h->variant, h->cfg.state_dim, h->cfg.hidden_dim, P.a6 != 0, h->cfg.csq_extra), P, h->smem_bwd, st);
    if (rc != RNDE_OK) return rc;
    if (a6_inkernel && h->detach == RNDE_DETACH_FIRST_TERM_ONLY) {      // diagnostic: only record 0 and the extra record keep their deltas
        CUDA_TRY(h, cudaMemsetAsync(h->tapeK + recD, 0, sizeof(float) * (size_t)6 * s.naccept * recD, st));
        CUDA_TRY(h, cudaMemsetAsync(h->tapeD1 + recH, 0, sizeof(float) * (size_t)6 * s.naccept * recH, st));
    }
    if (a6 && !a6_inkernel) {
        const int nparts = h->Q * h->G;
        auto reduce = [&](int slot) -> int {
            a6_reduce_kernel<<<1, 256, 0, st>>>(h->a6_part, nparts, h->a6_sum, slot);
            CUDA_TRY(h, cudaGetLastError());
            h->launches += 1;
            // reference-exact data parallel: the step sequence is shared, so dL/d(dt_1) is the sum over all ranks' columns
            if (h->cfg.nranks > 1 && h->cfg.dist_mode == RNDE_DIST_EXACT) return rnde_allreduce_grads(h, h->a6_sum + slot, 1, stream);
            return RNDE_OK;
        };
        KParams PA = P;
        PA.du = nullptr; PA.dusave = nullptr; PA.n_saveat = 0; PA.rec_x = wg_slot ? P.rec_init : 6 * s.naccept + 6;
        if (h->variant == RNDE_KERNEL_CLUSTER4) { PA.n_layers = 0; }
        if ((rc = reduce(0)) != RNDE_OK) return rc;
        if (h->detach == RNDE_DETACH_FIRST_TERM_ONLY) {      // diagnostic: drop what the sweep produced, keep only the first-dt term
            const size_t nr = (size_t)6 * s.naccept + 1;
            CUDA_TRY(h, cudaMemsetAsync(h->tapeK, 0, sizeof(float) * nr * recD, st));
            CUDA_TRY(h, cudaMemsetAsync(h->tapeD1, 0, sizeof(float) * nr * recH, st));
            if (dx_dev) CUDA_TRY(h, cudaMemsetAsync(dx_dev, 0, sizeof(float) * (size_t)h->cfg.state_dim * h->cfg.batch, st));
        }
        PA.a6_mode = 1;
        if ((rc = launch(h, a6_kernel_for(h->variant), PA, h->smem_a6, st)) != RNDE_OK) return rc;
        if ((rc = reduce(1)) != RNDE_OK) return rc;
        PA.a6_mode = 2;
        if ((rc = launch(h, a6_kernel_for(h->variant), PA, h->smem_a6, st)) != RNDE_OK) return rc;
    }
    const int nsteps_w = s.naccept + ((a6 && !wg_slot) ? 1 : 0);      // steps the weight-gradient contraction runs over
    if (h->cfg.csq_extra > 0) {      // FFJORD field (csq_bwd.cuh): every W receives two outer products per record, the bias / gate vectors a sum
        const int Dd = h->cfg.state_dim, Dz = Dd - h->cfg.csq_extra, Hh = h->cfg.hidden_dim;
        const CsqTapeRows R = csq_tape_rows(Dz, Hh);
        WgDesc desc; memset(&desc, 0, sizeof(desc));
        const int Ms[3] = {Hh, Hh, Dz}, Ks[3] = {Dz, Hh, Hh};
        const int linb[3] = {R.linb1, R.linb2, R.linb3}, us[3] = {R.u1, R.u2, R.u3}, vbs[3] = {R.vb1, R.vb2, R.vb3}, vecs[3] = {R.vec1, R.vec2, R.vec3};
        const int ain[3] = {0, R.a1, R.a2};
        int poff = 0;
        auto add = [&](const float* dptr, int dstride, int doff, const float* aptr, int astride, int aoff, int M, int K, int po, int nobias) {
            if (desc.nl >= 16) { desc.nl = 17; return; }      // reported below
            WgLayer& w = desc.l[desc.nl++];
            w.dptr = dptr; w.dstride = dstride; w.doff = doff; w.aptr = aptr; w.astride = astride; w.aoff = aoff; w.M = M; w.K = K; w.poff = po; w.nobias = nobias;
        };
        const int hs = R.total * h->NP, ds = Dd * h->NP;
        for (int l = 0; l < 3; ++l) {
            const int M = Ms[l], K = Ks[l];
            // the contraction kernel holds at most 512 4x4 output tiles per pseudo-layer: split the input index
            const int kmax = std::max(4, ((512 / ((M + 3) / 4)) * 4 - 4) / 4 * 4);
            for (int k0 = 0; k0 < K; k0 += kmax) {
                const int kc = std::min(kmax, K - k0);
                const bool lastc = k0 + kc >= K;
                // forward chain: dW += linb x^T (x = z on the state tape for layer 1), dB = sum linb through the bias row of the LAST chunk.
                // A chunk that is not the last must not add a bias row: it would land inside W.
                if (l == 0) add(h->tapeH, hs, linb[l], h->tapeZ, ds, k0, M, kc, poff + k0 * M, lastc ? 0 : 1);
                else add(h->tapeH, hs, linb[l], h->tapeH, hs, ain[l] + k0, M, kc, poff + k0 * M, lastc ? 0 : 1);
                // transposed chain: dW += u vbar^T
                add(h->tapeH, hs, us[l], h->tapeH, hs, vbs[l] + k0, M, kc, poff + k0 * M, 1);
            }
            // [bias_W | bias_B | gate_W] = sums of the three vectors (K = 0: only the bias row), in chunks that fit the kernel's shared memory
            for (int m0 = 0; m0 < 3 * M; m0 += 160) add(h->tapeH, hs, vecs[l] + m0, h->tapeH, hs, 0, std::min(160, 3 * M - m0), 0, poff + M * K + M + m0, 0);
            poff += M * K + 4 * M;
        }
        if (desc.nl > 16) return set_err(h, RNDE_ERR_UNSUPPORTED, "FFJORD weight gradients: layer too large for the contraction kernel");
        const int nrec_c = 1 + 6 * s.naccept;
        cudaError_t ce = launch_dense_wgrad(desc, h->NP, (long long)nrec_c * h->Q, (int)h->np, h->num_sms, reinterpret_cast<double*>(h->wg_ws), dp_dev, st, &h->launches);
        if (ce != cudaSuccess) return set_err(h, RNDE_ERR_CUDA, std::string("FFJORD wgrad launch: ") + cudaGetErrorString(ce));
        return RNDE_OK;
    }
    if (h->cfg.n_layers > 0) {      // chain field: per-layer contractions over the tape, FP64 across stages
        const int nrec_c = 1 + 6 * nsteps_w;
        WgDesc desc; memset(&desc, 0, sizeof(desc));
        desc.nl = h->cfg.n_layers;
        const int hrows = chain_hrows(h->cfg), Dd = h->cfg.state_dim;
        int K = Dd, poff = 0, hoff = 0;
        for (int l = 0; l < desc.nl; ++l) {
            WgLayer& w = desc.l[l];
            const bool last = (l == desc.nl - 1);
            w.M = h->cfg.layer_width[l]; w.K = K; w.poff = poff;
            w.aptr = h->tapeH; w.astride = hrows * h->NP; w.aoff = hoff;
            w.dptr = last ? h->tapeK : h->tapeD1; w.dstride = (last ? Dd : hrows) * h->NP; w.doff = last ? 0 : hoff + K;
            poff += w.M * K + w.M; hoff += K; K = w.M;
        }
        cudaError_t ce = launch_dense_wgrad(desc, h->NP, (long long)nrec_c * h->Q, (int)h->np, h->num_sms, reinterpret_cast<double*>(h->wg_ws), dp_dev, st, &h->launches);
        if (ce != cudaSuccess) return set_err(h, RNDE_ERR_CUDA, std::string("chain wgrad launch: ") + cudaGetErrorString(ce));
        return RNDE_OK;
    }
    // parameter gradients: two batched contractions over (record, column) -- see wgrad_kernel.cuh
    const int nrec = 1 + 6 * nsteps_w;
    // tensor-core (tcgen05 3xTF32) contraction for the 16-column tape layout; RNDE_WGRAD_FFMA=1 selects the FFMA kernel
    if (wg_tc) {
        rc = launch_wgrad_tc(h->cfg.state_dim, h->cfg.hidden_dim, h->cfg.time_dep ? 1 : 0, nrec, h->Q, h->tapeZ, h->tapeK, h->tapeH, h->tapeD1,
                             h->steps, h->cfg.t0, h->wg_ws, dp_dev, st, &h->launches, (a6 && wg_slot) ? P.rec_init : -1);
        if (rc != 0) return set_err(h, RNDE_ERR_CUDA, std::string("wgrad (tcgen05) launch: ") + cudaGetErrorString((cudaError_t)rc));
        return RNDE_OK;
    }
    rc = launch_wgrad(h->cfg.state_dim, h->cfg.hidden_dim, h->cfg.time_dep ? 1 : 0, nrec, h->Q, h->NP, h->cfg.batch,
                      h->tapeZ, h->tapeK, h->tapeH, h->tapeD1, h->steps, h->cfg.t0, h->wg_ws, dp_dev, st, &h->launches);
    if (rc != 0) return set_err(h, RNDE_ERR_CUDA, std::string("wgrad launch: ") + cudaGetErrorString((cudaError_t)rc));
    return RNDE_OK;
}

extern "C" int rnde_backward(rnde_handle* h, const float* du_dev, const float* dsaveval_dev, float* dp_dev, float* dx_dev, void* stream) {
    if (!du_dev) return RNDE_ERR_ARG;
    return backward_impl(h, du_dev, nullptr, dsaveval_dev, dp_dev, dx_dev, stream);
}

extern "C" int rnde_backward_saveat(rnde_handle* h, const float* du_dev, const float* dusave_dev, const float* dsaveval_dev, float* dp_dev,
                                    float* dx_dev, void* stream) {
    if (!dusave_dev) return RNDE_ERR_ARG;
    return backward_impl(h, du_dev, dusave_dev, dsaveval_dev, dp_dev, dx_dev, stream);
}

// ---- host-buffer (end-to-end) variants ------------------------------------------
static int ensure(rnde_handle* h, float** p, size_t n) {
    if (*p) return RNDE_OK;
    CUDA_TRY(h, cudaMalloc(p, sizeof(float) * n));
    return RNDE_OK;
}

extern "C" int rnde_forward_host(rnde_handle* h, const float* x_host, const float* p_host, float* u_out_host, float* saveval_host,
                                 rnde_stats* stats_host) {
    if (!h || !x_host || !p_host || !u_out_host) return RNDE_ERR_ARG;
    ON_HANDLE_DEVICE(h);
    const size_t n = (size_t)h->cfg.state_dim * h->cfg.batch;
    int rc;
    if ((rc = ensure(h, &h->hx, n)) || (rc = ensure(h, &h->hp, (size_t)h->np)) || (rc = ensure(h, &h->hu, n)) ||
        (rc = ensure(h, &h->hsv, (size_t)h->cfg.tape_capacity + 1))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->hx, x_host, sizeof(float) * n, cudaMemcpyHostToDevice, 0));
    CUDA_TRY(h, cudaMemcpyAsync(h->hp, p_host, sizeof(float) * h->np, cudaMemcpyHostToDevice, 0));
    rnde_stats st;
    rc = rnde_forward(h, h->hx, h->hp, h->hu, h->hsv, &st, 0);
    if (stats_host) *stats_host = st;
    if (rc != RNDE_OK) return rc;
    CUDA_TRY(h, cudaMemcpy(u_out_host, h->hu, sizeof(float) * n, cudaMemcpyDeviceToHost));
    if (saveval_host && st.n_saved > 0) CUDA_TRY(h, cudaMemcpy(saveval_host, h->hsv, sizeof(float) * st.n_saved, cudaMemcpyDeviceToHost));
    return RNDE_OK;
}

extern "C" int rnde_backward_host(rnde_handle* h, const float* du_host, const float* dsaveval_host, float* dp_host, float* dx_host) {
    if (!h || !du_host || !dp_host) return RNDE_ERR_ARG;
    ON_HANDLE_DEVICE(h);
    const size_t n = (size_t)h->cfg.state_dim * h->cfg.batch;
    int rc;
    if ((rc = ensure(h, &h->hdu, n)) || (rc = ensure(h, &h->hdsv, (size_t)h->cfg.tape_capacity + 1)) || (rc = ensure(h, &h->hdp, (size_t)h->np)) ||
        (rc = ensure(h, &h->hdx, n))) return rc;
    CUDA_TRY(h, cudaMemcpyAsync(h->hdu, du_host, sizeof(float) * n, cudaMemcpyHostToDevice, 0));
    const int nsv = h->last_stats.n_saved;
    if (dsaveval_host && nsv > 0) CUDA_TRY(h, cudaMemcpyAsync(h->hdsv, dsaveval_host, sizeof(float) * nsv, cudaMemcpyHostToDevice, 0));
    else CUDA_TRY(h, cudaMemsetAsync(h->hdsv, 0, sizeof(float) * ((size_t)h->cfg.tape_capacity + 1), 0));
    rc = rnde_backward(h, h->hdu, h->hdsv, h->hdp, h->hdx, 0);
    if (rc != RNDE_OK) return rc;
    CUDA_TRY(h, cudaMemcpy(dp_host, h->hdp, sizeof(float) * h->np, cudaMemcpyDeviceToHost));
    if (dx_host) CUDA_TRY(h, cudaMemcpy(dx_host, h->hdx, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return RNDE_OK;
}

extern "C" int rnde_allreduce_grads(rnde_handle* h, float* buf_dev, int64_t n, void* stream) {
    if (!h || !buf_dev || n <= 0) return RNDE_ERR_ARG;
    if (h->cfg.nranks <= 1) return RNDE_OK;
    if (h->cfg.dist_mode != RNDE_DIST_EXACT || !h->dist_ready || h->gcap == 0)
        return set_err(h, RNDE_ERR_STATE, "rnde_allreduce_grads needs a RNDE_DIST_EXACT handle after rnde_dist_import");
    if (n > h->gcap) return set_err(h, RNDE_ERR_ARG, "rnde_allreduce_grads: n exceeds num_params + 16384");
    ON_HANDLE_DEVICE(h);
    ArParams A; memset(&A, 0, sizeof(A));
    for (int i = 0; i < 8; ++i) A.peers[i] = h->peers[i];
    A.goff = h->goff_bytes; A.nranks = h->cfg.nranks; A.rank = h->cfg.rank; A.seq = ++h->ar_seq; A.n = n; A.gcap = h->gcap;
    const int blocks = (int)std::min<long long>(64, (n + 1023) / 1024);
    ar_push_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(A, buf_dev);
    ar_sum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(A, buf_dev);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 2;
    return RNDE_OK;
}

extern "C" int rnde_set_detach(rnde_handle* h, int32_t mode) {
    if (!h || mode < RNDE_DETACH_ALL || mode > RNDE_DETACH_FIRST_TERM_ONLY) return RNDE_ERR_ARG;
    h->detach = mode;
    h->have_tape = false;      // the forward tapes the initial-dt evaluation only in the all_but_first mode
    return RNDE_OK;
}

extern "C" int rnde_last_stats(rnde_handle* h, rnde_stats* out) {
    if (!h || !out) return RNDE_ERR_ARG;
    ON_HANDLE_DEVICE(h);
    if (h->ev_stats) CUDA_TRY(h, cudaEventSynchronize(h->ev_stats));
    const DevStats& s = *h->stats_pinned;
    out->nf = s.nf; out->naccept = s.naccept; out->nreject = s.nreject; out->n_saved = s.n_saved;
    out->retcode = s.retcode; out->t_final = s.t_final; out->dt_last = s.dt_last; out->dt_init = s.dt_init;
    h->last_stats = *out;
    return RNDE_OK;
}

extern "C" int rnde_reg_agg(rnde_handle* h, int32_t agg, float lam, float cot_scale, const float* saveval_dev, float* dsaveval_dev, float* reg_dev,
                            void* stream) {
    if (!h || agg < 0 || agg > 2 || !saveval_dev || !dsaveval_dev || !reg_dev) return RNDE_ERR_ARG;
    ON_HANDLE_DEVICE(h);
    reg_agg_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(h->stats, agg, lam, cot_scale, saveval_dev, dsaveval_dev, h->cfg.tape_capacity, reg_dev);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 1;
    return RNDE_OK;
}

extern "C" int rnde_head_loss_grad(rnde_handle* h, const float* u_dev, const float* p3_dev, const float* y_onehot_dev, int32_t n_classes,
                                   float loss_scale, float* loss_dev, float* logits_dev, float* du_dev, float* dp3_dev, void* stream) {
    if (!h || !u_dev || !p3_dev || !y_onehot_dev || !loss_dev || !du_dev || !dp3_dev || n_classes <= 0 || n_classes > 32) return RNDE_ERR_ARG;
    ON_HANDLE_DEVICE(h);
    const int D = h->cfg.state_dim, B = h->cfg.batch;
    if (!h->head_ws) CUDA_TRY(h, cudaMalloc(&h->head_ws, sizeof(float) * ((size_t)32 * B + B + 4)));
    int rc = launch_head(D, B, n_classes, u_dev, p3_dev, y_onehot_dev, loss_scale, loss_dev, logits_dev, du_dev, dp3_dev, h->head_ws,
                         (cudaStream_t)stream, &h->launches);
    if (rc != 0) return set_err(h, RNDE_ERR_CUDA, std::string("head launch: ") + cudaGetErrorString((cudaError_t)rc));
    return RNDE_OK;
}

extern "C" int rnde_opt_update(rnde_handle* h, float* p_dev, const float* g_dev, float* v_dev, int64_t n, float inv_decay_scale, float eta,
                               float rho, void* stream) {
    if (!p_dev || !g_dev || !v_dev || n < 0) return RNDE_ERR_ARG;
    if (n == 0) return RNDE_OK;   // update_parameters! skips empty parameter vectors (src/utils.jl:151)
    opt_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p_dev, g_dev, v_dev, (long long)n, inv_decay_scale, eta, rho);
    if (h) h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(h, RNDE_ERR_CUDA, std::string("opt_update: ") + cudaGetErrorString(e));
    return RNDE_OK;
}

extern "C" int rnde_adam_update(rnde_handle* h, float* p_dev, const float* g_dev, float* m_dev, float* v_dev, int64_t n, float eta, float beta1,
                                float beta2, float beta1_pow, float beta2_pow, float eps, float weight_decay, void* stream) {
    if (!p_dev || !g_dev || !m_dev || !v_dev || n < 0) return RNDE_ERR_ARG;
    if (n == 0) return RNDE_OK;
    adam_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p_dev, g_dev, m_dev, v_dev, (long long)n, eta, beta1, beta2,
                                                                                     beta1_pow, beta2_pow, eps, weight_decay);
    if (h) h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(h, RNDE_ERR_CUDA, std::string("adam_update: ") + cudaGetErrorString(e));
    return RNDE_OK;
}

// ---- test hooks (bit-level checks of the canonical device math against the CPU) -----------------
__global__ void canon_tanh_test_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = canon_tanhf(x[i]);
}
__global__ void canon_tanh_range_kernel(unsigned int first_bits, long long n, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = canon_tanhf(__uint_as_float(first_bits + (unsigned int)i));
}
__global__ void canon_pow_test_kernel(const float* __restrict__ x, float e, float* __restrict__ y, float* __restrict__ l10, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { y[i] = canon_powf(x[i], e); l10[i] = canon_log10f(x[i]); }
}
extern "C" int rnde_test_tanh(const float* x_dev, float* y_dev, int64_t n, void* stream) {
    canon_tanh_test_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_dev, y_dev, (long long)n);
    return cudaGetLastError() == cudaSuccess ? RNDE_OK : RNDE_ERR_CUDA;
}
extern "C" int rnde_test_tanh_bits(uint32_t first_bits, int64_t n, float* y_dev, void* stream) {
    canon_tanh_range_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(first_bits, (long long)n, y_dev);
    return cudaGetLastError() == cudaSuccess ? RNDE_OK : RNDE_ERR_CUDA;
}
__global__ void canon_unary_range_kernel(int fn, unsigned int first_bits, long long n, float* __restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = __uint_as_float(first_bits + (unsigned int)i);
    y[i] = fn == 0 ? canon_tanhf(x) : fn == 1 ? canon_sigmoidf(x) : fn == 2 ? canon_softplusf(x) : canon_expnegf(x);
}
extern "C" int rnde_test_unary_bits(int32_t fn, uint32_t first_bits, int64_t n, float* y_dev, void* stream) {
    if (fn < 0 || fn > 3 || n < 0 || !y_dev) return RNDE_ERR_ARG;
    if (n == 0) return RNDE_OK;
    canon_unary_range_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(fn, first_bits, (long long)n, y_dev);
    return cudaGetLastError() == cudaSuccess ? RNDE_OK : RNDE_ERR_CUDA;
}
extern "C" int rnde_test_csq_rhs(int32_t data_dim, int32_t hidden, int32_t extra, int32_t batch, const float* p_dev, const float* z_dev,
                                 const float* e_dev, float t, float* k_dev, void* stream) {
    if (data_dim <= 0 || hidden <= 0 || (extra != 1 && extra != 3) || batch <= 0 || !p_dev || !z_dev || !e_dev || !k_dev) return RNDE_ERR_ARG;
    const int D = data_dim + extra;
    const size_t smem = sizeof(float) * ((size_t)(2 * D + data_dim) * CSQ_TEST_NP + (size_t)csq_tile_floats(data_dim, hidden, CSQ_TEST_NP));
    if (smem > 200 * 1024) return RNDE_ERR_UNSUPPORTED;
    if (cudaFuncSetAttribute(csq_rhs_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return RNDE_ERR_CUDA; }
    csq_rhs_test_kernel<<<(batch + CSQ_TEST_NP - 1) / CSQ_TEST_NP, CSQ_TEST_NT, smem, (cudaStream_t)stream>>>(p_dev, data_dim, hidden, extra, batch, t, z_dev, e_dev, k_dev);
    return cudaGetLastError() == cudaSuccess ? RNDE_OK : RNDE_ERR_CUDA;
}
extern "C" int rnde_test_pow(const float* x_dev, float e, float* y_dev, float* l10_dev, int64_t n, void* stream) {
    canon_pow_test_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x_dev, e, y_dev, l10_dev, (long long)n);
    return cudaGetLastError() == cudaSuccess ? RNDE_OK : RNDE_ERR_CUDA;
}


// ---- Latent-ODE recognition RNN ------------------------------------------------------------------------
struct rnde_gru {
    rnde_gru_config cfg;
    GruOffsets off;
    int Q = 0, num_sms = 0, device = 0;
    size_t smem_fwd = 0, smem_bwd = 0;
    float* tapeA = nullptr; float* tapeD = nullptr; double* acc = nullptr;
    const float* last_p = nullptr;
    bool have_tape = false;
    int64_t launches = 0;
    std::string err;
};

extern "C" int64_t rnde_gru_num_params(const rnde_gru_config* c) {
    if (!c) return 0;
    return gru_offsets(c->in_dim, c->hidden_dim, c->latent_dim).np;
}
extern "C" const char* rnde_gru_last_error(const rnde_gru* g) { return g ? g->err.c_str() : "null handle"; }
extern "C" int64_t rnde_gru_launch_count(const rnde_gru* g) { return g ? g->launches : 0; }

extern "C" void rnde_gru_destroy(rnde_gru* g) {
    if (!g) return;
    DeviceScope scope(g->device);
    cudaFree(g->tapeA); cudaFree(g->tapeD); cudaFree(g->acc);
    delete g;
}

extern "C" int rnde_gru_create(const rnde_gru_config* cfg, rnde_gru** out) {
    if (!cfg || !out || cfg->struct_bytes != (int32_t)sizeof(rnde_gru_config)) return RNDE_ERR_ARG;
    *out = nullptr;
    if (cfg->in_dim <= 0 || cfg->hidden_dim <= 0 || cfg->latent_dim <= 0 || cfg->batch <= 0 || cfg->seq_len <= 0) return RNDE_ERR_ARG;
    if (rnde_device_count() <= 0) return RNDE_ERR_CUDA;
    rnde_gru* g = new rnde_gru();
    g->cfg = *cfg;
    const int I = cfg->in_dim, H = cfg->hidden_dim, L = cfg->latent_dim, C = 2 * L + 2 * I + 1;
    g->off = gru_offsets(I, H, L);
    g->Q = (cfg->batch + GRU_NP - 1) / GRU_NP;
    int dev = 0; cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { delete g; return RNDE_ERR_CUDA; }
    g->num_sms = prop.multiProcessorCount; g->device = dev;
    g->smem_fwd = sizeof(float) * ((size_t)round_up(g->off.np, 4) + (size_t)GRU_NP * (2 * C + 3 * H + 2 * L + 2 * L) + GRU_NP);
    g->smem_bwd = sizeof(float) * ((size_t)round_up(g->off.np, 4) + (size_t)GRU_NP * (2 * L * 4 + 2 * L + 3 * H));
    if (g->smem_fwd > prop.sharedMemPerBlockOptin || g->smem_bwd > prop.sharedMemPerBlockOptin) {
        fprintf(stderr, "regnde: GRU weights (%d floats) do not fit shared memory\n", g->off.np);
        delete g; return RNDE_ERR_UNSUPPORTED;
    }
    if (raise_smem_limit((const void*)gru_fwd_kernel, g->smem_fwd) != cudaSuccess ||
        raise_smem_limit((const void*)gru_bwd_kernel, g->smem_bwd) != cudaSuccess) { cudaGetLastError(); delete g; return RNDE_ERR_CUDA; }
    if (cfg->need_backward) {
        const size_t blocks = (size_t)cfg->seq_len * g->Q * GRU_NP;
        if (cudaMalloc(&g->tapeA, sizeof(float) * blocks * g->off.arows) != cudaSuccess ||
            cudaMalloc(&g->tapeD, sizeof(float) * blocks * g->off.drows) != cudaSuccess ||
            cudaMalloc(&g->acc, sizeof(double) * g->off.np) != cudaSuccess) { cudaGetLastError(); rnde_gru_destroy(g); return RNDE_ERR_CUDA; }
    }
    *out = g;
    return RNDE_OK;
}

static void gru_fill(const rnde_gru* g, GruParams& P) {
    memset(&P, 0, sizeof(P));
    const rnde_gru_config& c = g->cfg;
    P.I = c.in_dim; P.H = c.hidden_dim; P.L = c.latent_dim; P.X = 2 * c.in_dim + 1; P.C = 2 * c.latent_dim + P.X;
    P.T = c.seq_len; P.B = c.batch; P.Q = g->Q;
    P.tapeA = g->tapeA; P.tapeD = g->tapeD; P.arows = g->off.arows; P.drows = g->off.drows; P.need_tape = c.need_backward ? 1 : 0;
}

extern "C" int rnde_gru_forward(rnde_gru* g, const float* x_dev, const float* p_dev, float* out_dev, void* stream) {
    if (!g || !x_dev || !p_dev || !out_dev) return RNDE_ERR_ARG;
    DeviceScope scope(g->device);
    if (scope.err != cudaSuccess) { g->err = std::string("selecting the handle's device: ") + cudaGetErrorString(scope.err); return RNDE_ERR_CUDA; }
    GruParams P; gru_fill(g, P);
    P.x = x_dev; P.p = p_dev; P.out = out_dev;
    gru_fwd_kernel<<<g->Q, GRU_NT, g->smem_fwd, (cudaStream_t)stream>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g->err = std::string("gru_fwd_kernel: ") + cudaGetErrorString(e); return RNDE_ERR_CUDA; }
    g->launches += 1; g->last_p = p_dev; g->have_tape = g->cfg.need_backward != 0;
    return RNDE_OK;
}

extern "C" int rnde_gru_backward(rnde_gru* g, const float* dout_dev, float* dp_dev, void* stream) {
    if (!g || !dout_dev || !dp_dev) return RNDE_ERR_ARG;
    if (!g->have_tape) { g->err = "rnde_gru_backward needs a preceding rnde_gru_forward on a handle created with need_backward=1"; return RNDE_ERR_STATE; }
    DeviceScope scope(g->device);
    if (scope.err != cudaSuccess) { g->err = std::string("selecting the handle's device: ") + cudaGetErrorString(scope.err); return RNDE_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    GruParams P; gru_fill(g, P);
    P.p = g->last_p; P.dout = dout_dev;
    gru_bwd_kernel<<<g->Q, GRU_NT, g->smem_bwd, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g->err = std::string("gru_bwd_kernel: ") + cudaGetErrorString(e); return RNDE_ERR_CUDA; }
    g->launches += 1;
    const GruOffsets& O = g->off;
    const int H = P.H, L = P.L, C = P.C;
    WgDesc desc; memset(&desc, 0, sizeof(desc));
    desc.nl = 6;
    const int spec[6][5] = {   // delta row, input row, M, K, parameter offset
        {O.d_hu, O.a_yc, H, C, O.Wu1}, {O.d_u, O.a_hu, L, H, O.Wu2}, {O.d_hr, O.a_yc, H, C, O.Wr1},
        {O.d_r, O.a_hr, L, H, O.Wr2}, {O.d_hn, O.a_cc, H, C, O.Wn1}, {O.d_ns, O.a_hn, 2 * L, H, O.Wn2}};
    for (int l = 0; l < 6; ++l) {
        WgLayer& w = desc.l[l];
        w.dptr = g->tapeD; w.dstride = O.drows * GRU_NP; w.doff = spec[l][0];
        w.aptr = g->tapeA; w.astride = O.arows * GRU_NP; w.aoff = spec[l][1];
        w.M = spec[l][2]; w.K = spec[l][3]; w.poff = spec[l][4];
    }
    e = launch_dense_wgrad(desc, GRU_NP, (long long)P.T * g->Q, O.np, g->num_sms, g->acc, dp_dev, st, &g->launches);
    if (e != cudaSuccess) { g->err = std::string("gru wgrad: ") + cudaGetErrorString(e); return RNDE_ERR_CUDA; }
    return RNDE_OK;
}

// ---- Neural SDE (sde_kernel.cuh) -----------------------------------------------------------------------------------------
struct rnde_sde {
    rnde_sde_config cfg;
    int device = 0, NP = 4, Q = 0;
    size_t smem = 0;
    double* partial = nullptr; float* stacks = nullptr; unsigned* bar = nullptr; SdeStats* stats = nullptr; float* log = nullptr; float* saveval_int = nullptr;
    int log_cap = 4096;
    // reverse sweep (rnde_sde_enable_tape / rnde_sde_backward)
    float* tape = nullptr; float* tape_steps = nullptr; float* gpart = nullptr; int tape_cap = 0; size_t smem_bwd = 0;
    const float* last_p = nullptr; bool have_tape = false;
    int64_t launches = 0;
    std::string err;
};
static int sde_err(rnde_sde* s, int code, const std::string& msg) { if (s) s->err = msg; return code; }
static bool g_sri_init[64] = {false};

static void fill_sri(SriTableau& t, const double* v) {
    float* f = reinterpret_cast<float*>(&t);
    for (int i = 0; i < 44; ++i) f[i] = (float)v[i];
}
static int init_sri_tables(rnde_sde* s) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return RNDE_ERR_CUDA;
    std::lock_guard<std::mutex> lock(g_const_mutex);
    if (dev < 64 && g_sri_init[dev]) return RNDE_OK;
    // StochasticDiffEq constructSOSRI / constructSOSRI2 (the same Float64 literals as oracle/sde_oracle.py; order of SriTableau)
    static const double sosri[44] = {
        -0.04199224421316468, 2.842612915017106, -2.0527723684000727, 4.338237071435815, -2.8895936137439793, 2.3017575594644466,
        0.26204282091330466, 0.20903646383505375, -0.1502377115150942, 0.05836595312746999, 0.6149440396332373, 0.08535117634046772,
        -0.21641093549612528, 1.5336352863679572, 0.26066223492647056, -1.0536037558179159, 1.7015284721089472, -0.20725685784180017,
        -0.5119011827621657, 2.67767339866713, -4.9395031322250995, 0.15580956238299215, 3.2361551006624674, -1.4223118283355949,
        1.140099274172029, -0.6401334255743456, 0.4736296532772559, 0.026404498125060714,
        -1.8453464565104432, 2.688764531100726, -0.2523866501071323, 0.40896857551684956,
        0.4969658141589478, -0.5771202869753592, -0.12919702470322217, 0.2093514975196336,
        2.8453464565104425, -2.688764531100725, 0.2523866501071322, -0.40896857551684945,
        0.11522663875443433, -0.57877086147738, 0.2857851028163886, 0.17775911990655704};
    static const double sosri2[44] = {
        0.13804532298278663, 0.5818361298250374, 0.4181638701749618, 0.4670018408674211, 0.8046204792187386, -0.27162232008616016,
        0.45605532163856893, 0.7555807846451692, 0.24441921535482677, 0.6981181143266059, 0.3453277086024727, -0.04344582292908241,
        0.08852381537667678, 1.0317752458971061, 0.4563552922077882, 1.73078280444124, -0.46089678470929774, -0.9637509618944188,
        0.6753186815412179, -0.07452812525785148, -0.49783736486149366, -0.5591906709928903, 0.022696571806569924, -0.8984927888368557,
        -0.15036858140642623, 0.7545275856696072, 0.686995463807979, -0.2911544680711602,
        -0.45315689727309133, 0.8330937231303951, 0.3792843195533544, 0.24077885458934192,
        -0.4994383733810986, 0.9181786186154077, -0.25613778661003145, -0.16260245862427797,
        1.4531568972730915, -0.8330937231303933, -0.3792843195533583, -0.24077885458934023,
        -0.4976090683622265, 0.9148155835648892, -1.4102107084476505, 0.9930041932449877};
    SriTableau t[2];
    fill_sri(t[0], sosri); fill_sri(t[1], sosri2);
    if (cudaMemcpyToSymbol(c_SRI, t, sizeof(t)) != cudaSuccess) return RNDE_ERR_CUDA;
    if (dev < 64) g_sri_init[dev] = true;
    (void)s;
    return RNDE_OK;
}

extern "C" int64_t rnde_sde_num_params(const rnde_sde_config* c) {
    if (!c) return 0;
    const int64_t D = c->state_dim, H = c->hidden_dim;
    return H * D + H + D * H + D + D * D + D;
}
extern "C" const char* rnde_sde_last_error(const rnde_sde* s) { return s ? s->err.c_str() : "null handle"; }
extern "C" int64_t rnde_sde_launch_count(const rnde_sde* s) { return s ? s->launches : 0; }

template <int NP>
static int sde_capacity(size_t smem, int num_sms, int* out) {
    if (raise_smem_limit((const void*)sde_kernel<NP>, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sde_kernel<NP>, SDE_NT, smem) != cudaSuccess) { cudaGetLastError(); return 0; }
    *out = per_sm * num_sms;
    return 1;
}

extern "C" int rnde_sde_create(const rnde_sde_config* cfg, rnde_sde** out) {
    if (!cfg || !out) return RNDE_ERR_ARG;
    *out = nullptr;
    if (cfg->struct_bytes != (int32_t)sizeof(rnde_sde_config)) return RNDE_ERR_ARG;
    if (cfg->state_dim <= 0 || cfg->state_dim > 64 || cfg->hidden_dim <= 0 || cfg->hidden_dim > 256 || cfg->batch <= 0) return RNDE_ERR_ARG;
    if (cfg->alg < 0 || cfg->alg > 1 || !(cfg->t1 > cfg->t0) || !(cfg->abstol > 0.f) || !(cfg->reltol > 0.f)) return RNDE_ERR_ARG;
    if (cfg->reg_kind != RNDE_REG_NONE && cfg->reg_kind != RNDE_REG_ERR_DT && cfg->reg_kind != RNDE_REG_STIFF_SCALED) return RNDE_ERR_ARG;
    if (cfg->reg_kind == RNDE_REG_STIFF_SCALED && cfg->alg != RNDE_SDE_AUTO_SOSRI2) return RNDE_ERR_ARG;      // eigen_est exists for the composite algorithm only
    if (rnde_device_count() <= 0) return RNDE_ERR_CUDA;
    rnde_sde* s = new rnde_sde();
    s->cfg = *cfg;
    if (s->cfg.max_steps <= 0) s->cfg.max_steps = 100000;
    if (s->cfg.max_saved <= 0) s->cfg.max_saved = 1024;
    cudaDeviceProp prop;
    if (cudaGetDevice(&s->device) != cudaSuccess || cudaGetDeviceProperties(&prop, s->device) != cudaSuccess) { delete s; return RNDE_ERR_CUDA; }
    if (init_sri_tables(s) != RNDE_OK) { delete s; return RNDE_ERR_CUDA; }
    const int D = cfg->state_dim, H = cfg->hidden_dim, B = cfg->batch;
    const int np = (int)rnde_sde_num_params(cfg);
    // 4-column tiles while every CTA can be co-resident (the norm needs a grid barrier), else 8-column tiles
    int cap4 = 0, cap8 = 0, cap16 = 0;
    const size_t sm4 = sizeof(float) * sde_smem_floats(D, H, np, 4), sm8 = sizeof(float) * sde_smem_floats(D, H, np, 8), sm16 = sizeof(float) * sde_smem_floats(D, H, np, 16);
    const bool ok4 = sm4 <= prop.sharedMemPerBlockOptin && sde_capacity<4>(sm4, prop.multiProcessorCount, &cap4) && (B + 3) / 4 <= cap4;
    const bool ok8 = !ok4 && sm8 <= prop.sharedMemPerBlockOptin && sde_capacity<8>(sm8, prop.multiProcessorCount, &cap8) && (B + 7) / 8 <= cap8;
    const bool ok16 = !ok4 && !ok8 && sm16 <= prop.sharedMemPerBlockOptin && sde_capacity<16>(sm16, prop.multiProcessorCount, &cap16) && (B + 15) / 16 <= cap16;
    if (!ok4 && !ok8 && !ok16) { fprintf(stderr, "regnde: SDE batch %d exceeds the co-resident capacity (%d / %d / %d tiles of 4 / 8 / 16 columns)\n", B, cap4, cap8, cap16); delete s; return RNDE_ERR_UNSUPPORTED; }
    s->NP = ok4 ? 4 : (ok8 ? 8 : 16); s->smem = ok4 ? sm4 : (ok8 ? sm8 : sm16); s->Q = (B + s->NP - 1) / s->NP;
    auto fail = [&](const char* what) { s->err = what; rnde_sde_destroy(s); return RNDE_ERR_CUDA; };
    if (cudaMalloc(&s->partial, sizeof(double) * 2 * 3 * s->Q) != cudaSuccess) return fail("cudaMalloc partial");
    if (cudaMalloc(&s->stacks, sizeof(float) * (size_t)s->Q * 4 * SDE_MAXS * D * s->NP) != cudaSuccess) return fail("cudaMalloc stacks");
    if (cudaMalloc(&s->bar, sizeof(unsigned) * 4) != cudaSuccess) return fail("cudaMalloc bar");
    if (cudaMalloc(&s->stats, sizeof(SdeStats)) != cudaSuccess) return fail("cudaMalloc stats");
    if (cudaMalloc(&s->log, sizeof(float) * 3 * s->log_cap) != cudaSuccess) return fail("cudaMalloc log");
    if (cudaMalloc(&s->saveval_int, sizeof(float) * s->cfg.max_saved) != cudaSuccess) return fail("cudaMalloc saveval");
    *out = s;
    return RNDE_OK;
}

extern "C" void rnde_sde_destroy(rnde_sde* s) {
    if (!s) return;
    DeviceScope scope(s->device);
    cudaFree(s->partial); cudaFree(s->stacks); cudaFree(s->bar); cudaFree(s->stats); cudaFree(s->log); cudaFree(s->saveval_int);
    cudaFree(s->tape); cudaFree(s->tape_steps); cudaFree(s->gpart);
    delete s;
}

extern "C" int rnde_sde_forward(rnde_sde* s, const float* x_dev, const float* p_dev, const float* normals_dev, int32_t n_draws, float* u_out_dev,
                                float* saveval_dev, rnde_sde_stats* stats_host, void* stream) {
    if (!s || !x_dev || !p_dev || !normals_dev || n_draws <= 0 || !u_out_dev) return RNDE_ERR_ARG;
    DeviceScope scope(s->device);
    if (scope.err != cudaSuccess) return sde_err(s, RNDE_ERR_CUDA, "selecting the handle's device");
    cudaStream_t st = (cudaStream_t)stream;
    SdeParams P;
    memset(&P, 0, sizeof(P));
    const rnde_sde_config& c = s->cfg;
    P.D = c.state_dim; P.H = c.hidden_dim; P.B = c.batch; P.Q = s->Q; P.alg = c.alg; P.reg_kind = c.reg_kind; P.max_steps = c.max_steps;
    P.max_saved = c.max_saved; P.n_draws = n_draws; P.t0 = c.t0; P.t1 = c.t1; P.abstol = c.abstol; P.reltol = c.reltol;
    P.x = x_dev; P.p = p_dev; P.normals = normals_dev; P.u_out = u_out_dev; P.saveval = saveval_dev ? saveval_dev : s->saveval_int;
    P.partial = s->partial; P.stacks = s->stacks; P.bar = s->bar; P.stats = s->stats; P.log = s->log; P.log_cap = s->log_cap;
    P.tape = s->tape; P.tape_steps = s->tape_steps; P.tape_cap = s->tape_cap;
    s->last_p = p_dev; s->have_tape = s->tape != nullptr;
    if (cudaMemsetAsync(s->bar, 0, sizeof(unsigned) * 4, st) != cudaSuccess) return sde_err(s, RNDE_ERR_CUDA, "cudaMemsetAsync");
    if (s->NP == 4) sde_kernel<4><<<s->Q, SDE_NT, s->smem, st>>>(P);
    else if (s->NP == 8) sde_kernel<8><<<s->Q, SDE_NT, s->smem, st>>>(P);
    else sde_kernel<16><<<s->Q, SDE_NT, s->smem, st>>>(P);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sde_err(s, RNDE_ERR_CUDA, std::string("sde_kernel launch: ") + cudaGetErrorString(e));
    s->launches += 1;
    if (stats_host) {
        SdeStats h;
        if (cudaMemcpyAsync(&h, s->stats, sizeof(h), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
            return sde_err(s, RNDE_ERR_CUDA, std::string("sde stats: ") + cudaGetErrorString(cudaGetLastError()));
        stats_host->nfe1 = h.nfe1; stats_host->nfe2 = h.nfe2; stats_host->naccept = h.naccept; stats_host->nreject = h.nreject; stats_host->n_saved = h.n_saved;
        stats_host->draws = h.draws; stats_host->retcode = h.retcode; stats_host->reserved = 0;
        stats_host->t_final = h.t_final; stats_host->dt_init = h.dt_init; stats_host->dt_last = h.dt_last; stats_host->reserved2 = 0.f;
        if (h.retcode != RNDE_OK) return sde_err(s, h.retcode, h.retcode == RNDE_ERR_ARG ? "the supplied normals ran out" : rnde_status_string(h.retcode));
    }
    return RNDE_OK;
}

extern "C" int rnde_sde_enable_tape(rnde_sde* s, int32_t tape_capacity) {
    if (!s || tape_capacity <= 0) return RNDE_ERR_ARG;
    DeviceScope scope(s->device);
    if (scope.err != cudaSuccess) return sde_err(s, RNDE_ERR_CUDA, "selecting the handle's device");
    const rnde_sde_config& c = s->cfg;
    const int D = c.state_dim, H = c.hidden_dim;
    const int np = H * D + H + D * H + D + D * D + D;
    cudaFree(s->tape); cudaFree(s->tape_steps); cudaFree(s->gpart);
    s->tape = nullptr; s->tape_steps = nullptr; s->gpart = nullptr; s->tape_cap = 0; s->have_tape = false;
    const size_t smem = sizeof(float) * (size_t)sde_bwd_smem_floats(D, H, np, s->NP);
    int dev = 0, smem_limit = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_limit, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)smem_limit) return sde_err(s, RNDE_ERR_UNSUPPORTED, "reverse sweep: shared memory");
    const void* k = s->NP == 4 ? (const void*)sde_bwd_kernel<4> : (s->NP == 8 ? (const void*)sde_bwd_kernel<8> : (const void*)sde_bwd_kernel<16>);
    if (raise_smem_limit(k, smem) != cudaSuccess) { cudaGetLastError(); return sde_err(s, RNDE_ERR_CUDA, "cudaFuncSetAttribute(sde_bwd)"); }
    const size_t T = (size_t)D * s->NP;
    if (cudaMalloc(&s->tape, sizeof(float) * (size_t)tape_capacity * s->Q * 3 * T) != cudaSuccess ||
        cudaMalloc(&s->tape_steps, sizeof(float) * 4 * (size_t)tape_capacity) != cudaSuccess ||
        cudaMalloc(&s->gpart, sizeof(float) * (size_t)s->Q * np) != cudaSuccess) { cudaGetLastError(); return sde_err(s, RNDE_ERR_CUDA, "cudaMalloc (tape)"); }
    s->tape_cap = tape_capacity; s->smem_bwd = smem;
    return RNDE_OK;
}

extern "C" int rnde_sde_backward(rnde_sde* s, const float* du_dev, const float* dsaveval_dev, float* dp_dev, float* dx_dev, void* stream) {
    if (!s || !dp_dev || (!du_dev && !dsaveval_dev)) return RNDE_ERR_ARG;
    if (!s->have_tape) return sde_err(s, RNDE_ERR_STATE, "rnde_sde_backward needs rnde_sde_enable_tape and a forward solve on this handle first");
    DeviceScope scope(s->device);
    if (scope.err != cudaSuccess) return sde_err(s, RNDE_ERR_CUDA, "selecting the handle's device");
    cudaStream_t st = (cudaStream_t)stream;
    const rnde_sde_config& c = s->cfg;
    SdeBwdParams P; memset(&P, 0, sizeof(P));
    P.D = c.state_dim; P.H = c.hidden_dim; P.B = c.batch; P.Q = s->Q; P.alg = c.alg; P.reg_kind = c.reg_kind; P.tape_cap = s->tape_cap;
    P.abstol = c.abstol; P.reltol = c.reltol; P.p = s->last_p; P.tape = s->tape; P.tape_steps = s->tape_steps; P.stats = s->stats;
    P.du = du_dev; P.dsaveval = (c.reg_kind != RNDE_REG_NONE) ? dsaveval_dev : nullptr; P.dx = dx_dev; P.gpart = s->gpart;
    if (s->NP == 4) sde_bwd_kernel<4><<<s->Q, SDE_NT, s->smem_bwd, st>>>(P);
    else if (s->NP == 8) sde_bwd_kernel<8><<<s->Q, SDE_NT, s->smem_bwd, st>>>(P);
    else sde_bwd_kernel<16><<<s->Q, SDE_NT, s->smem_bwd, st>>>(P);
    const int np = c.hidden_dim * c.state_dim + c.hidden_dim + c.state_dim * c.hidden_dim + c.state_dim + c.state_dim * c.state_dim + c.state_dim;
    sde_grad_reduce_kernel<<<(np + 255) / 256, 256, 0, st>>>(s->gpart, s->Q, np, dp_dev);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return sde_err(s, RNDE_ERR_CUDA, std::string("sde_bwd_kernel launch: ") + cudaGetErrorString(e));
    s->launches += 2;
    return RNDE_OK;
}

extern "C" int rnde_sde_get_log(rnde_sde* s, float* log_host, int32_t cap) {
    if (!s || !log_host || cap <= 0 || cap > s->log_cap) return RNDE_ERR_ARG;
    DeviceScope scope(s->device);
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpy(log_host, s->log, sizeof(float) * 3 * cap, cudaMemcpyDeviceToHost) != cudaSuccess)
        return sde_err(s, RNDE_ERR_CUDA, "reading the attempt log");
    return RNDE_OK;
}
