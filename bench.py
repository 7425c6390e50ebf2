#!/usr/bin/env python
"""bench.py -- MNIST regularised neural-ODE TRAINING STEP throughput (BASELINE.json metric).

One "step" = one pass of the hot path over one synthetic batch, exactly the timed region of the
reference's training loop (/root/reference/experiments/mnist_node.jl:228-234):
    gs = Tracker.gradient((p1,p2,p3) -> loss_function(x, y, node, p1, p2, p3; λ), ps...)   # forward solve + loss + backward
    update_parameters!(ps, gs, opt)                                                          # InvDecay + Momentum
Workload (experiments/configs/mnist_node.yml shape): MLPDynamics(784,100), batch 512 per GPU, Tsit5,
reltol = abstol = 1.4e-8, error-estimate regulariser (type: error_est), agg = mean, λ = 100, Dense(784,10) head,
synthetic U[0,1) inputs and random one-hot labels, Glorot-uniform weights, seed 1999.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
    python bench.py --impl reference ...                     (CPU arm: the oracle port on the host cores)

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in HBM; `e2e` = the same
through the public API with pinned-host inputs copied H2D and the loss read back D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

D, H, NCLASS = 784, 100, 10
F_RHS = 2 * (H * (D + 1) + D * (H + 1))      # 315 368 FLOP per sample per field evaluation
SEED = 1999
LAMBDA = 1.0e2


def stepper_dram_traffic() -> float | None:
    """dram__bytes_read.sum + dram__bytes_write.sum of the forward stepper per launch, from the committed
    `ncu --set full` capture of the final round-1 build (profiles/r1x_fwd4_final_ncu_raw_metrics.json; r1h = earlier build)."""
    f = ROOT / "profiles" / "r1x_fwd4_final_ncu_raw_metrics.json"
    if not f.exists():
        f = ROOT / "profiles" / "r1h_steppers_ncu_raw_metrics.json"
    try:
        d = json.loads(f.read_text())
        k = next(v for n, v in d.items() if "fwd4" in n)
        tot = 0.0
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            val, unit = k[key]
            tot += float(val) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return tot
    except Exception:
        return None


def ffma_peak_tflops() -> tuple[float, str]:
    """FP32 FFMA peak of this pool's B200 (not in MEASURED_PEAKS.json): measured by tools/microbench.cu,
    committed in profiles/ffma_peak.json."""
    f = ROOT / "profiles" / "ffma_peak.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["ffma_tflops_sustained"]), "measured (profiles/ffma_peak.json, tools/microbench.cu)"
    return 72.0, "fallback (148 SM x 128 lanes x 2 x 1.9 GHz)"


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        self.gpu = gpu_index
        self.path = Path(f"/tmp/regnde_clocks_{os.getpid()}.csv")

    def start(self):
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.path.read_text().splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            self.path.unlink()
        except OSError:
            pass
        return out


def synth_batches(rng, nb, B):
    xs = [rng.random((D, B), dtype=np.float32) for _ in range(nb)]
    ys = []
    for _ in range(nb):
        lab = rng.integers(0, NCLASS, size=B)
        y = np.zeros((NCLASS, B), dtype=np.float32)
        y[lab, np.arange(B)] = 1.0
        ys.append(y)
    return xs, ys


def glorot(rng, out, inp):
    s = np.sqrt(6.0 / (inp + out))
    return rng.uniform(-s, s, size=(out, inp)).astype(np.float32)


def init_params(rng):
    W1, W2, W3 = glorot(rng, H, D + 1), glorot(rng, D, H + 1), glorot(rng, NCLASS, D)
    p2 = np.concatenate([W1.flatten(order="F"), np.zeros(H, np.float32), W2.flatten(order="F"), np.zeros(D, np.float32)])
    p3 = np.concatenate([W3.flatten(order="F"), np.zeros(NCLASS, np.float32)])
    return p2.astype(np.float32), p3.astype(np.float32)


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference itself needs Julia, absent from this image -- BASELINE.md 2)
# ----------------------------------------------------------------------------------------------
def cpu_train_step(orc, o, x, y, p2, p3, B):
    r = o.forward(x, p2)
    W3 = p3[: NCLASS * D].reshape(D, NCLASS).T
    b3 = p3[NCLASS * D:]
    logits = W3 @ r.u + b3[:, None]
    m = logits.max(axis=0, keepdims=True)
    lse = m + np.log(np.exp(logits - m).sum(axis=0, keepdims=True))
    ce = float(-(y * (logits - lse)).sum() / B)
    g = (np.exp(logits - lse) - y) / B
    du = (W3.T @ g).astype(np.float32)
    dsv = np.full(len(r.saveval), LAMBDA / max(len(r.saveval), 1), np.float32)
    dp2, _, _, _ = o.backward(du, dsv)
    dW3 = g @ r.u.T
    reg = LAMBDA * float(r.saveval.mean()) if len(r.saveval) else 0.0
    return ce + reg, dp2, np.concatenate([dW3.flatten(order="F"), g.sum(axis=1)]).astype(np.float32), r


def cpu_baseline(B: int, steps: int, warmup: int = 0, budget_s: float = 0.0):
    """Times `steps` CPU training steps after `warmup` untimed ones.  With budget_s > 0 the per-step sample (columns of the
    batch) is cut so that warmup + steps fit the budget, judged from one probe step on the full batch."""
    from oracle import orc
    orc.build()
    rng = np.random.default_rng(SEED)
    p2, p3 = init_params(rng)
    xs, ys = synth_batches(rng, 1, B)
    cores = os.cpu_count() or 1
    Bs = B
    if budget_s > 0:
        o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, kblock1=(D + 7) // 8, reg_kind=orc.REG_ERR_DT, nthreads=cores))
        tp = time.perf_counter()
        cpu_train_step(orc, o, xs[0], ys[0], p2, p3, B)
        probe = time.perf_counter() - tp
        need = probe * (steps + warmup)
        if need > budget_s:
            Bs = max(16, int(B * budget_s / need) // 16 * 16)
    x, y = np.ascontiguousarray(xs[0][:, :Bs]), np.ascontiguousarray(ys[0][:, :Bs])
    o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=Bs, kblock1=(D + 7) // 8, reg_kind=orc.REG_ERR_DT, nthreads=cores))
    for _ in range(warmup):
        cpu_train_step(orc, o, x, y, p2, p3, Bs)
    t0 = time.perf_counter()
    nf = 0
    for _ in range(steps):
        _, _, _, r = cpu_train_step(orc, o, x, y, p2, p3, Bs)
        nf = r.nf
    el = time.perf_counter() - t0
    what = f"the full {B}-sample batch" if Bs == B else f"the first {Bs} of the batch's {B} samples (cut to fit {budget_s:.0f} s)"
    return {"value": Bs * steps / el, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{steps} training step(s) of {what} (C oracle, OpenMP, forward+adjoint+head), nfe={nf}",
            "ms_per_step": 1e3 * el / steps}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch
    steps, warm = max(1, args.steps), max(0, args.warmup)
    cb = cpu_baseline(B, steps, warmup=warm, budget_s=150.0)      # exactly K timed steps after W warm-up steps, each a bounded sample
    line = {
        "impl": "reference", "metric": "mnist_reg_node_train_samples_per_sec", "value": cb["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"mnist_node reg(error_est) train step, MLPDynamics(784,100), batch {B}, Tsit5 tol 1.4e-8",
                   "note": "CPU restatement of the reference path (Julia toolchain unavailable); oracle/rnde_oracle.c on all host cores"},
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import regneuralde.jl_b200 as R
    from regneuralde.jl_b200 import _lib as L
    from regneuralde.jl_b200.parallel import average_gradients_, max_over_ranks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback in the product path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    B = args.batch
    rng = np.random.default_rng(SEED)                     # same weights on every rank
    p2_np, p3_np = init_params(rng)
    rng_data = np.random.default_rng(SEED + 1 + rank)     # each rank its own shard (weak scaling)
    NB = 8
    xs_np, ys_np = synth_batches(rng_data, NB, B)

    gen = torch.Generator().manual_seed(SEED)
    model = R.MLPDynamics(D, H, generator=gen)
    exact = world > 1 and args.dist == "exact"
    node = R.TrackedNeuralODE(model, [0.0, 1.0], True, True, R.Tsit5(), save_everystep=False, reltol=1.4e-8, abstol=1.4e-8,
                              save_start=False, tape_capacity=args.tape_capacity,
                              dist_mode=L.DIST_EXACT if exact else L.DIST_SINGLE, rank=rank if exact else 0, world=world if exact else 1)
    clf = R.ClassifierNODE(None, node, R.Dense(D, NCLASS, generator=gen))
    clf.p2.copy_(torch.from_numpy(p2_np)); clf.p3.copy_(torch.from_numpy(p3_np))
    node.p = clf.p2
    opt = R.Optimiser(1.0e-5, 0.1, 0.9)

    xs_dev = [torch.from_numpy(x).to(dev) for x in xs_np]            # (D,B) tensors resident in HBM
    ys_dev = [torch.from_numpy(y).to(dev) for y in ys_np]
    xs_pin = [torch.from_numpy(np.ascontiguousarray(x.T)).pin_memory() for x in xs_np]   # column-major D x B == row-major (B,D)
    ys_pin = [torch.from_numpy(np.ascontiguousarray(y.T)).pin_memory() for y in ys_np]
    x_stage = torch.empty(B, D, device=dev); y_stage = torch.empty(B, NCLASS, device=dev)
    loss_pin = torch.empty(1).pin_memory()

    gflat = None

    def train_step(x, y):
        if exact:      # global loss: CE mean over world*B samples, one shared regulariser; gradients are summed
            out = clf.loss_and_gradient(x, y, lam=LAMBDA, func=R.ERROR_ESTIMATE, agg="mean", ce_scale=1.0 / world)
            g2, g3 = out["g2"], out["g3"]
            node.allreduce_(g2, g3)        # one-shot push all-reduce over NVLink peer memory (rnde_allreduce_grads), no NCCL call
        else:
            out = clf.loss_and_gradient(x, y, lam=LAMBDA, func=R.ERROR_ESTIMATE, agg="mean")
            g2, g3 = out["g2"], out["g3"]
            average_gradients_([g2, g3], world)
        R.update_parameters_((clf.p1, clf.p2, clf.p3), (clf.p1, g2, g3), opt)
        return out

    def step_resident(i):
        return train_step(xs_dev[i % NB], ys_dev[i % NB])

    def step_e2e(i):
        x_stage.copy_(xs_pin[i % NB], non_blocking=True)
        y_stage.copy_(ys_pin[i % NB], non_blocking=True)
        out = train_step(x_stage.t(), y_stage.t())
        loss_pin.copy_(out["loss"].reshape(1), non_blocking=False)
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def launches():
        return node.launch_count()

    nfe_sum = [0]

    def timed(fn, K):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = launches()
        e0.record()
        last = None
        nfe_sum[0] = 0
        for i in range(K):
            last = fn(i)
            nfe_sum[0] += last["nfe"]
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        ms = max_over_ranks(ms, dev)
        return ms, last, launches() - l0

    W, K = max(args.warmup, 3), args.steps
    for i in range(W):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # both timed phases run the SAME K training steps: the state after warm-up is restored before the end-to-end phase,
    # so the step-size sequences (NFE drifts as the weights train) and hence the work are identical
    snap = (clf.p2.clone(), clf.p3.clone(), {k: (v["n"], v["v"].clone()) for k, v in opt.state.items()})
    ms, last, nlaunch = timed(step_resident, K)
    nfe_mean = nfe_sum[0] / K
    clf.p2.copy_(snap[0]); clf.p3.copy_(snap[1])
    for k, (n_upd, vel) in snap[2].items():
        opt.state[k]["n"] = n_upd; opt.state[k]["v"].copy_(vel)
    ms_e2e, last_e2e, _ = timed(step_e2e, K)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- roofline of the dominant kernel: the fused forward stepper, timed alone with CUDA events ----
    hd = node._handle(B, L.REG_ERR_DT, True)
    lib = hd.lib
    st = L.Stats()
    xb = R.colmajor(xs_dev[0]); ub = torch.empty(D * B, device=dev); svb = torch.zeros(args.tape_capacity + 1, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fw_ms, nfs = [], []
    for _ in range(5):
        torch.cuda.synchronize()
        e0.record()
        lib.rnde_forward(hd.h, xb.data_ptr(), clf.p2.data_ptr(), ub.data_ptr(), svb.data_ptr(), None, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        e1.record()
        torch.cuda.synchronize()
        lib.rnde_forward(hd.h, xb.data_ptr(), clf.p2.data_ptr(), ub.data_ptr(), svb.data_ptr(), C.byref(st), C.c_void_p(torch.cuda.current_stream().cuda_stream))
        fw_ms.append(e0.elapsed_time(e1)); nfs.append(st.nf)
    fwd_ms = statistics.median(fw_ms)
    nf = int(statistics.median(nfs))
    peak, peak_src = ffma_peak_tflops()
    achieved = nf * F_RHS * B / (fwd_ms * 1e-3) / 1e12
    variant = {1: "cta", 2: "stream", 3: "cluster8", 4: "cluster4"}.get(int(lib.rnde_kernel_variant(hd.h)), "?")
    kname = {"cluster4": "fwd4_kernel<100,98>", "cluster8": "fwd_kernel<8,32,4,true>", "cta": "fwd_kernel<1,32,4,true>", "stream": "fwd_kernel<1,4,1,false>"}.get(variant, "fwd_kernel")

    if rank == 0:
        total_samples = B * world * K
        value = total_samples / (ms * 1e-3)
        e2e_value = total_samples / (ms_e2e * 1e-3)
        h2d = (D * B + NCLASS * B) * 4
        cb = cpu_baseline(B, 2, warmup=1) if (world == 1 and not args.no_cpu_baseline) else None
        nacc = last["naccept"]
        flop_per_sample = (last["nfe"] + 2 + 12 * nacc) * F_RHS
        line = {
            "metric": "mnist_reg_node_train_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"mnist_node reg(error_est) train step, MLPDynamics(784,100), batch {B}/GPU, Tsit5 tol 1.4e-8, "
                                   "Dense(784,10) head, InvDecay+Momentum update",
                       "global_batch": B * world, "parallelism": (f"dp{world} " + ("reference-exact shared step sequence (in-kernel peer-memory norm exchange)" if exact else "independent-controller") + ", NCCL grad all-reduce") if world > 1 else "single",
                       "kernel_variant": variant, "nfe_per_step": last["nfe"], "nfe_mean": nfe_mean, "us_per_nfe": ms / K / nfe_mean * 1e3, "naccept": nacc, "nreject": last["nreject"],
                       "l2": "per-step tape working set (~3.4 MB x records) exceeds the 126 MB L2; inputs rotate over 8 resident batches",
                       "loss": float(last["loss"]), "flop_per_sample": flop_per_sample},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / K},
            "gpu_launches": int(nlaunch),
            "clocks": clocks,
            "roofline": {"bound": "fp32_ffma", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": stepper_dram_traffic(), "traffic_unit": "bytes/launch (ncu, tape writes; algorithmic = 3.416 MB x records)", "peak_source": peak_src, "kernel_ms": fwd_ms, "nfe": nf,
                         "flop_per_launch": nf * F_RHS * B,
                         "train_step_frac": flop_per_sample * B * world * K / (ms * 1e-3) / 1e12 / (peak * world)},
        }
        if cb is not None:
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="samples per GPU")
    ap.add_argument("--tape-capacity", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dist", default="independent", choices=["independent", "exact"],
                    help="N>1: independent step-size controllers per rank (default) or the reference-exact shared step sequence")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
