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

Both arms run the SAME work: the same initial weights (default_rng(1999)), the same 8 rotating batches of rank 0
(default_rng(2000)), the same canonical arithmetic (so the same step sequences), the optimiser update included, W warm-up
steps then K timed ones.  They print the same `config`.

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with inputs resident in HBM; `e2e` = the same through the
public API with pinned-host inputs copied H2D and the loss read back D2H inside the timed region.  At N > 1 `value` is the
REFERENCE-EXACT data-parallel mode (all ranks share the step sequence of the one batched solve); the independent-controller
mode and the strong-scaling split of one 512 batch are reported beside it (`independent`, `strong_scaling`).  `fixed_work`
repeats the timed steps with the controller replaying a recorded dt list (equal work per step whatever the weights).
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
NB = 8                                        # rotating batches
GAMMA, ETA, RHO = 1.0e-5, 0.1, 0.9            # Optimiser(InvDecay(1e-5), Momentum(0.1, 0.9)), mnist_node.jl:130
ARITH_SPLITK = 2                              # canonical arithmetic of the stepper both arms use (include/regnde.h)


def workload_config(B: int, world: int) -> dict:
    """The workload description both arms print (identical for identical flags)."""
    return {
        "workload": f"mnist_node reg(error_est) train step: MLPDynamics(784,100), batch {B}/GPU, Tsit5 reltol=abstol=1.4e-8, "
                    "error-estimate regulariser (agg mean, lambda 100), Dense(784,10) head + logitcrossentropy, InvDecay(1e-5)+Momentum(0.1,0.9) update",
        "global_batch": B * world, "per_gpu_batch": B, "n_ranks": world,
        "inputs": f"{NB} rotating synthetic batches per rank (numpy default_rng({SEED + 1}+rank)), U[0,1) pixels, random one-hot labels",
        "weights": f"Glorot-uniform from default_rng({SEED}), zero biases; trained through the warm-up and the timed steps",
        "arithmetic": "Float32, canonical order RNDE_ARITH_SPLITK (csrc/fwd4s_kernel.cuh == oracle arith 2)",
        "l2": "per-step tape working set (~3.4 MB x ~200 records) exceeds the 126 MB L2; inputs rotate over 8 batches",
    }


def stepper_dram_traffic() -> float | None:
    """dram__bytes_read.sum + dram__bytes_write.sum of the forward stepper per launch, from the newest committed
    `ncu --set full` capture under profiles/ (r2* = this round's fwd4s_kernel, r1x = round 1's fwd4_kernel)."""
    for name, pat in (("r2_fwd4s_ncu_raw_metrics.json", "fwd4s"), ("r1x_fwd4_final_ncu_raw_metrics.json", "fwd4")):
        f = ROOT / "profiles" / name
        try:
            d = json.loads(f.read_text())
            k = next(v for n, v in d.items() if pat in n)
            tot = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                val, unit = k[key]
                tot += float(val) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
            return tot
        except Exception:
            continue
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50",
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


def flop_per_sample(nfe: int, naccept: int) -> int:
    """forward nf evaluations, backward two GEMMs per taped evaluation (1 + 6 per accepted step)."""
    return (nfe + 2 + 12 * naccept) * F_RHS


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port (the reference itself needs Julia, absent from this image -- BASELINE.md 2)
# ----------------------------------------------------------------------------------------------
def cpu_loss_grad(o, x, y, p2, p3, B, hi=False):
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
    dp2, _, _, _ = o.backward(du, dsv, hi=hi)
    dW3 = g @ r.u.T
    reg = LAMBDA * float(r.saveval.mean()) if len(r.saveval) else 0.0
    return ce + reg, dp2, np.concatenate([dW3.flatten(order="F"), g.sum(axis=1)]).astype(np.float32), r


class CpuOptimiser:
    """Optimiser(InvDecay(gamma), Momentum(eta, rho)) on raw arrays (src/utils.jl:149-156; the arithmetic of rnde_opt_update)."""

    def __init__(self):
        self.n, self.v = {}, {}

    def update(self, key, p, g):
        n = self.n.get(key, 1)
        v = self.v.get(key, np.zeros_like(p))
        delta = g * np.float32(1.0 / (1.0 + GAMMA * n))
        v = np.float32(RHO) * v - np.float32(ETA) * delta
        p += v
        self.n[key], self.v[key] = n + 1, v


def cpu_run(B: int, steps: int, warmup: int, budget_s: float):
    """W warm-up + K timed CPU training steps on rank 0's batches, weights training like the GPU arm's.  When the full batch
    does not fit the time budget (judged from one probe step), every step takes the first Bs columns of its batch."""
    from oracle import orc
    orc.build()
    p2, p3 = init_params(np.random.default_rng(SEED))
    xs, ys = synth_batches(np.random.default_rng(SEED + 1), NB, B)
    cores = os.cpu_count() or 1
    mk = lambda b: orc.Oracle(orc.OracleConfig(D=D, H=H, B=b, kblock1=(D + 7) // 8, reg_kind=orc.REG_ERR_DT, nthreads=cores, arith=ARITH_SPLITK))
    Bs = B
    if budget_s > 0:
        o = mk(B)
        tp = time.perf_counter()
        cpu_loss_grad(o, xs[0], ys[0], p2, p3, B)
        probe = time.perf_counter() - tp
        need = probe * (steps + warmup)
        if need > budget_s:
            Bs = max(16, int(B * budget_s / need) // 16 * 16)
    o = mk(Bs)
    opt = CpuOptimiser()
    nfes = []

    def step(i):
        x, y = np.ascontiguousarray(xs[i % NB][:, :Bs]), np.ascontiguousarray(ys[i % NB][:, :Bs])
        _, g2, g3, r = cpu_loss_grad(o, x, y, p2, p3, Bs)
        opt.update("p2", p2, g2); opt.update("p3", p3, g3)
        return r

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        nfes.append(step(warmup + i).nf)
    el = time.perf_counter() - t0
    what = f"the full {B}-sample batches" if Bs == B else f"the first {Bs} of each batch's {B} samples (cut to fit {budget_s:.0f} s)"
    return {"value": Bs * steps / el, "unit": "samples/s", "cores": cores, "kind": "port",
            "sample": f"{steps} training step(s) after {warmup} warm-up on {what} (C oracle, OpenMP, forward + adjoint + head + optimiser update), "
                      f"nfe mean {statistics.mean(nfes):.1f}",
            "ms_per_step": 1e3 * el / steps, "nfe_mean": statistics.mean(nfes), "us_per_nfe": 1e6 * el / steps / statistics.mean(nfes)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = args.batch
    steps, warm = max(1, args.steps), max(0, args.warmup)
    cb = cpu_run(B, steps, warm, budget_s=150.0)      # exactly K timed steps after W warm-up steps, each a bounded sample
    line = {
        "impl": "reference", "metric": "mnist_reg_node_train_samples_per_sec", "value": cb["value"], "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, world),
        "note": "CPU restatement of the reference path (Julia toolchain unavailable): oracle/rnde_oracle.c on all host cores, rank 0's batches",
        "run": {"nfe_mean": cb["nfe_mean"], "us_per_nfe": cb["us_per_nfe"]},
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
    p2_np, p3_np = init_params(np.random.default_rng(SEED))                     # same weights on every rank
    xs_np, ys_np = synth_batches(np.random.default_rng(SEED + 1 + rank), NB, B)  # each rank its own shard (weak scaling)
    W, K = max(args.warmup, 3), max(1, args.steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Trainer:
        """One data-parallel mode: its node (handles, tape), parameters and optimiser state."""

        def __init__(self, mode: str, Bl: int, xs, ys):
            self.mode, self.B = mode, Bl
            exact = world > 1 and mode in ("exact", "strong")
            self.exact = exact
            gen = torch.Generator().manual_seed(SEED)
            self.node = R.TrackedNeuralODE(R.MLPDynamics(D, H, generator=gen), [0.0, 1.0], True, True, R.Tsit5(), save_everystep=False,
                                           reltol=1.4e-8, abstol=1.4e-8, save_start=False, tape_capacity=args.tape_capacity,
                                           dist_mode=L.DIST_EXACT if exact else L.DIST_SINGLE, rank=rank if exact else 0, world=world if exact else 1)
            self.clf = R.ClassifierNODE(None, self.node, R.Dense(D, NCLASS, generator=gen))
            self.clf.p2.copy_(torch.from_numpy(p2_np)); self.clf.p3.copy_(torch.from_numpy(p3_np))
            self.node.p = self.clf.p2
            self.opt = R.Optimiser(GAMMA, ETA, RHO)
            self.xs = [torch.from_numpy(np.ascontiguousarray(x[:, :Bl])).to(dev) for x in xs]            # (D,B) tensors resident in HBM
            self.ys = [torch.from_numpy(np.ascontiguousarray(y[:, :Bl])).to(dev) for y in ys]
            self.xs_pin = [torch.from_numpy(np.ascontiguousarray(x[:, :Bl].T)).pin_memory() for x in xs]   # column-major D x B == row-major (B,D)
            self.ys_pin = [torch.from_numpy(np.ascontiguousarray(y[:, :Bl].T)).pin_memory() for y in ys]
            self.x_stage = torch.empty(Bl, D, device=dev); self.y_stage = torch.empty(Bl, NCLASS, device=dev)
            self.loss_pin = torch.empty(1).pin_memory()

        def train_step(self, x, y):
            clf, node = self.clf, self.node
            if self.exact:      # global loss: CE mean over world*B samples, one shared regulariser; gradients are summed
                out = clf.loss_and_gradient(x, y, lam=LAMBDA, func=R.ERROR_ESTIMATE, agg="mean", ce_scale=1.0 / world)
                g2, g3 = out["g2"], out["g3"]
                node.allreduce_(g2, g3)        # one-shot push all-reduce over NVLink peer memory (rnde_allreduce_grads), no NCCL call
            else:
                out = clf.loss_and_gradient(x, y, lam=LAMBDA, func=R.ERROR_ESTIMATE, agg="mean")
                g2, g3 = out["g2"], out["g3"]
                average_gradients_([g2, g3], world)      # NCCL all-reduce (no-op at world 1)
            R.update_parameters_((clf.p1, clf.p2, clf.p3), (clf.p1, g2, g3), self.opt)
            return out

        def step_resident(self, i):
            return self.train_step(self.xs[i % NB], self.ys[i % NB])

        def step_e2e(self, i):
            self.x_stage.copy_(self.xs_pin[i % NB], non_blocking=True)
            self.y_stage.copy_(self.ys_pin[i % NB], non_blocking=True)
            out = self.train_step(self.x_stage.t(), self.y_stage.t())
            self.loss_pin.copy_(out["loss"].reshape(1), non_blocking=False)
            return out

        def snapshot(self):
            return (self.clf.p2.clone(), self.clf.p3.clone(), {k: (v["n"], v["v"].clone()) for k, v in self.opt.state.items()})

        def restore(self, snap):
            self.clf.p2.copy_(snap[0]); self.clf.p3.copy_(snap[1])
            for k, (n_upd, vel) in snap[2].items():
                self.opt.state[k]["n"] = n_upd; self.opt.state[k]["v"].copy_(vel)

        def timed(self, fn, first, n):
            """n steps fn(first), fn(first+1), ... bracketed by barrier + synchronize; device time, max over ranks."""
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = self.node.launch_count()
            e0.record()
            nfes, naccs, last = [], [], None
            for i in range(n):
                last = fn(first + i)
                nfes.append(last["nfe"]); naccs.append(last["naccept"])
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1), dev)
            flops = sum(flop_per_sample(a, b) for a, b in zip(nfes, naccs)) * self.B * world
            return {"ms": ms, "nfes": nfes, "naccs": naccs, "last": last, "launches": self.node.launch_count() - l0, "flops": flops}

        def handle(self):
            return self.node._handle(self.B, L.REG_ERR_DT, True)

    def summarise(tr, r, n):
        nfe_mean = statistics.mean(r["nfes"])
        return {"value": tr.B * world * n / (r["ms"] * 1e-3), "unit": "samples/s", "ms_per_step": r["ms"] / n, "nfe_mean": nfe_mean,
                "us_per_nfe": r["ms"] / n / nfe_mean * 1e3}

    peak, peak_src = ffma_peak_tflops()
    main_mode = "exact" if world > 1 else "single"
    tr = Trainer(main_mode, B, xs_np, ys_np)
    for i in range(W):
        tr.step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # every timed phase runs the SAME K training steps: the state after warm-up is restored before each of them, so the
    # step-size sequences (NFE drifts as the weights train) and hence the work are identical
    snap = tr.snapshot()
    res = tr.timed(tr.step_resident, W, K)
    tr.restore(snap)
    res_e2e = tr.timed(tr.step_e2e, W, K)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- fixed work: the controller replays the dt list recorded at the snapshot weights ----
    tr.restore(snap)
    hd = tr.handle()
    tr.step_resident(W)
    torch.cuda.synchronize()
    dts = tr.node.steps(B, L.REG_ERR_DT, True)[1]
    tr.restore(snap)
    arr = (C.c_float * len(dts))(*dts)
    hd.check(hd.lib.rnde_set_forced_steps(hd.h, arr, len(dts)), "rnde_set_forced_steps")
    res_fixed = tr.timed(tr.step_resident, W, K)
    hd.check(hd.lib.rnde_set_forced_steps(hd.h, None, 0), "rnde_set_forced_steps")
    tr.restore(snap)

    # ---- roofline of the dominant kernel: the fused forward stepper, timed alone with CUDA events (snapshot weights) ----
    lib = hd.lib
    st = L.Stats()
    xb = R.colmajor(tr.xs[0]); ub = torch.empty(D * B, device=dev); svb = torch.zeros(args.tape_capacity + 1, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fw_ms, nfs = [], []
    sptr = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for _ in range(7):
        barrier()
        e0.record()
        lib.rnde_forward(hd.h, xb.data_ptr(), tr.clf.p2.data_ptr(), ub.data_ptr(), svb.data_ptr(), None, sptr())
        e1.record()
        torch.cuda.synchronize()
        lib.rnde_last_stats(hd.h, C.byref(st))
        fw_ms.append(e0.elapsed_time(e1)); nfs.append(st.nf)
    fwd_ms = statistics.median(fw_ms)
    nf = int(statistics.median(nfs))
    achieved = nf * F_RHS * B / (fwd_ms * 1e-3) / 1e12
    variant = {1: "cta", 2: "stream", 3: "cluster8", 4: "cluster4"}.get(int(lib.rnde_kernel_variant(hd.h)), "?")
    arith = {0: "fma_chain", 1: "fixed24", 2: "splitk"}.get(tr.node.arith, "?")
    kname = {("cluster4", "splitk"): "fwd4s_kernel<100,784>", ("cluster4", "fma_chain"): "fwd4_kernel<100,98>", ("cluster8", "fma_chain"): "fwd_kernel<8,32,4,true>"}.get((variant, arith), "fwd_kernel")

    # ---- the other data-parallel modes (N > 1) ----
    extra = {}
    if world > 1:
        ti = Trainer("independent", B, xs_np, ys_np)
        for i in range(W):
            ti.step_resident(i)
        ri = ti.timed(ti.step_resident, W, K)
        extra["independent"] = dict(summarise(ti, ri, K), note="every rank its own step-size controller (a different numerical method from the one batched solve); NCCL gradient all-reduce")
        del ti
        if B % world == 0 and B // world >= 16:
            # strong scaling: the ONE 512 batch of rank 0's stream, rank r owns columns [r*B/R, (r+1)*B/R)
            xs0, ys0 = synth_batches(np.random.default_rng(SEED + 1), NB, B)
            Bl = B // world
            sl = slice(rank * Bl, (rank + 1) * Bl)
            ts = Trainer("strong", Bl, [x[:, sl] for x in xs0], [y[:, sl] for y in ys0])
            for i in range(W):
                ts.step_resident(i)
            rs = ts.timed(ts.step_resident, W, K)
            extra["strong_scaling"] = dict(summarise(ts, rs, K), global_batch=B, per_gpu_batch=Bl,
                                           note="reference-exact mode on the columns of one global batch of 512")
            del ts
            # what "reference-exact" claims, checked where the driver sees it: the same global batch of 512 solved once on rank 0
            # alone and by all ranks together (rank r: columns [r*B/R, (r+1)*B/R)), initial weights, one loss + gradient call each
            tc = Trainer("strong", Bl, [xs0[0][:, sl]], [ys0[0][:, sl]])
            oc = tc.clf.loss_and_gradient(tc.xs[0], tc.ys[0], lam=LAMBDA, func=R.ERROR_ESTIMATE, agg="mean", ce_scale=1.0 / world)
            tc.node.allreduce_(oc["g2"], oc["g3"])
            nfe_x, nacc_x = int(tc.node.last_stats.nf), int(tc.node.last_stats.naccept)
            lg = oc["logits"].t().contiguous()                          # (Bl, classes): this rank's columns
            lgs = [torch.empty_like(lg) for _ in range(world)]
            dist.all_gather(lgs, lg)
            if rank == 0:
                t1 = Trainer("single", B, [xs0[0]], [ys0[0]])
                o1 = t1.clf.loss_and_gradient(t1.xs[0], t1.ys[0], lam=LAMBDA, func=R.ERROR_ESTIMATE, agg="mean")
                rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
                extra["exact_check"] = {
                    "global_batch": B, "nfe": [nfe_x, int(t1.node.last_stats.nf)], "naccept": [nacc_x, int(t1.node.last_stats.naccept)],
                    "steps_identical": nfe_x == int(t1.node.last_stats.nf) and nacc_x == int(t1.node.last_stats.naccept),
                    "regulariser_bit_identical": bool(torch.equal(oc["reg"], o1["reg"])),
                    "logits_bit_identical": bool(torch.equal(torch.cat(lgs, 0), o1["logits"].t().contiguous())),
                    "grad_relerr_vs_single_solve": {"node": rel(oc["g2"], o1["g2"]), "head": rel(oc["g3"], o1["g3"])},
                    "note": "all ranks together (reference-exact mode, summed gradients) against rank 0 solving the 512 batch alone; the node gradient "
                            "differs by the Float32 summation noise of the regulariser part (grad_check.c_cpu32 is its size); "
                            "tools/dist_exact_check.py compares states and saved values bit by bit and the gradient with the oracle"}
                del t1
            del tc

    if rank == 0:
        main = summarise(tr, res, K)
        e2e = summarise(tr, res_e2e, K)
        fixed = summarise(tr, res_fixed, K)
        h2d = (D * B + NCLASS * B) * 4
        cfg = workload_config(B, world)
        par = "single" if world == 1 else f"dp{world} reference-exact: shared step sequence (in-kernel peer-memory norm exchange), peer-memory gradient all-reduce (rnde_allreduce_grads, no NCCL call)"
        line = {
            "metric": "mnist_reg_node_train_samples_per_sec", "value": main["value"], "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "run": {"parallelism": par, "kernel_variant": variant, "arith": arith, "nfe_mean": main["nfe_mean"], "us_per_nfe": main["us_per_nfe"],
                    "nfe_first": res["nfes"][0], "nfe_last": res["nfes"][-1], "naccept_last": res["naccs"][-1], "nreject_last": res["last"]["nreject"],
                    "loss_last": float(res["last"]["loss"]), "flop_per_sample_mean": res["flops"] / (B * world * K),
                    "backward": "tensor-core sweep (3-term BF16 cotangents)" if "RNDE_BWD_FFMA" not in os.environ else "FFMA sweep (RNDE_BWD_FFMA=1)"},
            "e2e": {"value": e2e["value"], "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": e2e["ms_per_step"]},
            "fixed_work": dict(fixed, n_forced_steps=len(dts), note="controller replays the dt list recorded at the post-warm-up weights: equal NFE every step"),
            "gpu_launches": int(res["launches"]),
            "clocks": clocks,
            "roofline": {"bound": "fp32_ffma", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak, "traffic": stepper_dram_traffic(),
                         "traffic_unit": "bytes/launch (ncu, tape writes; algorithmic = 3.416 MB x records)", "peak_source": peak_src,
                         "kernel_ms": fwd_ms, "nfe": nf, "flop_per_launch": nf * F_RHS * B,
                         "train_step_frac": res["flops"] / (res["ms"] * 1e-3) / 1e12 / (peak * world),
                         "train_step_frac_note": "sum over the timed steps of (nfe + 2 + 12 naccept) F_rhs B / their device time / FFMA peak"},
        }
        line.update(extra)
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_run(B, 2, 1, budget_s=0.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            line["grad_check"] = grad_check(tr, torch, R)
        if world == 1 and not args.no_secondary:
            line["secondary"] = secondary_workloads(torch, R, peak)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def secondary_workloads(torch, R, peak) -> dict:
    """The widened rows on the record: one Latent-ODE training step (N1: GRU encoder + chain-field solve with 49
    save times + decoder, PhysioNet-shaped synthetic batch 37 x 49 x 512, error-estimate regulariser), one Neural-SDE
    forward solve (N2: SOSRI, 32-dim state, batch 512, tol 1.4e-1, supplied noise) and one FFJORD training step (N4: MINIBOONE
    shape, batch 1024, forward + reverse sweep + WeightDecay/ADAM update); device times, CUDA events."""
    out = {}

    def timed(fn, iters=10, warm=3):
        tot = 0.0
        for it in range(iters + warm):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            if it >= warm:
                tot += e0.elapsed_time(e1)
        return tot / iters

    try:
        B, I, T = 512, 37, 49
        rng = np.random.default_rng(1234)
        times = np.unique(np.concatenate([[0.0], np.sort(rng.random(T - 2)), [1.0]]).astype(np.float32))
        S = len(times)
        data = rng.standard_normal((I, S, B)).astype(np.float32)
        mask = (rng.random((I, S, B)) < 0.15).astype(np.float32)
        trow = np.broadcast_to(times[None, :, None], (1, S, B)).astype(np.float32).copy()
        model = R.latent_ode_model(I, 40, 50, 20, 50, saveat=times.tolist(), regularize=True, solver=R.Tsit5(), generator=torch.Generator().manual_seed(5))
        ps = [p.clone().requires_grad_(True) for p in model.trainable()]
        d, m, t = (torch.from_numpy(a).cuda() for a in (data * mask, mask, trow))
        sample = torch.randn(20, B, device="cuda")
        info = {}

        def step():
            total, nfe, _ = R.loss_function(d, m, t, model, *ps, func=R.ERROR_ESTIMATE, regularize=True, lam_r=1.0e3, sample=sample)
            total.backward()
            info["nfe"] = nfe
            for p in ps:
                p.grad = None

        ms = timed(step)
        z0 = torch.randn(20, B, device="cuda")

        def fwd_only():
            with torch.no_grad():
                model.node(z0, ps[2].detach(), func=R.ERROR_ESTIMATE)

        ms_fwd = timed(fwd_only)
        f_rhs = 2 * 8 * 20 * 50        # FLOP per sample per evaluation of Chain(tanh, Dense(20,50,tanh), Dense(50,20,tanh), ... x8)
        out["latent_ode"] = {"workload": f"latent_ode train step, {I} features x {S} times x batch {B}, error_est", "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3),
                             "nfe": info["nfe"], "chain_solve_fwd_ms": ms_fwd, "chain_stepper_tflops": info["nfe"] * f_rhs * B / (ms_fwd * 1e-3) / 1e12,
                             "chain_stepper_frac_of_ffma_peak": info["nfe"] * f_rhs * B / (ms_fwd * 1e-3) / 1e12 / peak,
                             "note": "8 280 weights in shared memory, 128 CTAs x 4 columns: instruction-latency bound, not FFMA bound"}
    except Exception as ex:      # a secondary row must never take the headline down
        out["latent_ode"] = {"error": repr(ex)[:200]}
    try:
        D, H, B = 32, 64, 512
        rng = np.random.default_rng(SEED)
        node = R.TrackedNeuralDSDE(R.Chain(R.Dense(D, H, "tanh"), R.Dense(H, D)), R.Dense(D, D), [0.0, 1.0], True, R.SOSRI(), reltol=1.4e-1, abstol=1.4e-1)
        x = torch.from_numpy(rng.standard_normal((D, B)).astype(np.float32)).cuda()
        z = torch.randn(256, D, B, device="cuda")
        info = {}

        def solve():
            with torch.no_grad():
                _, n1, n2, _ = node(x, func=R.ERROR_ESTIMATE, noise=z)
            info["nfe"] = (n1, n2)

        ms = timed(solve)
        out["neural_sde"] = {"workload": f"mnist_nsde forward solve (SOSRI, {D}-dim state, batch {B}, tol 1.4e-1, supplied noise)", "ms_per_solve": ms,
                             "samples_per_s": B / (ms * 1e-3), "nfe1": info["nfe"][0], "nfe2": info["nfe"][1], "naccept": int(node.last_stats.naccept),
                             "nreject": int(node.last_stats.nreject)}
        # the training step of experiments/mnist_nsde.jl:89-110,191-204: Dense(784,32) -> SDE solve -> Dense(32,10), logitcrossentropy +
        # lam * mean(EEst dt), Tracker.gradient, Optimiser(InvDecay(1e-5), ADAM(0.01)); trajectories = 1 as in training
        gen = torch.Generator().manual_seed(SEED)
        clf = R.ClassifierNSDE(R.Dense(784, D, generator=gen), node, R.Dense(D, 10, generator=gen))
        xi = torch.rand(784, B, generator=gen).cuda()
        yi = torch.nn.functional.one_hot(torch.randint(0, 10, (B,), generator=gen), 10).T.float().cuda()
        opt = R.ADAMOptimiser(0.0, 0.01, inv_decay=1e-5)

        def train():
            o = clf.loss_and_gradient(xi, yi, lam=1.0e2, trajectories=1, func=R.ERROR_ESTIMATE, noise=z)
            R.update_parameters_((clf.p1, clf.p2, clf.p3), (o["g1"], o["g2"], o["g3"]), opt)

        out["neural_sde"]["train_ms_per_step"] = timed(train, 5, 2)
        out["neural_sde"]["train_samples_per_s"] = B / (out["neural_sde"]["train_ms_per_step"] * 1e-3)
        # algorithmic rate of the solve: drift Chain(Dense(D,H,tanh), Dense(H,D)) per f evaluation, diffusion Dense(D,D) per g evaluation
        fl = float((int(info["nfe"][0]) * 2 * (D * H + H * D) + int(info["nfe"][1]) * 2 * D * D) * B)
        out["neural_sde"]["solve_tflops"] = fl / (float(ms) * 1e-3) / 1e12
        out["neural_sde"]["solve_frac_of_ffma_peak"] = fl / (float(ms) * 1e-3) / 1e12 / float(peak)
        out["neural_sde"]["note"] = "5 248 weights in shared memory, 128 CTAs x 4 columns, 8 warps per SM: latency bound (profiles/r2zz_next_rows_ncu.txt)"
    except Exception as ex:
        out["neural_sde"] = {"error": repr(ex)[:200]}
    try:
        # ---- row N4: FFJORD training step (experiments/ffjord_tabular.jl:112-141: MLPDynamics(43, 100), batch 1024, Tsit5 1.4e-8,
        #      loss = -mean(logpx), Optimiser(WeightDecay(1e-5), ADAM(1e-2)); configs/ffjord_tabular.yml: regularize false) ----
        Dz, Hf, B = 43, 100, 1024
        g = torch.Generator().manual_seed(SEED)
        ff = R.TrackedFFJORD(R.CSQDynamics(Dz, Hf, generator=g), [0.0, 1.0], True, False, R.Tsit5(), reltol=1.4e-8, abstol=1.4e-8, tape_capacity=96)
        x = torch.randn(Dz, B, generator=g).cuda()
        e = torch.randn(Dz, B, generator=g).cuda()
        opt = R.ADAMOptimiser(1e-5, 1e-2)
        res = {}

        def train():
            o = R.ffjord.loss_and_gradient(ff, x, ff.p, e)
            R.update_parameters_((ff.p,), (o["g"],), opt)
            res.update(o)

        ms = timed(train, 5, 2)
        with torch.no_grad():
            ms_fwd = timed(lambda: ff(x, ff.p, e), 3, 1)
        out["ffjord"] = {"workload": f"ffjord_tabular train step (MINIBOONE shape: MLPDynamics({Dz},{Hf}), batch {B}, Tsit5 tol 1.4e-8, -mean(logpx), WeightDecay+ADAM)",
                         "ms_per_step": ms, "samples_per_s": B / (ms * 1e-3), "nfe": int(res["nfe"]), "nll": float(res["nll"]), "forward_ms": ms_fwd}
        # algorithmic rate of the forward solve: the three ConcatSquash layers and their e^T J product per evaluation
        fl = float(int(res["nfe"]) * 2 * 2 * (Dz * Hf + Hf * Hf + Hf * Dz) * B)
        out["ffjord"]["forward_tflops"] = fl / (float(ms_fwd) * 1e-3) / 1e12
        out["ffjord"]["forward_frac_of_ffma_peak"] = fl / (float(ms_fwd) * 1e-3) / 1e12 / float(peak)
        out["ffjord"]["note"] = "18.6 k weights in shared memory, 128 CTAs x 8 columns, 8 warps per SM: latency bound (profiles/r2zz_next_rows_ncu.txt)"
    except Exception as ex:      # secondary rows never break the headline line
        out["ffjord"] = {"error": repr(ex)[:200]}
    return out


def grad_check(tr, torch, R) -> dict:
    """Error of the device gradient of the flagship loss (batch 512, lambda 100) at the current weights against the
    Float64-cotangent adjoint over the same Float32 forward, beside the CPU Float32 adjoint's own error."""
    from oracle import orc
    B = tr.B
    p2 = tr.clf.p2.detach().cpu().numpy().copy(); p3 = tr.clf.p3.detach().cpu().numpy().copy()
    x, y = tr.xs[0], tr.ys[0]
    g2 = tr.clf.loss_and_gradient(x, y, lam=LAMBDA, func=R.ERROR_ESTIMATE, agg="mean")["g2"].cpu().numpy().copy()
    o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, kblock1=(D + 7) // 8, reg_kind=orc.REG_ERR_DT, arith=tr.node.arith))
    xn, yn = x.cpu().numpy(), y.cpu().numpy()
    _, hi, _, _ = cpu_loss_grad(o, xn, yn, p2, p3, B, hi=True)
    _, c32, _, _ = cpu_loss_grad(o, xn, yn, p2, p3, B, hi=False)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    return {"e_device": rel(g2, hi), "c_cpu32": rel(c32, hi), "what": "max-norm relative error of dL/dp2 vs the Float64-cotangent adjoint of the same Float32 forward"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="samples per GPU")
    ap.add_argument("--tape-capacity", type=int, default=128)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the Latent-ODE / Neural-SDE rows")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
