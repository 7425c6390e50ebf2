"""GPU parity tests: the CUDA path (through the C ABI / the host mirror of the Julia surface) against the
CPU oracle on identical inputs.  Shapes and kwargs follow the reference's test script (test/test_node.jl:4-89:
TDChain(Dense(3,10,tanh), Dense(11,2)), x 2xB, tspan [0,1], reltol = abstol = 1.4f-8, the three node variants)
and experiments/mnist_node.jl (MLPDynamics(784,100), batch 512, the three regulariser closures).

Bars (BASELINE.json north_star): accepted steps and NFE identical; trajectory / saved values / loss -- we get
BIT-identical Float32 (tolerance 0 is asserted); gradient relative error <= 1e-4 against the Float64-cotangent adjoint of the
same Float32 forward wherever a Float32 adjoint is that well conditioned (c_cpu32 <= 1e-4); where it is not (the
regulariser gradient cancels O(10) cotangents) the CUDA adjoint must be at most GRAD_BAR = 1.5 times
as far from that yardstick as the plain CPU Float32 adjoint is (tests/gradbar.py; on the toy shapes the CPU adjoint's
error is taken as its largest over 5 last-bit-equivalent cotangents, because a single draw of it moves by 2-3x)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import orc  # noqa: E402  (the checker)
from gradbar import GRAD_BAR, cpu32_noise  # noqa: E402


def R():
    import regneuralde.jl_b200 as r
    return r


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def make_node(D, H, act_out, regularize, solver, variant=0, kblock=0, cap=256):
    r = R()
    model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, "tanh" if act_out else None))
    return r.TrackedNeuralODE(model, [0.0, 1.0], True, regularize, solver, save_everystep=False, reltol=1.4e-8, abstol=1.4e-8,
                              save_start=False, kernel_variant=variant, kblock=kblock, tape_capacity=cap)


def oracle_cfg(D, H, B, act_out, alg, reg, kblock=0, arith=0):
    """arith: the canonical arithmetic the node resolved to (node.arith: FMA_CHAIN, or SPLITK on the cluster-4 stepper)."""
    kb = kblock if kblock else (D if D < 128 else (D + 7) // 8)
    return orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_TANH if act_out else orc.ACT_ID, alg=alg, reg_kind=reg, kblock1=kb, arith=arith)


FWD_CASES = [
    # name, D, H, B, act_out, auto, func, variant, kblock
    ("test_node unreg B=1", 2, 10, 1, 0, False, None, 0, 0),
    ("test_node errreg B=1", 2, 10, 1, 0, False, "ERROR_ESTIMATE", 0, 0),
    ("test_node stiffreg B=1", 2, 10, 1, 0, True, "STIFFNESS_ESTIMATE", 0, 0),
    ("toy B=7 ragged tile", 2, 10, 7, 0, False, "ERROR_ESTIMATE", 0, 0),
    ("toy B=512", 2, 10, 512, 0, False, "ERROR_ESTIMATE", 0, 0),
    ("toy B=33 streamed weights", 2, 10, 33, 0, False, "ERROR_ESTIMATE", 2, 0),
    ("latent-sized field K-blocked", 20, 50, 100, 1, True, "ERROR_PLUS_STIFFNESS", 0, 8),
    ("mnist B=32 cluster8", 784, 100, 32, 1, False, "ERROR_ESTIMATE", 3, 0),
    ("mnist B=40 cluster8 stiff_est", 784, 100, 40, 1, True, "STIFFNESS_SCALED", 3, 0),
    ("mnist B=17 cluster4 ragged", 784, 100, 17, 1, False, "ERROR_ESTIMATE", 4, 0),
    ("mnist B=48 cluster4 error_stiff_est", 784, 100, 48, 1, True, "ERROR_PLUS_STIFFNESS", 4, 0),
    ("mnist B=512 (auto variant)", 784, 100, 512, 1, False, "ERROR_ESTIMATE", 0, 0),
    ("mnist B=512 vanilla", 784, 100, 512, 1, False, None, 0, 0),
    ("D=200 H=37 cluster4 generic dims", 200, 37, 33, 0, True, "STIFFNESS_ESTIMATE", 4, 25),
]


@pytest.mark.parametrize("name,D,H,B,act_out,auto,func,variant,kblock", FWD_CASES, ids=[c[0] for c in FWD_CASES])
def test_forward_bit_identical(oracle_built, name, D, H, B, act_out, auto, func, variant, kblock):
    r = R()
    rng = np.random.default_rng(1999)        # seed of experiments/configs/mnist_node.yml:2
    p_np = orc.glorot_params(rng, D, H)
    x_np = rng.random((D, B), dtype=np.float32)
    regularize = func is not None
    solver = r.AutoTsit5() if auto else r.Tsit5()
    node = make_node(D, H, act_out, regularize, solver, variant, kblock)
    fobj = getattr(r, func) if func else None
    with torch.no_grad():
        res, nfe, sv = node(torch.from_numpy(x_np).cuda(), torch.from_numpy(p_np).cuda(), func=fobj)
    torch.cuda.synchronize()
    reg_kind = fobj.kind if fobj else orc.REG_NONE
    o = orc.Oracle(oracle_cfg(D, H, B, act_out, 1 if auto else 0, reg_kind, kblock, arith=node.arith))
    ref = o.forward(x_np, p_np)
    st = node.last_stats
    assert ref.retcode == 0 and st.retcode == 0
    assert (nfe, st.naccept, st.nreject) == (ref.nf, ref.naccept, ref.nreject)       # identical step count and NFE
    assert nfe == 3 + 6 * (st.naccept + st.nreject)
    assert np.array_equal(bits(res.cpu().numpy()), bits(ref.u)), "trajectory not bit-identical"
    if regularize:
        assert sv is not None and len(sv) == st.naccept + 1                            # SavingCallback semantics (A.7)
        assert np.array_equal(bits(sv.saveval.cpu().numpy()), bits(ref.saveval)), "saved values not bit-identical"
        assert abs(float(sv.t[-1]) - 1.0) < 1e-6 and float(sv.t[0]) == 0.0
    else:
        assert sv is None                                                               # neural_ode.jl:76
    t, dt, eest, eig = node.steps(B, reg_kind, False)
    steps = np.array(ref.steps)
    assert np.array_equal(bits(dt), bits(steps[:, 1])) and np.array_equal(bits(eest), bits(steps[:, 2]))


BWD_CASES = [
    # name, D, H, B, act_out, auto, func, variant, tol (None -> conditioning-aware only)
    ("test_node unreg grad", 2, 10, 5, 0, False, None, 0, 1e-4),
    ("test_node errreg grad", 2, 10, 5, 0, False, "ERROR_ESTIMATE", 0, None),
    ("test_node stiffreg grad", 2, 10, 5, 0, True, "STIFFNESS_ESTIMATE", 0, None),
    ("mid combined grad", 20, 50, 100, 1, True, "ERROR_PLUS_STIFFNESS", 0, None),
    ("mnist B=32 cluster8 grad", 784, 100, 32, 1, False, "ERROR_ESTIMATE", 3, 1e-4),
    ("mnist B=48 cluster4 grad", 784, 100, 48, 1, False, "ERROR_ESTIMATE", 4, 1e-4),
    ("mnist B=512 grad", 784, 100, 512, 1, False, "ERROR_ESTIMATE", 0, 1e-4),
    ("mnist B=512 stiff_est grad", 784, 100, 512, 1, True, "STIFFNESS_SCALED", 0, None),
    ("D=640 H=72 cluster4 (tensor-core sweep, generic dims)", 640, 72, 40, 1, True, "ERROR_PLUS_STIFFNESS", 4, None),
    ("D=800 H=96 cluster4 (tensor-core sweep, identity output)", 800, 96, 20, 0, False, None, 4, 1e-4),
]


@pytest.mark.parametrize("name,D,H,B,act_out,auto,func,variant,tol", BWD_CASES, ids=[c[0] for c in BWD_CASES])
def test_gradient_matches_oracle(oracle_built, name, D, H, B, act_out, auto, func, variant, tol):
    """Tracker.gradient(p -> sum(w .* res) + sum(ws .* sv.saveval), p) (test_node.jl:47-57 with random weights)."""
    r = R()
    rng = np.random.default_rng(7)
    p_np = orc.glorot_params(rng, D, H)
    x_np = rng.random((D, B), dtype=np.float32)
    regularize = func is not None
    fobj = getattr(r, func) if func else None
    node = make_node(D, H, act_out, regularize, r.AutoTsit5() if auto else r.Tsit5(), variant)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=fobj)
    reg_kind = fobj.kind if fobj else orc.REG_NONE
    o = orc.Oracle(oracle_cfg(D, H, B, act_out, 1 if auto else 0, reg_kind, arith=node.arith))
    ref = o.forward(x_np, p_np)
    assert np.array_equal(bits(res.detach().cpu().numpy()), bits(ref.u))
    w = rng.standard_normal((D, B)).astype(np.float32)
    ws = rng.standard_normal(max(len(ref.saveval), 1)).astype(np.float32)
    loss = (res * torch.from_numpy(w).cuda()).sum()
    if regularize:
        loss = loss + (sv.saveval * torch.from_numpy(ws[: len(ref.saveval)]).cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    dp_hi, dx_hi, _, _ = o.backward(w, ws, hi=True)       # FP64 cotangents over the FP32 forward: the yardstick
    dp_32, dx_32, _, _ = o.backward(w, ws)                 # plain CPU FP32 adjoint
    gp, gx = p.grad.cpu().numpy(), x.grad.cpu().numpy()
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    e_p, e_x = rel(gp, dp_hi), rel(gx, dx_hi)
    c_p, c_x = cpu32_noise(o, w, ws, n=5 if D * B <= 4096 else 1)
    if tol is not None:
        assert e_p <= tol and e_x <= tol, (e_p, e_x)
    assert e_x <= max(1e-4, GRAD_BAR * c_x), (e_x, c_x)
    # (until the end of round 2 the stiffness-estimate case needed an 8x bar here: the FFMA weight-gradient contraction kept a
    # Float32 partial per 64-entry stage, which mixes the cancelling records of a step; fixed in wgrad_kernel.cuh)
    assert e_p <= max(1e-4, GRAD_BAR * c_p), (e_p, c_p)


def test_mnist_training_step_against_oracle(oracle_built):
    """experiments/mnist_node.jl:132-152,229-232: loss = logitcrossentropy(model(x), y) + λ*mean(sv.saveval), gradient w.r.t.
    (p2, p3) through ClassifierNODE, at the full batch of 512."""
    r = R()
    D, H, B, Cn, lam = 784, 100, 512, 10, 100.0
    rng = np.random.default_rng(1999)
    p2 = orc.glorot_params(rng, D, H)
    s3 = np.sqrt(6.0 / (D + Cn))
    W3 = rng.uniform(-s3, s3, size=(Cn, D)).astype(np.float32)
    p3 = np.concatenate([W3.flatten(order="F"), np.zeros(Cn, np.float32)])
    x = rng.random((D, B), dtype=np.float32)
    lab = rng.integers(0, Cn, B)
    y = np.zeros((Cn, B), np.float32); y[lab, np.arange(B)] = 1
    node = make_node(D, H, 1, True, r.Tsit5())
    clf = r.ClassifierNODE(None, node, r.Dense(D, Cn))
    clf.p2.copy_(torch.from_numpy(p2)); clf.p3.copy_(torch.from_numpy(p3)); node.p = clf.p2
    out = clf.loss_and_gradient(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), lam=lam, func=r.ERROR_ESTIMATE, agg="mean")
    torch.cuda.synchronize()
    o = orc.Oracle(oracle_cfg(D, H, B, 1, 0, orc.REG_ERR_DT, arith=node.arith))
    ref = o.forward(x, p2)
    logits = W3 @ ref.u
    m = logits.max(0, keepdims=True)
    lse = m + np.log(np.exp(logits - m).sum(0, keepdims=True))
    ce = -(y * (logits - lse)).sum() / B
    reg = lam * ref.saveval.astype(np.float64).mean()
    assert (out["nfe"], out["naccept"]) == (ref.nf, ref.naccept)
    assert abs(float(out["ce"]) - ce) <= 1e-5 * abs(ce)
    assert abs(float(out["reg"]) - reg) <= 1e-5 * abs(reg)            # regulariser value: rel. err <= 1e-5
    assert abs(float(out["loss"]) - (ce + reg)) <= 1e-5 * abs(ce + reg)
    g = (np.exp(logits - lse) - y) / B
    du = (W3.T @ g).astype(np.float32)
    dsv = np.full(len(ref.saveval), lam / len(ref.saveval), np.float32)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    dp2_hi, _, _, _ = o.backward(du, dsv, hi=True)      # FP64 cotangents over the FP32 forward
    dp2_32, _, _, _ = o.backward(du, dsv)                # plain CPU FP32 adjoint
    dW3 = g @ ref.u.T
    e2 = rel(out["g2"].cpu().numpy(), dp2_hi)
    e3 = rel(out["g3"].cpu().numpy()[: Cn * D], dW3.flatten(order="F"))
    assert e3 <= 1e-4, e3
    # The regulariser part of dL/dp2 is ill-conditioned in ANY FP32 adjoint at reltol 1.4e-8 (the CPU FP32 adjoint
    # itself is ~0.5% off on this loss, ~8% on the regulariser part alone: DESIGN.md "gradient conditioning").
    assert e2 <= max(1e-4, GRAD_BAR * rel(dp2_32, dp2_hi)), (e2, rel(dp2_32, dp2_hi))
    # the cross-entropy part alone is well conditioned: <= 1e-4 (measured ~1e-6)
    out0 = clf.loss_and_gradient(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), lam=0.0, func=r.ERROR_ESTIMATE, agg="mean")
    dp2_ce, _, _, _ = o.backward(du, 0 * dsv, hi=True)
    assert rel(out0["g2"].cpu().numpy(), dp2_ce) <= 1e-4
    out = clf.loss_and_gradient(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda(), lam=lam, func=r.ERROR_ESTIMATE, agg="mean")
    # the autograd route (node(...) -> logits -> torch loss) gives the same gradient as the fused route
    pp2 = clf.p2.clone().requires_grad_(True); pp3 = clf.p3.clone().requires_grad_(True)
    lg, nfe, sv = clf(torch.from_numpy(x).cuda(), None, pp2, pp3, func=r.ERROR_ESTIMATE)
    l2 = -(torch.from_numpy(y).cuda() * torch.log_softmax(lg, dim=0)).sum() / B + lam * sv.saveval.mean()
    l2.backward()
    assert abs(float(l2) - float(out["loss"])) <= 1e-5 * abs(float(l2))
    # two FP32 evaluations with cotangents that differ in the last bit: equal up to the adjoint's rounding noise
    assert rel(pp2.grad.cpu().numpy(), dp2_hi) <= max(1e-4, GRAD_BAR * rel(dp2_32, dp2_hi))


SAVEAT_CASES = [
    # name, D, H, B, act_out, auto, func, variant, saveat
    ("toy unreg, t0 and t1 saved", 2, 10, 5, 0, False, None, 0, [0.0, 0.1, 0.25, 0.5, 0.75, 1.0]),
    ("toy errreg, interior only", 2, 10, 7, 0, False, "ERROR_ESTIMATE", 0, [0.05, 0.3, 0.31, 0.32, 0.9]),
    ("latent-sized stiffreg, 40 irregular times", 20, 50, 100, 1, True, "ERROR_PLUS_STIFFNESS", 0, None),
    ("mnist-sized B=32 cluster8", 784, 100, 32, 1, False, "ERROR_ESTIMATE", 3, [0.0, 0.2, 0.6, 1.0]),
    ("toy streamed weights", 2, 10, 33, 0, False, "ERROR_ESTIMATE", 2, [0.5, 1.0]),
]


@pytest.mark.parametrize("name,D,H,B,act_out,auto,func,variant,saveat", SAVEAT_CASES, ids=[c[0] for c in SAVEAT_CASES])
def test_saveat_multi_save_functors(oracle_built, name, D, H, B, act_out, auto, func, variant, saveat):
    """The {false,true} / {true,true} functors (neural_ode.jl:79-108,146-180; call site time_series.jl:51): res is
    feat x nsave x batch, saved states bit-identical to the oracle's dense output, gradient of a weighted sum of all
    saved states (+ saved regulariser values) against the oracle adjoint."""
    r = R()
    rng = np.random.default_rng(42)
    if saveat is None:      # irregular observation times like the PhysioNet grid (latent_ode.jl:137)
        saveat = np.unique(np.round(np.sort(rng.random(40)), 3)).astype(np.float32).tolist()
    sa32 = np.asarray(saveat, np.float32)
    p_np = orc.glorot_params(rng, D, H)
    x_np = rng.random((D, B), dtype=np.float32)
    regularize = func is not None
    solver = r.AutoTsit5() if auto else r.Tsit5()
    model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, "tanh" if act_out else None))
    node = r.TrackedNeuralODE(model, [0.0, 1.0], True, regularize, solver, saveat=sa32.tolist(), reltol=1.4e-8, abstol=1.4e-8,
                              kernel_variant=variant)
    fobj = getattr(r, func) if func else None
    reg_kind = fobj.kind if fobj else orc.REG_NONE
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=fobj)
    cfg = oracle_cfg(D, H, B, act_out, 1 if auto else 0, reg_kind, arith=node.arith)
    cfg.saveat = sa32.astype(np.float64)
    o = orc.Oracle(cfg)
    ref = o.forward(x_np, p_np)
    assert tuple(res.shape) == (D, len(sa32), B)                       # feat x nsave x batch
    assert (nfe, node.last_stats.naccept) == (ref.nf, ref.naccept)     # saveat adds no tstops: same steps as without it
    got = res.detach().permute(1, 0, 2).cpu().numpy()                  # -> (nsave, D, B) like the oracle
    assert ref.usave.shape == got.shape
    assert np.array_equal(bits(got), bits(ref.usave)), "saved states not bit-identical"
    if sa32[0] == 0.0:
        assert np.array_equal(bits(got[0]), bits(x_np))                # save_start: tspan[1] in saveat
    if regularize:
        assert np.array_equal(bits(sv.saveval.detach().cpu().numpy()), bits(ref.saveval))
    w = rng.standard_normal(ref.usave.shape).astype(np.float32)
    ws = rng.standard_normal(len(ref.saveval) if regularize else 1).astype(np.float32)
    loss = (res * torch.from_numpy(np.ascontiguousarray(w.transpose(1, 0, 2))).cuda()).sum()
    if regularize:
        loss = loss + (sv.saveval * torch.from_numpy(ws).cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    zeros = np.zeros((D, B), np.float32)
    dp_hi, dx_hi, _, _ = o.backward(zeros, ws if regularize else None, hi=True, dusave=w)
    dp_32, dx_32, _, _ = o.backward(zeros, ws if regularize else None, dusave=w)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    e_p, e_x = rel(p.grad.cpu().numpy(), dp_hi), rel(x.grad.cpu().numpy(), dx_hi)
    c_p, c_x = cpu32_noise(o, zeros, ws if regularize else None, n=5 if D * B <= 4096 else 1, dusave=w)
    assert e_p <= max(1e-4, GRAD_BAR * c_p) and e_x <= max(1e-4, GRAD_BAR * c_x), (e_p, c_p, e_x, c_x)
    if not regularize:
        assert e_p <= 1e-4 and e_x <= 1e-4, (e_p, e_x)
    # per-call saveat keyword (update_saveat!): a different grid on the same node, then the stored grid is back
    with torch.no_grad():
        res2, _, _ = node(x.detach(), p.detach(), func=fobj, saveat=[0.5])
        res3, _, _ = node(x.detach(), p.detach(), func=fobj)
    assert tuple(res2.shape) == (D, 1, B) and tuple(res3.shape) == tuple(res.shape)
    assert torch.equal(res3, res.detach())


LATENT_WIDTHS = (50, 20, 50, 20, 50, 20, 50, 20)       # gen_dynamics of experiments/latent_ode.jl:109-121

CHAIN_CASES = [
    # name, D, widths, acts, pre_act, B, auto, func, variant, n_saveat
    ("latent ODE field B=512, 49 save times, unregularised", 20, LATENT_WIDTHS, (1,) * 8, 1, 512, False, None, 0, 49),
    ("latent ODE field B=512, error_est", 20, LATENT_WIDTHS, (1,) * 8, 1, 512, False, "ERROR_ESTIMATE", 0, 49),
    ("latent ODE field B=130 ragged, error_stiff_est", 20, LATENT_WIDTHS, (1,) * 8, 1, 130, True, "ERROR_PLUS_STIFFNESS", 0, 49),
    ("3-layer chain, identity output, no pre-activation, 32-column tiles", 6, (9, 11, 6), (1, 1, 0), 0, 70, False, "ERROR_ESTIMATE", 1, 5),
    ("1-layer chain, final state only", 4, (4,), (1,), 1, 9, True, "STIFFNESS_ESTIMATE", 0, 0),
]


@pytest.mark.parametrize("name,D,widths,acts,pre_act,B,auto,func,variant,nsave", CHAIN_CASES, ids=[c[0] for c in CHAIN_CASES])
def test_chain_field_latent_ode(oracle_built, name, D, widths, acts, pre_act, B, auto, func, variant, nsave):
    """TrackedNeuralODE(gen_dynamics, [0,1], false, REGULARIZE, solver, saveat = saveat, reltol = abstol = 1.4f-8)
    (experiments/latent_ode.jl:137-147) called as in time_series.jl:51: forward bit-identical to the oracle
    (steps, NFE, every saved state, saved regulariser values), gradient against the oracle adjoint."""
    r = R()
    rng = np.random.default_rng(1234)
    p_np = orc.glorot_chain_params(rng, D, widths, bias_scale=0.05)
    x_np = rng.standard_normal((D, B)).astype(np.float32)           # z0 = sample * exp(logvar/2) + mu
    saveat = None
    if nsave:
        saveat = np.unique(np.concatenate([[0.0], np.sort(rng.random(nsave - 2)), [1.0]]).astype(np.float32))
    regularize = func is not None
    solver = r.AutoTsit5() if auto else r.Tsit5()
    layers = []
    K = D
    for M, a in zip(widths, acts):
        layers.append(r.Dense(K, M, "tanh" if a else None)); K = M
    model = r.Chain(*(["tanh"] if pre_act else []), *layers)
    kw = dict(saveat=saveat.tolist()) if saveat is not None else {}
    node = r.TrackedNeuralODE(model, [0.0, 1.0], False, regularize, solver, reltol=1.4e-8, abstol=1.4e-8, kernel_variant=variant, **kw)
    fobj = getattr(r, func) if func else None
    reg_kind = fobj.kind if fobj else orc.REG_NONE
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=fobj)
    cfg = orc.OracleConfig(D=D, H=max(widths), B=B, alg=1 if auto else 0, reg_kind=reg_kind, kblock1=D, widths=widths, acts=acts,
                           pre_act=pre_act, saveat=None if saveat is None else saveat.astype(np.float64))
    o = orc.Oracle(cfg)
    ref = o.forward(x_np, p_np)
    assert (nfe, node.last_stats.naccept, node.last_stats.nreject) == (ref.nf, ref.naccept, ref.nreject)
    if saveat is not None:
        got = res.detach().permute(1, 0, 2).cpu().numpy()
        assert got.shape == ref.usave.shape == (len(saveat), D, B)
        assert np.array_equal(bits(got), bits(ref.usave)), "saved states not bit-identical"
        w = rng.standard_normal(ref.usave.shape).astype(np.float32)
        loss = (res * torch.from_numpy(np.ascontiguousarray(w.transpose(1, 0, 2))).cuda()).sum()
        bw = dict(du=np.zeros((D, B), np.float32), dusave=w)
    else:
        assert np.array_equal(bits(res.detach().cpu().numpy()), bits(ref.u)), "final state not bit-identical"
        w = rng.standard_normal((D, B)).astype(np.float32)
        loss = (res * torch.from_numpy(w).cuda()).sum()
        bw = dict(du=w, dusave=None)
    ws = None
    if regularize:
        assert np.array_equal(bits(sv.saveval.detach().cpu().numpy()), bits(ref.saveval))
        ws = rng.standard_normal(len(ref.saveval)).astype(np.float32)
        loss = loss + (sv.saveval * torch.from_numpy(ws).cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    dp_hi, dx_hi, _, _ = o.backward(bw["du"], ws, hi=True, dusave=bw["dusave"])
    dp_32, dx_32, _, _ = o.backward(bw["du"], ws, dusave=bw["dusave"])
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    e_p, e_x = rel(p.grad.cpu().numpy(), dp_hi), rel(x.grad.cpu().numpy(), dx_hi)
    c_p, c_x = cpu32_noise(o, bw["du"], ws, n=5 if D * B <= 4096 else 1, dusave=bw["dusave"])
    assert e_p <= max(1e-4, GRAD_BAR * c_p) and e_x <= max(1e-4, GRAD_BAR * c_x), (e_p, c_p, e_x, c_x)
    if not regularize:
        assert e_p <= 1e-4 and e_x <= 1e-4, (e_p, e_x)


_SWEEP_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import regneuralde.jl_b200 as r
from oracle import orc
rng = np.random.default_rng(21)
D, H, B = 784, 100, 80
p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, B), dtype=np.float32)
w = rng.standard_normal((D, B)).astype(np.float32)
node = r.TrackedNeuralODE(r.MLPDynamics(D, H), [0.0, 1.0], True, True, r.AutoTsit5(), tape_capacity=64)
x = torch.from_numpy(x_np).cuda().requires_grad_(True); p = torch.from_numpy(p_np).cuda().requires_grad_(True)
res, nfe, sv = node(x, p, func=r.ERROR_PLUS_STIFFNESS)
ws = rng.standard_normal(len(sv)).astype(np.float32)
((res * torch.from_numpy(w).cuda()).sum() + (sv.saveval * torch.from_numpy(ws).cuda()).sum()).backward()
np.savez(sys.argv[1], dp=p.grad.cpu().numpy(), dx=x.grad.cpu().numpy(), w=w, ws=ws, x=x_np, p=p_np, nfe=nfe, arith=node.arith)
"""


def test_tensor_core_sweep_agrees_with_ffma_sweep(oracle_built, tmp_path):
    """bwd4tc_kernel (tcgen05, BF16 hi+lo split) against bwd4_kernel (FFMA, RNDE_BWD_FFMA=1) and the oracle on the same
    tape: MNIST-shaped field, ragged batch (80 = 5 clusters), AutoTsit5 with the combined regulariser (all cotangent paths)."""
    import os, subprocess, sys
    outs = {}
    for name, env in (("tc", {}), ("ffma", {"RNDE_BWD_FFMA": "1"})):
        f = tmp_path / f"{name}.npz"
        e = dict(os.environ); e.update(env)
        r_ = subprocess.run([sys.executable, "-c", _SWEEP_SCRIPT, str(f)], env=e, capture_output=True, text=True, cwd=os.path.dirname(os.path.dirname(__file__)))
        assert r_.returncode == 0, r_.stderr[-2000:]
        outs[name] = np.load(f)
    a, b = outs["tc"], outs["ffma"]
    assert int(a["nfe"]) == int(b["nfe"])
    rel = lambda u, v: np.abs(u - v).max() / np.abs(v).max()
    D, H, B = 784, 100, 80
    o = orc.Oracle(oracle_cfg(D, H, B, 1, 1, orc.REG_ERR_PLUS_STIFF, arith=int(a["arith"])))
    o.forward(a["x"], a["p"])
    dp_hi, dx_hi, _, _ = o.backward(a["w"], a["ws"], hi=True)
    dp_32, dx_32, _, _ = o.backward(a["w"], a["ws"])
    for name, g in outs.items():
        e_p, e_x = rel(g["dp"], dp_hi), rel(g["dx"], dx_hi)
        assert e_p <= max(1e-4, GRAD_BAR * rel(dp_32, dp_hi)) and e_x <= max(1e-4, GRAD_BAR * rel(dx_32, dx_hi)), (name, e_p, e_x)
    # the two sweeps see the same tape: they differ only by the rounding of their products
    assert rel(a["dx"], b["dx"]) <= max(1e-4, GRAD_BAR * rel(dx_32, dx_hi)) and rel(a["dp"], b["dp"]) <= max(1e-4, GRAD_BAR * rel(dp_32, dp_hi))


def test_flagship_gradient_both_sweeps(oracle_built):
    """The flagship training loss (experiments/mnist_node.jl:132-152: logitcrossentropy + 100 * mean(sv.saveval), batch 512)
    with BOTH reverse sweeps -- tensor cores (default) and FFMA (RNDE_BWD_FFMA=1) -- and both regularisers of the experiment:
    e = error of dL/dp2 against the Float64-cotangent adjoint of the same Float32 forward, c = the CPU Float32 adjoint's own.
    Bars: cross-entropy part <= 1e-4 (well conditioned); full gradient <= max(1e-4, 1.5 c).  The numbers are printed (-s)."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for sweep, env in (("tensor", {}), ("ffma", {"RNDE_BWD_FFMA": "1"})):
        e = dict(os.environ); e.update(env)
        r_ = subprocess.run([sys.executable, os.path.join(root, "tools", "grad_err.py"), "512"], env=e, capture_output=True, text=True, cwd=root)
        assert r_.returncode == 0, r_.stderr[-2000:]
        for line in r_.stdout.strip().splitlines()[-2:]:
            d = json.loads(line)
            print(f"flagship gradient [{d['func']}, {sweep} sweep]: e = {d['full']['e']:.2e}, c_cpu32 = {d['full']['c_cpu32']:.2e}; "
                  f"CE part e = {d['ce']['e']:.2e}; regulariser part e = {d['reg']['e']:.2e}, c_cpu32 = {d['reg']['c_cpu32']:.2e}")
            assert d["sweep"] == sweep
            assert d["ce"]["e"] <= 1e-4, d
            assert d["full"]["e"] <= max(1e-4, GRAD_BAR * d["full"]["c_cpu32"]), d
            assert d["reg"]["e"] <= max(1e-4, GRAD_BAR * d["reg"]["c_cpu32"]), d


FIXED24_CASES = [
    # name, D, H, B, act_out, auto, func
    ("mnist B=17 ragged", 784, 100, 17, 1, False, "ERROR_ESTIMATE"),
    ("mnist B=48 AutoTsit5 error_stiff_est", 784, 100, 48, 1, True, "ERROR_PLUS_STIFFNESS"),
    ("mnist B=512 vanilla", 784, 100, 512, 1, False, None),
    ("D=640 H=72 identity output", 640, 72, 40, 0, True, "STIFFNESS_ESTIMATE"),
]


@pytest.mark.parametrize("name,D,H,B,act_out,auto,func", FIXED24_CASES, ids=[c[0] for c in FIXED24_CASES])
def test_fixed24_tensor_core_forward_bit_identical(oracle_built, name, D, H, B, act_out, auto, func):
    """RNDE_ARITH_FIXED24: the layer products as exact integer tensor-core MMAs (csrc/fwd4x_kernel.cuh) against the oracle's
    arith = 1 -- states, saved values, step sizes and NFE bit for bit; gradient (tensor-core sweep over the same tape) against
    the oracle adjoint of the same arithmetic."""
    r = R()
    from regneuralde.jl_b200 import _lib as L
    rng = np.random.default_rng(1999)
    p_np = orc.glorot_params(rng, D, H)
    x_np = rng.random((D, B), dtype=np.float32)
    regularize = func is not None
    model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, "tanh" if act_out else None))
    node = r.TrackedNeuralODE(model, [0.0, 1.0], True, regularize, r.AutoTsit5() if auto else r.Tsit5(), reltol=1.4e-8, abstol=1.4e-8,
                              arith=L.ARITH_FIXED24, tape_capacity=96)
    fobj = getattr(r, func) if func else None
    x = torch.from_numpy(x_np).cuda().requires_grad_(True); p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=fobj)
    cfg = oracle_cfg(D, H, B, act_out, 1 if auto else 0, fobj.kind if fobj else orc.REG_NONE, arith=node.arith)
    cfg.arith = 1
    o = orc.Oracle(cfg)
    ref = o.forward(x_np, p_np)
    st = node.last_stats
    assert (nfe, st.naccept, st.nreject) == (ref.nf, ref.naccept, ref.nreject)
    assert np.array_equal(bits(res.detach().cpu().numpy()), bits(ref.u)), "trajectory not bit-identical"
    if regularize:
        assert np.array_equal(bits(sv.saveval.detach().cpu().numpy()), bits(ref.saveval)), "saved values not bit-identical"
    w = rng.standard_normal((D, B)).astype(np.float32)
    ws = rng.standard_normal(max(len(ref.saveval), 1)).astype(np.float32)
    loss = (res * torch.from_numpy(w).cuda()).sum()
    if regularize:
        loss = loss + (sv.saveval * torch.from_numpy(ws[: len(ref.saveval)]).cuda()).sum()
    loss.backward()
    dp_hi, dx_hi, _, _ = o.backward(w, ws, hi=True)
    dp_32, dx_32, _, _ = o.backward(w, ws)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    e_p, e_x = rel(p.grad.cpu().numpy(), dp_hi), rel(x.grad.cpu().numpy(), dx_hi)
    # unit-size random cotangents on every saved value are far harsher than the training loss (lambda/n each); the sweep's products
    # carry 16 mantissa bits (BF16 hi+lo), so on this ill-conditioned part it may sit at ~10x the CPU FP32 adjoint's own error
    assert e_p <= max(1e-4, GRAD_BAR * rel(dp_32, dp_hi)) and e_x <= max(1e-4, GRAD_BAR * rel(dx_32, dx_hi)), (e_p, e_x)
    if not regularize:
        assert e_p <= 1e-4 and e_x <= 1e-4


def test_backward_through_an_overwritten_tape_is_refused():
    """The tape lives in the handle (one per batch size): a second forward overwrites it, so differentiating the first
    result afterwards must fail loudly instead of returning gradients of the wrong solve."""
    r = R()
    node = make_node(2, 10, 0, False, r.Tsit5())
    p = r.track(node.p)
    x = torch.rand(2, 3, device="cuda")
    u1, _, _ = node(x, p)
    u2, _, _ = node(2 * x, p)
    with pytest.raises(RuntimeError, match="overwritten"):
        u1.sum().backward()
    u2.sum().backward()
    assert p.grad is not None and torch.isfinite(p.grad).all()
    with pytest.raises(ValueError, match="Float32"):
        node(x, p.detach().double())


def test_solution_object(oracle_built):
    """solution(n, x, p; solver, tspan, saveat) (neural_ode.jl:182-210): same solve without the callback, solver override."""
    r = R()
    rng = np.random.default_rng(3)
    D, H, B = 2, 10, 6
    p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, B), dtype=np.float32)
    node = make_node(D, H, 0, True, r.Tsit5())
    x = torch.from_numpy(x_np).cuda(); p = torch.from_numpy(p_np).cuda()
    sol = r.solution(node, x, p)
    ref = orc.Oracle(oracle_cfg(D, H, B, 0, 0, orc.REG_NONE, arith=node.arith)).forward(x_np, p_np)
    assert sol.retcode == "Success" and len(sol) == 1 and float(sol.t[-1]) == 1.0
    assert (sol.destats.nf, sol.destats.naccept, sol.destats.nreject) == (ref.nf, ref.naccept, ref.nreject)
    assert np.array_equal(bits(sol[0].cpu().numpy()), bits(ref.u))
    assert len(sol.step_t) == ref.naccept + 1 and abs(float(sol.step_t[-1]) - 1.0) < 1e-6
    sol2 = r.solution(node, x, p, solver=r.AutoTsit5(), tspan=[0.0, 0.5], saveat=[0.0, 0.25, 0.5])
    cfg = oracle_cfg(D, H, B, 0, 1, orc.REG_NONE, arith=node.arith); cfg.t1 = 0.5; cfg.saveat = np.array([0.0, 0.25, 0.5])
    ref2 = orc.Oracle(cfg).forward(x_np, p_np)
    assert len(sol2) == 3 and sol2.destats.nf == ref2.nf
    assert np.array_equal(bits(torch.stack(sol2.u).cpu().numpy()), bits(ref2.usave))


def test_update_parameters_matches_flux_momentum():
    """Optimiser(InvDecay(1e-5), Momentum(0.1, 0.9)) on raw arrays, empty parameter vectors skipped (src/utils.jl:149-156)."""
    r = R()
    rng = np.random.default_rng(0)
    p = torch.from_numpy(rng.standard_normal(1000).astype(np.float32)).cuda()
    ref = p.cpu().numpy().astype(np.float64); v = np.zeros(1000)
    opt = r.Optimiser(1e-5, 0.1, 0.9)
    empty = torch.zeros(0, device="cuda")
    for n in range(1, 6):
        g = torch.from_numpy(rng.standard_normal(1000).astype(np.float32)).cuda()
        r.update_parameters_((empty, p), (empty, g), opt)
        v = 0.9 * v - 0.1 * (g.cpu().numpy() / (1 + 1e-5 * n)); ref = ref + v
    assert np.abs(p.cpu().numpy() - ref).max() < 1e-5


def test_failure_codes_and_host_api(oracle_built):
    """retcodes surface as errors (never a silent fallback); the *_host entry points copy in/out themselves."""
    r = R()
    from regneuralde.jl_b200 import _lib as L
    rng = np.random.default_rng(3)
    D, H, B = 2, 10, 4
    p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, B), dtype=np.float32)
    node = make_node(D, H, 0, True, r.Tsit5(), cap=5)           # tape too small for ~24 steps
    with pytest.raises(r.RndeError) as ei:
        node(torch.from_numpy(x_np).cuda(), torch.from_numpy(p_np).cuda().requires_grad_(True))
    assert ei.value.code == L.ERR_TAPE_FULL
    bad = p_np.copy(); bad[0] = np.nan
    node = make_node(D, H, 0, False, r.Tsit5())
    with pytest.raises(r.RndeError) as ei:
        node(torch.from_numpy(x_np).cuda(), torch.from_numpy(bad).cuda())
    assert ei.value.code == L.ERR_NAN
    # host-buffer path == oracle
    lib = L.lib()
    cfg = L.Config(); cfg.struct_bytes = C.sizeof(L.Config)
    cfg.state_dim, cfg.hidden_dim, cfg.batch, cfg.act_hidden, cfg.act_out, cfg.time_dep = D, H, B, 1, 0, 1
    cfg.reg_kind, cfg.need_backward, cfg.t0, cfg.t1 = L.REG_ERR_DT, 1, 0.0, 1.0
    cfg.abstol = cfg.reltol = float(np.float32(1.4e-8))
    h = C.c_void_p()
    assert lib.rnde_create(C.byref(cfg), C.byref(h)) == 0
    xh = np.ascontiguousarray(x_np.T); u = np.zeros_like(xh); sv = np.zeros(300, np.float32); st = L.Stats()
    assert lib.rnde_forward_host(h, xh.ctypes.data, p_np.ctypes.data, u.ctypes.data, sv.ctypes.data, C.byref(st)) == 0
    ref = orc.Oracle(oracle_cfg(D, H, B, 0, 0, orc.REG_ERR_DT, arith=0)).forward(x_np, p_np)
    assert np.array_equal(bits(u.T), bits(ref.u)) and st.nf == ref.nf
    du = np.ones_like(xh); dsv = np.ones(300, np.float32); dp = np.zeros_like(p_np); dx = np.zeros_like(xh)
    assert lib.rnde_backward_host(h, du.ctypes.data, dsv.ctypes.data, dp.ctypes.data, dx.ctypes.data) == 0
    assert np.isfinite(dp).all() and np.abs(dp).max() > 0
    lib.rnde_destroy(h)
    # backward without a taped forward is refused
    node2 = make_node(D, H, 0, False, r.Tsit5())
    hd = node2._handle(B, 0, False)
    z = torch.zeros(D * B, device="cuda"); g = torch.zeros(64, device="cuda")
    assert lib.rnde_backward(hd.h, z.data_ptr(), None, g.data_ptr(), None, None) == L.ERR_STATE


def test_rejected_steps_path(oracle_built):
    """A stiff-ish field forces rejections: nreject > 0 and everything still bit-identical."""
    r = R()
    rng = np.random.default_rng(2)
    D, H, B = 4, 8, 6
    p_np = orc.glorot_params(rng, D, H) * np.float32(30.0)
    x_np = rng.random((D, B), dtype=np.float32)
    model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, "tanh"))
    node = r.TrackedNeuralODE(model, [0.0, 1.0], True, True, r.Tsit5(), reltol=1e-6, abstol=1e-6)
    with torch.no_grad():
        res, nfe, sv = node(torch.from_numpy(x_np).cuda(), torch.from_numpy(p_np).cuda())
    cfg = orc.OracleConfig(D=D, H=H, B=B, reg_kind=orc.REG_ERR_DT, abstol=float(np.float32(1e-6)), reltol=float(np.float32(1e-6)))
    ref = orc.Oracle(cfg).forward(x_np, p_np)
    assert ref.nreject > 0 and node.last_stats.nreject == ref.nreject and nfe == ref.nf
    assert np.array_equal(bits(res.cpu().numpy()), bits(ref.u))
    assert np.array_equal(bits(sv.saveval.cpu().numpy()), bits(ref.saveval))


def _exact_worker(rank, world, port, q):
    import os
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import regneuralde.jl_b200 as r
    from regneuralde.jl_b200 import _lib as L
    D, H, Bg = 784, 100, 96
    Bl = Bg // world
    rng = np.random.default_rng(1999)
    p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, Bg), dtype=np.float32)
    node = r.TrackedNeuralODE(r.MLPDynamics(D, H), [0.0, 1.0], True, True, r.Tsit5(), reltol=1.4e-8, abstol=1.4e-8,
                              dist_mode=L.DIST_EXACT, rank=rank, world=world)
    with torch.no_grad():
        res, nfe, sv = node(torch.from_numpy(np.ascontiguousarray(x_np[:, rank * Bl:(rank + 1) * Bl])).cuda(), torch.from_numpy(p_np).cuda())
    # the Latent-ODE solver half (chain field + saveat) in the same mode
    Dc, Bc = 20, 64
    rngc = np.random.default_rng(77)
    pc = orc.glorot_chain_params(rngc, Dc, LATENT_WIDTHS, bias_scale=0.05); xc = rngc.standard_normal((Dc, Bc)).astype(np.float32)
    layers, K = [], Dc
    for M in LATENT_WIDTHS:
        layers.append(r.Dense(K, M, "tanh")); K = M
    nodec = r.TrackedNeuralODE(r.Chain("tanh", *layers), [0.0, 1.0], False, True, r.AutoTsit5(), saveat=[0.0, 0.3, 0.55, 1.0], reltol=1.4e-8,
                               abstol=1.4e-8, dist_mode=L.DIST_EXACT, rank=rank, world=world)
    Blc = Bc // world
    with torch.no_grad():
        resc, nfec, svc = nodec(torch.from_numpy(np.ascontiguousarray(xc[:, rank * Blc:(rank + 1) * Blc])).cuda(), torch.from_numpy(pc).cuda(),
                                func=r.ERROR_PLUS_STIFFNESS)
    # the library's peer-memory all-reduce against NCCL's, twice (slots are double buffered by call parity), odd length
    ar_ok = True
    for rep in range(3):
        g = torch.from_numpy(np.random.default_rng(100 * rep + rank).standard_normal(158568 + 7 * rep).astype(np.float32)).cuda()
        ref_g = g.clone(); dist.all_reduce(ref_g)
        node.allreduce_(g)
        torch.cuda.synchronize()
        ar_ok = ar_ok and bool(torch.equal(g, ref_g))
    q.put((rank, res.cpu().numpy(), sv.saveval.cpu().numpy(), nfe, resc.cpu().numpy(), svc.saveval.cpu().numpy(), nfec, ar_ok))
    dist.destroy_process_group()


def test_exact_data_parallel_matches_single_solve(oracle_built):
    """SURVEY.md 8e reference-exact mode: 2 ranks sharing the step sequence == one batched solve, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (tools/dist_exact_check.py is the torchrun version; result in profiles/)")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    qq = ctx.Queue()
    procs = [ctx.Process(target=_exact_worker, args=(rk, 2, port, qq)) for rk in range(2)]
    for p in procs:
        p.start()
    outs = sorted((qq.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    D, H, Bg = 784, 100, 96
    rng = np.random.default_rng(1999)
    p_np = orc.glorot_params(rng, D, H); x_np = rng.random((D, Bg), dtype=np.float32)
    ref = orc.Oracle(oracle_cfg(D, H, Bg, 1, 0, orc.REG_ERR_DT, arith=2)).forward(x_np, p_np)
    u = np.concatenate([o[1] for o in outs], axis=1)
    assert outs[0][3] == ref.nf and outs[1][3] == ref.nf
    assert np.array_equal(bits(u), bits(ref.u))
    assert np.array_equal(bits(outs[0][2]), bits(ref.saveval)) and np.array_equal(bits(outs[1][2]), bits(ref.saveval))
    Dc, Bc = 20, 64
    rngc = np.random.default_rng(77)
    pc = orc.glorot_chain_params(rngc, Dc, LATENT_WIDTHS, bias_scale=0.05); xc = rngc.standard_normal((Dc, Bc)).astype(np.float32)
    cfg = orc.OracleConfig(D=Dc, H=50, B=Bc, alg=1, reg_kind=orc.REG_ERR_PLUS_STIFF, kblock1=Dc, widths=LATENT_WIDTHS, acts=(1,) * 8, pre_act=1,
                           saveat=np.array([0.0, 0.3, 0.55, 1.0], dtype=np.float32).astype(np.float64))
    refc = orc.Oracle(cfg).forward(xc, pc)
    usave = np.concatenate([o[4] for o in outs], axis=2).transpose(1, 0, 2)          # (D, S, B) shards -> (S, D, B)
    assert outs[0][6] == refc.nf and outs[1][6] == refc.nf
    assert np.array_equal(bits(usave), bits(refc.usave)), "chain field + saveat in exact mode not bit-identical"
    assert np.array_equal(bits(outs[0][5]), bits(refc.saveval))
    assert outs[0][7] and outs[1][7], "rnde_allreduce_grads differs from the NCCL all-reduce"


A6_CASES = [
    # name, D, H, B, act_out, auto, func, variant
    ("test_node errreg", 2, 10, 5, 0, False, "ERROR_ESTIMATE", 0),
    ("test_node stiffreg", 2, 10, 5, 0, True, "STIFFNESS_ESTIMATE", 0),
    ("toy B=33 streamed weights", 2, 10, 33, 0, False, "ERROR_ESTIMATE", 2),
    ("mid combined", 20, 50, 100, 1, True, "ERROR_PLUS_STIFFNESS", 0),
    ("mnist B=32 cluster8", 784, 100, 32, 1, False, "ERROR_ESTIMATE", 3),
    ("mnist B=48 cluster4 (split-K stepper, tensor-core sweep)", 784, 100, 48, 1, True, "ERROR_PLUS_STIFFNESS", 4),
    ("mnist B=512 (auto variant)", 784, 100, 512, 1, False, "ERROR_ESTIMATE", 0),
    ("D=200 H=37 cluster4 (FFMA sweep)", 200, 37, 33, 0, True, "STIFFNESS_ESTIMATE", 4),
]


@pytest.mark.parametrize("name,D,H,B,act_out,auto,func,variant", A6_CASES, ids=[c[0] for c in A6_CASES])
def test_first_dt_gradient_term(oracle_built, name, D, H, B, act_out, auto, func, variant):
    """`_convert_tspan` (utils.jl:21-23) makes tspan tracked, so the first step size
    (initial-dt heuristic) stays on the tape.  The extra term -- gradient with detach_dt = "all_but_first" minus the
    frozen-step gradient -- is 1e-7 ... 1e-3 of the gradient, below the Float32 noise of the regulariser part, so it is
    checked on its own: the library's diagnostic mode returns the term alone, and so does the oracle (Float64 cotangents
    over the Float32 forward).  The term is a scalar, dL/d(dt_1), times the direction d(dt_1)/d(theta, x):
      * the direction (two field VJPs through the Hairer-Wanner heuristic) must agree to 1e-3 after the best rescaling;
      * the scalar is a sum of the same cancelling O(10) cotangents as the regulariser gradient (a CPU Float32 adjoint gets it
        to 0.03 ... 27 % on these cases): within 10 %, and -- what matters for the gradient -- the term's absolute error must
        stay below 1e-5 of the gradient, a tenth of north_star's 1e-4 bar."""
    r = R()
    rng = np.random.default_rng(11)
    p_np = orc.glorot_params(rng, D, H)
    x_np = rng.random((D, B), dtype=np.float32)
    fobj = getattr(r, func)
    model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, "tanh" if act_out else None))
    node = r.TrackedNeuralODE(model, [0.0, 1.0], True, True, r.AutoTsit5() if auto else r.Tsit5(), reltol=1.4e-8, abstol=1.4e-8,
                              kernel_variant=variant, detach_dt="first_term_only")
    p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=fobj)
    o = orc.Oracle(oracle_cfg(D, H, B, act_out, 1 if auto else 0, fobj.kind, arith=node.arith))
    ref = o.forward(x_np, p_np)
    w = rng.standard_normal((D, B)).astype(np.float32)
    ws = rng.standard_normal(len(ref.saveval)).astype(np.float32)
    assert np.array_equal(bits(res.detach().cpu().numpy()), bits(ref.u))
    ((res * torch.from_numpy(w).cuda()).sum() + (sv.saveval * torch.from_numpy(ws).cuda()).sum()).backward()
    torch.cuda.synchronize()
    tp, tx, _, _ = o.backward(w, ws, hi=True, first_dt_tracked="term")
    cp, cx, _, _ = o.backward(w, ws, first_dt_tracked="term")
    full, _, _, _ = o.backward(w, ws, hi=True)
    assert np.abs(tp).max() > 0 and np.abs(tx).max() > 0
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    gp, gx = p.grad.cpu().numpy().astype(np.float64), x.grad.cpu().numpy().astype(np.float64)
    tp64, tx64 = tp.astype(np.float64), tx.astype(np.float64)
    scale = float((gp * tp64).sum() / (tp64 * tp64).sum())
    note = (scale, rel(gp, tp64), rel(gx, tx64), rel(cp, tp), np.abs(tp).max() / np.abs(full).max())
    assert rel(gp / scale, tp64) <= 1e-3 and rel(gx / scale, tx64) <= 1e-3, note       # direction
    assert abs(scale - 1.0) <= 0.1, note                                                # dL/d(dt_1)
    assert np.abs(gp - tp64).max() <= 1e-5 * np.abs(full).max(), note                   # what reaches the gradient


def test_first_dt_gradient_term_chain_saveat(oracle_built):
    """The same term on the Latent-ODE path (experiments/latent_ode.jl:137-147): Chain field without time input, saveat,
    AutoTsit5, error + stiffness regulariser.  Here the term also collects the explicit dt of states interpolated inside
    the first and the last step (u(theta) = uprev + dt * sum_j b_j(theta) k_j, theta frozen)."""
    r = R()
    rng = np.random.default_rng(77)
    D, widths, acts, B = 20, (50, 50, 20), (1, 1, 0), 70
    p_np = orc.glorot_chain_params(rng, D, widths, bias_scale=0.05)
    x_np = rng.standard_normal((D, B)).astype(np.float32)
    saveat = np.unique(np.concatenate([[0.0], np.sort(rng.random(11)), [1.0]]).astype(np.float32))
    layers, K = [], D
    for M, a in zip(widths, acts):
        layers.append(r.Dense(K, M, "tanh" if a else None)); K = M
    node = r.TrackedNeuralODE(r.Chain("tanh", *layers), [0.0, 1.0], False, True, r.AutoTsit5(), reltol=1.4e-8, abstol=1.4e-8,
                              saveat=saveat.tolist(), detach_dt="first_term_only")
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=r.ERROR_PLUS_STIFFNESS)
    o = orc.Oracle(orc.OracleConfig(D=D, H=max(widths), B=B, alg=1, reg_kind=r.ERROR_PLUS_STIFFNESS.kind, kblock1=D, widths=widths, acts=acts,
                                    pre_act=1, saveat=saveat.astype(np.float64)))
    ref = o.forward(x_np, p_np)
    assert np.array_equal(bits(res.detach().permute(1, 0, 2).cpu().numpy()), bits(ref.usave))
    w = rng.standard_normal(ref.usave.shape).astype(np.float32)
    ws = rng.standard_normal(len(ref.saveval)).astype(np.float32)
    ((res * torch.from_numpy(np.ascontiguousarray(w.transpose(1, 0, 2))).cuda()).sum() + (sv.saveval * torch.from_numpy(ws).cuda()).sum()).backward()
    torch.cuda.synchronize()
    zeros = np.zeros((D, B), np.float32)
    tp, tx, _, _ = o.backward(zeros, ws, hi=True, dusave=w, first_dt_tracked="term")
    full, _, _, _ = o.backward(zeros, ws, hi=True, dusave=w)
    gp, gx = p.grad.cpu().numpy().astype(np.float64), x.grad.cpu().numpy().astype(np.float64)
    tp64, tx64 = tp.astype(np.float64), tx.astype(np.float64)
    assert np.abs(tp64).max() > 0
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    scale = float((gp * tp64).sum() / (tp64 * tp64).sum())
    note = (scale, rel(gp, tp64), rel(gx, tx64), np.abs(tp).max() / np.abs(full).max())
    assert rel(gp / scale, tp64) <= 1e-3 and rel(gx / scale, tx64) <= 1e-3, note
    assert abs(scale - 1.0) <= 0.1, note
    assert np.abs(gp - tp64).max() <= 1e-5 * np.abs(full).max(), note


@pytest.mark.parametrize("seed,scale,tol", [(20, 6.0, 1e-3), (6, 8.0, 1e-4)])
def test_first_dt_gradient_term_after_a_rejected_first_attempt(oracle_built, seed, scale, tol):
    """A rejected attempt divides dt by a detached factor: after rejected FIRST attempts the first accepted step is
    kappa * initial_dt(theta, x), kappa != 1 (a6.cuh takes kappa = dt of the first accepted step / dt_init)."""
    r = R()
    rng = np.random.default_rng(seed)
    D, H, B = 3, 8, 2
    p_np = (orc.glorot_params(rng, D, H, dtype=np.float64) * scale).astype(np.float32)
    x_np = (rng.random((D, B)) * 2).astype(np.float32)
    model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, None))
    node = r.TrackedNeuralODE(model, [0.0, 1.0], True, True, r.Tsit5(), reltol=tol, abstol=tol, detach_dt="first_term_only")
    p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=r.ERROR_ESTIMATE)
    o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_ID, alg=0, reg_kind=orc.REG_ERR_DT, kblock1=D, abstol=tol, reltol=tol, arith=node.arith))
    ref = o.forward(x_np, p_np)
    assert ref.accept_log[0] == 0 and ref.nreject >= 1                      # the case this test exists for
    assert (nfe, node.last_stats.naccept, node.last_stats.nreject) == (ref.nf, ref.naccept, ref.nreject)
    assert np.array_equal(bits(res.detach().cpu().numpy()), bits(ref.u))
    w = rng.standard_normal((D, B)).astype(np.float32)
    ws = rng.standard_normal(len(ref.saveval)).astype(np.float32)
    ((res * torch.from_numpy(w).cuda()).sum() + (sv.saveval * torch.from_numpy(ws).cuda()).sum()).backward()
    torch.cuda.synchronize()
    tp, tx, _, _ = o.backward(w, ws, hi=True, first_dt_tracked="term")
    gp, tp64 = p.grad.cpu().numpy().astype(np.float64), tp.astype(np.float64)
    scale_ = float((gp * tp64).sum() / (tp64 * tp64).sum())
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(gp / scale_, tp64) <= 1e-3 and abs(scale_ - 1.0) <= 0.1, (scale_, rel(gp, tp64))
    assert rel(x.grad.cpu().numpy() / scale_, tx.astype(np.float64)) <= 1e-3


SHORT_CASES = [
    # name, D, H, B, act_out, variant, tol: two accepted steps (both instrumented by the first-dt term) and ONE (dt_1 = t1 - t0: a constant)
    ("toy, 2 steps", 2, 10, 5, 0, 0, 0.3),
    ("toy, 1 step", 2, 10, 5, 0, 0, 100.0),
    ("mnist cluster4, 2 steps", 784, 100, 16, 1, 4, 0.5),
    ("mnist cluster4, 1 step", 784, 100, 16, 1, 4, 1000.0),
]


@pytest.mark.parametrize("name,D,H,B,act_out,variant,tol", SHORT_CASES, ids=[c[0] for c in SHORT_CASES])
def test_gradient_of_very_short_solves(oracle_built, name, D, H, B, act_out, variant, tol):
    """Edge cases of the first-dt term (a6.cuh): the first step is also the last one, or the only one (then dt_1 is the whole
    interval, a constant, and the term vanishes); the extra record must still find its place in the weight-gradient kernels."""
    r = R()
    rng = np.random.default_rng(5)
    p_np = orc.glorot_params(rng, D, H)
    x_np = rng.random((D, B), dtype=np.float32)
    model = r.TDChain(r.Dense(D + 1, H, "tanh"), r.Dense(H + 1, D, "tanh" if act_out else None))
    node = r.TrackedNeuralODE(model, [0.0, 1.0], True, True, r.Tsit5(), reltol=tol, abstol=tol, kernel_variant=variant)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=r.ERROR_ESTIMATE)
    cfg = oracle_cfg(D, H, B, act_out, 0, r.ERROR_ESTIMATE.kind, arith=node.arith)
    cfg.abstol = cfg.reltol = tol
    o = orc.Oracle(cfg)
    ref = o.forward(x_np, p_np)
    assert ref.naccept == (1 if "1 step" in name else 2) and node.last_stats.naccept == ref.naccept
    assert np.array_equal(bits(res.detach().cpu().numpy()), bits(ref.u))
    w = rng.standard_normal((D, B)).astype(np.float32)
    ws = rng.standard_normal(len(ref.saveval)).astype(np.float32)
    ((res * torch.from_numpy(w).cuda()).sum() + (sv.saveval * torch.from_numpy(ws).cuda()).sum()).backward()
    torch.cuda.synchronize()
    dp_hi, dx_hi, _, _ = o.backward(w, ws, hi=True)
    term, _, _, _ = o.backward(w, ws, hi=True, first_dt_tracked="term")
    if "1 step" in name:
        assert np.abs(term).max() == 0.0
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    c_p, c_x = cpu32_noise(o, w, ws, n=3)
    e_p, e_x = rel(p.grad.cpu().numpy(), dp_hi), rel(x.grad.cpu().numpy(), dx_hi)
    assert e_p <= max(1e-4, GRAD_BAR * c_p) and e_x <= max(1e-4, GRAD_BAR * c_x), (e_p, c_p, e_x, c_x, np.abs(term).max() / np.abs(dp_hi).max())
