"""The hand-written adjoint of the C oracle against torch autograd through the same algorithm (FP64).
This is what pins the oracle's MATH (the reference's own tests pin nothing: test/test_node.jl is
@code_warntype only).  Shapes follow test/test_node.jl:4-6 (TDChain(Dense(3,10,tanh), Dense(11,2)), x 2x1)
plus larger ones for the canonical K-blocking and every regulariser closure of experiments/mnist_node.jl:62-103."""
import numpy as np
import pytest
import torch

from oracle import orc, torch_oracle as to

CASES = [
    # D, H, B, act2, alg, reg, kblock
    (2, 10, 1, orc.ACT_ID, orc.ALG_TSIT5, orc.REG_NONE, 0),            # test_node.jl:9-25
    (2, 10, 1, orc.ACT_ID, orc.ALG_TSIT5, orc.REG_ERR_DT, 0),          # test_node.jl:28-57
    (2, 10, 1, orc.ACT_ID, orc.ALG_AUTO_TSIT5, orc.REG_STIFF_DT_ABS, 0),  # test_node.jl:60-89
    (2, 10, 3, orc.ACT_ID, orc.ALG_TSIT5, orc.REG_ERR_DT, 0),
    (16, 12, 5, orc.ACT_TANH, orc.ALG_AUTO_TSIT5, orc.REG_ERR_PLUS_STIFF, 4),   # mnist_node.jl:82-99
    (16, 12, 5, orc.ACT_TANH, orc.ALG_AUTO_TSIT5, orc.REG_STIFF_SCALED, 4),     # mnist_node.jl:70-81
    (24, 9, 4, orc.ACT_TANH, orc.ALG_TSIT5, orc.REG_ERR_DT, 3),                 # mnist_node.jl:62-69
]


@pytest.mark.parametrize("detach", ["all", "all_but_first"])
@pytest.mark.parametrize("D,H,B,act2,alg,reg,kb", CASES)
def test_c_oracle_matches_autograd_fp64(oracle_built, D, H, B, act2, alg, reg, kb, detach):
    """detach: SURVEY.md Appendix A.6 -- "all" = every dt frozen; "all_but_first" = the initial-dt heuristic stays on the
    tape (utils.jl:21-23 makes tspan tracked; the recalled upstream behaviour and the default of both oracles' callers)."""
    torch.set_default_dtype(torch.float64)
    try:
        rng = np.random.default_rng(7 + D + 10 * reg)
        cfg = orc.OracleConfig(D=D, H=H, B=B, act2=act2, alg=alg, reg_kind=reg, kblock1=kb, abstol=1.4e-8, reltol=1.4e-8)
        p = orc.glorot_params(rng, D, H, dtype=np.float64) * 1.5 + 0.05 * rng.standard_normal(cfg.n_params)
        x = rng.random((D, B))
        o = orc.Oracle(cfg, f64=True)
        r = o.forward(x, p)
        pt = torch.tensor(p, requires_grad=True)
        xt = torch.tensor(x, requires_grad=True)
        tr = to.solve(xt, pt, D=D, H=H, act2_tanh=(act2 == 1), auto_tsit5=(alg == 1), reg_kind=reg, detach=detach)
        # identical step sequence / NFE accounting (nf = 3 + 6*(naccept+nreject))
        assert (r.nf, r.naccept, r.nreject) == (tr.nf, tr.naccept, tr.nreject)
        assert r.nf == 3 + 6 * (r.naccept + r.nreject)
        assert np.allclose(r.u, tr.u.detach().numpy(), rtol=0, atol=1e-12)
        assert np.allclose(r.dt_log, tr.dt_log, rtol=1e-5, atol=1e-6)   # the last step is tf - t: absolute, not relative
        sv_t = torch.stack(tr.saveval) if tr.saveval else torch.zeros(0)
        if reg:
            assert len(r.saveval) == r.naccept + 1          # SavingCallback: initial entry + one per accepted step (A.7)
            assert np.allclose(r.saveval, sv_t.detach().numpy(), rtol=2e-4, atol=1e-9)
        w = rng.standard_normal((D, B))
        ws = rng.standard_normal(max(len(r.saveval), 1))
        loss = (tr.u * torch.tensor(w)).sum()
        if reg:
            loss = loss + (sv_t * torch.tensor(ws[: len(r.saveval)])).sum()
        gp, gx = torch.autograd.grad(loss, [pt, xt])
        dp, dx, dtb, tb = o.backward(w, ws, first_dt_tracked=(detach == "all_but_first"))
        assert np.abs(dp - gp.numpy()).max() <= 1e-6 * np.abs(gp.numpy()).max()
        assert np.abs(dx - gx.numpy()).max() <= 1e-6 * np.abs(gx.numpy()).max()
        if detach != "all":
            return
        # scalar adjoints w.r.t. every accepted dt (direct + time-shift paths) via replay with leaf dts
        dts = torch.tensor(r.dt_log, requires_grad=True)
        tr2 = to.solve(xt, pt, D=D, H=H, act2_tanh=(act2 == 1), auto_tsit5=(alg == 1), reg_kind=reg, detach="all",
                       forced_dt=r.dt_log, forced_accept=r.accept_log, dt_leaf=dts)
        sv2 = torch.stack(tr2.saveval) if tr2.saveval else torch.zeros(0)
        loss2 = (tr2.u * torch.tensor(w)).sum() + ((sv2 * torch.tensor(ws[: len(r.saveval)])).sum() if reg else 0)
        gdt, = torch.autograd.grad(loss2, [dts])
        gdt = gdt.numpy()[r.accept_log == 1]
        mine = dtb + np.concatenate([np.cumsum(tb[::-1])[::-1][1:], [0.0]])
        assert np.abs(mine - gdt).max() <= 1e-5 * np.abs(gdt).max()
    finally:
        torch.set_default_dtype(torch.float32)


def test_first_saved_value_conventions(oracle_built):
    """Appendix A.7: the entry recorded at callback initialisation uses EEst=1, dt=0, eigen_est=1."""
    rng = np.random.default_rng(3)
    p = orc.glorot_params(rng, 2, 10); x = rng.random((2, 1), dtype=np.float32)
    first = {}
    for reg, alg in ((orc.REG_ERR_DT, 0), (orc.REG_STIFF_DT_ABS, 1), (orc.REG_STIFF_SCALED, 1), (orc.REG_ERR_PLUS_STIFF, 1)):
        r = orc.Oracle(orc.OracleConfig(D=2, H=10, B=1, act2=0, alg=alg, reg_kind=reg)).forward(x, p)
        first[reg] = r.saveval[0]
    assert first[orc.REG_ERR_DT] == 0 and first[orc.REG_STIFF_DT_ABS] == 0
    stab = np.float32(1) / np.float32(3.5068)
    assert first[orc.REG_STIFF_SCALED] == stab
    assert first[orc.REG_ERR_PLUS_STIFF] == np.float32(0.1) * stab


@pytest.mark.parametrize("detach", ["all", "all_but_first"])
@pytest.mark.parametrize("reg,alg", [(orc.REG_NONE, 0), (orc.REG_ERR_DT, 0), (orc.REG_ERR_PLUS_STIFF, 1)])
def test_saveat_dense_output_matches_autograd_fp64(oracle_built, reg, alg, detach):
    """Multi-save functors (neural_ode.jl:79-108,146-180): the C oracle's saved states (Tsit5 free interpolant,
    SURVEY.md Appendix A.9) and their adjoint against torch autograd in FP64; saveat must not change the steps."""
    torch.set_default_dtype(torch.float64)
    rng = np.random.default_rng(11)
    D, H, B = 5, 8, 3
    sa = np.array([0.0, 0.07, 0.4, 0.41, 0.93, 1.0])
    p = orc.glorot_params(rng, D, H, dtype=np.float64) * 1.5
    x = rng.random((D, B))
    cfg = orc.OracleConfig(D=D, H=H, B=B, reg_kind=reg, alg=alg, saveat=sa)
    o = orc.Oracle(cfg, f64=True)
    r = o.forward(x, p)
    plain = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, reg_kind=reg, alg=alg), f64=True).forward(x, p)
    assert (r.nf, r.naccept, r.nreject) == (plain.nf, plain.naccept, plain.nreject)      # no tstops added
    assert np.array_equal(r.usave[0], x) and np.array_equal(r.usave[-1], plain.u)          # t0 -> input, t1 -> copy of u
    pt = torch.tensor(p, requires_grad=True); xt = torch.tensor(x, requires_grad=True)
    tr = to.solve(xt, pt, D=D, H=H, reg_kind=reg, auto_tsit5=bool(alg), saveat=sa, detach=detach)
    us = torch.stack(tr.usave)
    assert np.abs(r.usave - us.detach().numpy()).max() < 1e-12
    w = rng.standard_normal(r.usave.shape)
    loss = (us * torch.tensor(w)).sum()
    ws = None
    if reg != orc.REG_NONE:
        ws = rng.standard_normal(len(r.saveval))
        loss = loss + (torch.stack(tr.saveval) * torch.tensor(ws)).sum()
    gp, gx = torch.autograd.grad(loss, [pt, xt])
    dp, dx, _, _ = o.backward(np.zeros((D, B)), ws, dusave=w, first_dt_tracked=(detach == "all_but_first"))
    assert np.abs(dp - gp.numpy()).max() <= 1e-7 * np.abs(gp.numpy()).max()
    assert np.abs(dx - gx.numpy()).max() <= 1e-7 * np.abs(gx.numpy()).max()
    torch.set_default_dtype(torch.float32)


def test_free_interpolant_endpoint_identities():
    """b_i(1) = a_7i (the interpolant ends on the step's new state) and b_i(0) = 0."""
    b1 = to.interp_weights(1.0)
    assert np.allclose(b1[:6], to.A[7], atol=1e-14) and abs(b1[6]) < 1e-14
    assert all(v == 0.0 for v in to.interp_weights(0.0))


@pytest.mark.parametrize("detach", ["all", "all_but_first"])
@pytest.mark.parametrize("reg,alg", [(orc.REG_NONE, 0), (orc.REG_ERR_DT, 0), (orc.REG_ERR_PLUS_STIFF, 1)])
def test_chain_field_matches_autograd_fp64(oracle_built, reg, alg, detach):
    """Chain fields (Latent-ODE generator dynamics, experiments/latent_ode.jl:109-121: tanh pre-activation, Dense
    layers, no time input) with saveat: C oracle against torch autograd in FP64 -- steps, saved states, gradients."""
    torch.set_default_dtype(torch.float64)
    rng = np.random.default_rng(5)
    D, B = 6, 4
    widths, acts = (9, 6, 11, 6), (1, 1, 1, 0)
    sa = np.array([0.0, 0.13, 0.5, 0.77, 1.0])
    cfg = orc.OracleConfig(D=D, H=11, B=B, reg_kind=reg, alg=alg, saveat=sa, widths=widths, acts=acts, pre_act=1)
    p = orc.glorot_chain_params(rng, D, widths, dtype=np.float64, bias_scale=0.1) * 1.5
    assert p.size == cfg.n_params == sum(m * k + m for m, k in zip(widths, (D,) + widths[:-1]))
    x = rng.random((D, B))
    o = orc.Oracle(cfg, f64=True)
    r = o.forward(x, p)
    pt = torch.tensor(p, requires_grad=True); xt = torch.tensor(x, requires_grad=True)
    tr = to.solve(xt, pt, D=D, H=11, reg_kind=reg, auto_tsit5=bool(alg), saveat=sa, chain=(widths, acts, 1), detach=detach)
    assert (r.nf, r.naccept) == (tr.nf, tr.naccept)
    us = torch.stack(tr.usave)
    assert np.abs(r.usave - us.detach().numpy()).max() < 1e-12
    w = rng.standard_normal(r.usave.shape)
    loss = (us * torch.tensor(w)).sum()
    ws = None
    if reg != orc.REG_NONE:
        ws = rng.standard_normal(len(r.saveval))
        loss = loss + (torch.stack(tr.saveval) * torch.tensor(ws)).sum()
    gp, gx = torch.autograd.grad(loss, [pt, xt])
    dp, dx, _, _ = o.backward(np.zeros((D, B)), ws, dusave=w, first_dt_tracked=(detach == "all_but_first"))
    assert np.abs(dp - gp.numpy()).max() <= 1e-7 * np.abs(gp.numpy()).max()
    assert np.abs(dx - gx.numpy()).max() <= 1e-7 * np.abs(gx.numpy()).max()
    torch.set_default_dtype(torch.float32)


def test_fixed24_arithmetic_is_a_rounding_of_the_same_field(oracle_built):
    """arith = 1 (exact truncated fixed-point layer products, the tensor-core forward stepper's arithmetic): one field
    evaluation agrees with the FMA-chain arithmetic and with Float64 to Float32 round-off, and the solve takes a comparable
    number of steps (the embedded error estimate is rounding noise at tol 1.4e-8: a noisier arithmetic would take more)."""
    rng = np.random.default_rng(5)
    D, H, B = 784, 100, 6
    p = orc.glorot_params(rng, D, H); x = rng.random((D, B), dtype=np.float32)
    k0, _ = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, kblock1=98)).rhs(p, x, 0.3)
    k1, _ = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, kblock1=98, arith=1)).rhs(p, x, 0.3)
    k64, _ = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, kblock1=98), f64=True).rhs(p.astype(np.float64), x.astype(np.float64), 0.3)
    assert np.abs(k1 - k64).max() <= 4e-7 and np.abs(k0 - k64).max() <= 4e-7
    r0 = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, reg_kind=orc.REG_ERR_DT, kblock1=98)).forward(x, p)
    r1 = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, reg_kind=orc.REG_ERR_DT, kblock1=98, arith=1)).forward(x, p)
    assert np.abs(r1.u - r0.u).max() <= 2e-6 * np.abs(r0.u).max()
    assert abs(r1.nf - r0.nf) <= 0.1 * r0.nf


@pytest.mark.parametrize("seed,scale,tol", [(20, 6.0, 1e-3), (6, 8.0, 1e-4), (34, 8.0, 1e-4)])
def test_first_dt_term_with_a_rejected_first_attempt(oracle_built, seed, scale, tol):
    """Appendix A.6: a rejected attempt divides dt by a detached factor, so after rejected FIRST attempts the first accepted step
    is kappa * initial_dt(theta, x) with a constant kappa != 1 -- the C oracle's first-dt term (kappa = dt_accepted / dt_init)
    against autograd through the torch oracle, which simply keeps the tracked dt through its retries."""
    torch.set_default_dtype(torch.float64)
    try:
        rng = np.random.default_rng(seed)
        D, H, B = 3, 8, 2
        cfg = orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_ID, alg=0, reg_kind=orc.REG_ERR_DT, abstol=tol, reltol=tol)
        p = orc.glorot_params(rng, D, H, dtype=np.float64) * scale
        x = rng.random((D, B)) * 2
        o = orc.Oracle(cfg, f64=True)
        r = o.forward(x, p)
        assert r.accept_log[0] == 0 and r.nreject >= 1                      # the case this test exists for
        pt = torch.tensor(p, requires_grad=True); xt = torch.tensor(x, requires_grad=True)
        tr = to.solve(xt, pt, D=D, H=H, act2_tanh=False, reg_kind=orc.REG_ERR_DT, abstol=tol, reltol=tol, detach="all_but_first")
        assert (r.nf, r.naccept, r.nreject) == (tr.nf, tr.naccept, tr.nreject)
        w = rng.standard_normal((D, B)); ws = rng.standard_normal(len(r.saveval))
        loss = (tr.u * torch.tensor(w)).sum() + (torch.stack(tr.saveval) * torch.tensor(ws)).sum()
        gp, gx = torch.autograd.grad(loss, [pt, xt])
        dp, dx, _, _ = o.backward(w, ws, first_dt_tracked=True)
        term, _, _, _ = o.backward(w, ws, first_dt_tracked="term")
        assert np.abs(term).max() > 1e-6 * np.abs(dp).max()                 # the term is not negligible here
        assert np.abs(dp - gp.numpy()).max() <= 1e-7 * np.abs(gp.numpy()).max()
        assert np.abs(dx - gx.numpy()).max() <= 1e-7 * np.abs(gx.numpy()).max()
    finally:
        torch.set_default_dtype(torch.float32)
