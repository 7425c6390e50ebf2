"""CPU pins of oracle/ffjord_oracle.py (no CUDA path for this row yet -- DESIGN.md section 9).
The reference holds no golden vectors for FFJORD, so the restatement is pinned against exact Jacobians (torch autograd), the
change-of-variables identity of a continuous normalising flow, and finite differences of the discrete solve."""
import math

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import ffjord_oracle as F  # noqa: E402


def setup(D=3, H=8, B=4, seed=0):
    rng = np.random.default_rng(seed)
    p = torch.from_numpy(F.glorot_params(rng, D, H, dtype=np.float64, bias_scale=0.3) * 1.5)
    x = torch.from_numpy(rng.standard_normal((D, B)))
    e = torch.from_numpy(rng.standard_normal((D, B)))
    return p, x, e


def test_parameter_count_and_order():
    # MLPDynamics(43, 100) of the tabular experiment (ffjord_tabular.jl:109): 3 ConcatSquash layers, W + 4 columns each
    assert F.n_params(43, 100) == (100 * 43 + 400) + (100 * 100 + 400) + (43 * 100 + 4 * 43)
    p = torch.arange(F.n_params(2, 3), dtype=torch.float64)
    W, B, bW, bB, G = F.unpack(p, 2, 3)[0]
    assert W[1, 0] == 1 and W[0, 1] == 3                      # column-major layer_W first
    assert B[0, 0] == 6 and bW[0, 0] == 9 and bB[0, 0] == 12 and G[0, 0] == 15


def test_stable_activations():
    x = torch.tensor([-800.0, -3.0, 0.0, 2.0, 800.0], dtype=torch.float64)
    assert torch.allclose(F.sigmoid(x)[1:4], torch.sigmoid(x)[1:4], rtol=1e-15, atol=0)
    assert torch.allclose(F.softplus(x), torch.nn.functional.softplus(x), rtol=1e-15, atol=1e-300)
    assert torch.isfinite(F.sigmoid(x)).all() and torch.isfinite(F.softplus(x)).all()


def test_forw_n_back_is_the_network_and_its_transposed_jacobian_product():
    D, H, B = 3, 8, 4
    p, x, e = setup(D, H, B)
    t = torch.tensor(0.37, dtype=torch.float64)
    mz, eJ = F.forw_n_back(p, x, t, e, D, H)
    z = x.clone().requires_grad_(True)
    y = F.dynamics(p, z, t, D, H)
    assert torch.equal(y.detach(), mz) or torch.allclose(y.detach(), mz, rtol=1e-15, atol=1e-15)
    vjp, = torch.autograd.grad((y * e).sum(), z)
    assert torch.allclose(vjp, eJ, rtol=1e-12, atol=1e-13)


def test_rhs_rows_and_hutchinson_trace_against_the_exact_trace():
    D, H, B = 3, 8, 4
    p, x, e = setup(D, H, B)
    t = torch.tensor(0.6, dtype=torch.float64)
    u = torch.cat([x, torch.zeros(3, B, dtype=torch.float64)], 0)
    r = F.ffjord_rhs(u, p, t, e, D, H, True)
    mz, eJ = F.forw_n_back(p, x, t, e, D, H)
    assert r.shape == (D + 3, B)
    assert torch.equal(r[:D], mz) and torch.allclose(r[D], -(eJ * e).sum(0)) and torch.allclose(r[D + 1], (mz ** 2).sum(0))
    assert torch.allclose(r[D + 2], (eJ ** 2).sum(0))
    # sum over the unit vectors of e^T J e is the trace: the estimator is exact in expectation
    exact = F.deterministic_rhs(torch.cat([x, torch.zeros(1, B, dtype=torch.float64)], 0), p, t, D, H)
    tr = torch.zeros(B, dtype=torch.float64)
    for i in range(D):
        ei = torch.zeros(D, B, dtype=torch.float64); ei[i] = 1
        tr = tr + F.ffjord_rhs(exact * 0 + torch.cat([x, torch.zeros(1, B, dtype=torch.float64)], 0), p, t, ei, D, H, False)[D]
    assert torch.allclose(tr, exact[D], rtol=1e-12, atol=1e-13)
    assert torch.allclose(exact[:D], mz, rtol=1e-14, atol=1e-14)


def test_change_of_variables_identity():
    """logp(x) = logN(z(1)) + log|det dz(1)/dx|: with e a unit vector the 'estimate' is one exact diagonal entry of J, so the
    sum of delta_logp over the D unit vectors is -log det of the flow map's Jacobian (taken by autograd through the solve)."""
    D, H, B = 2, 6, 1
    p, x, _ = setup(D, H, B, seed=3)
    total = 0.0
    for i in range(D):
        ei = torch.zeros(D, B, dtype=torch.float64); ei[i] = 1
        r = F.ffjord(x, p, ei, D=D, H=H, regularized_functor=False, abstol=1e-10, reltol=1e-10)
        total += float(r.sol.u[D, 0])
    xg = x.clone().requires_grad_(True)
    r = F.ffjord(xg, p, torch.zeros(D, B, dtype=torch.float64), D=D, H=H, regularized_functor=False, abstol=1e-10, reltol=1e-10)
    J = torch.stack([torch.autograd.grad(r.sol.u[i, 0], xg, retain_graph=True)[0][:, 0] for i in range(D)])
    assert abs(-total - math.log(abs(float(torch.det(J))))) < 1e-7
    # and the functor's log-density is the standard normal at z(1) minus delta_logp
    z = r.sol.u[:D, 0].detach()
    assert abs(float(r.logpx[0].detach()) - (float(-(math.log(2 * math.pi) + z * z).sum() / 2) - float(r.sol.u[D, 0].detach()))) < 1e-14


def test_regularised_functor_and_loss_gradient_against_finite_differences():
    D, H, B = 3, 6, 3
    p, x, e = setup(D, H, B, seed=5)
    pg = p.clone().requires_grad_(True)
    total, r = F.loss_function(x, pg, e, D=D, H=H, regularized_functor=True, lam_r=1.0, abstol=1e-5, reltol=1e-5)
    assert r.saveval.shape[0] == r.sol.naccept + 1 and float(r.saveval[0].detach()) == 0.0       # save_start entry: EEst*dt with dt = 0
    assert r.nfe == 3 + 6 * (r.sol.naccept + r.sol.nreject)
    assert float(r.lam1.abs().sum()) == 0 and float(r.lam2.abs().sum()) == 0               # ffjord.jl:135 returns _z
    g, = torch.autograd.grad(total, pg)
    # replaying the recorded attempts freezes the step sequence: central differences then see the same discrete map
    kw = dict(forced_dt=r.sol.dt_log, forced_accept=r.sol.accept_log)
    rng = np.random.default_rng(1)
    for k in rng.choice(p.numel(), size=6, replace=False):
        h = 1e-5      # EEst divides O(tol) differences by tol: its rounding noise (~1e-16/tol) bounds how small h may be
        pp, pm = p.clone(), p.clone()
        pp[k] += h; pm[k] -= h
        lp, _ = F.loss_function(x, pp, e, D=D, H=H, regularized_functor=True, lam_r=1.0, abstol=1e-5, reltol=1e-5, **kw)
        lm, _ = F.loss_function(x, pm, e, D=D, H=H, regularized_functor=True, lam_r=1.0, abstol=1e-5, reltol=1e-5, **kw)
        fd = float(lp - lm) / (2 * h)
        assert abs(fd - float(g[k])) <= 1e-6 * max(1.0, abs(fd)), (int(k), fd, float(g[k]))


def test_kinetic_rows_of_the_unregularised_functor():
    D, H, B = 3, 6, 2
    p, x, e = setup(D, H, B, seed=8)
    r = F.ffjord(x, p, e, D=D, H=H, regularized_functor=False, regularize=True, abstol=1e-8, reltol=1e-8)
    assert r.saveval is None and (r.lam1 > 0).all() and (r.lam2 > 0).all()
    r0 = F.ffjord(x, p, e, D=D, H=H, regularized_functor=False, regularize=False, abstol=1e-8, reltol=1e-8)
    assert torch.allclose(r.logpx, r0.logpx, rtol=0, atol=1e-6)                           # same flow, extra rows only


def test_golden_fixture_of_the_tabular_shape():
    """tests/golden/ffjord_tabular.npz (make_golden_ffjord.py): MLPDynamics(43, 100), both functors, loss gradient."""
    from pathlib import Path
    g = np.load(Path(__file__).resolve().parent / "golden" / "ffjord_tabular.npz")
    D, H = 43, 100
    p = torch.from_numpy(g["p"]).requires_grad_(True)
    x, e = torch.from_numpy(g["x"]), torch.from_numpy(g["e"])
    total, r = F.loss_function(x, p, e, D=D, H=H, regularized_functor=True, lam_r=100.0)
    assert [r.nfe, r.sol.naccept, r.sol.nreject] == g["counts"].tolist()
    assert np.allclose(r.logpx.detach().numpy(), g["logpx"], rtol=1e-11, atol=0)
    assert np.allclose(r.saveval.detach().numpy(), g["saveval"], rtol=1e-8, atol=1e-14)
    assert abs(float(total.detach()) - float(g["loss"])) <= 1e-10 * abs(float(g["loss"]))
    dp, = torch.autograd.grad(total, p)
    assert np.abs(dp.numpy() - g["dp"]).max() <= 1e-8 * np.abs(g["dp"]).max()
    r0 = F.ffjord(x, p.detach(), e, D=D, H=H, regularized_functor=False, regularize=True)
    assert [r0.nfe, r0.sol.naccept, r0.sol.nreject] == g["counts_kinetic"].tolist()
    assert np.allclose(r0.lam1.numpy(), g["lam1"], rtol=1e-10) and np.allclose(r0.lam2.numpy(), g["lam2"], rtol=1e-10)


# ---- the C restatement of the same field (oracle/rnde_oracle.c, csq_extra): the future bit-level specification ----
def _c_oracle(D, H, B, extra, e, f64, **kw):
    from oracle import orc
    orc.build()
    return orc.Oracle(orc.OracleConfig(D=D + extra, H=H, B=B, csq_extra=extra, csq_noise=e, kblock1=D + extra, **kw), f64=f64)


@pytest.mark.parametrize("extra", [1, 3])
def test_c_field_matches_the_torch_restatement(extra):
    D, H, B = 5, 9, 4
    p, x, e = setup(D, H, B, seed=11)
    u = torch.cat([x, torch.from_numpy(np.random.default_rng(2).standard_normal((extra, B)))], 0)
    ref = F.ffjord_rhs(u, p, torch.tensor(0.45, dtype=torch.float64), e, D, H, extra == 3).numpy()
    o = _c_oracle(D, H, B, extra, e.numpy(), True)
    assert o.cfg.n_params == F.n_params(D, H)
    k, _ = o.rhs(p.numpy(), u.numpy(), 0.45)
    assert np.abs(k - ref).max() <= 1e-13 * np.abs(ref).max()
    o32 = _c_oracle(D, H, B, extra, e.numpy().astype(np.float32), False)
    k32, _ = o32.rhs(p.numpy().astype(np.float32), u.numpy().astype(np.float32), 0.45)
    assert np.abs(k32 - ref).max() <= 2e-6 * np.abs(ref).max()          # canonical Float32 arithmetic (regnde_canon.h activations)


def test_c_solve_matches_the_torch_functors():
    """Same Tsit5 controller on the augmented state: identical step counts, log-densities to 1e-11 (Float64 builds)."""
    D, H, B = 4, 8, 3
    p, x, e = setup(D, H, B, seed=12)
    from oracle import orc
    for extra, regf in ((1, True), (3, False)):
        r = F.ffjord(x, p, e, D=D, H=H, regularized_functor=regf, regularize=not regf)
        o = _c_oracle(D, H, B, extra, e.numpy(), True, reg_kind=orc.REG_ERR_DT if regf else orc.REG_NONE, abstol=1.4e-8, reltol=1.4e-8)
        u0 = np.concatenate([x.numpy(), np.zeros((extra, B))], 0)
        c = o.forward(u0, p.numpy())
        assert (c.nf, c.naccept, c.nreject) == (r.nfe, r.sol.naccept, r.sol.nreject)
        assert np.abs(c.u - r.sol.u.numpy()).max() <= 1e-11 * np.abs(c.u).max()
        z = c.u[:D]
        logpx = (-(np.log(2 * np.pi) + z * z) / 2).sum(0) - c.u[D]
        assert np.abs(logpx - r.logpx.numpy()).max() <= 1e-10 * np.abs(logpx).max()
        if regf:
            # EEst cancels O(1) stage values down to O(tol): the two summation orders differ by ~1e-16 / 1.4e-8 relative per term
            assert np.allclose(c.saveval, r.saveval.numpy(), rtol=1e-3, atol=1e-14)


def test_c_float32_solve_is_deterministic_and_close():
    D, H, B = 43, 100, 5
    rng = np.random.default_rng(21)
    p = F.glorot_params(rng, D, H, dtype=np.float32, bias_scale=0.05)
    x = rng.standard_normal((D, B)).astype(np.float32); e = rng.standard_normal((D, B)).astype(np.float32)
    u0 = np.concatenate([x, np.zeros((1, B), np.float32)], 0)
    runs = []
    for threads in (1, 4):
        o = _c_oracle(D, H, B, 1, e, False, nthreads=threads)
        runs.append(o.forward(u0, p))
    assert runs[0].nf == runs[1].nf and np.array_equal(runs[0].u.view(np.uint32), runs[1].u.view(np.uint32))
    r = F.ffjord(torch.from_numpy(x).double(), torch.from_numpy(p).double(), torch.from_numpy(e).double(), D=D, H=H, regularized_functor=False)
    assert np.abs(runs[0].u - r.sol.u.numpy()).max() <= 5e-5 * np.abs(r.sol.u.numpy()).max()


@pytest.mark.parametrize("extra,regf", [(1, True), (3, False), (1, False)])
def test_c_adjoint_matches_autograd(extra, regf):
    """The hand-written reverse sweep through csq_column + the Tsit5 adjoint (rnde_oracle_bwd.inc) against torch autograd
    through the whole solve, Float64: gradients with respect to the parameters and the data."""
    from oracle import orc
    D, H, B = 4, 7, 3
    p, x, e = setup(D, H, B, seed=31 + extra)
    rng = np.random.default_rng(5)
    pg, xg = p.clone().requires_grad_(True), x.clone().requires_grad_(True)
    r = F.ffjord(xg, pg, e, D=D, H=H, regularized_functor=regf, regularize=(extra == 3))
    w = rng.standard_normal((D + extra, B))
    ws = rng.standard_normal(r.sol.naccept + 1)
    loss = (r.sol.u * torch.from_numpy(w)).sum()
    if regf:
        loss = loss + (r.saveval * torch.from_numpy(ws)).sum()
    gp, gx = torch.autograd.grad(loss, (pg, xg))
    o = _c_oracle(D, H, B, extra, e.numpy(), True, reg_kind=orc.REG_ERR_DT if regf else orc.REG_NONE, abstol=1.4e-8, reltol=1.4e-8)
    c = o.forward(np.concatenate([x.numpy(), np.zeros((extra, B))], 0), p.numpy())
    assert (c.nf, c.naccept) == (r.nfe, r.sol.naccept)
    dp, dx, _, _ = o.backward(w, ws, first_dt_tracked=False)      # oracle/ffjord_oracle.py solves with detach = all
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    # the regulariser's cotangents are O(1/tol): its gradient carries the 1e-16/1.4e-8 noise of EEst (see the solve test)
    tol = 1e-6 if regf else 1e-12         # observed 2.6e-8 / 5e-14
    assert rel(dp, gp.numpy()) <= tol, rel(dp, gp.numpy())
    assert rel(dx[:D], gx.numpy()) <= tol, rel(dx[:D], gx.numpy())
    if not regf and extra == 1:
        # Float32 build: canonical forward, cotangents in Float64 over it (the yardstick GPU adjoints are judged by) and in Float32
        o32 = _c_oracle(D, H, B, extra, e.numpy().astype(np.float32), False, abstol=1.4e-8, reltol=1.4e-8)
        o32.forward(np.concatenate([x.numpy(), np.zeros((extra, B))], 0).astype(np.float32), p.numpy().astype(np.float32))
        dp_hi, dx_hi, _, _ = o32.backward(w.astype(np.float32), ws.astype(np.float32), hi=True, first_dt_tracked=False)
        dp_32, _, _, _ = o32.backward(w.astype(np.float32), ws.astype(np.float32), first_dt_tracked=False)
        assert rel(dp_hi, gp.numpy()) <= 1e-4 and rel(dx_hi[:D], gx.numpy()) <= 1e-4 and rel(dp_32, dp_hi) <= 1e-4
