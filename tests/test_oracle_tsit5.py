"""Independent checks of the oracle's integrator (no Julia available): analytic linear ODE,
convergence order of the 5(4) pair, scipy cross-check, FP32 invariants and thread-count determinism."""
import numpy as np
import pytest
from scipy.integrate import solve_ivp
from scipy.linalg import expm

from oracle import orc


def linear_field_params(rng, D, H):
    """identity activations => u' = W2[:, :H] (W1[:, :D] u + w1t t + b1) + w2t t + b2 = A u + c0 + c1 t"""
    W1 = rng.standard_normal((H, D + 1)) * 0.4
    W2 = rng.standard_normal((D, H + 1)) * 0.4
    b1 = rng.standard_normal(H) * 0.1
    b2 = rng.standard_normal(D) * 0.1
    p = np.concatenate([W1.flatten(order="F"), b1, W2.flatten(order="F"), b2])
    A = W2[:, :H] @ W1[:, :D]
    c0 = W2[:, :H] @ b1 + b2
    c1 = W2[:, :H] @ W1[:, D] + W2[:, H]
    return p, A, c0, c1


def exact_linear(A, c0, c1, u0, T):
    # augment: d/dt [u; 1; t] = [[A, c0, c1], [0,0,0], [0,1,0]] [u; 1; t]
    D = len(u0)
    M = np.zeros((D + 2, D + 2))
    M[:D, :D] = A; M[:D, D] = c0; M[:D, D + 1] = c1; M[D + 1, D] = 1.0
    return (expm(M * T) @ np.concatenate([u0, [1.0, 0.0]]))[:D]


def test_fifth_order_convergence_and_estimator_order(oracle_built):
    rng = np.random.default_rng(5)
    D, H = 3, 4
    p, A, c0, c1 = linear_field_params(rng, D, H)
    u0 = rng.standard_normal((D, 1))
    ref = exact_linear(A, c0, c1, u0[:, 0], 1.0)
    errs, ests = [], []
    ns = [4, 8, 16, 32]
    for n in ns:
        dt = np.full(n, 1.0 / n)
        cfg = orc.OracleConfig(D=D, H=H, B=1, act1=orc.ACT_ID, act2=orc.ACT_ID, reg_kind=orc.REG_ERR_DT, forced_dt=dt,
                               forced_accept=np.ones(n, np.int32))
        r = orc.Oracle(cfg, f64=True).forward(u0, p)
        assert r.naccept == n and abs(r.t_final - 1.0) < 1e-12
        errs.append(np.abs(r.u[:, 0] - ref).max())
        ests.append(np.mean(r.eest_log))
    rate = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert np.all(rate > 4.6) and np.all(rate < 5.6), rate          # global order 5
    erate = np.log2(np.array(ests[:-1]) / np.array(ests[1:]))
    assert np.all(erate > 4.5) and np.all(erate < 5.6), erate       # local estimate of the embedded 4th-order solution ~ dt^5


def test_adaptive_solution_matches_scipy(oracle_built):
    rng = np.random.default_rng(11)
    D, H, B = 6, 16, 3
    p = orc.glorot_params(rng, D, H, dtype=np.float64) * 2.0
    x = rng.random((D, B))
    cfg = orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_TANH, abstol=1e-10, reltol=1e-10)
    o = orc.Oracle(cfg, f64=True)
    r = o.forward(x, p)
    assert r.retcode == 0 and r.nf == 3 + 6 * (r.naccept + r.nreject)

    def f(t, y):
        k, _ = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_TANH), f64=True).rhs(p, y.reshape(B, D).T, t)
        return k.T.reshape(-1)

    s = solve_ivp(f, (0.0, 1.0), x.T.reshape(-1), method="DOP853", rtol=1e-12, atol=1e-12)
    assert np.abs(s.y[:, -1].reshape(B, D).T - r.u).max() < 1e-8


def test_fp32_invariants_and_determinism(oracle_built):
    rng = np.random.default_rng(1999)
    D, H, B = 64, 20, 9
    p = orc.glorot_params(rng, D, H); x = rng.random((D, B), dtype=np.float32)
    base = None
    for nthreads in (1, 2, 8):
        cfg = orc.OracleConfig(D=D, H=H, B=B, kblock1=8, reg_kind=orc.REG_ERR_DT, nthreads=nthreads)
        r = orc.Oracle(cfg).forward(x, p)
        assert r.retcode == 0
        assert r.nf == 3 + 6 * (r.naccept + r.nreject)
        assert len(r.saveval) == r.naccept + 1
        assert r.t_final == 1.0
        key = (r.u.tobytes(), r.saveval.tobytes(), r.nf)
        base = base or key
        assert key == base      # bit-identical for any thread count: the order of operations is pinned
    # the canonical K-blocking is part of the arithmetic: a different kblock may change low-order bits
    r2 = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, kblock1=64, reg_kind=orc.REG_ERR_DT)).forward(x, p)
    assert np.abs(r2.u - np.frombuffer(base[0], np.float32).reshape(D, B)).max() < 1e-4


def test_rejections_and_failure_codes(oracle_built):
    rng = np.random.default_rng(2)
    D, H, B = 4, 8, 2
    p = orc.glorot_params(rng, D, H) * 30.0     # stiff-ish field: forces rejected steps
    x = rng.random((D, B), dtype=np.float32)
    r = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, reg_kind=orc.REG_ERR_DT, abstol=1e-6, reltol=1e-6)).forward(x, p)
    assert r.retcode == 0 and r.nf == 3 + 6 * (r.naccept + r.nreject) and r.nreject > 0
    assert np.sum(r.accept_log == 0) == r.nreject
    r = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, max_steps=3)).forward(x, p)
    assert r.retcode == 1      # maxiters
    pn = p.copy(); pn[0] = np.nan
    r = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B)).forward(x, pn)
    assert r.retcode == 3      # NaN


def test_free_interpolant_is_fourth_order_accurate(oracle_built):
    """Dense output behind `saveat`: on the analytic linear ODE the states interpolated INSIDE fixed
    steps converge with order >= 4 towards expm -- wrong interpolant coefficients would leave an O(1) error independent of dt."""
    rng = np.random.default_rng(9)
    D, H = 3, 4
    p, A, c0, c1 = linear_field_params(rng, D, H)
    u0 = rng.standard_normal((D, 1))
    errs = []
    for n in (4, 8, 16, 32):
        dt = np.full(n, 1.0 / n)
        # three interior points of every step, never a step end
        sa = np.sort(np.concatenate([(np.arange(n) + th) / n for th in (0.17, 0.5, 0.83)]))
        cfg = orc.OracleConfig(D=D, H=H, B=1, act1=orc.ACT_ID, act2=orc.ACT_ID, forced_dt=dt, forced_accept=np.ones(n, np.int32), saveat=sa)
        r = orc.Oracle(cfg, f64=True).forward(u0, p)
        ref = np.stack([exact_linear(A, c0, c1, u0[:, 0], t) for t in sa])
        errs.append(np.abs(r.usave[:, :, 0] - ref).max())
    rate = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert np.all(rate > 3.7), (rate, errs)
    assert errs[-1] < 1e-6


def test_initial_step_matches_the_hairer_wanner_heuristic_of_scipy(oracle_built):
    """SURVEY.md Appendix A.5 (initial dt, UNVERIFIED upstream detail): the same algorithm is implemented independently by
    scipy.integrate (select_initial_step, Hairer-Wanner II.4) -- same RMS norms, same 0.01*d0/d1, same (0.01/max(d1,d2))^(1/5)."""
    from scipy.integrate._ivp.common import select_initial_step
    rng = np.random.default_rng(21)
    D, H, B = 4, 7, 3
    p = orc.glorot_params(rng, D, H, dtype=np.float64) * 1.3
    x = rng.standard_normal((D, B))
    tol = 1e-6
    cfg = orc.OracleConfig(D=D, H=H, B=B, abstol=tol, reltol=tol)
    o = orc.Oracle(cfg, f64=True)
    r = o.forward(x, p)

    def fun(t, y):      # the batched field on the flattened (column-major) state
        k, _ = o.rhs(p, y.reshape(B, D).T, t)
        return np.ascontiguousarray(k.T).reshape(-1)

    y0 = np.ascontiguousarray(x.T).reshape(-1)
    h = select_initial_step(fun, 0.0, y0, 1.0, np.inf, fun(0.0, y0), 1, 4, tol, tol)
    assert abs(r.dt_init - h) <= 1e-9 * h, (r.dt_init, h)


def test_stiffness_estimate_is_a_rayleigh_type_quotient_of_the_jacobian(oracle_built):
    """eigen_est = ||k7 - k6|| / ||u_new - g6||.  For u' = A u + c(t) with c constant in t, k7 - k6 =
    A (u_new - g6) exactly, so every recorded estimate lies between the smallest and the largest singular value of A."""
    rng = np.random.default_rng(13)
    D, H, B = 5, 6, 2
    p, A, c0, c1 = linear_field_params(rng, D, H)
    # remove the explicit time dependence (time columns of both layers) so that c1 = 0
    W1 = p[: H * (D + 1)].reshape(D + 1, H).T.copy(); W1[:, D] = 0
    o2 = H * (D + 1) + H
    W2 = p[o2: o2 + D * (H + 1)].reshape(H + 1, D).T.copy(); W2[:, H] = 0
    p = np.concatenate([W1.flatten(order="F"), p[H * (D + 1): o2], W2.flatten(order="F"), p[o2 + D * (H + 1):]])
    sv = np.linalg.svd(W2[:, :H] @ W1[:, :D], compute_uv=False)
    cfg = orc.OracleConfig(D=D, H=H, B=B, act1=orc.ACT_ID, act2=orc.ACT_ID, alg=orc.ALG_AUTO_TSIT5, reg_kind=orc.REG_STIFF_DT_ABS,
                           abstol=1e-9, reltol=1e-9)
    r = orc.Oracle(cfg, f64=True).forward(rng.standard_normal((D, B)), p)
    eig = np.array([s[3] for s in r.steps])
    assert len(eig) > 5 and np.all(eig <= sv[0] * (1 + 1e-6)) and np.all(eig >= sv[-1] * (1 - 1e-6)), (eig.min(), eig.max(), sv)
