"""GPU: SURVEY.md 8f row N4 (FFJORD) against the C oracle -- one evaluation of the field (csrc/csq.cuh through
rnde_test_csq_rhs, oracle/rnde_oracle.c csq_column) and the augmented-state solve bit for bit; the gradient through the solve
(csrc/csq_bwd.cuh) against the oracle's Float64-cotangent adjoint; `sample` as a round trip; the WeightDecay + ADAM rule."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ffjord_oracle as F, orc  # noqa: E402  (the checker)


@pytest.mark.parametrize("Dz,H,B,extra", [(43, 100, 8, 1), (43, 100, 7, 3), (5, 9, 3, 1), (6, 64, 130, 3)])
def test_field_evaluation_bit_identical(oracle_built, Dz, H, B, extra):
    import regneuralde.jl_b200 as R
    lib = R.lib()
    rng = np.random.default_rng(17)
    p = F.glorot_params(rng, Dz, H, dtype=np.float32, bias_scale=0.2)
    z = rng.standard_normal((Dz + extra, B)).astype(np.float32)
    e = rng.standard_normal((Dz, B)).astype(np.float32)
    o = orc.Oracle(orc.OracleConfig(D=Dz + extra, H=H, B=B, csq_extra=extra, csq_noise=e, kblock1=Dz + extra))
    for t in (0.0, 0.37, 1.0):
        ref, _ = o.rhs(p, z, t)
        pd, zd, ed = (torch.from_numpy(np.ascontiguousarray(a.T if a.ndim == 2 else a)).cuda() for a in (p, z, e))
        kd = torch.empty(B, Dz + extra, device="cuda", dtype=torch.float32)
        assert lib.rnde_test_csq_rhs(Dz, H, extra, B, pd.data_ptr(), zd.data_ptr(), ed.data_ptr(), C.c_float(t), kd.data_ptr(), None) == 0
        got = kd.cpu().numpy().T
        assert np.array_equal(got.view(np.uint32), np.ascontiguousarray(ref, dtype=np.float32).view(np.uint32)), \
            (t, np.abs(got - ref).max())


# ---- the augmented-state solve through the C ABI / the TrackedFFJORD mirror (forward only) -----------------------------------
# First run on a B200 in round 2 (gpurun_out/r2a_ffjord.txt: 8 passed): bit-identical to the C oracle.


@pytest.mark.parametrize("Dz,H,B,regf,kinetic", [(43, 100, 8, False, False), (43, 100, 7, True, False), (43, 100, 6, False, True), (5, 9, 130, True, False)])
def test_ffjord_solve_bit_identical(oracle_built, Dz, H, B, regf, kinetic):
    """ffjord(x, p, e) -> (logpx, l1, l2, nfe, sv) (src/models/ffjord.jl:68-137): states, NFE and saved values of the augmented
    solve bit-identical to the C oracle; the log-density is host arithmetic on top (compared to 1e-6)."""
    import regneuralde.jl_b200 as R
    rng = np.random.default_rng(23)
    p = F.glorot_params(rng, Dz, H, dtype=np.float32, bias_scale=0.1)
    x = rng.standard_normal((Dz, B)).astype(np.float32)
    e = rng.standard_normal((Dz, B)).astype(np.float32)
    extra = 3 if kinetic else 1
    o = orc.Oracle(orc.OracleConfig(D=Dz + extra, H=H, B=B, csq_extra=extra, csq_noise=e, kblock1=Dz + extra,
                                    reg_kind=orc.REG_ERR_DT if regf else orc.REG_NONE))
    ref = o.forward(np.concatenate([x, np.zeros((extra, B), np.float32)], 0), p)
    model = R.CSQDynamics(Dz, H)
    ff = R.TrackedFFJORD(model, [0.0, 1.0], True, regf, R.Tsit5(), reltol=1.4e-8, abstol=1.4e-8)
    with torch.no_grad():
        logpx, l1, l2, nfe, sv = ff(torch.from_numpy(x).cuda(), torch.from_numpy(p).cuda(), torch.from_numpy(e).cuda(), regularize=kinetic)
    assert nfe == ref.nf and ff.last_stats.naccept == ref.naccept and ff.last_stats.nreject == ref.nreject
    z, dl = ref.u[:Dz], ref.u[Dz]
    want = (-(np.log(2 * np.pi) + z.astype(np.float64) ** 2) / 2).sum(0) - dl
    assert np.abs(logpx.cpu().numpy() - want).max() <= 1e-6 * np.abs(want).max()
    if kinetic:
        assert np.array_equal(l1.cpu().numpy().view(np.uint32), ref.u[Dz + 1].view(np.uint32))
        assert np.array_equal(l2.cpu().numpy().view(np.uint32), ref.u[Dz + 2].view(np.uint32))
    if regf:
        assert np.array_equal(sv.saveval.cpu().numpy().view(np.uint32), ref.saveval.view(np.uint32))
    else:
        assert sv is None


# ---- the gradient (round 2): reverse sweep through the hand-differentiated field (csrc/csq_bwd.cuh) -------------------------------


@pytest.mark.parametrize("Dz,H,B,regf,kinetic", [(5, 9, 10, False, False), (5, 9, 130, True, False), (43, 100, 8, False, True), (43, 100, 24, True, False),
                                                   (43, 100, 600, True, False)])     # 600 columns: the 8-column tile variant
def test_ffjord_gradient_matches_oracle(oracle_built, Dz, H, B, regf, kinetic):
    """Tracker.gradient of the tabular loss terms (experiments/ffjord_tabular.jl:137-141): random cotangents on logpx, on the kinetic
    regulariser rows and on the saved values; CUDA against the C oracle's adjoint with Float64 cotangents over the same Float32
    forward (which is itself checked against torch autograd through the same field, tests/test_ffjord_oracle.py).  The step sizes
    are frozen on both sides."""
    import regneuralde.jl_b200 as R
    sys_path_tests = __import__("gradbar")
    rng = np.random.default_rng(29)
    p_np = F.glorot_params(rng, Dz, H, dtype=np.float32, bias_scale=0.1)
    x_np = rng.standard_normal((Dz, B)).astype(np.float32)
    e_np = rng.standard_normal((Dz, B)).astype(np.float32)
    extra = 3 if kinetic else 1
    o = orc.Oracle(orc.OracleConfig(D=Dz + extra, H=H, B=B, csq_extra=extra, csq_noise=e_np, kblock1=Dz + extra,
                                    reg_kind=orc.REG_ERR_DT if regf else orc.REG_NONE))
    ref = o.forward(np.concatenate([x_np, np.zeros((extra, B), np.float32)], 0), p_np)
    ff = R.TrackedFFJORD(R.CSQDynamics(Dz, H), [0.0, 1.0], True, regf, R.Tsit5(), reltol=1.4e-8, abstol=1.4e-8, tape_capacity=64)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    x = torch.from_numpy(x_np).cuda().requires_grad_(True)
    logpx, l1, l2, nfe, sv = ff(x, p, torch.from_numpy(e_np).cuda(), regularize=kinetic)
    assert nfe == ref.nf and ff.last_stats.naccept == ref.naccept
    # loss = sum(w_l .* (logpz - delta_logp)) + sum(w1 .* l1) + sum(w2 .* l2) + sum(ws .* saveval)
    w_l = rng.standard_normal(B).astype(np.float32)
    w1, w2 = rng.standard_normal(B).astype(np.float32), rng.standard_normal(B).astype(np.float32)
    ws = rng.standard_normal(len(ref.saveval) if regf else 1).astype(np.float32)
    loss = (logpx * torch.from_numpy(w_l).cuda()).sum()
    if B > 592:
        hd = next(iter(ff._handles.values()))
        assert hd.lib.rnde_kernel_variant(hd.h) == 6      # RNDE_KERNEL_CHAIN8
    if kinetic:
        loss = loss + (l1 * torch.from_numpy(w1).cuda()).sum() + (l2 * torch.from_numpy(w2).cuda()).sum()
    if regf:
        loss = loss + (sv.saveval * torch.from_numpy(ws).cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    # the same cotangent on the augmented final state: d logpz/dz = -z
    du = np.zeros((Dz + extra, B), np.float32)
    du[:Dz] = -ref.u[:Dz] * w_l
    du[Dz] = -w_l
    if kinetic:
        du[Dz + 1], du[Dz + 2] = w1, w2
    dp_hi, dx_hi, _, _ = o.backward(du, ws if regf else None, hi=True, first_dt_tracked=False)
    dp_32, dx_32, _, _ = o.backward(du, ws if regf else None, first_dt_tracked=False)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    e_p, e_x = rel(p.grad.cpu().numpy(), dp_hi), rel(x.grad.cpu().numpy(), dx_hi[:Dz])
    c_p, c_x = rel(dp_32, dp_hi), rel(dx_32[:Dz], dx_hi[:Dz])
    print(f"ffjord grad Dz={Dz} H={H} B={B} regf={regf} kinetic={kinetic}: e_p {e_p:.2e} (cpu32 {c_p:.2e})  e_x {e_x:.2e} (cpu32 {c_x:.2e})  |dp| {np.abs(dp_hi).max():.2e}")
    assert np.abs(dp_hi).max() > 0 and np.abs(dx_hi).max() > 0
    assert e_p <= max(1e-4, sys_path_tests.GRAD_BAR * c_p) and e_x <= max(1e-4, sys_path_tests.GRAD_BAR * c_x), (e_p, c_p, e_x, c_x)


def test_sample_inverts_the_flow(oracle_built):
    """sample (src/models/ffjord.jl:160-167) integrates the flow backwards: x -> z(1) by the forward functor, then z(1) -> x by
    `sample` -- the round trip returns the data to the accuracy of two adaptive solves."""
    import regneuralde.jl_b200 as R
    rng = np.random.default_rng(31)
    Dz, H, B = 43, 100, 64
    p = torch.from_numpy(F.glorot_params(rng, Dz, H, dtype=np.float32, bias_scale=0.1)).cuda()
    x = torch.from_numpy(rng.standard_normal((Dz, B)).astype(np.float32)).cuda()
    ff = R.TrackedFFJORD(R.CSQDynamics(Dz, H), [0.0, 1.0], True, False, R.Tsit5(), reltol=1e-6, abstol=1e-6, tape_capacity=64)
    # z(1): the forward functor returns logpx only, so recover z from the oracle-checked stepper through the C ABI state
    with torch.no_grad():
        ff(x, p, torch.zeros(Dz, B, device="cuda"))
    o = orc.Oracle(orc.OracleConfig(D=Dz + 1, H=H, B=B, csq_extra=1, csq_noise=np.zeros((Dz, B), np.float32), kblock1=Dz + 1, abstol=1e-6, reltol=1e-6))
    ref = o.forward(np.concatenate([x.cpu().numpy(), np.zeros((1, B), np.float32)], 0), p.cpu().numpy())
    z1 = torch.from_numpy(np.ascontiguousarray(ref.u[:Dz])).cuda()
    xr = R.ffjord.sample(ff, Dz, p, z=z1)
    assert xr.shape == (Dz, B)
    assert float((xr - x).abs().max()) <= 1e-4 * float(x.abs().max()), float((xr - x).abs().max())
    assert float((z1 - x).abs().max()) > 1e-2          # the flow does move the points


def test_adam_weight_decay_update_matches_flux_rule():
    """Optimiser(WeightDecay(1e-5), ADAM(1e-2)) through update_parameters! (ffjord_tabular.jl:128): three updates against the
    Flux 0.11.6 apply! rules restated in numpy Float32."""
    import regneuralde.jl_b200 as R
    rng = np.random.default_rng(3)
    p0 = rng.standard_normal(1000).astype(np.float32)
    p = torch.from_numpy(p0.copy()).cuda()
    opt = R.ADAMOptimiser(1e-5, 1e-2)
    pr, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    bp = np.array([0.9, 0.999], np.float32)
    f = np.float32
    for it in range(3):
        g = rng.standard_normal(1000).astype(np.float32)
        R.update_parameters_((p,), (torch.from_numpy(g).cuda(),), opt)
        d = g + f(1e-5) * pr
        m = f(0.9) * m + (f(1) - f(0.9)) * d
        v = f(0.999) * v + (f(1) - f(0.999)) * d * d
        pr = pr - m / (f(1) - bp[0]) / (np.sqrt(v / (f(1) - bp[1])) + f(1e-8)) * f(1e-2)
        bp = bp * np.array([0.9, 0.999], np.float32)
        assert np.abs(p.cpu().numpy() - pr).max() <= 2e-6
