"""GPU parity tests of the Latent-ODE row: the recognition RNN kernel against the torch
restatement of experiments/latent_ode.jl:39-99, and the assembled LatentTimeSeriesModel training loss
(time_series.jl:36-70, latent_ode.jl:212-262) with its gradient chain GRU <- rec_to_gen <- ODE solve <- decoder.
Tolerances (BASELINE.json north_star): values <= 1e-5 relative, gradients <= 1e-4 relative."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

GRAD_BAR = 1.5      # at most 1.5x the CPU Float32 adjoint's own error where that exceeds 1e-4

from oracle import gru_oracle as G  # noqa: E402  (the checker)
from oracle import orc              # noqa: E402


def R():
    import regneuralde.jl_b200 as r
    return r


def physionet_like(rng, I, T, B):
    """data, mask, time row shaped like the PhysioNet batches (latent_ode.jl:226-233): sparse observations, some
    time points with no observation at all, time row = observation times in [0, 1]."""
    data = rng.standard_normal((I, T, B)).astype(np.float32)
    mask = (rng.random((I, T, B)) < 0.15).astype(np.float32)
    mask[:, rng.random(T) < 0.2, :] = 0.0
    times = np.sort(rng.random(T)).astype(np.float32); times[0] = 0.0
    trow = np.broadcast_to(times[None, :, None], (1, T, B)).astype(np.float32).copy()
    return data * mask, mask, trow, times


@pytest.mark.parametrize("I,H,L,T,B", [(37, 40, 50, 49, 512), (41, 40, 50, 12, 7), (3, 5, 4, 6, 9)])
def test_latent_gru_matches_restatement(I, H, L, T, B):
    r = R()
    rng = np.random.default_rng(3)
    p_np = G.glorot_params(rng, I, H, L, bias_scale=0.05)
    data, mask, trow, _ = physionet_like(rng, I, T, B)
    x_np = np.concatenate([data, mask, trow], 0)
    if T > 6:
        x_np[I:, 3, :] = 0.0            # a step whose mask+time rows are all zero: the state must pass through unchanged
    gru = r.LatentGRU(I, H, L)
    assert gru.p.numel() == p_np.size == G.n_params(I, H, L)
    p = torch.from_numpy(p_np).cuda().requires_grad_(True)
    out = gru(torch.from_numpy(x_np).cuda(), p)
    w = rng.standard_normal((2 * L, B)).astype(np.float32)
    (out * torch.from_numpy(w).cuda()).sum().backward()
    torch.cuda.synchronize()
    p64 = torch.tensor(p_np, dtype=torch.float64, requires_grad=True)
    ref = G.forward(p64, torch.tensor(x_np, dtype=torch.float64), I, H, L)
    (ref * torch.tensor(w, dtype=torch.float64)).sum().backward()
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert tuple(out.shape) == (2 * L, B)
    assert rel(out.detach().cpu().numpy(), ref.detach().numpy()) <= 1e-5
    assert rel(p.grad.cpu().numpy(), p64.grad.numpy()) <= 1e-4


def test_latent_ode_training_loss_and_gradients(oracle_built):
    r = R()
    rng = np.random.default_rng(11)
    I, T, B = 37, 49, 512
    data, mask, trow, times = physionet_like(rng, I, T, B)
    saveat = np.unique(times)
    gen = torch.Generator().manual_seed(5)
    model = r.latent_ode_model(I, 40, 50, 20, 50, saveat=saveat.tolist(), regularize=True, solver=r.Tsit5(), generator=gen)
    S = len(saveat)
    data, mask, trow = data[:, :S], mask[:, :S], trow[:, :S]          # np.unique may have merged equal times
    ps = [p.clone().requires_grad_(True) for p in model.trainable()]
    sample = torch.from_numpy(rng.standard_normal((20, B)).astype(np.float32)).cuda()
    d, m, t = (torch.from_numpy(a).cuda() for a in (data, mask, trow))
    lam = 1.0e3                                                       # lambda_r0 of error_est (latent_ode.jl:158)
    total, nfe, parts = r.loss_function(d, m, t, model, *ps, func=r.ERROR_ESTIMATE, regularize=True, lam_r=lam, sample=sample)
    total.backward()
    torch.cuda.synchronize()
    # ---- the same computation from the restatements (FP64 glue, C oracle for the solve on the CUDA path's own z0) ----
    p1, p2, p3, p4 = (p.detach().cpu().double().requires_grad_(True) for p in ps)
    x64 = torch.tensor(np.concatenate([data, mask, trow], 0), dtype=torch.float64)
    enc_in = G.forward(p1, x64, I, 40, 50)
    from regneuralde.jl_b200.latent import _dense_chain, kl_divergence, log_likelihood
    out = _dense_chain(p2, enc_in, model.enc)
    mu0, logvar = out[:20], out[20:]
    z0 = sample.cpu().double() * torch.exp(logvar / 2) + mu0
    with torch.no_grad():   # the CUDA path's own z0 (FP32) drives the oracle solve, so that the step sequence is the same
        out32 = _dense_chain(ps[1].detach(), model.rnn(torch.from_numpy(np.concatenate([data, mask, trow], 0)).cuda(), ps[0].detach()), model.enc)
        z0_32 = (sample * torch.exp(out32[20:] / 2) + out32[:20]).cpu().numpy()
    assert np.abs(z0_32 - z0.detach().numpy()).max() <= 1e-5 * np.abs(z0_32).max()
    W = (50, 20, 50, 20, 50, 20, 50, 20)
    cfg = orc.OracleConfig(D=20, H=50, B=B, reg_kind=orc.REG_ERR_DT, kblock1=20, widths=W, acts=(1,) * 8, pre_act=1, saveat=saveat.astype(np.float64))
    o = orc.Oracle(cfg)
    ref = o.forward(z0_32, ps[2].detach().cpu().numpy())
    assert nfe == ref.nf
    res = torch.tensor(ref.usave.transpose(1, 0, 2), dtype=torch.float64, requires_grad=True)          # D x S x B
    sv = torch.tensor(ref.saveval, dtype=torch.float64, requires_grad=True)
    result = _dense_chain(p4, res.reshape(20, S * B), model.dec).reshape(-1, S, B)
    m64, d64 = torch.tensor(mask, dtype=torch.float64), torch.tensor(data, dtype=torch.float64)
    ll = log_likelihood(result * m64 - d64 * m64, m64)
    kl = kl_divergence(mu0, logvar)
    total_ref = -(ll - kl).mean() + lam * sv.mean()
    assert abs(float(total) - float(total_ref)) <= 1e-5 * abs(float(total_ref))
    # gradient chain: dL/dres, dL/dsv -> oracle adjoint of the solve -> dL/dz0 -> rec_to_gen and the GRU
    g_res, g_sv = torch.autograd.grad(total_ref, [res, sv], retain_graph=True)
    dus = g_res.numpy().transpose(1, 0, 2).astype(np.float32)
    dp3_hi, dz0_hi, _, _ = o.backward(np.zeros((20, B), np.float32), g_sv.numpy().astype(np.float32), hi=True, dusave=dus)
    dp3_32, dz0_32, _, _ = o.backward(np.zeros((20, B), np.float32), g_sv.numpy().astype(np.float32), dusave=dus)
    total_ref.backward(retain_graph=True)
    z0.backward(torch.tensor(dz0_hi, dtype=torch.float64))
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    c3 = rel(dp3_32, dp3_hi)                                          # conditioning of the regulariser gradient in FP32
    e1, e2 = rel(ps[0].grad.cpu().numpy(), p1.grad.numpy()), rel(ps[1].grad.cpu().numpy(), p2.grad.numpy())
    e3, e4 = rel(ps[2].grad.cpu().numpy(), dp3_hi), rel(ps[3].grad.cpu().numpy(), p4.grad.numpy())
    assert e4 <= 1e-4, e4
    assert e3 <= max(1e-4, GRAD_BAR * c3), (e3, c3)
    cz = rel(dz0_32, dz0_hi)
    assert e1 <= max(1e-4, GRAD_BAR * cz) and e2 <= max(1e-4, GRAD_BAR * cz), (e1, e2, cz)
