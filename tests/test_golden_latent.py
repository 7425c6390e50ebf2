"""Golden fixtures of the Latent-ODE row (tests/golden/make_golden_latent.py).  CPU: the oracles reproduce them
(chain field + saveat: bit for bit; GRU restatement: to FP64 round-off).  GPU: the CUDA path reproduces the chain fixture bit for
bit (forward) and both gradients / the GRU within the stated tolerances."""
from pathlib import Path

import numpy as np
import pytest

from oracle import gru_oracle as G, orc

GOLD = Path(__file__).resolve().parent / "golden"
W = (50, 20, 50, 20, 50, 20, 50, 20)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def test_oracle_reproduces_chain_saveat_golden(oracle_built):
    g = np.load(GOLD / "latent_chain_saveat.npz")
    D, B = g["x"].shape
    o = orc.Oracle(orc.OracleConfig(D=D, H=50, B=B, reg_kind=orc.REG_ERR_DT, kblock1=D, widths=W, acts=(1,) * 8, pre_act=1,
                                    saveat=g["saveat"].astype(np.float64)))
    r = o.forward(g["x"], g["p"])
    assert [r.nf, r.naccept, r.nreject] == list(g["counts"])
    assert np.array_equal(bits(r.usave), bits(g["usave"])) and np.array_equal(bits(r.saveval), bits(g["saveval"]))
    dp, dx, _, _ = o.backward(np.zeros((D, B), np.float32), g["ws"], hi=True, dusave=g["w"])
    assert np.allclose(dp, g["dp_hi"], rtol=1e-6, atol=1e-9) and np.allclose(dx, g["dx_hi"], rtol=1e-6, atol=1e-9)


def test_gru_restatement_reproduces_golden():
    import torch
    g = np.load(GOLD / "latent_gru.npz")
    I, H, L, T, B = [int(v) for v in g["dims"]]
    p = torch.tensor(g["p"], dtype=torch.float64, requires_grad=True)
    out = G.forward(p, torch.tensor(g["x"], dtype=torch.float64), I, H, L)
    (out * torch.tensor(g["w"], dtype=torch.float64)).sum().backward()
    assert np.allclose(out.detach().numpy(), g["out64"], rtol=1e-12, atol=1e-14)
    assert np.allclose(p.grad.numpy(), g["dp64"], rtol=1e-10, atol=1e-13)


@pytest.mark.gpu
def test_cuda_reproduces_latent_goldens():
    torch = pytest.importorskip("torch")
    import regneuralde.jl_b200 as R
    g = np.load(GOLD / "latent_chain_saveat.npz")
    D, B = g["x"].shape
    layers, K = [], D
    for M in W:
        layers.append(R.Dense(K, M, "tanh")); K = M
    node = R.TrackedNeuralODE(R.Chain("tanh", *layers), [0.0, 1.0], False, True, R.Tsit5(), saveat=g["saveat"].tolist(), reltol=1.4e-8, abstol=1.4e-8)
    x = torch.from_numpy(g["x"]).cuda().requires_grad_(True); p = torch.from_numpy(g["p"]).cuda().requires_grad_(True)
    res, nfe, sv = node(x, p, func=R.ERROR_ESTIMATE)
    st = node.last_stats
    assert [nfe, st.naccept, st.nreject] == list(g["counts"])
    assert np.array_equal(bits(res.detach().permute(1, 0, 2).cpu().numpy()), bits(g["usave"]))
    assert np.array_equal(bits(sv.saveval.detach().cpu().numpy()), bits(g["saveval"]))
    ((res * torch.from_numpy(np.ascontiguousarray(g["w"].transpose(1, 0, 2))).cuda()).sum() + (sv.saveval * torch.from_numpy(g["ws"]).cuda()).sum()).backward()
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(x.grad.cpu().numpy(), g["dx_hi"]) <= 2e-2 and rel(p.grad.cpu().numpy(), g["dp_hi"]) <= 2e-2      # regularised: FP32 conditioning
    gg = np.load(GOLD / "latent_gru.npz")
    I, H, L, T, Bg = [int(v) for v in gg["dims"]]
    gru = R.LatentGRU(I, H, L)
    pg = torch.from_numpy(gg["p"]).cuda().requires_grad_(True)
    out = gru(torch.from_numpy(gg["x"]).cuda(), pg)
    (out * torch.from_numpy(gg["w"]).cuda()).sum().backward()
    assert rel(out.detach().cpu().numpy(), gg["out64"]) <= 1e-5 and rel(pg.grad.cpu().numpy(), gg["dp64"]) <= 1e-4
