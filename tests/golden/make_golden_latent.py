"""Golden fixtures of the Latent-ODE row (regression pins of the oracles, like make_golden.py -- the reference holds none):
  latent_chain_saveat.npz  generator dynamics Chain(tanh, Dense(20,50,tanh), ... x8) + 12 irregular saveat times, B=24,
                           error-estimate regulariser: saved states, saved values, counts, FP64-cotangent gradients
  latent_gru.npz           LatentGRU(5, 6, 4) on a 7-step sequence, B=5: output and gradient (FP64 torch restatement)
Run from the repo root:  python tests/golden/make_golden_latent.py"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import gru_oracle as G, orc  # noqa: E402

W = (50, 20, 50, 20, 50, 20, 50, 20)


def main():
    out = Path(__file__).resolve().parent
    rng = np.random.default_rng(1999)
    D, B = 20, 24
    p = orc.glorot_chain_params(rng, D, W, bias_scale=0.05)
    x = rng.standard_normal((D, B)).astype(np.float32)
    sa = np.unique(np.concatenate([[0.0], np.sort(rng.random(10)), [1.0]]).astype(np.float32))
    o = orc.Oracle(orc.OracleConfig(D=D, H=50, B=B, reg_kind=orc.REG_ERR_DT, kblock1=D, widths=W, acts=(1,) * 8, pre_act=1, saveat=sa.astype(np.float64)))
    r = o.forward(x, p)
    w = rng.standard_normal(r.usave.shape).astype(np.float32)
    ws = rng.standard_normal(len(r.saveval)).astype(np.float32)
    dp, dx, _, _ = o.backward(np.zeros((D, B), np.float32), ws, hi=True, dusave=w)
    np.savez_compressed(out / "latent_chain_saveat.npz", p=p, x=x, saveat=sa, usave=r.usave, saveval=r.saveval,
                        counts=np.array([r.nf, r.naccept, r.nreject]), w=w, ws=ws, dp_hi=dp, dx_hi=np.ascontiguousarray(dx))
    print("latent_chain_saveat nf", r.nf, "naccept", r.naccept)

    I, H, L, T, Bg = 5, 6, 4, 7, 5
    pg = G.glorot_params(rng, I, H, L, bias_scale=0.1)
    xg = rng.standard_normal((2 * I + 1, T, Bg)).astype(np.float32)
    xg[I:2 * I] = rng.random((I, T, Bg)) < 0.3
    xg[2 * I] = rng.random((T, Bg)) * 0.1
    xg[I:, 2, :] = 0.0
    wg = rng.standard_normal((2 * L, Bg)).astype(np.float32)
    p64 = torch.tensor(pg, dtype=torch.float64, requires_grad=True)
    ref = G.forward(p64, torch.tensor(xg, dtype=torch.float64), I, H, L)
    (ref * torch.tensor(wg, dtype=torch.float64)).sum().backward()
    np.savez_compressed(out / "latent_gru.npz", dims=np.array([I, H, L, T, Bg]), p=pg, x=xg, w=wg, out64=ref.detach().numpy(), dp64=p64.grad.numpy())
    print("latent_gru ok")


if __name__ == "__main__":
    main()
