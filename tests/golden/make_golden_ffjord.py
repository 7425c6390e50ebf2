"""Golden fixture of the FFJORD row (regression pin of oracle/ffjord_oracle.py -- the reference holds none,
and no kernel exists for this row yet: the file is what a future CUDA path will be compared with).
  ffjord_tabular.npz   MLPDynamics(43, 100) (ffjord_tabular.jl:109), B = 6, tspan [0, 1], both functors in Float64 at
                       reltol = abstol = 1.4e-8: log-densities, NFE, saved values, and the loss gradient of the regularised one
Run from the repo root:  python tests/golden/make_golden_ffjord.py"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ffjord_oracle as F  # noqa: E402


def main():
    out = Path(__file__).resolve().parent
    rng = np.random.default_rng(1999)
    D, H, B = 43, 100, 6
    p = F.glorot_params(rng, D, H, dtype=np.float64, bias_scale=0.05)
    x = rng.standard_normal((D, B))
    e = rng.standard_normal((D, B))
    tp, tx, te = torch.from_numpy(p).requires_grad_(True), torch.from_numpy(x), torch.from_numpy(e)
    total, r = F.loss_function(tx, tp, te, D=D, H=H, regularized_functor=True, lam_r=100.0)
    g, = torch.autograd.grad(total, tp)
    r0 = F.ffjord(tx, tp.detach(), te, D=D, H=H, regularized_functor=False, regularize=True)
    np.savez_compressed(out / "ffjord_tabular.npz", p=p, x=x, e=e, logpx=r.logpx.detach().numpy(), saveval=r.saveval.detach().numpy(),
                        counts=np.array([r.nfe, r.sol.naccept, r.sol.nreject]), loss=np.array(float(total.detach())), dp=g.numpy(),
                        logpx_kinetic=r0.logpx.numpy(), lam1=r0.lam1.numpy(), lam2=r0.lam2.numpy(),
                        counts_kinetic=np.array([r0.nfe, r0.sol.naccept, r0.sol.nreject]))
    print("ffjord_tabular: nfe", r.nfe, "naccept", r.sol.naccept, "loss", float(total.detach()), "| kinetic functor nfe", r0.nfe)


if __name__ == "__main__":
    main()
