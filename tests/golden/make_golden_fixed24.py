"""Golden fixture of the FIXED24 arithmetic (oracle arith = 1, the exact tensor-core forward stepper): a regression pin of the
oracle like the other fixtures (the reference holds none).  MNIST-shaped field, batch 12, AutoTsit5, combined regulariser.
Run from the repo root:  python tests/golden/make_golden_fixed24.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import orc  # noqa: E402


def main():
    D, H, B = 784, 100, 12
    rng = np.random.default_rng(1999)
    p = orc.glorot_params(rng, D, H)
    x = rng.random((D, B), dtype=np.float32)
    o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_TANH, alg=orc.ALG_AUTO_TSIT5, reg_kind=orc.REG_ERR_PLUS_STIFF, kblock1=98, arith=1))
    r = o.forward(x, p)
    k, _ = o.rhs(p, x, 0.25)
    steps = np.array(r.steps)
    np.savez_compressed(Path(__file__).resolve().parent / "fixed24_mnist_b12.npz", x=x, u=np.ascontiguousarray(r.u), saveval=r.saveval,
                        counts=np.array([r.nf, r.naccept, r.nreject]), dt=steps[:, 1].astype(np.float32), k_t025=np.ascontiguousarray(k),
                        p_sum=np.array([np.float64(p.astype(np.float64).sum()), np.float64(np.abs(p).astype(np.float64).sum())]))
    print("fixed24_mnist_b12 nf", r.nf, "naccept", r.naccept)


if __name__ == "__main__":
    main()
