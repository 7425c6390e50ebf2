"""Generates tests/golden/*.npz from the C oracle (FP32 canonical arithmetic).

The reference holds NO golden vectors (test/test_node.jl is @code_warntype on unseeded input) and Julia is
not available, so these are REGRESSION PINS of the oracle, not outputs of the reference: they freeze the
canonical arithmetic so that neither the oracle nor the CUDA kernels can drift unnoticed.
Run from the repo root:  python tests/golden/make_golden.py
Inputs use numpy.random.default_rng(1999) (seed of experiments/configs/mnist_node.yml:2) and are stored too.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import orc  # noqa: E402

CASES = {
    # name: D, H, B, act2, alg, reg, kblock [, arith]   (arith 2 = SPLITK, the order of csrc/fwd4s_kernel.cuh; default 0 = FMA_CHAIN)
    "test_node_unreg": (2, 10, 1, orc.ACT_ID, orc.ALG_TSIT5, orc.REG_NONE, 0),
    "test_node_errreg": (2, 10, 1, orc.ACT_ID, orc.ALG_TSIT5, orc.REG_ERR_DT, 0),
    "test_node_stiffreg": (2, 10, 1, orc.ACT_ID, orc.ALG_AUTO_TSIT5, orc.REG_STIFF_DT_ABS, 0),
    "toy_b7": (2, 10, 7, orc.ACT_ID, orc.ALG_TSIT5, orc.REG_ERR_DT, 0),
    "latent_sized": (20, 50, 37, orc.ACT_TANH, orc.ALG_AUTO_TSIT5, orc.REG_ERR_PLUS_STIFF, 8),
    "mnist_b16": (784, 100, 16, orc.ACT_TANH, orc.ALG_TSIT5, orc.REG_ERR_DT, 98),
    "mnist_b16_splitk": (784, 100, 16, orc.ACT_TANH, orc.ALG_AUTO_TSIT5, orc.REG_ERR_PLUS_STIFF, 98, 2),
}


def main():
    out = Path(__file__).resolve().parent
    only = sys.argv[1:]
    for name, case in CASES.items():
        if only and name not in only:
            continue
        D, H, B, act2, alg, reg, kb = case[:7]
        arith = case[7] if len(case) > 7 else 0
        rng = np.random.default_rng(1999)
        p = orc.glorot_params(rng, D, H)
        x = rng.random((D, B), dtype=np.float32)
        o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, act2=act2, alg=alg, reg_kind=reg, kblock1=kb, arith=arith))
        r = o.forward(x, p)
        w = rng.standard_normal((D, B)).astype(np.float32)
        ws = rng.standard_normal(max(len(r.saveval), 1)).astype(np.float32)
        dp, dx, _, _ = o.backward(w, ws, hi=True)
        steps = np.array(r.steps, dtype=np.float64)
        kw = dict(cfg=np.array([D, H, B, act2, alg, reg, kb]), x=x, u=np.ascontiguousarray(r.u), saveval=r.saveval,
                  counts=np.array([r.nf, r.naccept, r.nreject]), dt=steps[:, 1].astype(np.float32), eest=steps[:, 2].astype(np.float32),
                  w=w, ws=ws, dx_hi=np.ascontiguousarray(dx))
        if arith:
            kw["arith"] = np.array(arith)
        if D * H < 5000:
            kw["p"] = p
            kw["dp_hi"] = dp
        else:   # large parameter vectors are regenerated from the seed; a checksum guards the generator
            kw["p_sum"] = np.array([np.float64(p.astype(np.float64).sum()), np.float64(np.abs(p).astype(np.float64).sum())])
            kw["dp_hi_head"] = dp[:4096]
        np.savez_compressed(out / f"{name}.npz", **kw)
        print(name, "nf", r.nf, "naccept", r.naccept, "bytes", (out / f"{name}.npz").stat().st_size)


if __name__ == "__main__":
    main()
