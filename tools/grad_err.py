"""Gradient error of the reverse sweep at the FLAGSHIP training loss (experiments/mnist_node.jl:132-152: logitcrossentropy +
lambda * mean(sv.saveval), lambda = 100, batch 512 unless given): e = |g - g_hi|_max / |g_hi|_max for the CUDA sweep selected by
the environment (default tensor cores; RNDE_BWD_FFMA=1 the FFMA sweep) and c = the same for the CPU Float32 adjoint, g_hi = the
Float64-cotangent adjoint over the same Float32 forward (oracle/).  Also the regulariser part and the cross-entropy part alone.
Prints one JSON line (bench.py --grad-check and tests/test_gpu_parity.py use the same routine).

    python tools/grad_err.py [B] ; RNDE_BWD_FFMA=1 python tools/grad_err.py [B]
"""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import regneuralde.jl_b200 as r
from oracle import orc


def flagship_grad_errors(B=512, lam=100.0, seed=1999, func_name="ERROR_ESTIMATE", auto=False):
    D, H, Cn = 784, 100, 10
    rng = np.random.default_rng(seed)
    p2 = orc.glorot_params(rng, D, H)
    s3 = np.sqrt(6.0 / (D + Cn))
    W3 = rng.uniform(-s3, s3, size=(Cn, D)).astype(np.float32)
    p3 = np.concatenate([W3.flatten(order="F"), np.zeros(Cn, np.float32)])
    x = rng.random((D, B), dtype=np.float32)
    y = np.zeros((Cn, B), np.float32); y[rng.integers(0, Cn, B), np.arange(B)] = 1
    func = getattr(r, func_name)
    node = r.TrackedNeuralODE(r.MLPDynamics(D, H), [0.0, 1.0], True, True, r.AutoTsit5() if auto else r.Tsit5(), save_everystep=False,
                              reltol=1.4e-8, abstol=1.4e-8, save_start=False, tape_capacity=128)
    clf = r.ClassifierNODE(None, node, r.Dense(D, Cn))
    clf.p2.copy_(torch.from_numpy(p2)); clf.p3.copy_(torch.from_numpy(p3)); node.p = clf.p2
    xd, yd = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    agg = "maximum" if func_name == "STIFFNESS_SCALED" else "mean"
    o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, act2=orc.ACT_TANH, reg_kind=func.kind, kblock1=D // 8, alg=1 if auto else 0, arith=node.arith))
    ref = o.forward(x, p2)
    logits = W3 @ ref.u
    m = logits.max(0, keepdims=True)
    lse = m + np.log(np.exp(logits - m).sum(0, keepdims=True))
    du = (W3.T @ ((np.exp(logits - lse) - y) / B)).astype(np.float32)
    sv = ref.saveval.astype(np.float64)
    if agg == "mean":
        dsv = np.full(len(sv), lam / len(sv), np.float32)
    else:
        dsv = np.zeros(len(sv), np.float32); dsv[int(np.argmax(sv))] = lam
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    out = {"B": B, "lam": lam, "func": func_name, "sweep": "ffma" if "RNDE_BWD_FFMA" in os.environ else "tensor"}
    for part, (sdu, sdsv, l) in {"full": (du, dsv, lam), "ce": (du, 0 * dsv, 0.0), "reg": (0 * du, dsv, None)}.items():
        hi, _, _, _ = o.backward(sdu, sdsv, hi=True)
        c32, _, _, _ = o.backward(sdu, sdsv)
        if l is None:       # regulariser part alone: by linearity full - ce of the device gradients
            g = g_full - g_ce
        else:
            g = clf.loss_and_gradient(xd, yd, lam=l, func=func, agg=agg)["g2"].cpu().numpy()
            if part == "full": g_full = g
            else: g_ce = g
        out[part] = {"e": rel(g, hi), "c_cpu32": rel(c32, hi), "gmax": float(np.abs(hi).max())}
    out["nfe"] = int(ref.nf)
    return out


if __name__ == "__main__":
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    for fn, auto in (("ERROR_ESTIMATE", False), ("ERROR_PLUS_STIFFNESS", True)):
        print(json.dumps(flagship_grad_errors(B=B, func_name=fn, auto=auto)))
