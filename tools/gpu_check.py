"""Developer diagnostic (GPU box): CUDA path vs the C oracle, with verbose output.
Usage: python tools/gpu_check.py [case ...]   cases: toy toyB mnist32 mnist512 stream bwd timing
"""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import orc  # noqa: E402  (diagnostic tool: oracle as checker)
import regneuralde.jl_b200 as R  # noqa: E402
from regneuralde.jl_b200 import _lib as L  # noqa: E402

VAR = {"auto": 0, "cta": 1, "stream": 2, "cluster": 3, "cluster4": 4}


def make(D, H, B, seed, act2, scale=1.0):
    rng = np.random.default_rng(seed)
    p = orc.glorot_params(rng, D, H) * np.float32(scale)
    x = rng.random((D, B), dtype=np.float32)
    return x, p


def run_cuda(D, H, B, x, p, act2, alg, reg, variant, need_bwd=False, kblock=0, cap=256):
    cfg = L.Config()
    cfg.struct_bytes = C.sizeof(L.Config)
    cfg.state_dim, cfg.hidden_dim, cfg.batch = D, H, B
    cfg.act_hidden, cfg.act_out, cfg.time_dep = 1, act2, 1
    cfg.kblock = kblock
    cfg.alg, cfg.reg_kind = alg, reg
    cfg.tape_capacity = cap
    cfg.need_backward = 1 if need_bwd else 0
    cfg.kernel_variant = variant
    cfg.t0, cfg.t1 = 0.0, 1.0
    cfg.abstol = cfg.reltol = float(np.float32(1.4e-8))
    lib = L.lib()
    h = C.c_void_p()
    rc = lib.rnde_create(C.byref(cfg), C.byref(h))
    if rc != 0:
        raise RuntimeError(f"create rc={rc}")
    xd = torch.from_numpy(np.asfortranarray(x).T.copy()).cuda().view(-1)   # column-major buffer
    pd = torch.from_numpy(p).cuda()
    ud = torch.empty(D * B, device="cuda")
    sv = torch.zeros(cap + 1, device="cuda")
    st = L.Stats()
    torch.cuda.synchronize()
    t0 = time.time()
    rc = lib.rnde_forward(h, xd.data_ptr(), pd.data_ptr(), ud.data_ptr(), sv.data_ptr(), C.byref(st), None)
    torch.cuda.synchronize()
    el = time.time() - t0
    if rc != 0:
        print("forward rc", rc, lib.rnde_last_error(h).decode())
    n = st.naccept
    arrs = [(C.c_float * max(n, 1))() for _ in range(4)]
    lib.rnde_get_steps(h, *arrs, n)
    steps = np.array([list(a)[:n] for a in arrs], dtype=np.float32)
    u = ud.cpu().numpy().reshape(B, D).T
    return dict(u=u, st=st, sv=sv.cpu().numpy()[: st.n_saved], steps=steps, h=h, lib=lib, time=el, variant=lib.rnde_kernel_variant(h),
                keep=(xd, pd, ud, sv))


def compare_fwd(name, D, H, B, act2, alg, reg, variant, seed=1999, kblock=0):
    x, p = make(D, H, B, seed, act2)
    kb = kblock if kblock else (D if D < 128 else (D + 7) // 8)
    o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, act2=act2, alg=alg, reg_kind=reg, kblock1=kb))
    t0 = time.time(); r = o.forward(x, p); to = time.time() - t0
    c = run_cuda(D, H, B, x, p, act2, alg, reg, variant, kblock=kblock)
    st = c["st"]
    same_u = np.array_equal(r.u.view(np.uint32), c["u"].view(np.uint32))
    print(f"[{name}] variant={c['variant']} cuda {c['time']*1e3:.2f} ms oracle {to*1e3:.1f} ms | nf {st.nf}/{r.nf} acc {st.naccept}/{r.naccept} "
          f"rej {st.nreject}/{r.nreject} rc {st.retcode} | u bit-equal {same_u} maxabs {np.abs(r.u - c['u']).max():.3e}")
    osteps = np.array(r.steps, dtype=np.float64).T if r.steps else np.zeros((4, 0))
    n = min(osteps.shape[1], c["steps"].shape[1])
    if n:
        dt_eq = np.array_equal(osteps[1, :n].astype(np.float32), c["steps"][1, :n])
        ee_eq = np.array_equal(osteps[2, :n].astype(np.float32), c["steps"][2, :n])
        print(f"    dt bit-equal {dt_eq}  EEst bit-equal {ee_eq}  dt_init {st.dt_init:.9g}/{r.dt_init:.9g}")
        if not (dt_eq and ee_eq):
            print("    oracle dt  ", osteps[1, :6]); print("    cuda   dt  ", c["steps"][1, :6])
            print("    oracle EEst", osteps[2, :6]); print("    cuda   EEst", c["steps"][2, :6])
    if reg:
        sv_eq = np.array_equal(r.saveval.view(np.uint32), c["sv"].view(np.uint32)) if len(r.saveval) == len(c["sv"]) else False
        print(f"    saveval bit-equal {sv_eq}  sum {c['sv'].sum():.9g}/{r.saveval.sum():.9g}")
    c["lib"].rnde_destroy(c["h"])
    return same_u


def compare_bwd(name, D, H, B, act2, alg, reg, variant, seed=7):
    x, p = make(D, H, B, seed, act2)
    kb = D if D < 128 else (D + 7) // 8
    o = orc.Oracle(orc.OracleConfig(D=D, H=H, B=B, act2=act2, alg=alg, reg_kind=reg, kblock1=kb))
    r = o.forward(x, p)
    rng = np.random.default_rng(seed + 1)
    w = rng.standard_normal((D, B)).astype(np.float32)
    ws = rng.standard_normal(max(len(r.saveval), 1)).astype(np.float32)
    t0 = time.time(); dp, dx, _, _ = o.backward(w, ws); to = time.time() - t0
    c = run_cuda(D, H, B, x, p, act2, alg, reg, variant, need_bwd=True)
    lib, h = c["lib"], c["h"]
    du = torch.from_numpy(np.asfortranarray(w).T.copy()).cuda().view(-1)
    dsv = torch.zeros(257, device="cuda"); dsv[: len(ws)] = torch.from_numpy(ws).cuda()
    dpd = torch.zeros(p.size, device="cuda"); dxd = torch.zeros(D * B, device="cuda")
    torch.cuda.synchronize(); t0 = time.time()
    rc = lib.rnde_backward(h, du.data_ptr(), dsv.data_ptr(), dpd.data_ptr(), dxd.data_ptr(), None)
    torch.cuda.synchronize(); tc = time.time() - t0
    if rc != 0:
        print("backward rc", rc, lib.rnde_last_error(h).decode())
    gdp = dpd.cpu().numpy(); gdx = dxd.cpu().numpy().reshape(B, D).T
    e_p = np.abs(gdp - dp).max() / max(np.abs(dp).max(), 1e-30)
    e_x = np.abs(gdx - dx).max() / max(np.abs(dx).max(), 1e-30)
    print(f"[{name}] variant={c['variant']} bwd cuda {tc*1e3:.2f} ms oracle {to*1e3:.1f} ms | fwd bit-equal "
          f"{np.array_equal(r.u.view(np.uint32), c['u'].view(np.uint32))} | dp relerr {e_p:.3e} dx relerr {e_x:.3e} (|dp|max {np.abs(dp).max():.3e})")
    lib.rnde_destroy(h)


def timing(B=512, variant=0, reg=1, alg=0, reps=5):
    D, H = 784, 100
    x, p = make(D, H, B, 1999, 1)
    c = run_cuda(D, H, B, x, p, 1, alg, reg, variant, need_bwd=True)
    lib, h = c["lib"], c["h"]
    xd, pd, ud, sv = c["keep"]
    du = torch.randn(D * B, device="cuda"); dsv = torch.full((257,), 0.01, device="cuda")
    dpd = torch.zeros(p.size, device="cuda"); dxd = torch.zeros(D * B, device="cuda")
    st = L.Stats()
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for _ in range(reps):
        e[0].record()
        lib.rnde_forward(h, xd.data_ptr(), pd.data_ptr(), ud.data_ptr(), sv.data_ptr(), C.byref(st), None)
        e[1].record()
        lib.rnde_backward(h, du.data_ptr(), dsv.data_ptr(), dpd.data_ptr(), dxd.data_ptr(), None)
        e[2].record()
        torch.cuda.synchronize()
        print(f"[timing B={B} variant={c['variant']}] fwd {e[0].elapsed_time(e[1]):.3f} ms  bwd {e[1].elapsed_time(e[2]):.3f} ms  nf {st.nf} acc {st.naccept}")
    lib.rnde_destroy(h)


if __name__ == "__main__":
    cases = sys.argv[1:] or ["toy", "toyB", "mid", "mnist32", "stream", "bwd", "mnist512", "timing"]
    print("device:", torch.cuda.get_device_name(0), "lib version", L.lib().rnde_version())
    for cs in cases:
        try:
            if cs == "toy":
                compare_fwd("toy B=1 errreg", 2, 10, 1, 0, 0, 1, VAR["auto"])
                compare_fwd("toy B=1 unreg", 2, 10, 1, 0, 0, 0, VAR["auto"])
                compare_fwd("toy B=1 stiff", 2, 10, 1, 0, 1, 2, VAR["auto"])
            elif cs == "toyB":
                compare_fwd("toy B=7", 2, 10, 7, 0, 0, 1, VAR["auto"])
                compare_fwd("toy B=512", 2, 10, 512, 0, 0, 1, VAR["auto"])
                compare_fwd("toy B=512 stream", 2, 10, 512, 0, 0, 1, VAR["stream"])
            elif cs == "mid":
                compare_fwd("mid D=20 H=50 B=100 auto/combined", 20, 50, 100, 1, 1, 4, VAR["auto"])
                compare_fwd("mid D=20 H=50 B=100 kblock=8", 20, 50, 100, 1, 0, 1, VAR["auto"], kblock=8)
            elif cs == "mnist32":
                compare_fwd("mnist B=32 cluster", 784, 100, 32, 1, 0, 1, VAR["cluster"])
                compare_fwd("mnist B=40 cluster auto-tsit5", 784, 100, 40, 1, 1, 4, VAR["cluster"])
            elif cs == "stream":
                compare_fwd("mnist B=32 stream", 784, 100, 32, 1, 0, 1, VAR["stream"])
            elif cs == "mnist512":
                compare_fwd("mnist B=512 cluster", 784, 100, 512, 1, 0, 1, VAR["cluster"])
            elif cs == "bwd":
                compare_bwd("bwd toy B=3", 2, 10, 3, 0, 0, 1, VAR["auto"])
                compare_bwd("bwd mid auto/combined", 20, 50, 100, 1, 1, 4, VAR["auto"])
                compare_bwd("bwd mid stiff", 20, 50, 37, 1, 1, 2, VAR["auto"])
                compare_bwd("bwd mnist B=32 cluster", 784, 100, 32, 1, 0, 1, VAR["cluster"])
                compare_bwd("bwd mnist B=32 stream", 784, 100, 32, 1, 0, 1, VAR["stream"])
            elif cs == "c4":
                compare_fwd("mnist B=16 cluster4", 784, 100, 16, 1, 0, 1, VAR["cluster4"])
                compare_fwd("mnist B=40 cluster4 auto-tsit5 combined", 784, 100, 40, 1, 1, 4, VAR["cluster4"])
                compare_fwd("mnist B=512 cluster4", 784, 100, 512, 1, 0, 1, VAR["cluster4"])
                compare_fwd("D=64 H=20 B=100 cluster4", 64, 20, 100, 1, 0, 1, VAR["cluster4"], kblock=8)
                compare_fwd("D=200 H=37 B=33 cluster4 stiff", 200, 37, 33, 0, 1, 2, VAR["cluster4"], kblock=25)
            elif cs == "c4time":
                x, p = make(784, 100, 512, 1999, 1)
                for _ in range(3):
                    c = run_cuda(784, 100, 512, x, p, 1, 0, 1, VAR["cluster4"])
                    print("cluster4 fwd B=512", c["time"] * 1e3, "ms nf", c["st"].nf, "variant", c["variant"])
                    c["lib"].rnde_destroy(c["h"])
            elif cs == "timeline":
                import os
                os.environ["RNDE_DEBUG_TIMELINE"] = "1"
                for (D_, H_, kb_) in ((784, 100, 0), (64, 20, 8)):
                    x, p = make(D_, H_, 512, 1999, 1)
                    c = run_cuda(D_, H_, 512, x, p, 1, 0, 1, VAR["cluster4"], kblock=kb_)
                    buf = (C.c_longlong * 8000)()
                    c["lib"].rnde_debug_timeline(c["h"], buf, 8000)
                    a = np.array(list(buf)).reshape(-1, 2)
                    a = a[: np.nonzero(a[:, 1])[0].max() + 1]
                    ids, ts = a[:, 0], a[:, 1]
                    # per-eval phase durations: average over evals 5..60
                    starts = np.nonzero(ids == 0)[0]
                    names = {1: "stageZ+sync", 2: "phaseA", 3: "syncA", 4: "pair+scatter", 5: "sync", 6: "waitP", 7: "phaseB", 8: "sync", 9: "waitH", 10: "phaseC"}
                    durs = {k: [] for k in names}
                    between = []
                    for si in range(5, min(60, len(starts) - 1)):
                        s0 = starts[si]
                        for k in range(1, 11):
                            durs[k].append(ts[s0 + k] - ts[s0 + k - 1])
                        between.append(ts[starts[si + 1]] - ts[s0 + 10])
                    print(f"timeline D={D_} H={H_}: total/eval {np.mean([ts[starts[i+1]]-ts[starts[i]] for i in range(5, min(60, len(starts)-1))]):.0f} cycles")
                    for k in range(1, 11):
                        print(f"   {names[k]:14s} {np.mean(durs[k]):8.0f}")
                    print(f"   between evals  {np.mean(between):8.0f}  (median {np.median(between):.0f})")
                    c["lib"].rnde_destroy(c["h"])
            elif cs == "prof":
                x, p = make(784, 100, 480, 1999, 1)
                for _ in range(2):
                    c = run_cuda(784, 100, 480, x, p, 1, 0, 1, VAR["cluster"])
                    print("prof fwd", c["time"] * 1e3, "ms nf", c["st"].nf)
                    c["lib"].rnde_destroy(c["h"])
            elif cs == "c4bwd":
                compare_bwd("bwd mnist B=32 cluster4", 784, 100, 32, 1, 0, 1, VAR["cluster4"])
                compare_bwd("bwd mnist B=40 cluster4 auto combined", 784, 100, 40, 1, 1, 4, VAR["cluster4"])
                compare_bwd("bwd mnist B=24 cluster4 stiff identity-out", 784, 100, 24, 0, 1, 2, VAR["cluster4"])
                compare_bwd("bwd mnist B=512 cluster4", 784, 100, 512, 1, 0, 1, VAR["cluster4"])
                timing(512, VAR["cluster4"], reps=4)
            elif cs == "prof4":
                timing(512, VAR["cluster4"], reps=2)
            elif cs == "timing":
                timing(512, VAR["cluster"])
                timing(512, VAR["stream"], reps=2)
        except Exception as ex:  # keep going: this is a diagnostic sweep
            import traceback
            traceback.print_exc()
            print(f"[{cs}] FAILED: {ex}")
