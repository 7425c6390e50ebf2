"""CPU restatement of the reference's Neural-SDE path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED: the reference's tests hold no golden vectors and no Julia exists here.  What this file restates:

* src/models/neural_sde.jl:84-146          TrackedNeuralDSDE functors: SDEProblem{false}(drift, diffusion, x, tspan, p),
                                           two NFE counters (:46,:50), SavingCallback(func, sv), return tuple
* experiments/mnist_nsde.jl:44-84          drift Chain(Dense(32,64,tanh), Dense(64,32)), DIAGONAL diffusion Dense(32,32),
                                           SOSRI() / AutoSOSRI2(SOSRI2()), reltol = abstol = 1.4f-1, func = EEst*dt
* src/models/supervised_classification.jl:50-103   ClassifierNSDE: trajectory replication, pre/post nets, mean over trajectories
* StochasticDiffEq 6.30.1 (Manifest.toml:1262-1266, author's fork, un-vendored), recalled:
    - the four-stage Roessler SRI step (FourStageSRIConstantCache perform_step!) with the SOSRI / SOSRI2 tableaus of
      Rackauckas & Nie, "Stability-optimized high order methods ... for stiff SDEs" (2018); the coefficients below are
      checked against Roessler's strong-order-1.5 conditions in tests/test_sde_oracle.py, which any wrong digit breaks;
    - EEst = norm((delta*E1 + E2) / (abstol + max(|uprev|,|u|)*reltol)), delta = 1/26, RMS norm over all entries;
    - PI controller with beta2 = 2/(5*order), beta1 = 7/(10*order), order = 3/2, gamma = 9/10, qmin = 1/5, qmax = 9/8,
      qoldinit = 1e-4; dtnew detached on accept (the DiffEqBase.value(...) of A.6);
    - sde_determine_initdt (Hairer-style with f0 +- 3*g0);
* DiffEqNoiseProcess 5.5.1 (Manifest.toml:268), recalled: WienerProcess with the RSwM3 rejection-sampling-with-memory
  stacks (Rackauckas & Nie 2017, Algorithm 3): dW = sqrt(dt)*xi, bridge(q) = q*W + sqrt((1-q)*q*h)*xi, futures stack S1,
  re-use stack S2.  UNVERIFIED details are marked below.

The reference's random stream (Julia's MersenneTwister through randn!) cannot be reproduced, so the noise is INJECTED: a
`NormalStream` hands out one (D, B) array of standard normals per draw, in the order the solver asks for them; the CUDA
stepper consumes the same array in the same order (parity "with supplied noise").
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ---- tableaus (StochasticDiffEq constructSOSRI / constructSOSRI2) ---------------------------------------------------
SOSRI = dict(
    a021=-0.04199224421316468, a031=2.842612915017106, a032=-2.0527723684000727, a041=4.338237071435815, a042=-2.8895936137439793,
    a043=2.3017575594644466, a121=0.26204282091330466, a131=0.20903646383505375, a132=-0.1502377115150942, a141=0.05836595312746999,
    a142=0.6149440396332373, a143=0.08535117634046772, b021=-0.21641093549612528, b031=1.5336352863679572, b032=0.26066223492647056,
    b041=-1.0536037558179159, b042=1.7015284721089472, b043=-0.20725685784180017, b121=-0.5119011827621657, b131=2.67767339866713,
    b132=-4.9395031322250995, b141=0.15580956238299215, b142=3.2361551006624674, b143=-1.4223118283355949,
    al1=1.140099274172029, al2=-0.6401334255743456, al3=0.4736296532772559, al4=0.026404498125060714,
    c02=-0.04199224421316468, c03=0.7898405466170333, c04=3.7504010171562823, c11=0.0, c12=0.26204282091330466, c13=0.05879875232001766,
    c14=0.758661169101175,
    beta11=-1.8453464565104432, beta12=2.688764531100726, beta13=-0.2523866501071323, beta14=0.40896857551684956,
    beta21=0.4969658141589478, beta22=-0.5771202869753592, beta23=-0.12919702470322217, beta24=0.2093514975196336,
    beta31=2.8453464565104425, beta32=-2.688764531100725, beta33=0.2523866501071322, beta34=-0.40896857551684945,
    beta41=0.11522663875443433, beta42=-0.57877086147738, beta43=0.2857851028163886, beta44=0.17775911990655704)
SOSRI2 = dict(
    a021=0.13804532298278663, a031=0.5818361298250374, a032=0.4181638701749618, a041=0.4670018408674211, a042=0.8046204792187386,
    a043=-0.27162232008616016, a121=0.45605532163856893, a131=0.7555807846451692, a132=0.24441921535482677, a141=0.6981181143266059,
    a142=0.3453277086024727, a143=-0.04344582292908241, b021=0.08852381537667678, b031=1.0317752458971061, b032=0.4563552922077882,
    b041=1.73078280444124, b042=-0.46089678470929774, b043=-0.9637509618944188, b121=0.6753186815412179, b131=-0.07452812525785148,
    b132=-0.49783736486149366, b141=-0.5591906709928903, b142=0.022696571806569924, b143=-0.8984927888368557,
    al1=-0.15036858140642623, al2=0.7545275856696072, al3=0.686995463807979, al4=-0.2911544680711602,
    c02=0.13804532298278663, c03=0.9999999999999992, c04=0.9999999999999994, c11=0.0, c12=0.45605532163856893, c13=0.999999999999996,
    c14=0.9999999999999962,
    beta11=-0.45315689727309133, beta12=0.8330937231303951, beta13=0.3792843195533544, beta14=0.24077885458934192,
    beta21=-0.4994383733810986, beta22=0.9181786186154077, beta23=-0.25613778661003145, beta24=-0.16260245862427797,
    beta31=1.4531568972730915, beta32=-0.8330937231303933, beta33=-0.3792843195533583, beta34=-0.24077885458934023,
    beta41=-0.4976090683622265, beta42=0.9148155835648892, beta43=-1.4102107084476505, beta44=0.9930041932449877)

ALG_SOSRI, ALG_AUTO_SOSRI2 = 0, 1
REG_NONE, REG_ERR_DT, REG_STIFF_SCALED = 0, 1, 3
SOSRI2_STABILITY_SIZE = 10.6          # StochasticDiffEq.alg_stability_size(SOSRI2())  -- UNVERIFIED (recalled)


def tableau_matrices(tab):
    """(A0, B0, A1, B1, alpha, beta1..4, c0, c1) in Roessler's notation (strictly lower triangular 4 x 4)."""
    A0 = np.zeros((4, 4)); B0 = np.zeros((4, 4)); A1 = np.zeros((4, 4)); B1 = np.zeros((4, 4))
    for i in range(2, 5):
        for j in range(1, i):
            A0[i - 1, j - 1] = tab[f"a0{i}{j}"]; A1[i - 1, j - 1] = tab[f"a1{i}{j}"]
            B0[i - 1, j - 1] = tab[f"b0{i}{j}"]; B1[i - 1, j - 1] = tab[f"b1{i}{j}"]
    al = np.array([tab[f"al{i}"] for i in range(1, 5)])
    be = [np.array([tab[f"beta{k}{i}"] for i in range(1, 5)]) for k in range(1, 5)]
    c0 = np.array([0.0, tab["c02"], tab["c03"], tab["c04"]]); c1 = np.array([tab["c11"], tab["c12"], tab["c13"], tab["c14"]])
    return A0, B0, A1, B1, al, be, c0, c1


class NormalStream:
    """Injected noise: draw k returns normals[k] (an array shaped like the state)."""

    def __init__(self, normals: np.ndarray):
        self.z = normals
        self.k = 0

    def draw(self):
        if self.k >= len(self.z):
            raise RuntimeError("noise stream exhausted")
        v = self.z[self.k]
        self.k += 1
        return v


class RSwM3:
    """Wiener increments (dW, dZ) for the step the solver is about to take, with rejection sampling with memory
    (DiffEqNoiseProcess RSwM3, recalled).  S1: futures (pieces of already-sampled path AHEAD of the current step, top = next
    in time); S2: the pieces the CURRENT step was assembled from (top = last in time)."""

    def __init__(self, stream: NormalStream, dtype):
        self.s, self.dtype = stream, dtype
        self.S1, self.S2 = [], []
        self.dt = None; self.dW = None; self.dZ = None
        self.discard = 1e-15

    def _fresh(self, h):
        sq = self.dtype(math.sqrt(abs(float(h))))
        return sq * self.s.draw().astype(self.dtype), sq * self.s.draw().astype(self.dtype)

    def _bridge(self, q, h, W, Z):
        sq = self.dtype(math.sqrt((1.0 - float(q)) * float(q) * abs(float(h))))
        qd = self.dtype(q)
        return qd * W + sq * self.s.draw().astype(self.dtype), qd * Z + sq * self.s.draw().astype(self.dtype)

    def setup(self, dt):
        """increments for a step of size dt starting where the last accepted step ended (accept_step! -> calculate_step!)"""
        self.S2 = []
        dt = float(dt)
        if not self.S1:
            self.dW, self.dZ = self._fresh(dt)
            self.S2.append((dt, self.dW, self.dZ))
        else:
            dttmp = 0.0
            dW = None; dZ = None
            add = lambda a, b: b if a is None else a + b
            while self.S1:
                L1, L2, L3 = self.S1.pop()
                qtmp = (dt - dttmp) / L1
                if qtmp > 1:
                    dttmp += L1; dW = add(dW, L2); dZ = add(dZ, L3)
                    self.S2.append((L1, L2, L3))
                else:       # popped too far: bridge inside the piece, keep its remainder as a future
                    bW, bZ = self._bridge(qtmp, L1, L2, L3)
                    dW = add(dW, bW); dZ = add(dZ, bZ)
                    if (1 - qtmp) * L1 > self.discard:
                        self.S1.append(((1 - qtmp) * L1, L2 - bW, L3 - bZ))
                    if qtmp * L1 > self.discard:
                        self.S2.append((qtmp * L1, bW, bZ))
                    dttmp = dt
                    break
            left = dt - dttmp
            if left > 0:      # the stack ran out before dt was covered
                fW, fZ = self._fresh(left)
                dW = add(dW, fW); dZ = add(dZ, fZ)
                self.S2.append((left, fW, fZ))
            self.dW, self.dZ = dW, dZ
        self.dt = dt

    def reject(self, dtnew):
        """the attempt of size self.dt was rejected; shrink to dtnew re-using the sampled path (reject_step!)"""
        dtnew = float(dtnew)
        q = dtnew / self.dt
        dttmp = 0.0; dWtmp = None; dZtmp = None
        add = lambda a, b: b if a is None else a + b
        # move whole pieces from the end of the step to the futures while they lie entirely beyond dtnew
        while self.S2:
            L1, L2, L3 = self.S2.pop()
            if dttmp + L1 < (1 - q) * self.dt:
                dttmp += L1; dWtmp = add(dWtmp, L2); dZtmp = add(dZtmp, L3)
                self.S1.append((L1, L2, L3))
            else:
                self.S2.append((L1, L2, L3))
                break
        dtK = self.dt - dttmp
        K2 = self.dW if dWtmp is None else self.dW - dWtmp
        K3 = self.dZ if dZtmp is None else self.dZ - dZtmp
        qK = q * self.dt / dtK
        bW, bZ = self._bridge(qK, dtK, K2, K3)
        cut = (1 - qK) * dtK
        if cut > self.discard:
            self.S1.append((cut, K2 - bW, K3 - bZ))
        # UNVERIFIED: upstream keeps the leading pieces of S2; the re-use stack of the shrunken step is rebuilt here as one piece
        self.S2 = [(dtnew, bW, bZ)]
        self.dt, self.dW, self.dZ = dtnew, bW, bZ


@dataclass
class SdeResult:
    u: np.ndarray
    nfe1: int
    nfe2: int
    naccept: int
    nreject: int
    saveval: np.ndarray
    dts: list = field(default_factory=list)          # dt of every attempt
    accepted: list = field(default_factory=list)
    eests: list = field(default_factory=list)
    draws: int = 0
    dt_init: float = 0.0
    steps: list = field(default_factory=list)        # accepted steps: (dt, dW, dZ, EEst) -- what the discrete adjoint replays


def rms(x):
    return math.sqrt(float(np.mean(np.square(x.astype(np.float64))))) if x.dtype == np.float64 else float(np.sqrt(np.mean(np.square(x), dtype=x.dtype)))


def drift_diffusion(p, dtype, D=32, H=64):
    """the experiment's networks (mnist_nsde.jl:73-74) from the concatenated parameter vector vcat(p1, p2) (neural_sde.jl:16-18)"""
    p = np.asarray(p, dtype=dtype)
    o = 0
    W1 = p[o:o + H * D].reshape(D, H).T; o += H * D
    b1 = p[o:o + H]; o += H
    W2 = p[o:o + D * H].reshape(H, D).T; o += D * H
    b2 = p[o:o + D]; o += D
    Wg = p[o:o + D * D].reshape(D, D).T; o += D * D
    bg = p[o:o + D]; o += D
    assert o == p.size
    f = lambda u: W2 @ np.tanh(W1 @ u + b1[:, None]) + b2[:, None]
    g = lambda u: Wg @ u + bg[:, None]
    return f, g


def solve(x, f, g, normals, *, alg=ALG_SOSRI, reg_kind=REG_NONE, t0=0.0, t1=1.0, abstol=0.14, reltol=0.14, dtype=np.float32,
          max_steps=100000, forced_dt=None):
    """Adaptive SOSRI / SOSRI2 solve of du = f(u) dt + g(u) dW (diagonal noise) from x over [t0, t1].
    normals: (ndraws, D, B) standard normals.  Returns the final state and the bookkeeping of the reference's functor."""
    tab = SOSRI if alg == ALG_SOSRI else SOSRI2
    T = {k: dtype(v) for k, v in tab.items()}
    c = dtype
    u = np.asarray(x, dtype=dtype)
    stream = NormalStream(np.asarray(normals))
    W = RSwM3(stream, dtype)
    nfe = [0, 0]

    def F(v):
        nfe[0] += 1
        return f(v).astype(dtype)

    def G(v):
        nfe[1] += 1
        return g(v).astype(dtype)

    order = 1.5
    beta2, beta1 = 2.0 / (5.0 * order), 7.0 / (10.0 * order)
    gamma, qmin, qmax, qoldinit, delta = 0.9, 0.2, 9.0 / 8.0, 1e-4, 1.0 / 26.0
    dtmax = t1 - t0
    t = float(t0)
    saveval = [0.0] if reg_kind != REG_NONE else []          # SavingCallback at initialisation: EEst = 1, dt = 0 (A.7)
    if reg_kind == REG_STIFF_SCALED:
        saveval = [1.0 / SOSRI2_STABILITY_SIZE]

    # ---- sde_determine_initdt ----
    if forced_dt is None:
        sk = c(abstol) + np.abs(u) * c(reltol)
        d0 = rms(u / sk)
        f0 = F(u); g0 = c(3) * G(u)
        d1 = rms(np.maximum(np.abs(f0 + g0), np.abs(f0 - g0)) / sk)
        dt0 = 1e-6 if (d0 < 1e-5 or d1 < 1e-5) else 0.01 * (d0 / d1)
        dt0 = min(dt0, dtmax)
        u1 = u + c(dt0) * f0
        f1 = F(u1); g1 = c(3) * G(u1)
        dgmax = np.maximum(np.abs(g0 - g1), np.abs(g0 + g1))
        d2 = rms(np.maximum(np.abs(f1 - f0 + dgmax), np.abs(f1 - f0 - dgmax)) / sk) / dt0
        md = max(d1, d2)
        dt1 = max(1e-6, dt0 * 1e-3) if md <= 1e-15 else 10.0 ** (-(2 + math.log10(md)) / (order + 0.5))
        dt = min(100 * dt0, dt1, dtmax)
    else:
        dt = float(forced_dt[0])
    dt = float(dtype(dt))
    res = SdeResult(u=u, nfe1=0, nfe2=0, naccept=0, nreject=0, saveval=None, dt_init=dt)
    qold, q11 = qoldinit, 1.0
    dt = min(dt, t1 - t)
    W.setup(dt)
    it = 0
    while t < t1:
        if it >= max_steps:
            raise RuntimeError("maxiters")
        it += 1
        dtc, sqdt = c(dt), c(math.sqrt(dt))
        dW, dZ = W.dW, W.dZ
        chi1 = (dW * dW - dtc) / (c(2) * sqdt)
        chi2 = (dW + dZ / c(math.sqrt(3.0))) / c(2)
        chi3 = (dW * dW * dW - c(3) * dW * dtc) / (c(6) * dtc)
        k1 = F(u); g1 = G(u)
        H01 = u + dtc * T["a021"] * k1 + T["b021"] * chi2 * g1
        H11 = u + dtc * T["a121"] * k1 + sqdt * T["b121"] * g1
        k2 = F(H01); g2 = G(H11)
        H02 = u + dtc * (T["a031"] * k1 + T["a032"] * k2) + chi2 * (T["b031"] * g1 + T["b032"] * g2)
        H12 = u + dtc * (T["a131"] * k1 + T["a132"] * k2) + sqdt * (T["b131"] * g1 + T["b132"] * g2)
        k3 = F(H02); g3 = G(H12)
        H03 = u + dtc * (T["a041"] * k1 + T["a042"] * k2 + T["a043"] * k3) + chi2 * (T["b041"] * g1 + T["b042"] * g2 + T["b043"] * g3)
        H13 = u + dtc * (T["a141"] * k1 + T["a142"] * k2 + T["a143"] * k3) + sqdt * (T["b141"] * g1 + T["b142"] * g2 + T["b143"] * g3)
        k4 = F(H03); g4 = G(H13)
        E2 = chi2 * (T["beta31"] * g1 + T["beta32"] * g2 + T["beta33"] * g3 + T["beta34"] * g4) + \
            chi3 * (T["beta41"] * g1 + T["beta42"] * g2 + T["beta43"] * g3 + T["beta44"] * g4)
        unew = u + dtc * (T["al1"] * k1 + T["al2"] * k2 + T["al3"] * k3 + T["al4"] * k4) + E2 + \
            dW * (T["beta11"] * g1 + T["beta12"] * g2 + T["beta13"] * g3 + T["beta14"] * g4) + \
            chi1 * (T["beta21"] * g1 + T["beta22"] * g2 + T["beta23"] * g3 + T["beta24"] * g4)
        E1 = dtc * (k1 + k2 + k3 + k4)
        resid = (c(delta) * E1 + E2) / (c(abstol) + np.maximum(np.abs(u), np.abs(unew)) * c(reltol))
        EEst = rms(resid)
        eig = 1.0
        if alg == ALG_AUTO_SOSRI2:          # UNVERIFIED (recalled): stiffness estimate of the composite algorithm, c03 = c04 = 1
            eig = rms(k4 - k3) / max(rms(H03 - H02), 1e-300)
        if math.isnan(EEst):
            raise FloatingPointError("NaN EEst")
        if EEst == 0:
            q = 1 / qmax
        else:
            q11 = EEst ** beta1
            q = q11 / (qold ** beta2)
            q = max(1 / qmax, min(1 / qmin, q / gamma))
        accept = True if forced_dt is not None else EEst <= 1.0
        res.dts.append(dt); res.accepted.append(accept); res.eests.append(EEst)
        if accept:
            res.naccept += 1
            res.steps.append((float(dt), np.array(dW, copy=True), np.array(dZ, copy=True), float(EEst)))
            qold = max(EEst, qoldinit)
            t = t + dt
            if abs(t1 - t) < 10 * np.finfo(dtype).eps * max(abs(t1), 1.0):
                t = t1
            u = unew
            if reg_kind == REG_ERR_DT:
                saveval.append(EEst * dt)
            elif reg_kind == REG_STIFF_SCALED:
                a = abs(eig)
                saveval.append((0.0 if (a == 0 or math.isnan(a)) else a) / SOSRI2_STABILITY_SIZE)
            if not (t < t1):
                break
            dtnew = dt / q if forced_dt is None else float(forced_dt[min(it, len(forced_dt) - 1)])
            dt = float(dtype(min(dtmax, dtnew)))
            dt = min(dt, t1 - t)
            W.setup(dt)
        else:
            res.nreject += 1
            dtnew = dt / min(1 / qmin, q11 / gamma)
            dtnew = float(dtype(dtnew))
            W.reject(dtnew)
            dt = dtnew
    res.u = u
    res.nfe1, res.nfe2 = nfe
    res.saveval = np.asarray(saveval, dtype=dtype)
    res.draws = stream.k
    return res


def classifier_nsde(x, p1, p2, p3, normals, *, trajectories=1, **kw):
    """ClassifierNSDE (supervised_classification.jl:82-103): replicate the batch `trajectories` times, Dense(784,32) pre-net,
    the SDE solve, Dense(32,10) post-net, mean over the trajectories.  x: (784, B).  Returns logits (10, B) and the solve."""
    dtype = kw.get("dtype", np.float32)
    B = x.shape[1]
    xr = np.tile(np.asarray(x, dtype=dtype), (1, trajectories))
    Wp = np.asarray(p1[: 32 * 784], dtype=dtype).reshape(784, 32).T; bp = np.asarray(p1[32 * 784:], dtype=dtype)
    f, g = drift_diffusion(p2, dtype)
    r = solve(Wp @ xr + bp[:, None], f, g, normals, **kw)
    Wq = np.asarray(p3[: 10 * 32], dtype=dtype).reshape(32, 10).T; bq = np.asarray(p3[10 * 32:], dtype=dtype)
    z = Wq @ r.u + bq[:, None]
    z = z.reshape(10, trajectories, B).mean(axis=1)
    return z, r


def replay_torch(x, p, steps, *, alg=ALG_SOSRI, reg_kind=REG_NONE, abstol=0.14, reltol=0.14, D=32, H=64):
    """The accepted steps of a solve replayed in torch (Float64) with the step sizes and the noise increments frozen: what
    Tracker.gradient through solve(SDEProblem, SOSRI(); sensealg = SensitivityADPassThrough()) differentiates
    (src/models/neural_sde.jl:84-146, experiments/mnist_nsde.jl:201-204; the proposed dt are detached like in the ODE path, the
    Wiener increments come from the untracked RNG).  x: (D, B) tensor, p: flat tensor (vcat(p_drift, p_diffusion)), both may
    require grad.  Returns (u_final, saved values as a tensor incl. the initial entry)."""
    import torch
    tab = SOSRI if alg == ALG_SOSRI else SOSRI2
    T = {k: float(v) for k, v in tab.items()}
    o = 0
    W1 = p[o:o + H * D].reshape(D, H).T; o += H * D
    b1 = p[o:o + H]; o += H
    W2 = p[o:o + D * H].reshape(H, D).T; o += D * H
    b2 = p[o:o + D]; o += D
    Wg = p[o:o + D * D].reshape(D, D).T; o += D * D
    bg = p[o:o + D]
    F = lambda u: W2 @ torch.tanh(W1 @ u + b1[:, None]) + b2[:, None]
    G = lambda u: Wg @ u + bg[:, None]
    u = x
    saved = [torch.zeros((), dtype=x.dtype)] if reg_kind != REG_NONE else []
    delta = 1.0 / 26.0
    for dt, dW, dZ, _ in steps:
        dW = torch.as_tensor(np.asarray(dW, dtype=np.float64)); dZ = torch.as_tensor(np.asarray(dZ, dtype=np.float64))
        sqdt = math.sqrt(dt)
        chi1 = (dW * dW - dt) / (2 * sqdt)
        chi2 = (dW + dZ / math.sqrt(3.0)) / 2
        chi3 = (dW * dW * dW - 3 * dW * dt) / (6 * dt)
        k1 = F(u); g1 = G(u)
        H01 = u + dt * T["a021"] * k1 + T["b021"] * chi2 * g1
        H11 = u + dt * T["a121"] * k1 + sqdt * T["b121"] * g1
        k2 = F(H01); g2 = G(H11)
        H02 = u + dt * (T["a031"] * k1 + T["a032"] * k2) + chi2 * (T["b031"] * g1 + T["b032"] * g2)
        H12 = u + dt * (T["a131"] * k1 + T["a132"] * k2) + sqdt * (T["b131"] * g1 + T["b132"] * g2)
        k3 = F(H02); g3 = G(H12)
        H03 = u + dt * (T["a041"] * k1 + T["a042"] * k2 + T["a043"] * k3) + chi2 * (T["b041"] * g1 + T["b042"] * g2 + T["b043"] * g3)
        H13 = u + dt * (T["a141"] * k1 + T["a142"] * k2 + T["a143"] * k3) + sqdt * (T["b141"] * g1 + T["b142"] * g2 + T["b143"] * g3)
        k4 = F(H03); g4 = G(H13)
        E2 = chi2 * (T["beta31"] * g1 + T["beta32"] * g2 + T["beta33"] * g3 + T["beta34"] * g4) + \
            chi3 * (T["beta41"] * g1 + T["beta42"] * g2 + T["beta43"] * g3 + T["beta44"] * g4)
        unew = u + dt * (T["al1"] * k1 + T["al2"] * k2 + T["al3"] * k3 + T["al4"] * k4) + E2 + \
            dW * (T["beta11"] * g1 + T["beta12"] * g2 + T["beta13"] * g3 + T["beta14"] * g4) + \
            chi1 * (T["beta21"] * g1 + T["beta22"] * g2 + T["beta23"] * g3 + T["beta24"] * g4)
        if reg_kind == REG_ERR_DT:
            E1 = dt * (k1 + k2 + k3 + k4)
            resid = (delta * E1 + E2) / (abstol + torch.maximum(torch.abs(u), torch.abs(unew)) * reltol)
            saved.append(torch.sqrt(torch.mean(resid * resid)) * dt)
        elif reg_kind == REG_STIFF_SCALED:
            eig = torch.sqrt(torch.mean((k4 - k3) ** 2)) / torch.sqrt(torch.mean((H03 - H02) ** 2))
            saved.append(torch.abs(eig) / SOSRI2_STABILITY_SIZE)
        u = unew
    return u, (torch.stack(saved) if saved else None)
