"""torch-autograd restatement of the reference path (TEST INFRASTRUCTURE ONLY).

PARITY UNPINNED (no Julia here; the reference's tests hold no golden vectors).

What it restates, with plain tensor ops so that ``torch.autograd`` differentiates
straight through every solver operation -- the analogue of Tracker +
``SensitivityADPassThrough()`` (src/models/neural_ode.jl:67,98,134,170):

* src/models/basic.jl:16-28, experiments/mnist_node.jl:51-54 -- time-concatenated
  2-layer field;
* src/models/neural_ode.jl:110-144 -- solve + SavingCallback + (res, nfe, sv);
* OrdinaryDiffEq 5.50.0 Tsit5 / PI controller / initial dt / AutoSwitch as
  recalled in SURVEY.md Appendix A.1-A.8 (un-vendored; Manifest.toml:964).

Its job is to validate the hand-written adjoint of oracle/rnde_oracle.c: in
float64 the two must agree on trajectory, step sequence, saved values and all
gradients (tests/test_oracle_agreement.py).  ``detach`` selects Appendix A.6:
"all" (frozen step sequence) or "all_but_first" (initial dt tracked).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch

A = {
    2: [0.161],
    3: [-0.008480655492356989, 0.335480655492357],
    4: [2.8971530571054935, -6.359448489975075, 4.3622954328695815],
    5: [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525],
    6: [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383],
    7: [0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774],
}
BT = [-0.00178001105222577714, -0.0008164344596567469, 0.007880878010261995, -0.1447110071732629,
      0.5823571654525552, -0.45808210592918697, 0.015151515151515152]
CS = {2: 0.161, 3: 0.327, 4: 0.9, 5: 0.9800255409045097, 6: 1.0, 7: 1.0}
STAB = 3.5068

REG_NONE, REG_ERR_DT, REG_STIFF_DT_ABS, REG_STIFF_SCALED, REG_ERR_PLUS_STIFF = range(5)


@dataclass
class TorchResult:
    u: torch.Tensor
    nf: int
    naccept: int
    nreject: int
    saveval: list
    dt_log: list
    accept_log: list
    eest_log: list
    dt_list: list      # tensors: dt of every accepted step
    t_list: list


def unpack(p, D, H, time_dep=True):
    td = 1 if time_dep else 0
    o = 0
    W1 = p[o:o + H * (D + td)].reshape(D + td, H).T; o += H * (D + td)
    b1 = p[o:o + H]; o += H
    W2 = p[o:o + D * (H + td)].reshape(H + td, D).T; o += D * (H + td)
    b2 = p[o:o + D]
    return W1, b1, W2, b2


def field(p, z, t, D, H, act2_tanh=True, time_dep=True):
    W1, b1, W2, b2 = unpack(p, D, H, time_dep)
    s = W1[:, :D] @ z
    if time_dep:
        s = s + W1[:, D:D + 1] * t
    h = torch.tanh(s + b1[:, None])
    y = W2[:, :H] @ h
    if time_dep:
        y = y + W2[:, H:H + 1] * t
    y = y + b2[:, None]
    return torch.tanh(y) if act2_tanh else y


def chain_field(p, z, D, widths, acts, pre_act):
    """Chain([x -> tanh.(x),] Dense(D,w0,act0), ...)(z) with p in Flux.destructure order (latent_ode.jl:109-121)."""
    a = torch.tanh(z) if pre_act else z
    o, K = 0, D
    for M, act in zip(widths, acts):
        W = p[o:o + M * K].reshape(K, M).T; o += M * K
        b = p[o:o + M]; o += M
        a = W @ a + b[:, None]
        if act:
            a = torch.tanh(a)
        K = M
    return a


R = [[1.0, -2.763706197274826, 2.9132554618219126, -1.0530884977290216],
     [0.13169999999999998, -0.2234, 0.1017], [3.9302962368947516, -5.941033872131505, 2.490627285651253],
     [-12.411077166933676, 30.33818863028232, -16.548102889244902], [37.50931341651104, -88.1789048947664, 47.37952196281928],
     [-27.896526289197286, 65.09189467479366, -34.87065786149661], [1.5, -4.0, 2.5]]


def interp_weights(th):
    """Tsit5 free interpolant b_i(theta)."""
    b = [th * (R[0][0] + th * (R[0][1] + th * (R[0][2] + th * R[0][3])))]
    for i in range(1, 7):
        b.append(th * th * (R[i][0] + th * (R[i][1] + th * R[i][2])))
    return b


def _f(v):
    """value of a tensor (or number) for control flow -- never part of the autograd graph"""
    return float(v.detach()) if torch.is_tensor(v) else float(v)


def rms(x):
    return torch.sqrt(torch.sum(x * x) / x.numel())


def saved_value(kind, EEst, eig, dt, dtype):
    stab = 1.0 / float(np.float32(STAB))
    if kind == REG_ERR_DT:
        return EEst * dt
    if kind == REG_STIFF_DT_ABS:
        return torch.abs(eig * dt)
    if kind == REG_STIFF_SCALED:
        s = torch.abs(eig)
        return stab * (s * 0 if (_f(s) == 0 or math.isnan(_f(s))) else s)
    if kind == REG_ERR_PLUS_STIFF:
        e = EEst * dt
        a = e * 0 if (_f(e) == 0 or math.isnan(_f(e))) else e
        b = eig * 0 if (_f(eig) == 0 or math.isnan(_f(eig))) else eig
        return (a + (float(np.float32(0.1)) * stab) * b) * 1.0
    raise ValueError(kind)


def solve(x, p, *, D, H, act2_tanh=True, time_dep=True, t0=0.0, t1=1.0, abstol=1.4e-8, reltol=1.4e-8,
          auto_tsit5=False, reg_kind=REG_NONE, detach="all", forced_dt=None, forced_accept=None,
          dt_leaf=None, max_steps=100000, saveat=None, chain=None, rhs=None) -> TorchResult:
    """x: (D,B) tensor, p: flat parameter tensor (Flux.destructure order).

    forced_dt/forced_accept replay a recorded attempt sequence (controller
    bypassed).  dt_leaf (tensor, one entry per attempt) lets a test take the
    gradient with respect to every dt: when given it supplies the attempt dts and
    t advances by them, so autograd sees both the direct and the time-shift paths.
    """
    dtype = x.dtype
    f = lambda z, t: field(p, z, t, D, H, act2_tanh, time_dep)
    if chain is not None:       # (widths, acts, pre_act): field without time input
        f = lambda z, t: chain_field(p, z, D, chain[0], chain[1], chain[2])
    if rhs is not None:         # any (z, t) -> dz/dt on the stepper's state (oracle/ffjord_oracle.py: the augmented CNF state)
        f = rhs
    c = lambda v: torch.tensor(v, dtype=dtype)
    t = c(t0)
    tf = c(t1)
    dtmax = t1 - t0
    gamma, qmin, qmax, beta1, beta2, qoldinit = 0.9, 0.2, 10.0, 7.0 / 50.0, 2.0 / 25.0, 1e-4
    qold, q11 = qoldinit, 1.0
    u = x
    saveval = []
    if reg_kind != REG_NONE:
        saveval.append(saved_value(reg_kind, c(1.0), c(1.0), c(0.0), dtype))
    usave = []
    save_idx = 0
    if saveat is not None:
        while save_idx < len(saveat) and float(saveat[save_idx]) <= t0:
            usave.append(u); save_idx += 1
    k1 = f(u, t)
    nf = 1
    forced = forced_dt is not None
    if forced:
        dt = dt_leaf[0] if dt_leaf is not None else c(float(forced_dt[0]))
    else:
        # Hairer-Wanner initial dt (Appendix A.5)
        sk = abstol + torch.abs(u) * reltol
        d0 = rms(u / sk)
        d1 = rms(k1 / sk)
        if float(d0.detach()) < 1e-5 or float(d1.detach()) < 1e-5:
            dt0 = c(1e-6)
        else:
            dt0 = (d0 / d1) / 100
        dt0 = torch.minimum(dt0, c(dtmax))
        u1 = u + dt0 * k1
        f1 = f(u1, t + dt0)
        d2 = rms((f1 - k1) / sk) / dt0
        md = torch.maximum(d1, d2)
        if _f(md.detach()) <= 1e-15:
            dt1 = torch.maximum(c(1e-6), dt0 * 1e-3)
        else:
            dt1 = 10.0 ** (-(2 + torch.log10(md)) / 5)
        dt = torch.minimum(torch.minimum(100 * dt0, dt1), c(dtmax))
        if detach == "all":
            dt = dt.detach()
    nf += 2
    dt_init = _f(dt)
    dt_log, accept_log, eest_log, dt_list, t_list = [], [], [], [], []
    naccept = nreject = 0
    as_count, as_stiff = 0, False
    eig_prev = 1.0
    it = 0
    accept_prev = True
    dtpropose = dt
    while _f(t) < t1:
        if it >= max_steps:
            raise RuntimeError("maxiters")
        if it > 0:
            if accept_prev:
                dt = dtpropose
            elif not forced:
                dt = dt / min(1 / qmin, q11 / gamma)
        it += 1
        if auto_tsit5 and not forced:
            stiffness = abs(eig_prev * _f(dt) / STAB)
            stiff = stiffness > 0.9
            as_count = (1 if as_count < 0 else as_count + 1) if stiff else (-1 if as_count > 0 else as_count - 1)
            if (not as_stiff) and as_count > 10:
                dt = dt * 2; as_stiff = True; nf += 1
            elif as_stiff and as_count < -3:
                dt = dt / 2; as_stiff = False; nf += 1
        if forced:
            dt = dt_leaf[it - 1] if dt_leaf is not None else c(float(forced_dt[it - 1]))
        else:
            dt = torch.minimum(dt, c(dtmax))
            rem = tf - t
            if _f(rem) < _f(dt):
                dt = rem if detach != "all" else rem.detach()
        ks = [None, k1]
        zs = {}
        for i in range(2, 8):
            acc = A[i][0] * ks[1]
            for j in range(2, i):
                acc = acc + A[i][j - 1] * ks[j]
            z = u + dt * acc
            zs[i] = z
            ti = t + CS[i] * dt
            ks.append(f(z, ti))
        unew = zs[7]
        if auto_tsit5:
            eig = rms(ks[7] - ks[6]) / rms(unew - zs[6])
        else:
            eig = c(1.0)
        acc = BT[0] * ks[1]
        for j in range(2, 8):
            acc = acc + BT[j - 1] * ks[j]
        utilde = dt * acc
        atmp = utilde / (abstol + torch.maximum(torch.abs(u), torch.abs(unew)) * reltol)
        EEst = rms(atmp)
        nf += 6
        ee = float(EEst.detach())
        if math.isnan(ee):
            raise FloatingPointError("NaN EEst")
        if ee == 0:
            q = 1 / qmax
        else:
            q11 = ee ** beta1
            q = q11 / (qold ** beta2)
            q = max(1 / qmax, min(1 / qmin, q / gamma))
        accept = bool(forced_accept[it - 1]) if forced else ee <= 1.0
        dt_log.append(_f(dt)); accept_log.append(int(accept)); eest_log.append(ee)
        if auto_tsit5:
            eig_prev = _f(eig)
        if accept:
            naccept += 1
            qold = max(ee, qoldinit)
            dtnew = _f(dt) / q
            dt_list.append(dt); t_list.append(t)
            tprev = t
            t = t + dt
            while saveat is not None and save_idx < len(saveat) and _f(saveat[save_idx]) <= _f(t):
                tau = float(saveat[save_idx])
                if tau == _f(t):
                    usave.append(unew)
                else:
                    th = (tau - _f(tprev)) / _f(dt)       # theta is not differentiated (frozen step sequence)
                    bw = interp_weights(th)
                    acc = bw[0] * ks[1]
                    for j in range(2, 8):
                        acc = acc + bw[j - 1] * ks[j]
                    usave.append(u + dt * acc)
                save_idx += 1
            dtpropose = c(min(dtmax, dtnew))     # DiffEqBase.value(...): detached (Appendix A.6)
            u = unew
            k1 = ks[7]
            if reg_kind != REG_NONE:
                saveval.append(saved_value(reg_kind, EEst, eig, dt, dtype))
        else:
            nreject += 1
        accept_prev = accept
    res = TorchResult(u=u, nf=nf, naccept=naccept, nreject=nreject, saveval=saveval, dt_log=dt_log,
                      accept_log=accept_log, eest_log=eest_log, dt_list=dt_list, t_list=t_list)
    res.dt_init = dt_init
    res.usave = usave
    return res
