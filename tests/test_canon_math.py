"""Canonical arithmetic (include/regnde_canon.h) against libm, and the Tsit5 tableau identities
(SURVEY.md Appendix A.1 / A.9 self-checks).  CPU only."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent

SRC = r'''
#include "regnde_canon.h"
void t_tanh(const float* x, float* y, long n) { for (long i = 0; i < n; ++i) y[i] = canon_tanhf(x[i]); }
void t_powf(const float* x, float e, float* y, long n) { for (long i = 0; i < n; ++i) y[i] = canon_powf(x[i], e); }
void t_log10f(const float* x, float* y, long n) { for (long i = 0; i < n; ++i) y[i] = canon_log10f(x[i]); }
void t_exp10(const double* x, double* y, long n) { for (long i = 0; i < n; ++i) y[i] = canon_exp10(x[i]); }
void t_sigmoid(const float* x, float* y, long n) { for (long i = 0; i < n; ++i) y[i] = canon_sigmoidf(x[i]); }
void t_softplus(const float* x, float* y, long n) { for (long i = 0; i < n; ++i) y[i] = canon_softplusf(x[i]); }
void t_expneg(const float* x, float* y, long n) { for (long i = 0; i < n; ++i) y[i] = canon_expnegf(x[i]); }
'''


@pytest.fixture(scope="module")
def canon(tmp_path_factory):
    d = tmp_path_factory.mktemp("canon")
    (d / "c.c").write_text(SRC)
    so = d / "c.so"
    subprocess.run(["/usr/bin/gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", "-mfma", f"-I{ROOT / 'include'}", str(d / "c.c"),
                    "-o", str(so), "-lm"], check=True)
    return C.CDLL(str(so))


def _call(fn, x, out_dtype, *extra):
    y = np.zeros(x.shape, dtype=out_dtype)
    fn(x.ctypes.data_as(C.c_void_p), *extra, y.ctypes.data_as(C.c_void_p), C.c_long(x.size))
    return y


def test_tanh_within_3ulp_and_odd(canon):
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-10, 10, 400000), rng.normal(0, 1e-3, 100000), np.linspace(-0.7, 0.7, 100001),
                        [0.0, -0.0, 9.0, 9.01, 20.0, -20.0, 1e-30, 1e-40]]).astype(np.float32)
    y = _call(canon.t_tanh, x, np.float32)
    ref = np.tanh(x.astype(np.float64))
    ulp = np.spacing(np.abs(ref).astype(np.float32)).astype(np.float64)
    err = np.abs(y.astype(np.float64) - ref) / ulp
    assert err.max() < 3.0, err.max()
    assert np.array_equal(_call(canon.t_tanh, -x, np.float32), -y)
    assert np.all(np.abs(y) <= 1.0)
    assert np.isnan(_call(canon.t_tanh, np.array([np.nan], np.float32), np.float32))[0]


def test_tanh_monotone_near_zero(canon):
    x = np.linspace(0, 0.01, 20001).astype(np.float32)
    y = _call(canon.t_tanh, x, np.float32)
    assert np.all(np.diff(y) >= 0)


def test_pow_log_match_libm_rounded(canon):
    rng = np.random.default_rng(1)
    x = np.exp(rng.uniform(-25, 5, 300000)).astype(np.float32)
    for e in (np.float32(0.14), np.float32(0.08)):
        y = _call(canon.t_powf, x, np.float32, C.c_float(float(e)))
        ref = np.power(x.astype(np.float64), np.float64(e)).astype(np.float32)
        assert np.array_equal(y, ref)
    y = _call(canon.t_log10f, x, np.float32)
    assert np.array_equal(y, np.log10(x.astype(np.float64)).astype(np.float32))
    z = rng.uniform(-8, 2, 100000)
    ye = _call(canon.t_exp10, z, np.float64)
    assert np.max(np.abs(ye / 10.0 ** z - 1)) < 1e-13


def test_tableau_identities():
    from oracle import torch_oracle as T
    c = {2: 0.161, 3: 0.327, 4: 0.9, 5: 0.9800255409045097, 6: 1.0, 7: 1.0}
    for i, row in T.A.items():
        assert abs(sum(row) - c[i]) < 1e-14, i
    assert abs(sum(T.BT)) < 1e-15
    b = np.array(T.A[7] + [0.0])
    cs = np.array([0.0] + [c[i] for i in range(2, 8)])
    for k in range(1, 5):      # quadrature conditions of the 5th-order solution
        assert abs(np.sum(b * cs ** k) - 1.0 / (k + 1)) < 1e-13
    bhat = b - np.array(T.BT)  # embedded 4th-order weights
    for k in range(0, 4):
        assert abs(np.sum(bhat * cs ** k) - 1.0 / (k + 1)) < 1e-13


def _ulp_err(y32, ref64):
    """error of Float32 results against a Float64 reference in units of the reference's Float32 ulp"""
    ref32 = ref64.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    return np.abs(y32.astype(np.float64) - ref64) / ulp


def _grid():
    rng = np.random.default_rng(3)
    mags = np.concatenate([np.exp(rng.uniform(np.log(1e-30), np.log(120.0), 1 << 21)), rng.uniform(0, 20, 1 << 21),
                           np.array([0.0, 1e-45, 1e-38, 0.41421354, 0.41421357, 0.6931472, 1.0, 87.0, 88.0, 103.9, 104.0, 150.0, 1e30])])
    x = np.concatenate([mags, -mags]).astype(np.float32)
    return x


def test_expneg_sigmoid_softplus_against_libm(canon):
    """FFJORD activations (ffjord_tabular.jl:39-45) in canonical arithmetic: within 2.5 ulp of Float64 libm over 8 M points
    covering every binade, the branch points of the reductions and the subnormal tail."""
    x = _grid()
    xd = x.astype(np.float64)
    a = np.abs(x)
    e = _call(canon.t_expneg, a, np.float32)
    ok = a <= 87.0                                                      # normal results
    assert _ulp_err(e[ok], np.exp(-a[ok].astype(np.float64))).max() <= 1.5
    sub = (a > 87.0) & (a <= 103.9)
    assert np.all(np.abs(e[sub].astype(np.float64) - np.exp(-a[sub].astype(np.float64))) <= 1.5 * 1.4012984643e-45)   # subnormal: absolute
    assert np.all(e[a > 104.0] == e[a == 104.0][0]) and e[a == 104.0][0] == 0.0
    s = _call(canon.t_sigmoid, x, np.float32)
    sref = np.where(xd >= 0, 1 / (1 + np.exp(-np.abs(xd))), np.exp(-np.abs(xd)) / (1 + np.exp(-np.abs(xd))))
    big = a <= 87.0
    assert _ulp_err(s[big], sref[big]).max() <= 2.5            # exp (<= 1.5) + the rounding of 1 + t + the division
    assert np.all((s >= 0) & (s <= 1)) and s[x == 0][0] == 0.5
    p = _call(canon.t_softplus, x, np.float32)
    pref = np.maximum(xd, 0) + np.log1p(np.exp(-np.abs(xd)))
    assert _ulp_err(p[big], pref[big]).max() <= 2.0
    assert np.all(p >= 0) and np.all(p[x > 104] == x[x > 104])
    # NaN propagates, infinities saturate
    sp = np.array([np.nan, np.inf, -np.inf], np.float32)
    s3, p3 = _call(canon.t_sigmoid, sp, np.float32), _call(canon.t_softplus, sp, np.float32)
    assert np.isnan(s3[0]) and np.isnan(p3[0]) and s3[1] == 1.0 and s3[2] == 0.0 and p3[1] == np.inf and p3[2] == 0.0


def test_sigmoid_softplus_monotone(canon):
    x = np.sort(np.random.default_rng(4).uniform(-30, 30, 1 << 20).astype(np.float32))
    s = _call(canon.t_sigmoid, x, np.float32); p = _call(canon.t_softplus, x, np.float32)
    # correctly-rounded-ish, not exactly monotone: no inversion larger than 2 ulp
    ds, dp = np.diff(s.astype(np.float64)), np.diff(p.astype(np.float64))
    assert ds.min() >= -2 * np.spacing(np.float32(1.0)) and np.all(dp >= -2 * np.spacing(np.abs(p[1:])).astype(np.float64))
