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
