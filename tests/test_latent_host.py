"""CPU checks of the Latent-ODE host glue that needs no GPU: the experiment's likelihood / KL terms
(experiments/latent_ode.jl:212-224) and the Flux.destructure layout of the dense chains (rec_to_gen, gen_to_data)."""
import math

import numpy as np
import pytest
import torch

from regneuralde.jl_b200.latent import _dense_chain, kl_divergence, log_likelihood
from regneuralde.jl_b200.node import Chain, Dense


def test_log_likelihood_matches_the_experiment_formula():
    rng = np.random.default_rng(0)
    F, T, B = 5, 7, 3
    mask = torch.tensor((rng.random((F, T, B)) < 0.4).astype(np.float64))
    mask[0, 0, :] = 1.0                                   # at least one observation per sample
    d = torch.tensor(rng.standard_normal((F, T, B))) * mask
    ll = log_likelihood(d, mask)
    s = 0.01
    ref = [(-(d[:, :, b] ** 2) / (2 * s * s) - math.log(s) - math.log(2 * math.pi) / 2).sum() / mask[:, :, b].sum() for b in range(B)]
    assert ll.shape == (B,) and torch.allclose(ll, torch.stack(ref))


def test_kl_divergence_is_the_standard_gaussian_kl_mean_over_latent_dims():
    rng = np.random.default_rng(1)
    mu, lv = torch.tensor(rng.standard_normal((20, 4))), torch.tensor(rng.standard_normal((20, 4)))
    kl = kl_divergence(mu, lv)
    ref = 0.5 * (torch.exp(lv) + mu ** 2 - 1 - lv).mean(0)
    assert torch.allclose(kl, ref) and torch.allclose(kl_divergence(torch.zeros(3, 2), torch.zeros(3, 2)), torch.zeros(2))


def test_dense_chain_uses_flux_destructure_order():
    gen = torch.Generator().manual_seed(3)
    layers = (Dense(4, 6, "tanh", generator=gen), Dense(6, 3, None, generator=gen))
    p = torch.cat([l.destructure() for l in layers])
    assert p.numel() == 6 * 4 + 6 + 3 * 6 + 3
    x = torch.randn(4, 5, generator=gen)
    ref = layers[1].W @ torch.tanh(layers[0].W @ x + layers[0].b[:, None]) + layers[1].b[:, None]
    assert torch.allclose(_dense_chain(p, x, layers), ref, atol=1e-6)
    # vec(W) is column-major: the first `out` entries of a layer's block are W[:, 0]
    assert torch.equal(p[:6], layers[0].W[:, 0])


def test_chain_model_shapes_and_validation():
    m = Chain("tanh", Dense(20, 50, "tanh"), Dense(50, 20, "tanh"))
    assert (m.D, m.H, m.pre_act) == (20, 50, 1) and m.destructure().numel() == 50 * 20 + 50 + 20 * 50 + 20
    for bad in (lambda: Chain(Dense(20, 50), Dense(40, 20)), lambda: Chain(Dense(20, 50), Dense(50, 21)), lambda: Chain()):
        try:
            bad()
        except (ValueError, NotImplementedError):
            continue
        raise AssertionError("invalid chain accepted")


def test_ffjord_mirror_parameter_order_and_validation():
    """CSQDynamics / ConcatSquashLinear destructure in the order the oracle unpacks (Flux.destructure of MLPDynamics,
    ffjord_tabular.jl:49-55,76-80); the library's count agrees; unsupported constructions are refused."""
    import ctypes as C
    import regneuralde.jl_b200 as R
    from regneuralde.jl_b200 import _lib as L
    from oracle import ffjord_oracle as F
    m = R.CSQDynamics(43, 100)
    p = m.destructure()
    assert p.numel() == F.n_params(43, 100) == 19572
    W, B, bW, bB, G = F.unpack(p.double(), 43, 100)[0]
    l0 = m.layers[0]
    assert torch.equal(W.float(), l0.layer_W) and torch.equal(bW.float(), l0.bias_W) and torch.equal(G.float(), l0.gate_W)
    W3 = F.unpack(p.double(), 43, 100)[2][0]
    assert torch.equal(W3.float(), m.layers[2].layer_W)
    cfg = L.Config(); cfg.struct_bytes = C.sizeof(L.Config); cfg.state_dim, cfg.hidden_dim, cfg.batch, cfg.csq_extra = 44, 100, 8, 1
    assert L.lib().rnde_num_params(C.byref(cfg)) == 19572
    # argument validation happens before any device work
    cfg.t0, cfg.t1, cfg.abstol, cfg.reltol = 0.0, 1.0, 1e-6, 1e-6
    h = C.c_void_p()
    cfg.csq_extra = 2
    assert L.lib().rnde_create(C.byref(cfg), C.byref(h)) == L.ERR_ARG
    cfg.csq_extra, cfg.dist_mode, cfg.nranks = 1, L.DIST_EXACT, 2      # the FFJORD field has no reference-exact data-parallel mode
    assert L.lib().rnde_create(C.byref(cfg), C.byref(h)) == L.ERR_UNSUPPORTED
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            R.TrackedFFJORD(m, [0.0, 1.0], True, False)
