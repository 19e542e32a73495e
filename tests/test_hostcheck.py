"""CPU tests of the kernels' fp32 element math: the __host__ __device__ bodies of
spmf_b200/csrc/spmf_model.cuh / spmf_math.cuh, compiled for the host (tests/hostcheck) and
compared with the float64 oracle and the golden fixtures.  The data-term upstream gradients are
computed here in float64 numpy from the kernels' own fp32 operands, so what is checked is exactly
what backward_dk_kernel / backward_feat_kernel / backward_lat_kernel compute."""
import glob
import os

import numpy as np
import pytest
import torch
from scipy import special, stats

import tests.hostcheck as hc
from oracle import spmf_oracle as O
from tests.util import make_counts, make_oracle, perturbed_params, rel_err

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_philox_known_answer():
    """Random123 known-answer vectors for philox4x32-10."""
    assert [hex(v) for v in hc.philox((0, 0, 0, 0), (0, 0))] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    f = 0xffffffff
    assert [hex(v) for v in hc.philox((f, f, f, f), (f, f))] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(v) for v in hc.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))] == \
        ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_normal_and_gamma_samplers_pass_ks():
    n = hc.normals(200000, stream=2, step=5, seed=42)
    assert stats.kstest(n, 'norm').pvalue > 1e-3
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1) < 0.01
    for a in (0.3, 1.0, 2.0, 3.0, 7.5):
        g = hc.gammas(100000, a, stream=4, seed=7)
        assert (g > 0).all()
        assert stats.kstest(g, 'gamma', args=(a,)).pvalue > 1e-3, a


def test_fp32_gamma_gradient_and_digamma():
    rng = np.random.default_rng(0)
    for a in (0.3, 1.0, 2.0, 3.0, 10.0, 30.0):
        x = np.maximum(rng.gamma(a, size=20000), 1e-30).astype(np.float32)
        got = hc.gamma_grad(np.full_like(x, a), x)
        ref = O.gamma_sample_der_alpha(torch.full((x.size,), a, dtype=torch.float64),
                                       torch.tensor(x, dtype=torch.float64)).numpy()
        assert np.max(np.abs(got - ref) / np.abs(ref)) < 2e-5, a
        got4 = hc.gamma_grad4(np.full_like(x, a), x)           # the 4-draws-per-thread form the kernel uses
        assert np.max(np.abs(got4 - ref) / np.abs(ref)) < 2e-5, a
    x = np.linspace(0.01, 60, 5000).astype(np.float32)
    ref = special.digamma(x.astype(np.float64))
    assert np.max(np.abs(hc.digamma(x) - ref) / np.maximum(1, np.abs(ref))) < 2e-6


def _data_term_float64(x, r, Ap, EV, PH):
    """Float64 data term from fp32 operands: upstream gradients and Lx, Lz per draw."""
    S = Ap.shape[0]
    out = []
    for s in range(S):
        z = r[:, None] * (x @ Ap[s].astype(np.float64))
        lam = z @ EV[s].astype(np.float64).T + PH[s][None, :]
        gq = np.where(x > 0, x / lam, 0.0)
        Lx = (np.where(x > 0, x * np.log(lam), 0) - lam - special.gammaln(x + 1)).sum()
        Lz = (O.HALF_LOG_2_OVER_PI - 0.5 * z ** 2).sum()
        G = gq - 1.0
        dz = G @ EV[s].astype(np.float64) - z
        out.append(((x * r[:, None]).T @ dz, G.T @ z, gq.sum(0), Lx, Lz))
    return [np.stack([o[i] for o in out]) for i in range(5)]


def _check_against_oracle(m, params, noise, x, ref_loss, ref_grads, ref_parts, tol=2e-5):
    D, K = m.feature_dim, m.latent_dim
    S = noise['u'].shape[0]
    P = hc.pack_params(params, D, K, S)
    N = hc.pack_noise(noise, D, K, S)
    eta = m.eta_i.numpy().reshape(-1)
    Ap, EV, PH = hc.draw_operands(P, N, eta, D, K, S)
    r = x.sum(1) / float(m.xi_u_global)
    GAp, GEV, Gph, Lx, Lz = _data_term_float64(x.astype(np.float64), r, Ap, EV, PH)
    grads, parts = hc.backward_params(P, N, eta, D, K, S, GAp, GEV, Gph, float(x.shape[0]), m.u_tau_scale,
                                      m.s_tau_scale, m.symmetry_breaking_decay)
    gd = hc.unpack_grads(grads, D, K, S, m.var_shapes())
    for k, g in ref_grads.items():
        assert rel_err(gd[k], g) < tol, (k, rel_err(gd[k], g))
    for i, name in enumerate(O.VAR_LIST + ['logq']):
        assert rel_err(parts[:, i], ref_parts[name]) < tol, name
    loss = np.mean(parts[:, 12] - parts[:, :12].sum(1) - Lz - Lx)
    assert abs(loss - ref_loss) < tol * abs(ref_loss)


@pytest.mark.parametrize("D,K,B,S,perturb", [(7, 3, 5, 2, 0.3), (40, 8, 30, 4, 0.3), (33, 2, 20, 3, 0.0),
                                            (20, 50, 10, 1, 0.2), (48, 128, 12, 2, 0.1)])
def test_fp32_bodies_match_oracle(D, K, B, S, perturb):
    x = make_counts(B, D, seed=D)
    m = make_oracle(D, K, 1000, x)
    params = perturbed_params(m, perturb, seed=K)
    noise = O.draw_noise(m, params, S, seed=B)
    loss, grads, parts = m.loss_and_grads(params, noise, {'counts': torch.tensor(x, dtype=torch.float64)})
    _check_against_oracle(m, params, noise, x, loss, {k: v.numpy() for k, v in grads.items()},
                          {k: v.numpy() for k, v in parts.items()})


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_fp32_bodies_match_golden(path):
    g = np.load(path)
    D, K, B, S, N = (int(v) for v in g["meta"])
    m = make_oracle(D, K, N, g["x"])
    params = {k[6:]: torch.tensor(g[k]) for k in g.files if k.startswith("param:")}
    noise = {k[6:]: torch.tensor(g[k]).double() for k in g.files if k.startswith("noise:")}
    _check_against_oracle(m, params, noise, g["x"], float(g["loss"]),
                          {k[5:]: g[k] for k in g.files if k.startswith("grad:")},
                          {k[5:]: g[k] for k in g.files if k.startswith("part:")})


def test_world_size_scaling_of_replicated_terms():
    """rep_scale: summing `world` ranks' gradients (each on its row shard) reproduces the
    single-rank gradient -- the prior/entropy share of v,w,u,s is split, the data part adds up."""
    D, K, B, S = 12, 3, 8, 2
    x = make_counts(B, D, seed=3)
    m = make_oracle(D, K, 100, x)
    params = perturbed_params(m, 0.2, seed=1)
    noise = O.draw_noise(m, params, S, seed=2)
    P, N = hc.pack_params(params, D, K, S), hc.pack_noise(noise, D, K, S)
    eta = m.eta_i.numpy().reshape(-1)
    Ap, EV, PH = hc.draw_operands(P, N, eta, D, K, S)
    xi = float(m.xi_u_global)

    def grads_for(rows, world):
        xs = x[rows].astype(np.float64)
        GAp, GEV, Gph, _, _ = _data_term_float64(xs, xs.sum(1) / xi, Ap, EV, PH)
        return hc.backward_params(P, N, eta, D, K, S, GAp, GEV, Gph, float(len(rows)), m.u_tau_scale,
                                  m.s_tau_scale, m.symmetry_breaking_decay, world=world)[0]
    full = grads_for(np.arange(B), 1)
    a, b = grads_for(np.arange(0, 5), 2), grads_for(np.arange(5, B), 2)
    toff, _ = hc.layout(D, K, S)
    nblock = toff[8] - 1024                     # data-touched tensors (v, w, u, s)
    assert rel_err((a + b)[:nblock], full[:nblock]) < 2e-6
    # the other 16 tensors do not depend on the data: identical on every rank
    np.testing.assert_array_equal(a[toff[8]:], full[toff[8]:])
    np.testing.assert_array_equal(b[toff[8]:], full[toff[8]:])
