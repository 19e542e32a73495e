"""CPU tests of the oracle itself (no GPU, no product code): densities against scipy.stats, the
implicit Gamma gradient against mpmath and finite differences of scipy's inverse incomplete gamma,
autograd against the independent analytic restatement, the closed-form identities of SURVEY 3.4,
hand-checkable invariants, and the committed golden fixtures."""
import glob
import math
import os

import numpy as np
import pytest
import torch
from scipy import special, stats

from oracle import spmf_oracle as O
from oracle.analytic import analytic_loss_and_grads
from tests.util import make_counts, make_oracle, perturbed_params, rel_err

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def test_densities_match_scipy():
    y = torch.tensor([0.05, 0.3, 1.0, 2.5, 9.0], dtype=torch.float64)
    for sc in (0.1, 1.0, 3.0):
        ref = stats.halfnorm(scale=sc).logpdf(y.numpy())
        assert rel_err(O.halfnormal_log_prob(y, torch.tensor(sc, dtype=torch.float64)).numpy(), ref) < 1e-13
    for c, b in ((0.5, 1.0), (2.0, 3.0), (3.0, 1e4)):
        ref = stats.invgamma(c, scale=b).logpdf(y.numpy())
        assert rel_err(O.inverse_gamma_log_prob(y, c, b).numpy(), ref) < 1e-12
        # SqrtInverseGamma: density of sqrt(X), X ~ InvGamma  =>  p(y) = p_X(y^2) 2y
        ref = stats.invgamma(c, scale=b).logpdf(y.numpy() ** 2) + np.log(2 * y.numpy())
        assert rel_err(O.sqrt_inverse_gamma_log_prob(y, c, b).numpy(), ref) < 1e-12
    x = torch.tensor([0., 1., 2., 7., 30.], dtype=torch.float64)
    for lam in (0.01, 1.0, 12.5):
        ref = stats.poisson(lam).logpmf(x.numpy())
        assert rel_err(O.poisson_log_prob(x, torch.full_like(x, lam)).numpy(), ref) < 1e-12


def test_gamma_implicit_gradient_against_mpmath():
    import mpmath as mp
    mp.mp.dps = 40

    def ref(a, x):
        a, x = mp.mpf(a), mp.mpf(x)
        dP = mp.diff(lambda aa: mp.gammainc(aa, 0, x, regularized=True), a)
        return float(-dP / (x ** (a - 1) * mp.e ** (-x) / mp.gamma(a)))

    for a in (0.05, 0.5, 1.0, 2.0, 3.0, 20.0):
        for x in (1e-6, 0.1, 1.0, a + 0.99, a + 1.01, 10.0, 40.0):
            got = float(O.gamma_sample_der_alpha(torch.tensor([a], dtype=torch.float64),
                                                 torch.tensor([x], dtype=torch.float64)))
            r = ref(a, x)
            assert abs(got - r) <= 1e-10 * abs(r), (a, x, got, r)


def test_gamma_implicit_gradient_is_quantile_derivative():
    """dg/dalpha at fixed quantile == d/dalpha gammaincinv(alpha, u) (central difference)."""
    for a in (0.7, 2.0, 5.0):
        for u in (0.05, 0.5, 0.95):
            g = special.gammaincinv(a, u)
            h = 1e-5 * a
            fd = (special.gammaincinv(a + h, u) - special.gammaincinv(a - h, u)) / (2 * h)
            got = float(O.gamma_sample_der_alpha(torch.tensor([a], dtype=torch.float64),
                                                 torch.tensor([g], dtype=torch.float64)))
            assert abs(got - fd) < 1e-6 * abs(fd)


@pytest.mark.parametrize("D,K,B,S,kind", [(7, 3, 5, 2, "noise"), (30, 4, 20, 3, "sparse"), (12, 9, 8, 1, "linear")])
def test_autograd_matches_analytic_restatement(D, K, B, S, kind):
    x = make_counts(B, D, seed=1, kind=kind)
    m = make_oracle(D, K, 100, x)
    p = perturbed_params(m, 0.3, seed=2)
    nz = O.draw_noise(m, p, S, seed=3)
    l1, g1, parts1 = m.loss_and_grads(p, nz, {'counts': torch.tensor(x, dtype=torch.float64)})
    l2, g2, parts2 = analytic_loss_and_grads(m, p, nz, x)
    assert abs(l1 - l2) < 1e-10 * abs(l1)
    for k in g1:
        assert rel_err(g2[k], g1[k].numpy()) < 1e-9, k
    for k in parts1:
        assert rel_err(parts2[k], parts1[k].numpy()) < 1e-10, k


def test_initial_values_follow_reference():
    """poisson.py:404-539: softplus of the stored raw values gives back the listed constants."""
    m = O.OraclePoissonFactorization(3, 5, u_tau_scale=0.02, s_tau_scale=0.5)
    p = m.init_params()
    sp = torch.nn.functional.softplus
    assert torch.all(p['v/loc'] == -6) and torch.allclose(sp(p['v/scale_raw']), torch.tensor(5e-4, dtype=torch.float64))
    assert p['s/loc'][0, 0] == -2 and p['s/loc'][1, 0] == -1
    assert torch.allclose(sp(p['s/scale_raw']), torch.tensor(1e-3, dtype=torch.float64))
    assert torch.allclose(sp(p['u_eta/conc_raw']), torch.tensor(3., dtype=torch.float64))
    assert torch.allclose(sp(p['s_tau/conc_raw']), torch.tensor(1., dtype=torch.float64))
    assert torch.allclose(sp(p['u_tau_a/scale_raw']), torch.tensor(1 / 0.02 ** 2, dtype=torch.float64))
    assert torch.allclose(sp(p['s_tau_a/scale_raw']), torch.tensor(1 / 0.5 ** 2, dtype=torch.float64))
    assert m.param_names()[:4] == ['v/loc', 'v/scale_raw', 'w/loc', 'w/scale_raw']
    assert m.var_list == ['v', 'w', 'u', 'u_eta', 'u_tau', 's_eta', 's_tau', 's', 'u_eta_a', 'u_tau_a', 's_eta_a', 's_tau_a']


def test_compute_scales_known_answer():
    """poisson.py:113-154 on a hand-computable matrix."""
    x = np.array([[0, 2, 5], [0, 4, 0], [3, 0, 1]], dtype=np.float64)
    m = O.OraclePoissonFactorization(2, 3)
    m.compute_scales([{'counts': torch.tensor(x[:2])}, {'counts': torch.tensor(x[2:])}])
    # column sums 3, 6, 6 ; nonzero counts 1, 2, 2 -> means 3, 3, 3 ; xi = 9
    assert torch.equal(m.eta_i, torch.tensor([[3., 3., 3.]], dtype=torch.float64))
    assert float(m.xi_u_global) == 9.0
    x2 = np.array([[1, 0], [1, 0]], dtype=np.float64)          # mean 1 is NOT > 1 -> eta 1 ; empty column -> nan xi
    m2 = O.OraclePoissonFactorization(1, 2)
    m2.compute_scales([{'counts': torch.tensor(x2)}])
    assert m2.eta_i[0, 0] == 1.0 and math.isnan(float(m2.xi_u_global))


def test_encode_and_rate_by_hand():
    """z = (x/eta) A rowsum/xi and rate = eta (z v) + eta (s1/(s0+s1)) w on a 2x2 example."""
    m = O.OraclePoissonFactorization(1, 2)
    m.eta_i = torch.tensor([[2., 4.]], dtype=torch.float64)
    m.xi_u_global = torch.tensor(5., dtype=torch.float64)
    x = torch.tensor([[2., 8.], [0., 4.]], dtype=torch.float64)
    u = torch.tensor([[[1.], [3.]]], dtype=torch.float64)          # (S=1,D=2,K=1)
    s = torch.tensor([[[1., 1.], [3., 1.]]], dtype=torch.float64)   # a = [1/4, 1/2]
    v = torch.tensor([[[0.5, 2.0]]], dtype=torch.float64)          # (1,K,D)
    w = torch.tensor([[[1.0, 2.0]]], dtype=torch.float64)
    A = m.encoding_matrix(u, s)
    assert torch.allclose(A[0, :, 0], torch.tensor([0.25, 1.5], dtype=torch.float64))
    z = m.encode(x, u, s)
    # row0: (1*0.25 + 2*1.5) * 10/5 = 6.5 ; row1: (0 + 1*1.5) * 4/5 = 1.2
    assert torch.allclose(z[0, :, 0], torch.tensor([6.5, 1.2], dtype=torch.float64))
    comp = m.log_likelihood_components(s=s, u=u, v=v, w=w, data={'counts': x})
    phi = torch.tensor([2 * 0.75 * 1.0, 4 * 0.5 * 2.0], dtype=torch.float64)
    rate = torch.tensor([[2 * 6.5 * 0.5, 4 * 6.5 * 2.0], [2 * 1.2 * 0.5, 4 * 1.2 * 2.0]], dtype=torch.float64) + phi
    assert torch.allclose(comp['rate'][0], rate)
    assert torch.allclose(comp['log_likelihood'][0], torch.tensor(stats.poisson(rate.numpy()).logpmf(x.numpy())))


def test_guard_is_identity_when_finite_and_repairs_nonfinite():
    """poisson.py:606-616."""
    D, K, B, S = 6, 2, 4, 2
    x = make_counts(B, D, seed=4)
    m = make_oracle(D, K, 50, x)
    p = perturbed_params(m, 0.2, seed=5)
    nz = O.draw_noise(m, p, S, seed=6)
    theta, logq = m.sample(p, nz)
    parts = m.unormalized_log_prob_parts({'counts': torch.tensor(x, dtype=torch.float64)}, **theta)
    ll = m.log_likelihood_components(data={'counts': torch.tensor(x, dtype=torch.float64)}, **theta)['log_likelihood']
    assert torch.allclose(parts['x'], ll.sum((-1, -2)))
    assert set(parts) == set(O.VAR_LIST) | {'z', 'x'}
    bad = {k: v.clone() for k, v in theta.items()}
    bad['v'][0] = 0.0
    bad['w'][0] = 0.0                                  # rate 0 with x > 0 -> -inf entries for draw 0
    parts_bad = m.unormalized_log_prob_parts({'counts': torch.tensor(x, dtype=torch.float64)}, **bad)
    assert torch.isfinite(parts_bad['x']).all()


def test_closed_form_sum_of_rates_and_invariants():
    D, K, B, S = 15, 3, 10, 2
    x = make_counts(B, D, seed=8, kind="sparse")
    m = make_oracle(D, K, 80, x)
    p = perturbed_params(m, 0.3, seed=9)
    nz = O.draw_noise(m, p, S, seed=10)
    theta, _ = m.sample(p, nz)
    xt = torch.tensor(x, dtype=torch.float64)
    comp = m.log_likelihood_components(data={'counts': xt}, **theta)
    z = m.encode(xt, theta['u'], theta['s'])
    assert (z >= 0).all()
    vsum = (m.eta_i * theta['v']).sum(-1)                                     # (S,K)
    closed = (z * vsum[:, None, :]).sum((-1, -2)) + B * m.intercept_matrix(theta['w'], theta['s']).sum((-1, -2))
    assert torch.allclose(comp['rate'].sum((-1, -2)), closed, rtol=1e-12)
    # permuting rows leaves the energy unchanged; duplicating the batch doubles 'x' and 'z'
    e0 = m.unormalized_log_prob_parts({'counts': xt}, **theta)
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    e1 = m.unormalized_log_prob_parts({'counts': xt[perm]}, **theta)
    e2 = m.unormalized_log_prob_parts({'counts': torch.cat([xt, xt])}, **theta)
    for k in e0:
        assert torch.allclose(e0[k], e1[k], rtol=1e-12)
    assert torch.allclose(e2['x'], 2 * e0['x'], rtol=1e-12) and torch.allclose(e2['z'], 2 * e0['z'], rtol=1e-12)


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    D, K, B, S, N = (int(v) for v in g["meta"])
    m = make_oracle(D, K, N, g["x"])
    assert np.array_equal(m.eta_i.numpy(), g["eta"]) and float(m.xi_u_global) == float(g["xi"])
    params = {k[6:]: torch.tensor(g[k]) for k in g.files if k.startswith("param:")}
    noise = {k[6:]: torch.tensor(g[k]).double() for k in g.files if k.startswith("noise:")}
    loss, grads, parts = m.loss_and_grads(params, noise, {'counts': torch.tensor(g["x"], dtype=torch.float64)})
    assert abs(loss - float(g["loss"])) <= 1e-12 * abs(loss)
    for k, v in grads.items():
        assert rel_err(v.numpy(), g["grad:" + k]) < 1e-11, k
    for k, v in parts.items():
        assert rel_err(v.numpy(), g["part:" + k]) < 1e-12, k
    assert len(GOLDEN) >= 3


def test_callable_links_plug_into_the_oracle_like_the_reference():
    """poisson.py:94-97 assigns user callables over encoder_function / decoder_function.  The oracle takes them the
    same way (instance attributes): the built-in pair handed in as callables reproduces the built-in energy and
    gradients exactly, and a different pair changes them (this is the checker the GPU custom-link tests rely on)."""
    import torch
    from oracle.spmf_oracle import OraclePoissonFactorization, draw_noise
    D, K, B, S = 12, 3, 9, 2
    rng = np.random.default_rng(5)
    x = rng.poisson(1.5, size=(B, D)).astype(np.float64)
    x[:, 0] = np.maximum(x[:, 0], 1)
    data = {'counts': torch.tensor(x)}

    def build():
        m = OraclePoissonFactorization(K, D, u_tau_scale=1.0 / np.sqrt(100 * D))
        m.compute_scales([data])
        return m
    base = build()
    params = base.init_params()
    noise = draw_noise(base, params, S, seed=1)
    l0, g0, p0 = base.loss_and_grads(params, noise, data)
    same = build()
    eta = same.eta_i
    same.encoder_function, same.decoder_function = (lambda t: t / eta), (lambda y: y * eta)
    l1, g1, p1 = same.loss_and_grads(params, noise, data)
    assert l1 == l0 and all(torch.equal(g1[k], g0[k]) for k in g0)
    other = build()
    other.encoder_function = lambda t: torch.log1p(t)
    other.decoder_function = lambda y: torch.nn.functional.softplus(y)
    l2, g2, _ = other.loss_and_grads(params, noise, data)
    assert np.isfinite(l2) and abs(l2 - l0) > 1e-6 * abs(l0)
    assert all(bool(torch.isfinite(v).all()) for v in g2.values())
