"""The rest of the PoissonFactorization surface (SURVEY 8 rows a10, a16, a17 and the crumbs of 8b):
log_likelihood_components / predictive_distribution (poisson.py:156-210), unormalized_log_prob_list and
reconstitute (:703-717), prior_distribution (:400-401), sample / save / load [EXT]."""
import numpy as np
import pytest
import torch

from tests.util import make_counts, make_oracle, rel_err

pytestmark = pytest.mark.gpu


def _model(D=30, K=4, B=40, **kw):
    import spmf_b200
    dev = torch.device("cuda:0")
    x = make_counts(B, D, seed=8)
    N = 10 * B
    m = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D), device=dev,
                                       seed=5, **kw)
    m.compute_scales(lambda: [{'counts': x}])
    return m, make_oracle(D, K, N, x, **{k: v for k, v in kw.items() if k == 'log_transform'}), x


def test_log_likelihood_components_and_predictive_distribution():
    m, oracle, x = _model()
    th = m.sample(3, seed=2)
    thc = {k: v.cpu().double() for k, v in th.items()}
    data = {'counts': torch.tensor(x, dtype=torch.float64)}
    got = m.log_likelihood_components(data={'counts': x}, **{k: th[k] for k in ('s', 'u', 'v', 'w')})
    ref = oracle.log_likelihood_components(data=data, **{k: thc[k] for k in ('s', 'u', 'v', 'w')})
    assert set(got) == {'log_likelihood', 'rate'} and got['rate'].shape == (3,) + x.shape
    assert rel_err(got['rate'].cpu().numpy(), ref['rate'].numpy()) < 1e-4
    assert rel_err(got['log_likelihood'].cpu().numpy(), ref['log_likelihood'].numpy()) < 1e-4
    pred = m.predictive_distribution(data={'counts': x}, **{k: th[k] for k in ('s', 'u', 'v', 'w')})
    # the reference reduces a key 'll' that does not exist (poisson.py:205-208); here it is the sum
    assert rel_err(pred['ll'].cpu().numpy(), ref['log_likelihood'].sum((-1, -2)).numpy()) < 1e-4
    # no sample axis -> no reduction (reduce_dim == 0, poisson.py:204-205)
    one = {k: th[k][0] for k in ('s', 'u', 'v', 'w')}
    assert 'll' not in m.predictive_distribution(data={'counts': x}, **one)


def test_unormalized_log_prob_list_follows_var_list_order():
    m, oracle, x = _model()
    th = m.sample(2, seed=3)
    import functools
    m.unormalized_log_prob = functools.partial(m.unormalized_log_prob, {'counts': x})     # bind data (poisson.py:703-709)
    got = m.unormalized_log_prob_list(*[th[v] for v in m.var_list])
    ref = oracle.unormalized_log_prob({'counts': torch.tensor(x, dtype=torch.float64)},
                                      **{k: v.cpu().double() for k, v in th.items()})
    assert m.var_list == oracle.var_list
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-4


def test_prior_distribution_parts_and_samples():
    m, oracle, x = _model()
    th = m.sample(2, seed=4)
    got = m.prior_distribution.log_prob_parts(th)
    ref = oracle.prior_log_prob_parts({k: v.cpu().double() for k, v in th.items()})
    assert set(got) == set(ref) == set(m.var_list)
    for k in ref:
        assert rel_err(got[k].cpu().numpy(), ref[k].numpy()) < 1e-6, k
    pr = m.prior_distribution.sample(5, seed=1)
    from spmf_b200.variables import var_shapes
    for k, shp in var_shapes(m.feature_dim, m.latent_dim).items():
        assert pr[k].shape == (5,) + shp and bool((pr[k] > 0).all()) and bool(torch.isfinite(pr[k]).all()), k
    assert bool(torch.isfinite(m.prior_distribution.log_prob(pr)).all())


def test_save_load_reconstitute_round_trip(tmp_path):
    import spmf_b200
    m, _, x = _model()
    m.fit(lambda: [{'counts': x}], num_steps=3, sample_size=2, verbose=False)
    f = str(tmp_path / "model.pkl")
    m.save(f)
    again = spmf_b200.PoissonFactorization.load(f)
    for a, b in zip(m.surrogate_vars, again.surrogate_vars):
        assert torch.equal(a.cpu(), b.cpu())
    assert torch.equal(m.eta_i, again.eta_i) and m.xi_u_global == again.xi_u_global
    z0, z1 = m.encode(x), again.encode(x)
    assert rel_err(z1.cpu().numpy(), z0.cpu().numpy()) < 1e-6
    import pickle
    st = pickle.load(open(f, 'rb'))                      # the dill file is a plain pickle of tensors
    m2, _, _ = _model()
    m2.reconstitute(st)                                  # poisson.py:711-717: assign in surrogate_vars order
    for a, b in zip(m.surrogate_vars, m2.surrogate_vars):
        assert torch.equal(a.cpu(), b.cpu())


def test_streaming_fit_with_non_hybrid_sample_count():
    """K=8, S=2 has no tensor-core record shape; host-resident compact batches must still train."""
    import spmf_b200
    from spmf_b200.data import CsrShard, HostCsr
    dev = torch.device("cuda:0")
    x = make_counts(256, 96, seed=1)
    shard = CsrShard.from_dense(torch.from_numpy(x), dev)
    m = spmf_b200.PoissonFactorization(latent_dim=8, feature_dim=96, u_tau_scale=1e-3, device=dev)
    m.compute_scales(shard)
    host = HostCsr.from_shard(shard)
    losses = m.fit(lambda: ({'counts': b} for b in host.iter_batches(64)), num_steps=2, sample_size=2, verbose=False)
    assert np.isfinite(losses).all()
    losses = m.fit(lambda: ({'counts': b} for b in host.iter_batches(64)), num_steps=2, sample_size=4, verbose=False)
    assert np.isfinite(losses).all()
