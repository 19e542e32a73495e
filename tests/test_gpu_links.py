"""GPU parity of (1) the exact non-finite guard of poisson.py:606-616 and (2) the log_transform link of
poisson.py:41-42, 52-53 -- both evaluated by the dense CUDA-core kernels of csrc/spmf_dense.cu.

Guard: one feature is forced to rate exactly 0 (v and w of that column at loc = -1000: softplus
underflows to 0 in float32 AND float64), so every nonzero count in that column has log-likelihood
-inf.  The reference replaces those entries by min(finite entries of the whole (S,B,D) tensor) - 10.
  * value: `unormalized_log_prob_parts` vs the oracle (oracle/spmf_oracle.py:340-352);
  * training step: loss, parts and all 24 gradients vs autograd of the same guarded energy.  The
    reference's OWN autograd gradient is NaN in this regime (0 * d log(rate)/d rate at rate = 0, in TF
    as in torch), so the comparison uses a restatement that differs from the oracle in exactly one
    respect: log() of a non-finite entry's rate is not back-propagated (tests/util-free, below).
"""
import numpy as np
import pytest
import torch

from tests.util import make_counts, make_oracle, perturbed_params, rel_err

pytestmark = pytest.mark.gpu

TOL, TOL_IG = 1e-4, 5e-4


def _load(eng, params):
    views = eng.layout.views(eng.params)
    for k, v in params.items():
        views[k].copy_(v.to(device=eng.device, dtype=torch.float32))


def _kill_column(params, d0):
    """rate of feature d0 == 0 for every row and draw (in fp32 and fp64)."""
    p = {k: v.clone() for k, v in params.items()}
    p['v/loc'][:, d0] = -1000.0
    p['w/loc'][:, d0] = -1000.0
    return p


def _guarded_loss_safe(oracle, params, noise, x):
    """mean_s[log q - energy] with the guard of poisson.py:606-616, autograd-safe: identical to
    OraclePoissonFactorization.loss except that log(rate) is only formed where it is used (finite
    entries with x > 0), so that the zero-weight branches do not inject 0 * inf = NaN into the gradient."""
    from oracle.spmf_oracle import halfnormal_log_prob
    theta, logq = oracle.sample(params, noise)
    parts = oracle.prior_log_prob_parts(theta)
    xt = torch.as_tensor(x, dtype=torch.float64)
    z = oracle.encode(xt, theta['u'], theta['s'])
    rate = oracle.decoder_function(torch.matmul(z, theta['v'])) + oracle.intercept_matrix(theta['w'], theta['s'])
    bad = ~torch.isfinite(torch.where(xt == 0, torch.zeros_like(rate), xt * torch.log(rate)) - rate)
    rs = torch.where(bad, torch.ones_like(rate), rate)
    rlog = torch.where(bad | (xt == 0), torch.ones_like(rate), rate)     # log() only where it is used
    ll = torch.where(xt == 0, torch.zeros_like(rs), xt * torch.log(rlog)) - torch.lgamma(xt + 1.0) - rs
    min_val = torch.where(bad, torch.zeros_like(ll), ll).min() - 10.0
    ll = torch.where(bad, torch.ones_like(ll) * min_val, ll)
    parts['x'] = ll.sum((-1, -2))
    parts['z'] = halfnormal_log_prob(z, torch.ones_like(z)).sum((-1, -2))
    return (logq - sum(parts.values())).mean(), logq, parts, int(bad.sum())


@pytest.mark.parametrize("D,K,B,S,kind,hot", [
    (40, 4, 64, 4, "noise", False),        # gather path (K=4 has no tensor-core form)
    (160, 16, 192, 4, "noise", True),      # every column hot: tcgen05 tile kernel raises the flag
    (300, 32, 96, 4, "sparse", True),      # hot + cold columns; the dead column is a hot one
    (48, 64, 40, 2, "noise", False),       # wide latent space: 2 lanes per row in the dense kernels
])
def test_guard_training_step_matches_guarded_autograd(D, K, B, S, kind, hot):
    import spmf_b200
    from oracle.spmf_oracle import draw_noise
    dev = torch.device("cuda:0")
    x = make_counts(B, D, seed=4, kind=kind)
    N = 10 * B
    oracle = make_oracle(D, K, N, x)
    d0 = 0 if kind == "sparse" else D // 3          # a populated column (column 0 is forced non-empty)
    assert (x[:, d0] > 0).sum() > 0
    params = _kill_column(perturbed_params(oracle, 0.3, seed=1), d0)
    noise = draw_noise(oracle, params, S, seed=2)
    leaves = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    loss, logq, parts, nbad = _guarded_loss_safe(oracle, leaves, noise, x)
    assert nbad == S * int((x[:, d0] > 0).sum())
    names = oracle.param_names()
    ref_grads = dict(zip(names, torch.autograd.grad(loss, [leaves[n] for n in names])))
    assert all(bool(torch.isfinite(g).all()) for g in ref_grads.values())
    # the oracle proper agrees on the VALUE (its gradient is NaN here, like the reference's)
    _, _, oparts = oracle.loss_parts(params, noise, {'counts': torch.tensor(x, dtype=torch.float64)})
    np.testing.assert_allclose(oparts['x'].numpy(), parts['x'].detach().numpy(), rtol=1e-12)

    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D), device=dev)
    model.compute_scales(lambda: [{'counts': x}])
    eng = model._engine_for(S)
    assert (eng.hot_cols > 0 and eng.hybrid_ok) == hot
    _load(eng, params)
    eng.set_noise_from(noise)
    batch = spmf_b200.as_device_batch(x, dev)
    p = eng.loss_and_grad(batch, fresh_noise=False)
    torch.cuda.synchronize()
    got_loss = float(eng.loss_value(p).item())
    ref_loss = float(loss.detach())
    assert abs(got_loss - ref_loss) <= TOL * abs(ref_loss), (got_loss, ref_loss)
    pd = eng.parts_dict()
    for name in ('x', 'z'):
        ref = parts[name].detach().numpy()
        assert np.abs(pd[name].numpy() - ref).max() <= TOL * np.abs(ref).max(), (name, pd[name].numpy(), ref)
    grads = eng.layout.views(eng.grads)
    for k, g in ref_grads.items():
        tol = TOL if k.split('/')[0] in ('v', 'w', 'u', 's') else TOL_IG
        e = rel_err(grads[k].cpu().numpy(), g.numpy())
        assert e <= tol, (k, e)
    # the guard re-armed itself: a clean step afterwards takes the fast path and matches the oracle
    clean = perturbed_params(oracle, 0.3, seed=1)
    _load(eng, clean)
    p = eng.loss_and_grad(batch, fresh_noise=False)
    ref_loss2, ref_grads2, _ = oracle.loss_and_grads(clean, noise, {'counts': torch.tensor(x, dtype=torch.float64)})
    assert abs(float(eng.loss_value(p).item()) - ref_loss2) <= TOL * abs(ref_loss2)
    assert rel_err(eng.layout.views(eng.grads)['u/loc'].cpu().numpy(), ref_grads2['u/loc'].numpy()) <= TOL


def test_guard_value_through_the_energy_api():
    """unormalized_log_prob_parts (poisson.py:582-621) with non-finite entries: all 14 parts vs the oracle."""
    import spmf_b200
    from oracle.spmf_oracle import draw_noise
    dev = torch.device("cuda:0")
    D, K, B, S = 50, 8, 70, 3
    x = make_counts(B, D, seed=9)
    N = 10 * B
    oracle = make_oracle(D, K, N, x)
    params = _kill_column(perturbed_params(oracle, 0.3, seed=3), 7)
    noise = draw_noise(oracle, params, S, seed=5)
    theta, _ = oracle.sample(params, noise)
    ref = oracle.unormalized_log_prob_parts({'counts': torch.tensor(x, dtype=torch.float64)}, **theta)
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D), device=dev)
    model.compute_scales(lambda: [{'counts': x}])
    got = model.unormalized_log_prob_parts({'counts': x}, **{k: v.to(dev) for k, v in theta.items()})
    assert set(got) == set(ref)
    for k in ref:
        r = ref[k].detach().numpy()
        assert np.abs(got[k].cpu().numpy() - r).max() <= TOL * max(np.abs(r).max(), 1.0), (k, got[k], r)
    flag, nbad, min_val = model._engine_for(S).guard_report()
    assert flag & 1 and nbad == S * int((x[:, 7] > 0).sum()) and min_val < -10.0
    # with exact_guard=False the non-finite entries are dropped and counted (round-1 behaviour)
    m2 = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D), device=dev,
                                        exact_guard=False)
    m2.compute_scales(lambda: [{'counts': x}])
    g2 = m2.unormalized_log_prob_parts({'counts': x}, **{k: v.to(dev) for k, v in theta.items()})
    assert bool(torch.isfinite(g2['x']).all()) and float((g2['x'].cpu() - ref['x']).abs().max()) > 1.0


# ------------------------------------------------------------------------------------------------
# log_transform=True (poisson.py:41-42, 52-53)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,K,B,S,kind", [
    (64, 8, 96, 4, "noise"),
    (100, 2, 130, 4, "noise"),        # C1-shaped, K=2
    (90, 16, 70, 2, "linear"),
    (40, 64, 33, 1, "sparse"),        # wide latent space
])
def test_log_transform_step_matches_oracle(D, K, B, S, kind):
    import spmf_b200
    from oracle.spmf_oracle import draw_noise
    dev = torch.device("cuda:0")
    x = make_counts(B, D, seed=6, kind=kind)
    N = 10 * B
    oracle = make_oracle(D, K, N, x, log_transform=True)
    params = perturbed_params(oracle, 0.3, seed=2)
    noise = draw_noise(oracle, params, S, seed=3)
    data = {'counts': torch.tensor(x, dtype=torch.float64)}
    ref_loss, ref_grads, ref_parts = oracle.loss_and_grads(params, noise, data)
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D),
                                           log_transform=True, device=dev)
    model.compute_scales(lambda: [{'counts': x}])
    assert rel_err(model.eta_i.numpy(), oracle.eta_i.numpy()) < 1e-12
    eng = model._engine_for(S)
    assert eng.link == 1 and not (eng.hot_cols > 0 and eng.hybrid_ok)
    _load(eng, params)
    eng.set_noise_from(noise)
    batch = spmf_b200.as_device_batch(x, dev)
    p = eng.loss_and_grad(batch, fresh_noise=False)
    torch.cuda.synchronize()
    loss = float(eng.loss_value(p).item())
    assert abs(loss - ref_loss) <= TOL * abs(ref_loss), (loss, ref_loss)
    pd = eng.parts_dict()
    for name in ref_parts:
        ref = ref_parts[name].numpy()
        assert np.abs(pd[name].numpy() - ref).max() <= TOL * max(np.abs(ref).max(), 1.0), (name, pd[name].numpy(), ref)
    grads = eng.layout.views(eng.grads)
    for k, g in ref_grads.items():
        tol = TOL if k.split('/')[0] in ('v', 'w', 'u', 's') else TOL_IG
        e = rel_err(grads[k].cpu().numpy(), g.numpy())
        assert e <= tol, (k, e)
    # encoder / energy / likelihood surface with the log link
    th = model.surrogate_distribution.sample(2, seed=4)
    thc = {k: v.cpu().double() for k, v in th.items()}
    z = model.encode(x, th['u'], th['s']).cpu().double().numpy()
    assert rel_err(z, oracle.encode(data['counts'], thc['u'], thc['s']).numpy()) < 1e-5
    got = model.unormalized_log_prob_parts({'counts': x}, **th)
    ref = oracle.unormalized_log_prob_parts(data, **thc)
    for k in ref:
        r = ref[k].numpy()
        assert np.abs(got[k].cpu().numpy() - r).max() <= TOL * max(np.abs(r).max(), 1.0), (k,)
    llc = model.log_likelihood_components(data={'counts': x}, **{k: th[k] for k in ('s', 'u', 'v', 'w')})
    ollc = oracle.log_likelihood_components(data=data, **{k: thc[k] for k in ('s', 'u', 'v', 'w')})
    assert rel_err(llc['rate'].cpu().double().numpy(), ollc['rate'].numpy()) < 1e-4


def test_log_transform_fit_decreases_loss_and_waic_runs():
    import spmf_b200
    from spmf_b200.data import synth_linear_dense
    dev = torch.device("cuda:0")
    x = synth_linear_dense(512, 60, seed=1)
    x[0, :] = x[0, :].clip(min=1)
    model = spmf_b200.PoissonFactorization(latent_dim=4, feature_dim=60, u_tau_scale=1.0 / np.sqrt(512 * 60),
                                           log_transform=True, device=dev, seed=3)
    factory = lambda: [{'counts': x[i:i + 128]} for i in range(0, 512, 128)]
    model.compute_scales(factory)
    losses = model.fit(factory, num_steps=30, learning_rate=0.05, sample_size=4, verbose=False, rel_tol=None)
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    w = model.waic(factory, sample_size=8)
    assert np.isfinite(w['waic']) and w['pwaic'] >= 0


# ------------------------------------------------------------------------------------------------
# BernoulliFactorization (bernoulli.py): Bernoulli-logit link, Identity bijectors / Normal priors on v, w
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("D,K,B,S,log_transform", [
    (64, 8, 96, 4, False),
    (50, 3, 70, 2, False),
    (40, 16, 48, 4, True),
])
def test_bernoulli_step_matches_oracle(D, K, B, S, log_transform):
    import spmf_b200
    from oracle.spmf_oracle import OracleBernoulliFactorization, draw_noise
    dev = torch.device("cuda:0")
    x = (make_counts(B, D, seed=12) > 1).astype(np.float32)            # binary observations
    x[:, 0] = 1.0
    x[0, :] = 1.0
    N = 10 * B
    oracle = OracleBernoulliFactorization(K, D, u_tau_scale=1.0 / np.sqrt(N * D), log_transform=log_transform)
    params = perturbed_params(oracle, 0.3, seed=4)
    # v, w are unconstrained here: move them off the Poisson initialisation so that logits of both signs occur
    g = torch.Generator().manual_seed(9)
    params['v/loc'] = (0.3 * torch.randn(params['v/loc'].shape, generator=g, dtype=torch.float64)).float().double()
    params['w/loc'] = (0.5 * torch.randn(params['w/loc'].shape, generator=g, dtype=torch.float64)).float().double()
    params['u/loc'] = params['u/loc'] + 4.0
    noise = draw_noise(oracle, params, S, seed=5)
    data = {'counts': torch.tensor(x, dtype=torch.float64)}
    ref_loss, ref_grads, ref_parts = oracle.loss_and_grads(params, noise, data)
    model = spmf_b200.BernoulliFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D),
                                             log_transform=log_transform, device=dev)
    assert model.bijectors['v'] == 'identity' and model.var_list == oracle.var_list
    eng = model._engine_for(S)
    assert eng.link in (2, 3) and eng.model == 1 and not eng.scale_rows
    _load(eng, params)
    eng.set_noise_from(noise)
    batch = spmf_b200.as_device_batch(x, dev)
    p = eng.loss_and_grad(batch, fresh_noise=False)
    torch.cuda.synchronize()
    loss = float(eng.loss_value(p).item())
    assert abs(loss - ref_loss) <= TOL * abs(ref_loss), (loss, ref_loss)
    pd = eng.parts_dict()
    for name in ref_parts:
        ref = ref_parts[name].numpy()
        assert np.abs(pd[name].numpy() - ref).max() <= TOL * max(np.abs(ref).max(), 1.0), (name, pd[name].numpy(), ref)
    grads = eng.layout.views(eng.grads)
    for k, gr in ref_grads.items():
        tol = TOL if k.split('/')[0] in ('v', 'w', 'u', 's') else TOL_IG
        e = rel_err(grads[k].cpu().numpy(), gr.numpy())
        assert e <= tol, (k, e)
    # draws of v, w are signed; the energy API and the likelihood surface agree with the oracle
    th = model.sample(3, seed=6)
    assert bool((th['v'] < 0).any()) and bool((th['u'] > 0).all())
    thc = {k: v.cpu().double() for k, v in th.items()}
    got = model.unormalized_log_prob_parts({'counts': x}, **th)
    ref = oracle.unormalized_log_prob_parts(data, **thc)
    for k in ref:
        r = ref[k].numpy()
        assert np.abs(got[k].cpu().numpy() - r).max() <= TOL * max(np.abs(r).max(), 1.0), (k,)
    llc = model.log_likelihood_components(data={'counts': x}, **{k: th[k] for k in ('s', 'u', 'v', 'w')})
    ollc = oracle.log_likelihood_components(data=data, **{k: thc[k] for k in ('s', 'u', 'v', 'w')})
    assert rel_err(llc['log_likelihood'].cpu().double().numpy(), ollc['log_likelihood'].numpy()) < 1e-4


def test_bernoulli_fit_decreases_loss():
    import spmf_b200
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    z = rng.random((400, 3)) < 0.3
    w = rng.random((3, 40)) < 0.4
    x = ((z.astype(np.float32) @ w.astype(np.float32)) > 0).astype(np.float32)
    x[0, :] = 1.0
    m = spmf_b200.BernoulliFactorization(latent_dim=3, feature_dim=40, u_tau_scale=1e-2, device=dev, seed=2)
    factory = lambda: [{'counts': x[i:i + 100]} for i in range(0, 400, 100)]
    losses = m.fit(factory, num_steps=40, learning_rate=0.05, sample_size=4, verbose=False, rel_tol=None)
    assert np.isfinite(losses).all() and losses[-1] < losses[0]


def test_guard_on_a_dense_ingested_batch():
    """A batch that arrived as a dense slab (HostDense) has no feature-order CSR on the device: the guard's slow
    path rebuilds its dense working copy from the raw slab (spmf_guard_rows_fix_dense) and must give the same
    step as the CSR-uploaded batch (which test_guard_training_step_matches_guarded_autograd pins to autograd)."""
    import spmf_b200
    from spmf_b200.data import BatchUploader, HostDense
    dev = torch.device("cuda:0")
    D, K, B, S = 160, 16, 192, 4
    x = make_counts(B, D, seed=4, kind="noise")
    oracle = make_oracle(D, K, 10 * B, x)
    params = _kill_column(perturbed_params(oracle, 0.3, seed=1), D // 3)
    from oracle.spmf_oracle import draw_noise
    noise = draw_noise(oracle, params, S, seed=2)
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(10 * B * D), device=dev)
    model.compute_scales(lambda: [{'counts': x}])
    eng = model._engine_for(S)
    assert eng.hot_mode == 2
    _load(eng, params)
    eng.set_noise_from(noise)
    p_ref = eng.loss_and_grad(spmf_b200.as_device_batch(x, dev), fresh_noise=False).clone()
    g_ref = eng.grads.clone()
    assert bool(torch.isfinite(p_ref).all()) and bool(torch.isfinite(g_ref).all())
    db = BatchUploader(dev, D, hot=model._hot_spec(eng)).upload(HostDense(x).batch(0, B))
    assert db.dense_raw is not None
    p = eng.loss_and_grad(db, fresh_noise=False).clone()
    torch.cuda.synchronize()
    assert abs(float(eng.loss_value(p).item()) - float(eng.loss_value(p_ref).item())) <= 1e-6 * abs(float(eng.loss_value(p_ref).item()))
    assert rel_err(eng.grads.cpu().numpy(), g_ref.cpu().numpy()) < 2e-5


def _softplus_pair():
    enc = lambda x: torch.log1p(x)                                   # g(x): acts on the raw counts
    dec = lambda y: torch.nn.functional.softplus(y)                  # f(y) >= 0
    return enc, dec


@pytest.mark.parametrize("D,K,B,S", [(64, 8, 96, 4), (90, 16, 70, 2)])
def test_custom_encoder_decoder_step_matches_oracle(D, K, B, S):
    """poisson.py:94-97: user-supplied encoder / decoder callables.  The data term runs in torch (autograd down to
    the operand tables), everything else in the native kernels: one step == the float64 oracle with the same
    functions plugged in, loss and all 24 gradients."""
    import spmf_b200
    from oracle.spmf_oracle import draw_noise
    dev = torch.device("cuda:0")
    enc, dec = _softplus_pair()
    x = make_counts(B, D, seed=6, kind="linear")
    N = 10 * B
    oracle = make_oracle(D, K, N, x)
    oracle.encoder_function, oracle.decoder_function = enc, dec     # as the reference assigns them, :94-97
    params = perturbed_params(oracle, 0.3, seed=2)
    noise = draw_noise(oracle, params, S, seed=3)
    data = {'counts': torch.tensor(x, dtype=torch.float64)}
    ref_loss, ref_grads, ref_parts = oracle.loss_and_grads(params, noise, data)
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D),
                                           encoder_function=enc, decoder_function=dec, device=dev)
    model.compute_scales(lambda: [{'counts': x}])
    eng = model._engine_for(S)
    assert eng.custom_link is not None and not (eng.hot_cols > 0 and eng.hybrid_ok)
    _load(eng, params)
    eng.set_noise_from(noise)
    batch = spmf_b200.as_device_batch(x, dev)
    p = eng.loss_and_grad(batch, fresh_noise=False)
    torch.cuda.synchronize()
    loss = float(eng.loss_value(p).item())
    assert abs(loss - ref_loss) <= TOL * abs(ref_loss), (loss, ref_loss)
    pd = eng.parts_dict()
    for name in ref_parts:
        ref = ref_parts[name].numpy()
        assert np.abs(pd[name].numpy() - ref).max() <= TOL * max(np.abs(ref).max(), 1.0), (name, pd[name].numpy(), ref)
    grads = eng.layout.views(eng.grads)
    for k, g in ref_grads.items():
        tol = TOL if k.split('/')[0] in ('v', 'w', 'u', 's') else TOL_IG
        e = rel_err(grads[k].cpu().numpy(), g.numpy())
        assert e <= tol, (k, e)
    # surface: encode / energy parts / likelihood components / per-row log-likelihood
    th = model.surrogate_distribution.sample(2, seed=4)
    thc = {k: v.cpu().double() for k, v in th.items()}
    z = model.encode(x, th['u'], th['s']).cpu().double().numpy()
    assert rel_err(z, oracle.encode(data['counts'], thc['u'], thc['s']).numpy()) < 1e-5
    got = model.unormalized_log_prob_parts({'counts': x}, **th)
    ref = oracle.unormalized_log_prob_parts(data, **thc)
    for k in ref:
        r = ref[k].numpy()
        assert np.abs(got[k].cpu().numpy() - r).max() <= TOL * max(np.abs(r).max(), 1.0), (k,)
    llc = model.log_likelihood_components(data={'counts': x}, **{k: th[k] for k in ('s', 'u', 'v', 'w')})
    ollc = oracle.log_likelihood_components(data=data, **{k: thc[k] for k in ('s', 'u', 'v', 'w')})
    assert rel_err(llc['rate'].cpu().double().numpy(), ollc['rate'].numpy()) < 1e-4
    rll = model.row_log_likelihood({'counts': x}, **th).cpu().numpy()
    assert rel_err(rll, ollc['log_likelihood'].sum(-1).numpy()) < 1e-5


def test_custom_callables_equal_to_the_builtin_pair_reproduce_the_native_step():
    """The reference's own linear pair handed in as callables must give what the CUDA kernels give: the torch
    data term and the native one are two evaluations of the same function (one step, then a short fit)."""
    import spmf_b200
    dev = torch.device("cuda:0")
    D, K, B, S = 80, 8, 128, 4
    x = make_counts(4 * B, D, seed=9, kind="linear")
    kw = dict(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(x.size), device=dev, seed=11)
    native = spmf_b200.PoissonFactorization(hot_density=0.0, **kw)
    native.compute_scales(lambda: [{'counts': x}])
    eta = native.eta_i.reshape(-1).to(dev, torch.float32)
    custom = spmf_b200.PoissonFactorization(encoder_function=lambda t: t / eta, decoder_function=lambda y: y * eta, **kw)
    custom.compute_scales(lambda: [{'counts': x}])
    e0, e1 = native._engine_for(S), custom._engine_for(S)
    assert e1.custom_link is not None and e0.custom_link is None
    b = spmf_b200.as_device_batch(x[:B], dev)
    p0 = e0.loss_and_grad(b).clone()
    p1 = e1.loss_and_grad(b).clone()
    assert rel_err(p1.cpu().numpy(), p0.cpu().numpy()) < 1e-5
    assert rel_err(e1.grads.cpu().numpy(), e0.grads.cpu().numpy()) < 5e-5
    fac = lambda: ({'counts': x[i * B:(i + 1) * B]} for i in range(4))
    l0 = native.fit(fac, num_steps=3, learning_rate=0.05, sample_size=S, verbose=False, rel_tol=None)
    l1 = custom.fit(fac, num_steps=3, learning_rate=0.05, sample_size=S, verbose=False, rel_tol=None)
    assert l1[-1] < l1[0] and np.allclose(l0, l1, rtol=1e-4), (l0, l1)
    # a saved model carries its callables (dill) ... when they can be pickled
    import os
    import tempfile
    m2 = spmf_b200.PoissonFactorization(encoder_function=_softplus_pair()[0], **kw)
    with tempfile.TemporaryDirectory() as td:
        fn = os.path.join(td, "m.pkl")
        m2.save(fn)
        m3 = spmf_b200.PoissonFactorization.load(fn, device=dev)
        assert m3._custom_link is not None
        t = torch.rand(3, D, device=dev)
        assert torch.equal(m3.encoder_function(t), torch.log1p(t))
