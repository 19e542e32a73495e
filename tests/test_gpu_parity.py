"""GPU parity: the CUDA ADVI step (through the C ABI) vs the float64 CPU oracle on identical
parameters, identical base noise and identical counts.  Tolerance: 1e-4 relative on the loss and
on every gradient tensor (max-abs error over max-abs reference), as BASELINE.json's north_star
states for fp32; the InverseGamma-family gradients carry the fp32 implicit-gradient error of the
Gamma draw and are held to 5e-4."""
import numpy as np
import pytest
import torch

from tests.util import make_counts, make_oracle, perturbed_params, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-4
TOL_IG = 5e-4


def _load(eng, params):
    views = eng.layout.views(eng.params)
    for k, v in params.items():
        views[k].copy_(v.to(device=eng.device, dtype=torch.float32))


def _run_case(D, K, B, S, kind="noise", seed=0, perturb=0.3, scale_rows=True, via_model=True):
    import spmf_b200
    from oracle.spmf_oracle import draw_noise
    dev = torch.device("cuda:0")
    x = make_counts(B, D, seed=seed, kind=kind)
    N = 10 * B
    oracle = make_oracle(D, K, N, x, scale_rows=scale_rows)
    params = perturbed_params(oracle, perturb, seed)
    noise = draw_noise(oracle, params, S, seed=seed + 1)
    ref_loss, ref_grads, ref_parts = oracle.loss_and_grads(params, noise, {'counts': torch.tensor(x, dtype=torch.float64)})

    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D),
                                           scale_rows=scale_rows, device=dev)
    model.compute_scales(lambda: [{'counts': x}])
    assert rel_err(model.eta_i.numpy(), oracle.eta_i.numpy()) < 1e-12
    if scale_rows:
        assert abs(model.xi_u_global - float(oracle.xi_u_global)) < 1e-9 * abs(float(oracle.xi_u_global))
    eng = model._engine_for(S)
    _load(eng, params)
    eng.set_noise_from(noise)
    batch = spmf_b200.as_device_batch(x, dev)
    parts = eng.loss_and_grad(batch, fresh_noise=False)
    torch.cuda.synchronize()
    loss = float(eng.loss_value(parts).item())
    assert abs(loss - ref_loss) <= TOL * abs(ref_loss), (loss, ref_loss)
    pd = eng.parts_dict()
    for name in list(ref_parts):
        ref = ref_parts[name].numpy()
        got = pd[name].numpy()
        assert np.abs(got - ref).max() <= TOL * max(np.abs(ref).max(), 1.0), (name, got, ref)
    grads = eng.layout.views(eng.grads)
    for k, g in ref_grads.items():
        tol = TOL if k.split('/')[0] in ('v', 'w', 'u', 's') else TOL_IG
        e = rel_err(grads[k].cpu().numpy(), g.numpy())
        assert e <= tol, (k, e)
    return eng


@pytest.mark.parametrize("D,K,B,S,kind", [
    (7, 3, 5, 2, "noise"),          # hand-checkable known-answer shape
    (100, 2, 256, 4, "noise"),      # C1-shaped tile: K=2, S=4
    (96, 8, 200, 4, "linear"),      # C2-shaped tile: K=8
    (64, 16, 128, 4, "linear"),     # C3-shaped tile: K=16
    (300, 32, 96, 4, "sparse"),     # C4-shaped tile: K=32, sparse counts
    (50, 50, 40, 1, "noise"),       # tests/spmf_test.py latent dim (P=50), non power of two, S=1
    (40, 128, 33, 2, "sparse"),     # K=128 (C5 upper end), S=2
    (33, 5, 17, 3, "noise"),        # odd everything, S=3 -> scalar draw lanes
    (24, 64, 20, 8, "noise"),       # two draw quads
])
def test_step_matches_oracle(D, K, B, S, kind):
    _run_case(D, K, B, S, kind)


def test_step_at_initial_params():
    _run_case(60, 8, 64, 4, perturb=0.0)


def test_step_without_row_scaling():
    _run_case(40, 4, 32, 4, scale_rows=False)


def test_philox_noise_is_deterministic_and_matches_host():
    import spmf_b200
    import tests.hostcheck as hc
    dev = torch.device("cuda:0")
    D, K, S = 37, 6, 4
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, device=dev, seed=1234)
    eng = model._engine_for(S)
    eng.fill_noise(step=7)
    a = eng.noise.clone()
    eng.fill_noise(step=7)
    assert torch.equal(a, eng.noise)
    eng.fill_noise(step=8)
    assert not torch.equal(a, eng.noise)
    # the N(0,1) block of u equals the host run of the same Philox stream (stream id = var index 2)
    eng.fill_noise(step=7)
    u = eng.layout.noise_view(eng.noise, 'u').reshape(-1).cpu().numpy()
    host = hc.normals(u.size, stream=2, step=7, seed=1234)
    np.testing.assert_allclose(u, host, rtol=2e-5, atol=2e-6)


def test_sampler_distributions():
    """Normal and Gamma draws pass a KS test against scipy."""
    import spmf_b200
    from scipy import stats
    dev = torch.device("cuda:0")
    D, K, S = 500, 16, 4
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, device=dev, seed=3)
    eng = model._engine_for(S)
    eng.fill_noise(step=0)
    u = eng.layout.noise_view(eng.noise, 'u').reshape(-1).cpu().numpy()
    assert stats.kstest(u, 'norm').pvalue > 1e-3
    g = eng.layout.noise_view(eng.noise, 'u_eta').reshape(-1).cpu().numpy()     # Gamma(3,1) at init
    assert stats.kstest(g, 'gamma', args=(3.0,)).pvalue > 1e-3
    g = eng.layout.noise_view(eng.noise, 's_eta').reshape(-1).cpu().numpy()     # Gamma(1,1) at init
    assert stats.kstest(g, 'gamma', args=(1.0,)).pvalue > 1e-3


def test_encode_matches_oracle():
    import spmf_b200
    dev = torch.device("cuda:0")
    D, K, B = 80, 8, 50
    x = make_counts(B, D, seed=3)
    oracle = make_oracle(D, K, 500, x)
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(500 * D), device=dev)
    model.compute_scales(lambda: [{'counts': x}])
    th = model.surrogate_distribution.sample(3, seed=5)
    z = model.encode(x, th['u'], th['s']).cpu().double().numpy()
    zr = oracle.encode(torch.tensor(x, dtype=torch.float64), th['u'].cpu().double(), th['s'].cpu().double()).numpy()
    assert rel_err(z, zr) < 1e-5
    z1 = model.encode(x).cpu().double().numpy()          # calibrated expectations, no sample axis
    ce = model.calibrated_expectations
    zr1 = oracle.encode(torch.tensor(x, dtype=torch.float64), ce['u'].cpu().double(), ce['s'].cpu().double()).numpy()
    assert z1.shape == (B, K) and rel_err(z1, zr1) < 1e-5


def test_unormalized_log_prob_matches_oracle():
    import spmf_b200
    dev = torch.device("cuda:0")
    D, K, B, S = 60, 4, 40, 4
    x = make_counts(B, D, seed=4)
    oracle = make_oracle(D, K, 400, x)
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(400 * D), device=dev)
    model.compute_scales(lambda: [{'counts': x}])
    th = model.surrogate_distribution.sample(S, seed=9)
    got = model.unormalized_log_prob(data={'counts': x}, **th).cpu().numpy()
    th64 = {k: v.cpu().double() for k, v in th.items()}
    ref = oracle.unormalized_log_prob(data={'counts': torch.tensor(x, dtype=torch.float64)}, **th64).numpy()
    assert got.shape == (S,) and rel_err(got, ref) < 1e-5
    parts = model.unormalized_log_prob_parts({'counts': x}, **th)
    assert set(parts) == set(oracle.var_list) | {'z', 'x'}          # poisson.py:582-621 dict keys


def test_row_log_likelihood_and_waic_match_oracle():
    """per-row log-likelihood of every draw from the CUDA row pass vs the oracle's dense
    (S,B,D) log-likelihood (poisson.py:156-184) summed over features; WAIC over the rows from it."""
    import spmf_b200
    from spmf_b200.poisson import waic_terms
    dev = torch.device("cuda:0")
    D, K, B, S = 60, 4, 48, 8
    x = make_counts(B, D, seed=6)
    oracle = make_oracle(D, K, 400, x)
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(400 * D), device=dev)
    model.compute_scales(lambda: [{'counts': x[:24]}, {'counts': x[24:]}])
    th = model.surrogate_distribution.sample(S, seed=12345)
    th64 = {k: v.cpu().double() for k, v in th.items()}
    ref = oracle.log_likelihood_components(data={'counts': torch.tensor(x, dtype=torch.float64)},
                                           **{k: th64[k] for k in 'suvw'})['log_likelihood'].sum(-1).numpy()
    got = model.row_log_likelihood({'counts': x}, **th).cpu().numpy()
    assert got.shape == (S, B)
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-3)
    # WAIC with the same draws (seed 12345 is waic()'s default), streamed over the two batches
    w = model.waic(sample_size=S)
    lppd = np.log(np.exp(ref - ref.max(0)).mean(0)) + ref.max(0)
    pw = ref.var(0, ddof=1)
    elpd = lppd - pw
    exp = {'waic': -2 * elpd.sum(), 'se': 2 * np.sqrt(B * elpd.var(ddof=1)), 'lppd': lppd.sum(), 'pwaic': pw.sum()}
    assert set(w) == set(exp)                       # the reference's keys (notebook output)
    for k in exp:
        assert abs(w[k] - exp[k]) <= 1e-4 * abs(exp[k]) + 1e-3, (k, w[k], exp[k])
    a, b = waic_terms(torch.tensor(ref))
    np.testing.assert_allclose(a.numpy(), lppd, rtol=1e-10)
    np.testing.assert_allclose(b.numpy(), pw, rtol=1e-8)


def test_csr_csc_roundtrip_and_dense_compaction():
    import spmf_b200
    dev = torch.device("cuda:0")
    x = make_counts(70, 90, seed=5, kind="sparse")
    sh = spmf_b200.CsrShard.from_dense(x, dev)
    import scipy.sparse as sp
    ref = sp.csr_matrix(x)
    assert np.array_equal(sh.rowptr.cpu().numpy(), ref.indptr)
    assert np.array_equal(sh.cols.cpu().numpy(), ref.indices)
    assert np.array_equal(sh.vals.cpu().numpy(), ref.data)
    b = sh.batch(10, 40).ensure_csc()
    torch.cuda.synchronize()
    sub = sp.csc_matrix(x[10:50])
    assert np.array_equal(b.colptr.cpu().numpy(), sub.indptr)
    rows, vals, cp = b.crows.cpu().numpy(), b.cvals.cpu().numpy(), sub.indptr
    for d in range(90):                       # scatter order inside a column is not fixed
        o = np.argsort(rows[cp[d]:cp[d + 1]])
        assert np.array_equal(rows[cp[d]:cp[d + 1]][o], sub.indices[cp[d]:cp[d + 1]])
        assert np.array_equal(vals[cp[d]:cp[d + 1]][o], sub.data[cp[d]:cp[d + 1]])
    from scipy.special import gammaln
    np.testing.assert_allclose(sh.rowsum.cpu().numpy(), x.sum(1), rtol=1e-6)
    np.testing.assert_allclose(sh.lgam.cpu().numpy(), gammaln(x + 1).sum(1), rtol=1e-5, atol=1e-5)


def test_properties_row_permutation_and_duplication():
    """Size-independent properties: permuting rows leaves loss/gradients unchanged;
    duplicating the batch doubles the data parts."""
    import spmf_b200
    dev = torch.device("cuda:0")
    D, K, B, S = 120, 8, 300, 4
    x = make_counts(B, D, seed=6, kind="sparse")
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1e-3, device=dev, seed=11)
    model.compute_scales(lambda: [{'counts': x}])
    eng = model._engine_for(S)
    eng.fill_noise(step=0)
    p0 = eng.loss_and_grad(spmf_b200.as_device_batch(x, dev), fresh_noise=False).clone()
    g0 = eng.grads.clone()
    perm = np.random.default_rng(0).permutation(B)
    p1 = eng.loss_and_grad(spmf_b200.as_device_batch(x[perm], dev), fresh_noise=False).clone()
    g1 = eng.grads.clone()
    assert rel_err(p1.cpu().numpy(), p0.cpu().numpy()) < 1e-6
    assert rel_err(g1.cpu().numpy(), g0.cpu().numpy()) < 2e-5
    p2 = eng.loss_and_grad(spmf_b200.as_device_batch(np.concatenate([x, x]), dev), fresh_noise=False)
    np.testing.assert_allclose(p2[:, 13:15].cpu().numpy(), 2 * p0[:, 13:15].cpu().numpy(), rtol=1e-6)
    np.testing.assert_allclose(p2[:, :13].cpu().numpy(), p0[:, :13].cpu().numpy(), rtol=1e-12)


def test_fit_reduces_loss():
    """Qualitative acceptance mirroring notebooks/factorize_linear_structure.ipynb: the loss goes
    down and signal columns (every 3rd) carry more encoding weight than noise columns."""
    import spmf_b200
    from spmf_b200.data import synth_linear_dense
    dev = torch.device("cuda:0")
    N, D, K = 4000, 30, 3
    x = synth_linear_dense(N, D, seed=0)
    sh = spmf_b200.CsrShard.from_dense(x, dev)
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D), device=dev)
    model.compute_scales(sh)
    factory = lambda: ({'counts': b} for b in sh.iter_batches(1000))
    losses = model.fit(factory, num_steps=60, learning_rate=0.05, sample_size=8, rel_tol=None, verbose=False)
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    A = model.encoding_matrix()
    assert A.shape == (D, K) and bool(torch.isfinite(A).all())


import glob as _glob
import os as _os

_GOLDEN = sorted(_glob.glob(_os.path.join(_os.path.dirname(__file__), "golden", "*.npz")))


@pytest.mark.parametrize("path", _GOLDEN, ids=[_os.path.basename(p) for p in _GOLDEN])
def test_step_matches_committed_golden(path):
    """CUDA step vs the committed oracle fixtures (tests/golden/make_golden.py)."""
    import spmf_b200
    g = np.load(path)
    D, K, B, S, N = (int(v) for v in g["meta"])
    dev = torch.device("cuda:0")
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1.0 / np.sqrt(N * D), device=dev)
    model.compute_scales(lambda: [{'counts': g["x"]}])
    assert rel_err(model.eta_i.numpy(), g["eta"]) < 1e-12
    eng = model._engine_for(S)
    _load(eng, {k[6:]: torch.tensor(g[k]) for k in g.files if k.startswith("param:")})
    eng.set_noise_from({k[6:]: torch.tensor(g[k]) for k in g.files if k.startswith("noise:")})
    parts = eng.loss_and_grad(spmf_b200.as_device_batch(g["x"], dev), fresh_noise=False)
    loss = float(eng.loss_value(parts).item())
    assert abs(loss - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    grads = eng.layout.views(eng.grads)
    for k in (f for f in g.files if f.startswith("grad:")):
        name = k[5:]
        tol = TOL if name.split('/')[0] in ('v', 'w', 'u', 's') else TOL_IG
        assert rel_err(grads[name].cpu().numpy(), g[k]) <= tol, name


def test_streamed_host_batches_match_resident():
    """HostCsr -> H2D -> device CSC build gives the same step as the resident shard."""
    import spmf_b200
    from spmf_b200.data import HostCsr
    dev = torch.device("cuda:0")
    D, K, S, B = 200, 8, 4, 128
    x = make_counts(3 * B, D, seed=12, kind="sparse")
    sh = spmf_b200.CsrShard.from_dense(x, dev)
    host = HostCsr.from_shard(sh)
    model = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=D, u_tau_scale=1e-3, device=dev, seed=5)
    model.compute_scales(sh)
    eng = model._engine_for(S)
    eng.fill_noise(step=0)
    p0 = eng.loss_and_grad(sh.batch(B, B), fresh_noise=False).clone()
    g0 = eng.grads.clone()
    assert host.cols.dtype == torch.uint16 and host.vals.dtype == torch.uint16      # compact transfer format
    p1 = eng.loss_and_grad(spmf_b200.as_device_batch(host.batch(B, B), dev), fresh_noise=False).clone()
    assert rel_err(p1.cpu().numpy(), p0.cpu().numpy()) < 1e-6
    assert rel_err(eng.grads.cpu().numpy(), g0.cpu().numpy()) < 2e-5
    # prefetched stream of host batches: same batches, same order, same results
    from spmf_b200.data import prefetch_to_device
    wide = HostCsr.from_shard(sh, compact=False)
    xb = x.copy()
    xb[5, 7] = 300.0                       # a count above 254: travels through the overflow list of the 2-byte format
    xb[9, :] = 0
    xb[9, 199] = 2.0                       # a gap wider than 256 columns needs no bridge here (D = 200) ...
    shb = spmf_b200.CsrShard.from_dense(xb, dev)
    bytes8 = HostCsr.from_shard(shb, compact="u8")
    assert bytes8.u8 and bytes8.batch(0, B).nbytes() < 0.6 * HostCsr.from_shard(shb).batch(0, B).nbytes()
    for hsrc, ref_shard in ((host, sh), (wide, sh), (bytes8, shb)):
        got = []
        for db in prefetch_to_device(hsrc.iter_batches(B), dev):
            got.append(eng.loss_and_grad(db, fresh_noise=False)[:, 13:15].clone())
        ref = [eng.loss_and_grad(ref_shard.batch(i * B, B), fresh_noise=False)[:, 13:15].clone() for i in range(3)]
        assert len(got) == 3
        for a, b in zip(got, ref):
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-6


def test_data_parallel_two_gpus():
    """Row-sharded step on 2 GPUs == single-GPU step on the whole batch (skipped with < 2 GPUs)."""
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29517",
                          _os.path.join(root, "tests", "dp_check.py")], capture_output=True, text=True, timeout=600)
    assert "DP_CHECK_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
