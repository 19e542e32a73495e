"""Shared helpers for the parity tests (oracle on one side, CUDA path on the other)."""
import numpy as np
import torch

from oracle.spmf_oracle import OraclePoissonFactorization, draw_noise


def make_counts(B, D, seed=0, kind="noise", rate=1.0):
    rng = np.random.default_rng(seed)
    if kind == "noise":
        x = rng.poisson(rate, size=(B, D)).astype(np.float32)
    elif kind == "linear":
        from spmf_b200.data import synth_linear_dense
        x = synth_linear_dense(B, D, seed=seed)
    elif kind == "sparse":
        g = np.exp(1.5 * rng.standard_normal(D)) * 0.05
        c = np.exp(0.5 * rng.standard_normal(B))
        x = rng.poisson(c[:, None] * g[None, :]).astype(np.float32)
    else:
        raise ValueError(kind)
    x[:, 0] = np.maximum(x[:, 0], 1)      # no empty row
    x[0, :] = np.maximum(x[0, :], 1)      # no empty column (reference's xi would be NaN)
    return x


def make_oracle(D, K, N, x=None, **kw):
    m = OraclePoissonFactorization(K, D, u_tau_scale=1.0 / np.sqrt(N * D), **kw)
    if x is not None:
        m.compute_scales([{'counts': torch.tensor(x, dtype=torch.float64)}])
    return m


def perturbed_params(oracle, scale=0.3, seed=0):
    g = torch.Generator().manual_seed(seed)
    p = oracle.init_params()
    return {k: (v + scale * torch.randn(v.shape, generator=g, dtype=torch.float64)).float().double()
            for k, v in p.items()}


def rel_err(got, ref):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.abs(got - ref).max() / (np.abs(ref).max() + 1e-300))
