"""`BernoulliFactorization` -- host-side mirror of mederrata_spmf/bernoulli.py::BernoulliFactorization
(class at bernoulli.py:31-649): the sparse-encoder factorisation with a Bernoulli-logit likelihood.

Differences from `PoissonFactorization` in the reference, all restated here:
  * likelihood: tfd.Bernoulli(logits=rate).log_prob(x) = x*rate - softplus(rate)   (bernoulli.py:148-156);
  * v and w are unconstrained: Identity bijectors (bernoulli.py:186-195) and Normal(0, 0.1) / Normal(0, 1)
    priors (bernoulli.py:200-215) -- a model flag of the parameter-side kernels (SPMF_MODEL_BERNOULLI);
  * encode has no row scaling (bernoulli.py:580-593), eta_i = 1 unless column_norms is given
    (bernoulli.py:105-109; the constructor has no scale_columns / scale_rows switches).
There is no closed form for sum_d softplus(rate), so the data term runs the dense CUDA-core kernels of
csrc/spmf_dense.cu (links SPMF_LINK_BERNOULLI / _BERNOULLI_LOG), exactly like log_transform Poisson.
"""
from __future__ import annotations

import math

import torch

from . import _abi
from .poisson import PoissonFactorization


class BernoulliFactorization(PoissonFactorization):
    """Sparse (horseshoe) Bernoulli matrix factorisation, ADVI on B200."""

    _model_id = _abi.MODEL_BERNOULLI

    def __init__(self, latent_dim=None, feature_dim=None, u_tau_scale=0.01, s_tau_scale=1.0,
                 symmetry_breaking_decay=0.99, strategy=None, encoder_function=None, decoder_function=None,
                 log_transform=False, horshoe_plus=True, column_norms=None, count_key="counts",
                 dtype=torch.float32, **kwargs):                       # bernoulli.py:64-79
        kwargs.pop("scale_rows", None)
        kwargs.pop("scale_columns", None)
        super().__init__(latent_dim=latent_dim, feature_dim=feature_dim, u_tau_scale=u_tau_scale,
                         s_tau_scale=s_tau_scale, symmetry_breaking_decay=symmetry_breaking_decay,
                         strategy=strategy, encoder_function=encoder_function, decoder_function=decoder_function,
                         scale_columns=True, scale_rows=False, log_transform=log_transform,
                         horshoe_plus=horshoe_plus, column_norms=column_norms, count_key=count_key, dtype=dtype,
                         **kwargs)

    @staticmethod
    def _link_id(log_transform):
        return _abi.LINK_BERNOULLI_LOG if log_transform else _abi.LINK_BERNOULLI

    def create_distributions(self):
        super().create_distributions()
        self.bijectors = dict(self.bijectors, v='identity', w='identity')     # bernoulli.py:186-195

    @staticmethod
    def _vw_prior(y, sc):
        """Normal(0, scale) log-density of v and w (bernoulli.py:200-215)."""
        return -0.5 * math.log(2.0 * math.pi) - torch.log(sc) - 0.5 * (y / sc) ** 2

    @staticmethod
    def _log_prob(x, rate):
        """tfd.Bernoulli(logits=rate).log_prob(x)  (bernoulli.py:148-156)."""
        return x * rate - torch.nn.functional.softplus(rate)
