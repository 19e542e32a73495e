"""CPU oracle for the SPMF ADVI step -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A float64 torch-CPU restatement of the reference's ``PoissonFactorization``
energy (``/root/reference/mederrata_spmf/poisson.py``) plus the slices of its
un-vendored dependencies (TensorFlow-Probability, ``bayesianquilts``) that the
ADVI step needs: mean-field Softplus surrogate, reparameterised draws, log q,
prior log-densities.  Gradients come from torch autograd so this file contains
*no* hand-derived backward; ``oracle/analytic.py`` restates the analytic
backward independently and the tests check one against the other.

PARITY UNPINNED: the reference ships no golden vector / known-answer test for
this path (``tests/spmf_test.py`` prints, never asserts, and is unseeded), and
TF / TFP / bayesianquilts cannot be installed here (no network), so nothing in
this file could be checked against outputs of the reference itself.  What pins
it instead: scipy.stats densities, mpmath for the implicit Gamma gradient,
autograd-vs-analytic gradients, and hand-computable invariants (tests/).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this module.  The product package
``spmf_b200`` never does.

Every function cites the reference lines it follows (paths relative to
/root/reference).  Items marked [EXT] restate documented TFP / bayesianquilts
semantics that are not in the reference tree.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional

import numpy as np
import torch

DTYPE = torch.float64

# mederrata_spmf/poisson.py:403-572 -- dict insertion order of surrogate_dict
VAR_LIST = ['v', 'w', 'u', 'u_eta', 'u_tau', 's_eta', 's_tau', 's',
            'u_eta_a', 'u_tau_a', 's_eta_a', 's_tau_a']
NORMAL_VARS = ('v', 'w', 'u', 's')
IG_VARS = ('u_eta', 'u_tau', 's_eta', 's_tau',
           'u_eta_a', 'u_tau_a', 's_eta_a', 's_tau_a')

HALF_LOG_2_OVER_PI = 0.5 * math.log(2.0 / math.pi)
HALF_LOG_2PI = 0.5 * math.log(2.0 * math.pi)


def softplus_inverse(y: torch.Tensor) -> torch.Tensor:
    """[EXT] tfp.util.TransformedVariable(x, Softplus()) stores softplus^-1(x)."""
    return y + torch.log(-torch.expm1(-y))


# ---------------------------------------------------------------------------
# Implicit reparameterisation gradient of a standard Gamma draw  [EXT]
# tf.random.gamma's gradient is  dg/dalpha = -(dP(alpha,g)/dalpha) / p(g;alpha)
# (Figurnov et al. 2018), evaluated by Eigen's igamma_der_a/gamma_sample_der_alpha
# with a power series for small g and a continued fraction otherwise.
# ---------------------------------------------------------------------------
def gamma_sample_der_alpha(alpha: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    """dg/dalpha at fixed uniform quantile, float64, elementwise.

    series (g <= alpha+1 or g <= 1):  P = g^a e^-g / Gamma(a+1) * sum_n T_n,
      T_n = prod_{i<=n} g/(a+i)  =>  dg/da = (g/a) [sum T_n H_n - (log g - psi(a+1)) sum T_n],
      H_n = sum_{i<=n} 1/(a+i).
    continued fraction (otherwise), Cephes igamc recurrences differentiated in a:
      Q = ans * g^a e^-g / Gamma(a)   =>  dg/da = g [dans/da + ans (log g - psi(a))].
    """
    a = alpha.detach().to(DTYPE).contiguous()
    x = g.detach().to(DTYPE).contiguous()
    a, x = torch.broadcast_tensors(a, x)
    a = a.reshape(-1).clone()
    x = x.reshape(-1).clone()
    out = torch.empty_like(x)
    use_series = (x <= 1.0) | (x <= a + 1.0)

    # --- power series --------------------------------------------------
    if use_series.any():
        aa, xx = a[use_series], x[use_series]
        T = torch.ones_like(xx)
        H = torch.zeros_like(xx)
        sT = torch.ones_like(xx)
        sTH = torch.zeros_like(xx)
        for n in range(1, 2000):
            T = T * xx / (aa + n)
            H = H + 1.0 / (aa + n)
            sT = sT + T
            sTH = sTH + T * H
            if float((T * (1.0 + H)).max()) < 1e-18 * float(sT.min()):
                break
        out[use_series] = (xx / aa) * (sTH - (torch.log(xx) - torch.digamma(aa + 1.0)) * sT)

    # --- continued fraction ---------------------------------------------
    cf = ~use_series
    if cf.any():
        aa, xx = a[cf], x[cf]
        y = 1.0 - aa
        z = xx + y + 1.0
        dy = -torch.ones_like(aa)
        dz = -torch.ones_like(aa)
        pkm2 = torch.ones_like(xx)
        qkm2 = xx.clone()
        pkm1 = xx + 1.0
        qkm1 = z * xx
        dpkm2 = torch.zeros_like(xx)
        dqkm2 = torch.zeros_like(xx)
        dpkm1 = torch.zeros_like(xx)
        dqkm1 = dz * xx
        ans = pkm1 / qkm1
        dans = (dpkm1 - ans * dqkm1) / qkm1
        for c in range(1, 5000):
            y = y + 1.0
            z = z + 2.0
            yc = y * c
            dyc = dy * c
            pk = pkm1 * z - pkm2 * yc
            qk = qkm1 * z - qkm2 * yc
            dpk = dpkm1 * z + pkm1 * dz - dpkm2 * yc - pkm2 * dyc
            dqk = dqkm1 * z + qkm1 * dz - dqkm2 * yc - qkm2 * dyc
            new_ans = pk / qk
            new_dans = (dpk - new_ans * dqk) / qk
            delta = (new_dans - dans).abs().max()
            ans, dans = new_ans, new_dans
            pkm2, pkm1, qkm2, qkm1 = pkm1, pk, qkm1, qk
            dpkm2, dpkm1, dqkm2, dqkm1 = dpkm1, dpk, dqkm1, dqk
            big = pk.abs() > 1e150
            if big.any():
                sc = torch.where(big, torch.full_like(pk, 1e-150), torch.ones_like(pk))
                pkm2, pkm1, qkm2, qkm1 = pkm2 * sc, pkm1 * sc, qkm2 * sc, qkm1 * sc
                dpkm2, dpkm1, dqkm2, dqkm1 = dpkm2 * sc, dpkm1 * sc, dqkm2 * sc, dqkm1 * sc
            if float(delta) < 1e-17 and c > 4:
                break
        out[cf] = xx * (dans + ans * (torch.log(xx) - torch.digamma(aa)))
    return out.reshape(torch.broadcast_shapes(alpha.shape, g.shape))


# Timing mode (bench.py CPU baseline only): use torch's native C++ implicit-gradient approximation
# instead of the exact-but-slow python series above, so the CPU baseline is not dominated by an
# artefact of this restatement.  Never enabled in parity tests.
FAST_GAMMA_GRAD = False


class _GammaDraw(torch.autograd.Function):
    """g(alpha): value supplied from outside (shared with the GPU), gradient implicit."""

    @staticmethod
    def forward(ctx, alpha, g):
        ctx.save_for_backward(alpha, g)
        return g.clone()

    @staticmethod
    def backward(ctx, grad_out):
        alpha, g = ctx.saved_tensors
        if FAST_GAMMA_GRAD:
            return grad_out * torch._standard_gamma_grad(alpha.contiguous(), g.contiguous()), None
        return grad_out * gamma_sample_der_alpha(alpha, g), None


# ---------------------------------------------------------------------------
# Log densities  [EXT: TFP definitions]
# ---------------------------------------------------------------------------
def halfnormal_log_prob(y, scale):
    return HALF_LOG_2_OVER_PI - torch.log(scale) - 0.5 * (y / scale) ** 2


def inverse_gamma_log_prob(y, concentration, scale):
    c = torch.as_tensor(concentration, dtype=DTYPE)
    b = torch.as_tensor(scale, dtype=DTYPE)
    return c * torch.log(b) - torch.lgamma(c) - (c + 1.0) * torch.log(y) - b / y


def sqrt_inverse_gamma_log_prob(y, concentration, scale):
    """[EXT] bayesianquilts.distributions.SqrtInverseGamma: Y = sqrt(X), X ~ InvGamma."""
    return inverse_gamma_log_prob(y * y, concentration, scale) + torch.log(2.0 * y)


def poisson_log_prob(x, rate):
    """[EXT] tfd.Poisson(rate).log_prob(x) = multiply_no_nan(log rate, x) - lgamma(1+x) - rate."""
    xlogr = torch.where(x == 0, torch.zeros_like(rate), x * torch.log(rate))
    return xlogr - torch.lgamma(x + 1.0) - rate


class OraclePoissonFactorization:
    """Float64 restatement of mederrata_spmf/poisson.py::PoissonFactorization (poisson.py:25-717)."""

    def __init__(self, latent_dim, feature_dim, u_tau_scale=0.01, s_tau_scale=1.,
                 symmetry_breaking_decay=0.99, scale_columns=True, scale_rows=True,
                 log_transform=False, column_norms=None, count_key='counts'):
        # poisson.py:85-107
        self.latent_dim = int(feature_dim if latent_dim is None else latent_dim)
        self.feature_dim = int(feature_dim)
        self.u_tau_scale = float(u_tau_scale)
        self.s_tau_scale = float(s_tau_scale)
        self.symmetry_breaking_decay = float(symmetry_breaking_decay)
        self.scale_columns = scale_columns
        self.scale_rows = scale_rows
        self.log_transform = log_transform
        self.count_key = count_key
        self.eta_i = torch.ones(1, self.feature_dim, dtype=DTYPE)      # poisson.py:88
        self.xi_u_global = torch.tensor(1.0, dtype=DTYPE)             # poisson.py:89
        if column_norms is not None:
            self.eta_i = torch.as_tensor(column_norms, dtype=DTYPE).reshape(1, -1)
        self.var_list = list(VAR_LIST)

    # ----- shapes & initial variational parameters (poisson.py:403-539) -----
    def var_shapes(self):
        D, K = self.feature_dim, self.latent_dim
        return {'v': (K, D), 'w': (1, D), 'u': (D, K), 'u_eta': (D, K), 'u_tau': (1, K),
                's_eta': (2, D), 's_tau': (1, D), 's': (2, D), 'u_eta_a': (D, K),
                'u_tau_a': (1, K), 's_eta_a': (2, D), 's_tau_a': (1, D)}

    def init_params(self) -> Dict[str, torch.Tensor]:
        """24 unconstrained tensors: '<var>/loc','<var>/scale_raw' (Normal) or
        '<var>/conc_raw','<var>/scale_raw' (InverseGamma).  [EXT] build_trainable_normal_dist
        keeps loc raw and scale behind a Softplus TransformedVariable; build_trainable_
        InverseGamma_dist keeps both concentration and scale behind Softplus."""
        sh = self.var_shapes()
        ones = lambda k: torch.ones(sh[k], dtype=DTYPE)
        spi = softplus_inverse
        p = {}
        p['v/loc'] = -6. * ones('v');  p['v/scale_raw'] = spi(5e-4 * ones('v'))        # :404-414
        p['w/loc'] = -6. * ones('w');  p['w/scale_raw'] = spi(5e-4 * ones('w'))        # :415-422
        p['u/loc'] = -6. * ones('u');  p['u/scale_raw'] = spi(5e-4 * ones('u'))        # :427-437
        p['s/loc'] = ones('s') * torch.tensor([[-2.], [-1.]], dtype=DTYPE)             # :479-489
        p['s/scale_raw'] = spi(1e-3 * ones('s'))
        ig_init = {
            'u_eta': (3., 1.), 'u_tau': (3., 1.),                                       # :438-459
            's_eta': (1., 1.), 's_tau': (1., 1.),                                       # :463-478
            'u_eta_a': (2., 1.), 'u_tau_a': (2., 1. / self.u_tau_scale ** 2),           # :493-516
            's_eta_a': (2., 1.), 's_tau_a': (2., 1. / self.s_tau_scale ** 2),           # :520-539
        }
        for k, (c, b) in ig_init.items():
            p[k + '/conc_raw'] = spi(c * ones(k))
            p[k + '/scale_raw'] = spi(b * ones(k))
        return p

    @staticmethod
    def param_names():
        names = []
        for k in VAR_LIST:
            names += [k + '/loc', k + '/scale_raw'] if k in NORMAL_VARS else \
                     [k + '/conc_raw', k + '/scale_raw']
        return names

    # ----- compute_scales (poisson.py:113-154) -----
    def compute_scales(self, batches: Iterable[dict]):
        if not self.scale_columns:
            return
        colsum = torch.zeros(1, self.feature_dim, dtype=DTYPE)
        col_nnz = torch.zeros(1, self.feature_dim, dtype=torch.float32)   # fp32 counter, :128
        for batch in batches:
            x = torch.as_tensor(batch[self.count_key])
            colsum += x.to(DTYPE).sum(0, keepdim=True)
            col_nnz += (x > 0).to(torch.float32).sum(0, keepdim=True)
        colmeans_nonzero = colsum / col_nnz.to(DTYPE)                    # :136-138 (0/0 -> nan)
        rowmean_nonzero = colmeans_nonzero.sum()                          # :139-140
        self.eta_i = torch.where(colmeans_nonzero > 1, colmeans_nonzero,
                                 torch.ones_like(colmeans_nonzero))       # :142-149
        self.xi_u_global = rowmean_nonzero if self.scale_rows else torch.tensor(1.0, dtype=DTYPE)

    # variables behind an Identity bijector (none here; v, w in OracleBernoulliFactorization)
    IDENTITY_VARS = ()

    # ----- surrogate draws + log q  [EXT L3] -----
    def sample(self, params, noise):
        """noise[var]: (S,*shape) -- N(0,1) draws for Normal vars, standard-Gamma(alpha) draws
        for InverseGamma vars.  Returns (theta dict with leading S axis, log q of shape (S,)).
        Softplus bijector: y = softplus(t), log q(y) = log base(t) - log sigmoid(t)."""
        theta, logq = {}, 0.
        for k in VAR_LIST:
            if k in NORMAL_VARS:
                loc, sig = params[k + '/loc'], torch.nn.functional.softplus(params[k + '/scale_raw'])
                eps = noise[k].to(DTYPE)
                t = loc + sig * eps
                base = -0.5 * eps ** 2 - torch.log(sig) - HALF_LOG_2PI
            else:
                conc = torch.nn.functional.softplus(params[k + '/conc_raw'])
                beta = torch.nn.functional.softplus(params[k + '/scale_raw'])
                S = noise[k].shape[0]
                g = _GammaDraw.apply(conc.expand(S, *conc.shape), noise[k].to(DTYPE))
                t = beta / g                              # tfd.InverseGamma sample = 1/Gamma(c, rate=scale)
                base = inverse_gamma_log_prob(t, conc, beta)
            if k in self.IDENTITY_VARS:                   # tfb.Identity: y = t, no Jacobian term
                theta[k] = t
                logq = logq + base.sum((-1, -2))
                continue
            theta[k] = torch.nn.functional.softplus(t)
            logq = logq + (base - torch.nn.functional.logsigmoid(t)).sum((-1, -2))
        return theta, logq

    # ----- priors (poisson.py:225-377, horshoe_plus=True branch) -----
    def prior_log_prob_parts(self, th):
        K = self.latent_dim
        c_k = self.symmetry_breaking_decay ** torch.arange(K, dtype=DTYPE)[None, :]   # :225-226
        red = lambda t: t.sum((-1, -2))                                               # Independent(...,2)
        parts = {}
        parts['v'] = red(halfnormal_log_prob(th['v'], torch.tensor(0.1, dtype=DTYPE)))      # :229-235
        parts['w'] = red(halfnormal_log_prob(th['w'], torch.tensor(1.0, dtype=DTYPE)))      # :236-242
        parts['u'] = red(halfnormal_log_prob(th['u'], th['u_eta'] * th['u_tau'] * c_k))     # :247-251
        parts['s'] = red(halfnormal_log_prob(th['s'], th['s_eta'] * th['s_tau']))           # :273-277
        parts['u_eta'] = red(sqrt_inverse_gamma_log_prob(th['u_eta'], 0.5, 1.0 / th['u_eta_a']))   # :303-311
        parts['u_eta_a'] = red(inverse_gamma_log_prob(th['u_eta_a'], 0.5, 1.0))                    # :312-322
        parts['u_tau'] = red(sqrt_inverse_gamma_log_prob(th['u_tau'], 0.5, 1.0 / th['u_tau_a']))   # :323-331
        parts['u_tau_a'] = red(inverse_gamma_log_prob(th['u_tau_a'], 0.5, 1.0 / self.u_tau_scale ** 2))  # :332-341
        parts['s_eta'] = red(sqrt_inverse_gamma_log_prob(th['s_eta'], 0.5, 1.0 / th['s_eta_a']))   # :343-351
        parts['s_eta_a'] = red(inverse_gamma_log_prob(th['s_eta_a'], 0.5, 1.0))                    # :352-359
        parts['s_tau'] = red(sqrt_inverse_gamma_log_prob(th['s_tau'], 0.5, 1.0 / th['s_tau_a']))   # :360-367
        parts['s_tau_a'] = red(inverse_gamma_log_prob(th['s_tau_a'], 0.5, 1.0 / self.s_tau_scale ** 2))  # :368-377
        return parts

    # ----- encoder / decoder (poisson.py:34-54, 623-701) -----
    def encoder_function(self, x):
        if self.log_transform:
            return torch.log(x / self.eta_i + 1.)
        return x / self.eta_i

    def decoder_function(self, y):
        if self.log_transform:
            return torch.exp(y * self.eta_i) - 1.
        return y * self.eta_i

    def encoding_matrix(self, u, s):
        weights = s / s.sum(-2, keepdim=True)                    # :661-662
        return weights[..., 0, :].unsqueeze(-1) * u              # :663-665

    def decoding_matrix(self, v):
        return v                                                 # :677-678

    def intercept_matrix(self, w, s):
        weights = s / s.sum(-2, keepdim=True)                    # :694-696
        return self.eta_i * weights[..., 1, :].unsqueeze(-2) * w  # :697-701  -> (...,1,D)

    def encode(self, x, u, s):
        x = torch.as_tensor(x).to(DTYPE)
        z = torch.matmul(self.encoder_function(x), self.encoding_matrix(u, s))   # :640-643
        if self.scale_rows:
            z = z * (x.sum(-1, keepdim=True) / self.xi_u_global)                  # :644-649
        return z

    @staticmethod
    def observation_log_prob(x, rate):
        return poisson_log_prob(x, rate)                         # :178-183

    def log_likelihood_components(self, s, u, v, w, data, **_):
        x = torch.as_tensor(data[self.count_key]).to(DTYPE)
        theta_u = self.encode(x, u, s)                           # :170
        phi = self.intercept_matrix(w, s)                        # :171
        B = self.decoding_matrix(v)                              # :172
        theta_beta = self.decoder_function(torch.matmul(theta_u, B))   # :174-175
        rate = theta_beta + phi                                  # :177
        return {'log_likelihood': self.observation_log_prob(x, rate), 'rate': rate}   # :178-184

    def unormalized_log_prob_parts(self, data, prior_weight=1., **params):
        parts = self.prior_log_prob_parts(params)                # :590
        parts = {k: v * prior_weight for k, v in parts.items()}  # :591
        ll = self.log_likelihood_components(data=data, **params)['log_likelihood']   # :592-593
        theta = self.encode(data[self.count_key], params['u'], params['s'])          # :598 (2nd encode)
        parts['z'] = halfnormal_log_prob(theta, torch.ones_like(theta)).sum((-1, -2))  # :599-604
        finite = torch.isfinite(ll)
        finite_portion = torch.where(finite, ll, torch.zeros_like(ll))               # :606-608
        min_val = finite_portion.min() - 10.                                          # :609 (global min, differentiable)
        # :611 tf.clip_by_value = maximum(minimum(t, max), min): the gradient of a clipped / NaN entry
        # flows to min_val (and through reduce_min to the entry attaining the minimum)
        ll = torch.maximum(torch.minimum(ll, torch.zeros_like(ll)), min_val)
        ll = torch.where(torch.isfinite(ll), ll, torch.ones_like(ll) * min_val)       # :612-616
        parts['x'] = ll.sum(-1).sum(-1)                                               # :617-619
        return parts

    def unormalized_log_prob(self, data=None, prior_weight=1., **params):
        # :575-580 -- the caller's prior_weight is discarded, literal 1. is passed on.
        parts = self.unormalized_log_prob_parts(data, prior_weight=1., **params)
        return sum(parts.values())

    # ----- the ADVI step integrand [EXT L3]: mean_s[log q - target] -----
    def loss_parts(self, params, noise, data):
        theta, logq = self.sample(params, noise)
        parts = self.unormalized_log_prob_parts(data, **theta)
        return theta, logq, parts

    def loss(self, params, noise, data, entropy_weight=1.0):
        _, logq, parts = self.loss_parts(params, noise, data)
        target = sum(parts.values())
        return (entropy_weight * logq - target).mean()

    def loss_and_grads(self, params, noise, data):
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
        theta, logq, parts = self.loss_parts(leaves, noise, data)
        target = sum(parts.values())
        loss = (logq - target).mean()
        names = self.param_names()
        grads = torch.autograd.grad(loss, [leaves[n] for n in names])
        return (float(loss.detach()), {n: g for n, g in zip(names, grads)},
                {'logq': logq.detach(), **{k: v.detach() for k, v in parts.items()}})


def normal_log_prob(y, scale):
    return -HALF_LOG_2PI - torch.log(scale) - 0.5 * (y / scale) ** 2


class OracleBernoulliFactorization(OraclePoissonFactorization):
    """Float64 restatement of mederrata_spmf/bernoulli.py::BernoulliFactorization (bernoulli.py:31-649):
    Bernoulli-logit likelihood (:148-156), v and w behind Identity bijectors (:186-195) with Normal(0, 0.1) /
    Normal(0, 1) priors (:200-215), encode without row scaling (:580-593), eta_i = 1 unless column norms are
    given (:105-109).  Everything else is inherited, as in the reference."""

    IDENTITY_VARS = ('v', 'w')

    def __init__(self, latent_dim, feature_dim, u_tau_scale=0.01, s_tau_scale=1., symmetry_breaking_decay=0.99,
                 log_transform=False, column_norms=None, count_key='counts'):
        super().__init__(latent_dim, feature_dim, u_tau_scale=u_tau_scale, s_tau_scale=s_tau_scale,
                         symmetry_breaking_decay=symmetry_breaking_decay, scale_columns=True, scale_rows=False,
                         log_transform=log_transform, column_norms=column_norms, count_key=count_key)

    def prior_log_prob_parts(self, th):
        parts = super().prior_log_prob_parts(th)
        red = lambda t: t.sum((-1, -2))
        parts['v'] = red(normal_log_prob(th['v'], torch.tensor(0.1, dtype=DTYPE)))       # bernoulli.py:200-208
        parts['w'] = red(normal_log_prob(th['w'], torch.tensor(1.0, dtype=DTYPE)))       # bernoulli.py:209-215
        return parts

    @staticmethod
    def observation_log_prob(x, rate):
        """[EXT] tfd.Bernoulli(logits).log_prob(x) = -sigmoid_cross_entropy(labels=x, logits) = x*l - softplus(l)."""
        return x * rate - torch.nn.functional.softplus(rate)


def draw_noise(model: OraclePoissonFactorization, params, S, seed=0):
    """N(0,1) / standard-Gamma(alpha) draws for every variable, float32-representable
    (so the GPU, which stores noise in fp32, consumes bit-identical values)."""
    gen = torch.Generator().manual_seed(seed)
    sh = model.var_shapes()
    noise = {}
    for k in VAR_LIST:
        if k in NORMAL_VARS:
            n = torch.randn((S,) + sh[k], generator=gen, dtype=DTYPE)
        else:
            conc = torch.nn.functional.softplus(params[k + '/conc_raw']).expand(S, *sh[k])
            n = torch._standard_gamma(conc.contiguous(), generator=gen)
        noise[k] = n.to(torch.float32).to(DTYPE)
    return noise
