"""Bookkeeping of the 12 latent variables / 24 variational tensors.

Mirrors what `create_distributions` sets up in the reference
(mederrata_spmf/poisson.py:403-573): `var_list` order (the dict insertion order at :572),
shapes, Normal vs InverseGamma families, initial values -- and maps them onto the flat fp32
device buffer the CUDA kernels use (offsets come from spmf_layout() in the C ABI).
"""
from __future__ import annotations

import math

import torch

from . import _abi

# poisson.py:572 -- list(surrogate_dict.keys())
VAR_LIST = ['v', 'w', 'u', 'u_eta', 'u_tau', 's_eta', 's_tau', 's',
            'u_eta_a', 'u_tau_a', 's_eta_a', 's_tau_a']
NORMAL_VARS = ('v', 'w', 'u', 's')
# order of the variables inside the flat buffers (data-touched block first; see include/spmf_b200.h)
INTERNAL_ORDER = ['v', 'w', 'u', 's', 'u_eta', 'u_tau', 's_eta', 's_tau',
                  'u_eta_a', 'u_tau_a', 's_eta_a', 's_tau_a']
PART_NAMES = VAR_LIST + ['logq', 'z', 'x', 'loss']


def var_shapes(D, K):
    """Reference shapes (poisson.py:404-539)."""
    return {'v': (K, D), 'w': (1, D), 'u': (D, K), 'u_eta': (D, K), 'u_tau': (1, K),
            's_eta': (2, D), 's_tau': (1, D), 's': (2, D), 'u_eta_a': (D, K),
            'u_tau_a': (1, K), 's_eta_a': (2, D), 's_tau_a': (1, D)}


def _softplus_inverse(y):
    return y + math.log(-math.expm1(-y))


class VariableLayout:
    """Offsets of every tensor in the flat parameter / gradient / noise buffers."""

    def __init__(self, D, K, S):
        self.D, self.K, self.S = int(D), int(K), int(S)
        self.toff, self.noff = _abi.layout(self.D, self.K, self.S)
        self.n_params = self.toff[-1]
        self.n_noise = self.noff[-1]
        self.shapes = var_shapes(self.D, self.K)
        # number of floats in the all-reduced (data-touched) block: v, w, u, s
        self.n_data_block = self.toff[2 * 4]
        self.comm_slack = 1024                      # kCommSlack in spmf_model.cuh
        self.comm_off = self.n_data_block - self.comm_slack

    # --- internal (device) shape: v is stored transposed as (D,K) ---
    def internal_shape(self, name):
        return (self.D, self.K) if name == 'v' else self.shapes[name]

    def tensor_names(self, name):
        return ((name + '/loc', name + '/scale_raw') if name in NORMAL_VARS
                else (name + '/conc_raw', name + '/scale_raw'))

    def param_names(self):
        """24 names in the reference's variable order (2 per var_list entry)."""
        out = []
        for v in VAR_LIST:
            out += list(self.tensor_names(v))
        return out

    def view(self, flat, name, which):
        """View of tensor `which` (0/1) of variable `name` inside a flat buffer, reference shape."""
        vi = INTERNAL_ORDER.index(name)
        off = self.toff[2 * vi + which]
        shp = self.internal_shape(name)
        n = shp[0] * shp[1]
        t = flat[off:off + n].view(*shp)
        return t.t() if name == 'v' else t

    def views(self, flat):
        """dict '<var>/<tensor>' -> view, reference shapes, reference order."""
        out = {}
        for v in VAR_LIST:
            a, b = self.tensor_names(v)
            out[a] = self.view(flat, v, 0)
            out[b] = self.view(flat, v, 1)
        return out

    def noise_view(self, flat_noise, name, S=None):
        """(S, *reference shape) view of the noise (or sample) buffer for variable `name`."""
        S = self.S if S is None else S
        vi = INTERNAL_ORDER.index(name)
        shp = self.internal_shape(name)
        n = shp[0] * shp[1]
        t = flat_noise[self.noff[vi]:self.noff[vi] + S * n].view(S, *shp)
        return t.transpose(-1, -2) if name == 'v' else t

    def init_values(self, u_tau_scale, s_tau_scale, horseshoe_plus=True):
        """Initial raw values: dict name -> (first, second) python floats or (2,1) columns.

        poisson.py:404-539 (horshoe_plus=True).  [EXT] build_trainable_normal_dist keeps `loc` raw and
        `scale` behind a Softplus TransformedVariable; build_trainable_InverseGamma_dist keeps
        concentration and scale behind Softplus.
        """
        spi = _softplus_inverse
        return {
            'v': (-6., spi(5e-4)), 'w': (-6., spi(5e-4)), 'u': (-6., spi(5e-4)),
            's': ((-2., -1.), spi(1e-3)),
            'u_eta': (spi(3.), spi(1.)), 'u_tau': (spi(3.), spi(1.)),
            's_eta': (spi(1.), spi(1.)), 's_tau': (spi(1.), spi(1.)),
            'u_eta_a': (spi(2.), spi(1.)), 'u_tau_a': (spi(2.), spi(1. / u_tau_scale ** 2)),
            's_eta_a': (spi(2.), spi(1.)), 's_tau_a': (spi(2.), spi(1. / s_tau_scale ** 2)),
        }

    def fill_initial(self, flat, u_tau_scale, s_tau_scale):
        flat.zero_()
        init = self.init_values(u_tau_scale, s_tau_scale)
        for name in VAR_LIST:
            a, b = init[name]
            va, vb = self.view(flat, name, 0), self.view(flat, name, 1)
            if isinstance(a, tuple):
                va[0].fill_(a[0])
                va[1].fill_(a[1])
            else:
                va.fill_(a)
            vb.fill_(b)
        return flat
