"""spmf_b200 -- B200-native ADVI step for sparse Poisson matrix factorisation.

Drop-in for the `PoissonFactorization` path of mederrata/spmf (mederrata_spmf/__init__.py:1
exports the same name).  Importing this package loads the CUDA C-ABI library built from
spmf_b200/csrc; there is no CPU fallback.
"""
from . import _abi
from ._abi import SpmfError, version
from .data import CsrShard, DeviceBatch, HostCsr, HostDense, as_device_batch
from .engine import AdviEngine
from .poisson import PoissonFactorization
from .bernoulli import BernoulliFactorization
from .variables import VAR_LIST, VariableLayout

__all__ = ["PoissonFactorization", "BernoulliFactorization", "AdviEngine", "CsrShard", "DeviceBatch", "HostCsr", "HostDense", "as_device_batch",
           "VariableLayout", "VAR_LIST", "SpmfError", "version"]
