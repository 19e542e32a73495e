"""World-size-2 tests of the data-parallel host logic on CPU (gloo): row sharding, the packed
all-reduce block (gradients of v,w,u,s + hi/lo split loss parts) and its unpacking."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeEngine:
    """The slice of AdviEngine that parallel.py touches, on CPU tensors."""

    def __init__(self, D, K, S):
        from spmf_b200.variables import VariableLayout
        self.layout = VariableLayout(D, K, S)
        self.S = S
        self.grads = torch.zeros(self.layout.n_params)
        self.entropy_weight = 1.0
        self.prior_weight = 1.0


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from spmf_b200 import parallel
    D, K, S = 9, 4, 3
    eng = _FakeEngine(D, K, S)
    L = eng.layout
    g = torch.Generator().manual_seed(100 + rank)
    eng.grads[:L.comm_off] = torch.randn(L.comm_off, generator=g)
    eng.grads[L.n_data_block:] = 5.0                       # replicated block: must stay untouched
    parts = torch.zeros(S, 16, dtype=torch.float64)
    parts[:, :13] = torch.arange(13, dtype=torch.float64)  # identical prior/logq parts on every rank
    z_local = torch.tensor([1e9 + 0.125 * rank + s for s in range(S)], dtype=torch.float64)
    x_local = torch.tensor([-3e8 - 0.5 * rank - s for s in range(S)], dtype=torch.float64)
    for s in range(S):                                     # what finalize_parts_kernel writes
        for j, val in ((0, z_local[s]), (2, x_local[s])):
            hi = np.float32(val.item())
            eng.grads[L.comm_off + 4 * s + j] = float(hi)
            eng.grads[L.comm_off + 4 * s + j + 1] = float(np.float32(val.item() - float(hi)))
    local = eng.grads.clone()
    loss = parallel.allreduce_step(eng, parts, None)
    out[rank] = (local, eng.grads.clone(), parts.clone(), float(loss), z_local, x_local)
    dist.destroy_process_group()


def test_shard_rows_partition():
    from spmf_b200.parallel import shard_rows
    for n, w in ((10, 2), (11, 4), (5, 8), (1000003, 8)):
        spans = [shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_allreduce_step_world2():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    (l0, g0, p0, loss0, z0, x0), (l1, g1, p1, loss1, z1, x1) = out[0], out[1]
    from spmf_b200.variables import VariableLayout
    L = VariableLayout(9, 4, 3)
    # summed data-touched gradients, identical on both ranks
    assert torch.equal(g0, g1)
    assert torch.allclose(g0[:L.comm_off], l0[:L.comm_off] + l1[:L.comm_off])
    # replicated block untouched, scalar slack cleared
    assert torch.equal(g0[L.n_data_block:], l0[L.n_data_block:])
    assert float(g0[L.comm_off:L.n_data_block].abs().sum()) == 0.0
    # hi/lo recombination keeps float64 sums (fp32 alone would lose the fractional parts at 1e9)
    assert torch.allclose(p0[:, 13], z0 + z1, rtol=0, atol=1e-4)
    assert torch.allclose(p0[:, 14], x0 + x1, rtol=0, atol=1e-4)
    expect = (p0[:, 12] - p0[:, :12].sum(1) - p0[:, 13] - p0[:, 14]).mean()
    assert loss0 == loss1 == float(expect)
