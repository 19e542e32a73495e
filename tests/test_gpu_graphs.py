"""CUDA-graph replay of the native step (spmf_step_graph_*): a resident batch's step replayed as ONE graph
launch must reproduce the eagerly launched sequence -- same kernels, same streams / dependencies,
per-step scalars (Philox step, Adam step and rates) read from the device step state.  Losses and
parameters agree to fp32 atomics' re-association noise (the tensor-core kernels accumulate split ranges
with floating-point atomics, so two eager runs differ by the same amount)."""
import numpy as np
import pytest
import torch

from tests.util import make_counts

pytestmark = pytest.mark.gpu


def _train(x, K, S, graphs, log_transform=False, epochs=4, lr=0.03):
    import spmf_b200
    from spmf_b200.data import CsrShard
    dev = torch.device("cuda:0")
    shard = CsrShard.from_dense(torch.from_numpy(x), dev)
    m = spmf_b200.PoissonFactorization(latent_dim=K, feature_dim=x.shape[1], u_tau_scale=1e-3, device=dev, seed=11,
                                       log_transform=log_transform)
    m.compute_scales(shard)
    eng = m._engine_for(S)
    eng.use_graphs = graphs
    losses = []
    for ep in range(epochs):
        cur = lr * (0.5 if ep >= 2 else 1.0)              # a learning-rate change must reach the replayed graph
        for b in shard.iter_batches(64):
            losses.append(m.elbo_step({'counts': b}, S, learning_rate=cur, clip_value=2.0))
    torch.cuda.synchronize()
    return eng, torch.stack(losses).cpu()


@pytest.mark.parametrize("D,K,S,kind,log_transform", [
    (192, 16, 4, "linear", False),      # tcgen05 tile-hybrid step: five streams in the graph
    (40, 4, 4, "noise", False),         # gather step
    (300, 32, 4, "sparse", False),      # hot + cold columns
    (48, 8, 2, "noise", True),          # dense link
])
def test_graph_replay_is_bit_identical_to_eager(D, K, S, kind, log_transform):
    x = make_counts(256, D, seed=3, kind=kind)
    e_graph, l_graph = _train(x, K, S, True, log_transform)
    e_eager, l_eager = _train(x, K, S, False, log_transform)
    assert e_graph.graph_launches >= 8 and e_eager.graph_launches == 0
    assert torch.allclose(l_graph, l_eager, rtol=1e-7, atol=0)
    upd = float((e_eager.params - e_graph.params).abs().max())
    assert upd <= 1e-4 * 0.03 * 16, upd              # a fraction of one Adam step
    assert torch.allclose(e_graph.adam_v, e_eager.adam_v, rtol=1e-3, atol=1e-12)
    assert e_graph.opt_step == e_eager.opt_step == 16 and e_graph.rng_step == e_eager.rng_step
    assert bool(torch.isfinite(l_graph).all()) and float(l_graph[-1]) < float(l_graph[0])
