#!/usr/bin/env python
"""Generates tests/golden/*.npz from the float64 CPU oracle (oracle/spmf_oracle.py).

The reference ships no golden vectors for this path (tests/spmf_test.py prints, never asserts) and
its TF/TFP/bayesianquilts stack cannot be installed here, so these fixtures pin the ORACLE, not the
reference: they freeze today's oracle outputs so that (a) an accidental change of the oracle is
caught on the CPU and (b) the GPU path is checked against committed numbers as well as against a
live oracle run.  Re-run only when the oracle is changed on purpose:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.spmf_oracle import VAR_LIST, draw_noise  # noqa: E402
from tests.util import make_counts, make_oracle, perturbed_params  # noqa: E402

CASES = {
    # name: (D, K, B, S, kind, perturb)
    "tiny_D7_K3_B5_S2": (7, 3, 5, 2, "noise", 0.3),
    "c1tile_D100_K2_B64_S4": (100, 2, 64, 4, "noise", 0.2),
    "sparse_D150_K32_B48_S4": (150, 32, 48, 4, "sparse", 0.2),
}


def main():
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (D, K, B, S, kind, pert) in CASES.items():
        x = make_counts(B, D, seed=7, kind=kind)
        oracle = make_oracle(D, K, 10 * B, x)
        params = perturbed_params(oracle, pert, seed=7)
        noise = draw_noise(oracle, params, S, seed=8)
        loss, grads, parts = oracle.loss_and_grads(params, noise, {'counts': torch.tensor(x, dtype=torch.float64)})
        blob = {"x": x, "eta": oracle.eta_i.numpy(), "xi": np.float64(oracle.xi_u_global), "loss": np.float64(loss),
                "meta": np.array([D, K, B, S, 10 * B])}
        for k, v in params.items():
            blob["param:" + k] = v.numpy()
        for k, v in noise.items():
            blob["noise:" + k] = v.numpy().astype(np.float32)
        for k, v in grads.items():
            blob["grad:" + k] = v.numpy()
        for k, v in parts.items():
            blob["part:" + k] = v.numpy()
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), **blob)
        print(name, "loss", loss)


if __name__ == "__main__":
    main()
