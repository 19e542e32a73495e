"""The CSV driver (bin/factorize_csv.py): argument contract and output files of the reference's
bin/factorize_csv.py:20-200 -- parser defaults and file names on the CPU, an end-to-end run on the GPU."""
import csv
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cli():
    spec = importlib.util.spec_from_file_location("factorize_csv", os.path.join(ROOT, "bin", "factorize_csv.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_parser_matches_reference_flags_and_defaults():
    a = _cli().build_parser().parse_args([])
    # bin/factorize_csv.py:23-55 of the reference
    assert (a.csv_file, a.epoch, a.dimension, a.batch_size, a.learning_rate, a.clip_value) == (None, 300, 2, 5000, 0.01, 3.0)
    assert a.log_transform is False and a.row_normalize is False
    a = _cli().build_parser().parse_args(["-f", "x.csv", "-e", "7", "-d", "4", "-b", "99", "-lr", "0.5", "-c", "1", "-rn"])
    assert (a.csv_file, a.epoch, a.dimension, a.batch_size, a.learning_rate, a.clip_value, a.row_normalize) == \
        ("x.csv", 7, 4, 99, 0.5, 1.0, True)


def test_output_names_and_loader(tmp_path):
    m = _cli()
    enc, model, rep = m.output_names("data.csv", 3, False, True)
    assert enc == "data.csv_3D_encoding_lt_False_rn_True.csv"            # bin/factorize_csv.py:128-129
    assert model == "data.csv_3D_model_lt_False_rn_True.pkl"             # :136-137
    assert rep == "data.csv_3D_representation_lt_False_rn_True.csv"      # :186-187
    p = tmp_path / "c.csv"
    p.write_text("1,0,2\n0,3,0\n")
    x = m.load_counts(str(p))
    assert x.shape == (2, 3) and x.dtype == np.float32 and x[1, 1] == 3


@pytest.mark.gpu
def test_cli_end_to_end(tmp_path):
    m = _cli()
    rng = np.random.default_rng(0)
    x = rng.poisson(1.5, size=(64, 12))
    x[:, 0] = np.maximum(x[:, 0], 1)
    x[0, :] = np.maximum(x[0, :], 1)
    p = tmp_path / "counts.csv"
    with open(p, "w") as f:
        csv.writer(f).writerows(x.tolist())
    enc, model, rep = m.main(["-f", str(p), "-e", "3", "-d", "2", "-b", "32", "-rn", "--sample-size", "4"])
    E = np.loadtxt(enc, delimiter=",", ndmin=2)
    R = np.loadtxt(rep, delimiter=",", ndmin=2)
    assert E.shape == (2, 12) and np.isfinite(E).all() and (E >= 0).all()
    assert R.shape == (64, 3) and np.array_equal(R[:, 0], np.arange(64)) and np.isfinite(R).all()
    assert os.path.getsize(model) > 0
    import spmf_b200
    again = spmf_b200.PoissonFactorization.load(model)
    assert again.latent_dim == 2 and again.feature_dim == 12
