"""CPU checks of the host-visible pieces of the tensor-core path: the operand tiling index functions
exported by the C ABI, and the arithmetic facts the bf16-split scheme rests on (restated with
torch.bfloat16 on the CPU -- no kernel runs here)."""
import numpy as np
import torch


def test_tiled_a_index_is_a_bijection_in_core_matrix_order():
    from spmf_b200 import _abi
    L = _abi._lib
    rows, kd = 200, 192
    n = L.spmf_umma_tiled_a_elems(rows, kd)
    assert n == 256 * 192                       # rows padded to whole 128-row tiles
    seen = np.full(n, -1, np.int64)
    for r in range(rows):
        for k in range(kd):
            i = L.spmf_umma_tiled_a_index(r, k, kd)
            assert 0 <= i < n and seen[i] < 0
            seen[i] = r * kd + k
    # tile (mt, kc) is contiguous: 128 x 64 elements; inside, 8-row x 8-element core matrices
    assert L.spmf_umma_tiled_a_index(0, 0, kd) == 0
    assert L.spmf_umma_tiled_a_index(0, 7, kd) == 7            # 8 consecutive k: one 16-byte chunk
    assert L.spmf_umma_tiled_a_index(1, 0, kd) == 8            # next row of the core matrix
    assert L.spmf_umma_tiled_a_index(0, 8, kd) == 64           # next core matrix along k (+128 B)
    assert L.spmf_umma_tiled_a_index(8, 0, kd) == 512          # next 8-row group (+1 KiB)
    assert L.spmf_umma_tiled_a_index(0, 64, kd) == 128 * 64    # next k-chunk tile
    assert L.spmf_umma_tiled_a_index(128, 0, kd) == 3 * 128 * 64   # next row tile (3 k-chunks per row tile)
    assert L.spmf_umma_tiled_b_elems(128, 100) == 3 * 128 * 128    # three terms, k padded to 128


def test_counts_up_to_256_are_exact_in_bf16():
    x = torch.arange(0, 300, dtype=torch.float32)
    exact = x.to(torch.bfloat16).float() == x
    assert bool(exact[:257].all())              # the coverage rule of spmf_hot_split
    assert not bool(exact[257])                 # 257 needs 9 significant bits
    assert bool(exact[258]) and bool(exact[264]) and not bool(exact[259])


def _split(x, terms):
    parts, r = [], x.clone()
    for _ in range(terms):
        p = r.to(torch.bfloat16)
        parts.append(p)
        r = r - p.float()
    return parts, r


def test_three_bf16_terms_carry_an_fp32_value_and_two_carry_16_bits():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(100000, generator=g) * torch.exp(3 * torch.randn(100000, generator=g))
    p3, r3 = _split(x, 3)
    rec = sum(p.double() for p in p3)
    assert float(((rec - x.double()).abs() / x.double().abs()).max()) < 2.0 ** -22    # 24 mantissa bits in 3 x 8
    assert float((r3.abs() / x.abs()).max()) < 2.0 ** -22
    p2, r2 = _split(x, 2)
    assert float((r2.abs() / x.abs()).max()) < 2.0 ** -15                             # 16 bits in 2 x 8
    # the products the kernels keep: hi.hi + hi.lo + lo.hi drop only lo.lo ~ 2^-16 * 2^-16... of the product
    y = torch.randn(100000, generator=g)
    (xh, xl), _ = _split(x, 2)
    (yh, yl), _ = _split(y, 2)
    kept = xh.double() * yh.double() + xh.double() * yl.double() + xl.double() * yh.double()
    err = (kept - x.double() * y.double()).abs() / (x.double() * y.double()).abs().clamp_min(1e-300)
    assert float(err.max()) < 3 * 2.0 ** -16


def test_waic_terms_formula():
    """lppd_i = log mean_s exp(ll_si), pwaic_i = var_s(ll_si) (the pieces of PoissonFactorization.waic)."""
    import math
    import numpy as np
    import torch
    from spmf_b200.poisson import waic_terms
    rng = np.random.default_rng(0)
    ll = rng.normal(-50.0, 3.0, size=(6, 11))
    lppd, pw = waic_terms(torch.tensor(ll, dtype=torch.float32).double())
    exp_lppd = np.array([math.log(np.mean(np.exp(ll[:, i]))) for i in range(11)])
    np.testing.assert_allclose(lppd.numpy(), exp_lppd, rtol=1e-6)
    np.testing.assert_allclose(pw.numpy(), ll.var(0, ddof=1), rtol=1e-5)
    one, zero = waic_terms(torch.tensor(ll[:1]))
    assert np.allclose(one.numpy(), ll[0]) and (zero == 0).all()


def test_two_byte_transfer_format_round_trip():
    """encode_u8 (host side of spmf_csr_unpack8): gaps / counts as bytes, wide gaps bridged by zero entries,
    large counts through the overflow list -- decoded here by the definition of the format."""
    import scipy.sparse as sp
    from spmf_b200.data import encode_u8
    rng = np.random.default_rng(3)
    X = sp.random(40, 5000, density=0.004, random_state=2, format='csr')
    X.data = np.ceil(X.data * 600).astype(np.float32)          # counts up to 600: some above 254
    X = X.tolil(); X[7, :] = 0; X[8, 4999] = 3; X = X.tocsr(); X.eliminate_zeros()
    rp, g8, v8, oi, ov = encode_u8(X.indptr, X.indices, X.data)
    assert g8.dtype == np.uint8 and v8.dtype == np.uint8 and rp[0] == 0 and rp[-1] == g8.size == v8.size
    assert g8.size > X.nnz                                       # bridges were needed (mean gap ~250 columns)
    vals = v8.astype(np.float32)
    assert (vals[oi] == 255).all()
    vals[oi] = ov
    D = np.zeros(X.shape, np.float32)
    for r in range(X.shape[0]):
        c = np.cumsum(g8[rp[r]:rp[r + 1]].astype(np.int64) + 1) - 1
        assert (np.diff(c) > 0).all() and (c.size == 0 or c[-1] < X.shape[1])
        D[r, c] = vals[rp[r]:rp[r + 1]]
    assert np.array_equal(D, X.toarray())
