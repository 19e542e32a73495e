"""Count-matrix containers on the device: CSR shard, per-batch CSC copy, row constants.

The reference streams dense `(B,D)` minibatches through tf.data as dicts keyed by `count_key`
(tests/spmf_test.py:17-27, bin/factorize_csv.py:75-112).  Here the dataset (or this rank's row
shard of it) is resident in HBM as CSR: int64 row pointers, int32 column indices, fp32 counts --
8 B per nonzero -- and every minibatch is a contiguous row range whose CSC copy (for the
column-owned gradients) is built once on the device and cached.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import os

import numpy as np
import torch

from . import _abi


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ceil64(n):
    return (int(n) + 63) // 64 * 64


class StepGraphs:
    """CUDA-graph replays of the native step for ONE resident batch (spmf_step_graph_*): key -> handle.
    Handles are destroyed with the batch."""

    def __init__(self):
        self.handles = {}

    def drop(self, keep_gen):
        for key in [k for k in self.handles if k[1] != keep_gen]:
            _abi._lib.spmf_step_graph_destroy(self.handles.pop(key))

    def __del__(self):
        try:
            for h in self.handles.values():
                _abi._lib.spmf_step_graph_destroy(h)
        except Exception:
            pass


@dataclass
class HotSplit:
    """Hybrid form of a minibatch (spmf_hot_split): ranked + partitioned CSR, its CSC copy, and the
    dense bf16 hot-column block with its transpose -- the operands of the tcgen05 GEMMs."""
    H: int
    rowptr: torch.Tensor          # int64 [nrows+1], zero-based
    cols: torch.Tensor            # int32 [nnz]   column RANKS
    vals: torch.Tensor            # fp32 [nnz]    negative = covered by the tensor-core products
    rowmid: torch.Tensor          # int32 [nrows] first uncovered entry of each row (row-local)
    xhot: torch.Tensor            # bf16, UMMA-tiled X[nrows][ceil64(H)]   (see include/spmf_b200.h)
    xthot: Optional[torch.Tensor] # bf16, UMMA-tiled X^T[H][ceil64(nrows)] (only on request; the step reads xhot)
    colptr: torch.Tensor          # CSC of the entries NOT covered by the GEMMs: int32 [D+1]
    crows: torch.Tensor
    cvals: torch.Tensor
    hcolptr: torch.Tensor         # CSC of the covered entries (valid only if has_hot_csc)
    hcrows: torch.Tensor
    hcvals: torch.Tensor
    has_hot_csc: bool = True
    version: int = 0              # column-ordering version this form was built for


@dataclass
class DeviceBatch:
    """One minibatch on the device.  `rowptr` has nrows+1 entries indexing `cols`/`vals`."""
    rowptr: torch.Tensor          # int64 [nrows+1] (absolute offsets into cols / vals)
    cols: torch.Tensor            # int32 [>= rowptr[-1]]
    vals: torch.Tensor            # fp32
    rowsum: torch.Tensor          # fp32 [nrows]
    lgam: torch.Tensor            # fp32 [nrows]  sum_d lgamma(x+1)
    nrows: int
    nnz: int
    D: int
    colptr: Optional[torch.Tensor] = None   # int32 [D+1]
    crows: Optional[torch.Tensor] = None    # int32 [nnz] batch-local rows
    cvals: Optional[torch.Tensor] = None    # fp32 [nnz]
    hot: Optional[HotSplit] = None          # hybrid form (built for one (rank, H) ordering)
    dense_raw: Optional[torch.Tensor] = None   # dense-ingested batch: the raw [nrows][D] upload (feature order)
    dense_dtype: int = 0                       # SPMF_DENSE_* of dense_raw

    def ensure_hot(self, rank, H, bufs=None, hot_csc=True, row_consts=False, build_xt=False, packed=None,
                   version=0):
        """Build (once) the hybrid form for the column ordering `rank` (int32 [D] device tensor) with
        H hot columns.  `bufs` may supply reusable staging (the streaming uploader).  `hot_csc`: also
        build the CSC copy of the covered entries (only the GEMM-only hybrid mode reads it; the tile
        mode gets those gradients from the tensor-core kernel).  `packed` = (cols16, vals16): read the
        compact upload format instead of self.cols / self.vals (either may be None = use the wide array)."""
        if (self.hot is not None and self.hot.H == H and self.hot.version == version
                and (self.hot.has_hot_csc or not hot_csc)):
            return self.hot
        dev = self.rowptr.device
        n, nnz = self.nrows, self.nnz
        m = max(nnz, 1) + 8
        na = _abi._lib.spmf_umma_tiled_a_elems(n, _ceil64(H))
        nt = _abi._lib.spmf_umma_tiled_a_elems(H, _ceil64(n))
        if bufs is None:
            bufs = dict(rowptr=torch.empty(n + 1, dtype=torch.int64, device=dev),
                        cols=torch.empty(m, dtype=torch.int32, device=dev),
                        vals=torch.empty(m, dtype=torch.float32, device=dev),
                        rowmid=torch.empty(n, dtype=torch.int32, device=dev),
                        xhot=torch.empty(na, dtype=torch.bfloat16, device=dev),
                        xthot=torch.empty(nt, dtype=torch.bfloat16, device=dev) if build_xt else None,
                        colptr=torch.empty(self.D + 1, dtype=torch.int32, device=dev),
                        crows=torch.empty(m, dtype=torch.int32, device=dev),
                        cvals=torch.empty(m, dtype=torch.float32, device=dev),
                        hcolptr=torch.empty(self.D + 1, dtype=torch.int32, device=dev),
                        hcrows=torch.empty(m, dtype=torch.int32, device=dev),
                        hcvals=torch.empty(m, dtype=torch.float32, device=dev),
                        scratch=torch.empty(_abi._lib.spmf_csc_scratch_ints(self.D), dtype=torch.int32, device=dev))
        st = _stream()
        rs, lg = (_ptr(self.rowsum), _ptr(self.lgam)) if row_consts else (None, None)
        if packed is not None and len(packed) == 5:
            # 2-byte transfer format: (gaps8, vals8, ovf_idx, ovf_val, novf) -- expanded inside the split
            g8, v8, oi, ov, novf = packed
            if bufs.get("xthot") is not None:
                raise _abi.SpmfError("the 2-byte split does not build the transposed block")
            _abi.call("spmf_hot_split_u8", _ptr(self.rowptr), _ptr(g8), _ptr(v8), _ptr(oi) if novf else None,
                      _ptr(ov) if novf else None, int(novf), n, _ptr(rank), H, _ptr(bufs["rowptr"]), _ptr(bufs["cols"]),
                      _ptr(bufs["vals"]), _ptr(bufs["rowmid"]), _ptr(bufs["xhot"]), rs, lg, st)
        else:
            c16, v16 = packed if packed is not None else (None, None)
            _abi.call("spmf_hot_split_packed", _ptr(self.rowptr), None if c16 is not None else _ptr(self.cols), _ptr(c16),
                      None if v16 is not None else _ptr(self.vals), _ptr(v16), n, nnz, _ptr(rank), H,
                      _ptr(bufs["rowptr"]), _ptr(bufs["cols"]), _ptr(bufs["vals"]), _ptr(bufs["rowmid"]),
                      _ptr(bufs["xhot"]), _ptr(bufs.get("xthot")), rs, lg, st)
        for part, pre in ((0, "h"), (1, "")):       # covered entries / the rest
            if part == 0 and not hot_csc:
                continue
            _abi.call("spmf_csr_to_csc_part", _ptr(bufs["rowptr"]), _ptr(bufs["rowmid"]), part, _ptr(bufs["cols"]),
                      _ptr(bufs["vals"]), n, self.D, _ptr(bufs[pre + "colptr"]), _ptr(bufs[pre + "crows"]),
                      _ptr(bufs[pre + "cvals"]), _ptr(bufs["scratch"]), st)
        self.hot = HotSplit(H=H, rowptr=bufs["rowptr"], cols=bufs["cols"], vals=bufs["vals"],
                            rowmid=bufs["rowmid"], xhot=bufs["xhot"], xthot=bufs.get("xthot"),
                            colptr=bufs["colptr"], crows=bufs["crows"], cvals=bufs["cvals"],
                            hcolptr=bufs["hcolptr"], hcrows=bufs["hcrows"], hcvals=bufs["hcvals"],
                            has_hot_csc=bool(hot_csc), version=version)
        return self.hot

    def ensure_csc(self):
        if self.colptr is None:
            dev = self.vals.device
            self.colptr = torch.empty(self.D + 1, dtype=torch.int32, device=dev)
            self.crows = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)
            self.cvals = torch.empty(max(self.nnz, 1), dtype=torch.float32, device=dev)
            cursor = torch.empty(_abi._lib.spmf_csc_scratch_ints(self.D), dtype=torch.int32, device=dev)
            _abi.call("spmf_csr_to_csc", _ptr(self.rowptr), _ptr(self.cols), _ptr(self.vals),
                      self.nrows, self.D, _ptr(self.colptr), _ptr(self.crows), _ptr(self.cvals),
                      _ptr(cursor), _stream())
        return self


class CsrShard:
    """This rank's rows of the count matrix, resident on the device."""

    def __init__(self, rowptr, cols, vals, D, device=None):
        device = torch.device(device) if device is not None else torch.device("cuda")
        self.rowptr = torch.as_tensor(rowptr).to(device=device, dtype=torch.int64).contiguous()
        self.cols = torch.as_tensor(cols).to(device=device, dtype=torch.int32).contiguous()
        self.vals = torch.as_tensor(vals).to(device=device, dtype=torch.float32).contiguous()
        self.D = int(D)
        self.nrows = self.rowptr.numel() - 1
        self.nnz = int(self.vals.numel())
        self.rowsum = torch.empty(self.nrows, dtype=torch.float32, device=device)
        self.lgam = torch.empty(self.nrows, dtype=torch.float32, device=device)
        if self.nrows > 0:
            _abi.call("spmf_csr_row_consts", _ptr(self.rowptr), _ptr(self.vals), self.nrows,
                      _ptr(self.rowsum), _ptr(self.lgam), _stream())
        self._rowptr_host = None
        self._batches: Dict[tuple, DeviceBatch] = {}

    # ---- constructors -------------------------------------------------
    @classmethod
    def from_dense(cls, x, device=None):
        """Dense (N,D) counts (torch / numpy, host or device) -> CSR on the device, compacted by a kernel."""
        device = torch.device(device) if device is not None else torch.device("cuda")
        xd = torch.as_tensor(x).to(device=device, dtype=torch.float32).contiguous()
        n, D = xd.shape
        rowptr = torch.empty(n + 1, dtype=torch.int64, device=device)
        _abi.call("spmf_dense_count", _ptr(xd), n, D, _ptr(rowptr), _stream())
        nnz = int(rowptr[-1].item())
        cols = torch.empty(max(nnz, 1), dtype=torch.int32, device=device)
        vals = torch.empty(max(nnz, 1), dtype=torch.float32, device=device)
        _abi.call("spmf_dense_fill", _ptr(xd), n, D, _ptr(rowptr), _ptr(cols), _ptr(vals), _stream())
        return cls(rowptr, cols[:nnz] if nnz else cols[:0], vals[:nnz] if nnz else vals[:0], D, device)

    @classmethod
    def from_scipy(cls, m, device=None):
        m = m.tocsr()
        return cls(torch.from_numpy(m.indptr.astype(np.int64)), torch.from_numpy(m.indices.astype(np.int32)),
                   torch.from_numpy(m.data.astype(np.float32)), m.shape[1], device)

    # ---- batches --------------------------------------------------------
    def rowptr_host(self):
        if self._rowptr_host is None:
            self._rowptr_host = self.rowptr.cpu()
        return self._rowptr_host

    def batch(self, row0, nrows, cache=True) -> DeviceBatch:
        key = (int(row0), int(nrows))
        if cache and key in self._batches:
            return self._batches[key]
        rp = self.rowptr_host()
        nnz = int(rp[row0 + nrows] - rp[row0])
        b = DeviceBatch(rowptr=self.rowptr[row0:row0 + nrows + 1], cols=self.cols, vals=self.vals,
                        rowsum=self.rowsum[row0:row0 + nrows], lgam=self.lgam[row0:row0 + nrows],
                        nrows=int(nrows), nnz=nnz, D=self.D)
        if cache:
            self._batches[key] = b
            b._resident = True          # same object / device arrays every epoch: its step may be replayed as a graph
        return b

    def num_batches(self, batch_rows, drop_remainder=False):
        if drop_remainder:
            return self.nrows // batch_rows
        return (self.nrows + batch_rows - 1) // batch_rows

    def iter_batches(self, batch_rows, order=None, drop_remainder=False):
        nb = self.num_batches(batch_rows, drop_remainder)
        order = range(nb) if order is None else order
        for i in order:
            r0 = i * batch_rows
            yield self.batch(r0, min(batch_rows, self.nrows - r0))

    def column_stats(self):
        """colsum (float64 [D]) and col_nnz (float32 [D]) of this shard -- compute_scales, poisson.py:118-134."""
        dev = self.vals.device
        colsum = torch.zeros(self.D, dtype=torch.float64, device=dev)
        colnnz = torch.zeros(self.D, dtype=torch.float32, device=dev)
        _abi.call("spmf_csr_colstats", _ptr(self.cols), _ptr(self.vals), self.nnz, self.D,
                  _ptr(colsum), _ptr(colnnz), _stream())
        return colsum, colnnz


_UPLOADERS = {}


def as_device_batch(counts, device, D=None) -> DeviceBatch:
    """Accept what a reference-style data factory may yield under `count_key`: a DeviceBatch, a
    dense torch/numpy array, or a scipy.sparse matrix."""
    if isinstance(counts, DeviceBatch):
        return counts
    if isinstance(counts, (HostCsrBatch, HostDenseBatch)):
        up = _UPLOADERS.get((str(device), counts.D))
        if up is None:
            up = _UPLOADERS[(str(device), counts.D)] = BatchUploader(device, counts.D)
        return up.upload(counts)
    if isinstance(counts, CsrShard):
        return counts.batch(0, counts.nrows)
    if hasattr(counts, "tocsr"):
        sh = CsrShard.from_scipy(counts, device)
    else:
        sh = CsrShard.from_dense(counts, device)
    return sh.batch(0, sh.nrows, cache=False)


# ---------------------------------------------------------------------------
# Synthetic generators.  Generated on the host with numpy for small parity
# cases and on the device for the bench shapes.
# ---------------------------------------------------------------------------
def synth_noise_dense(N, D, rate=1.0, seed=0):
    """notebooks/factorizing_random_noise.ipynb:51-62 generator: iid Poisson(rate)."""
    rng = np.random.default_rng(seed)
    return rng.poisson(rate, size=(N, D)).astype(np.float32)


def synth_linear_dense(N, D, k_true=3, seed=0):
    """notebooks/factorize_linear_structure.ipynb:53-67 generator: every 3rd column is
    Poisson(Z.V) with V=|N(1.5,0.5)|, Z=|N(0,1)|, the rest Poisson(1) noise."""
    rng = np.random.default_rng(seed)
    n_sig = len(range(0, D, 3))
    V = np.abs(rng.normal(1.5, 0.5, size=(k_true, n_sig)))
    Z = np.abs(rng.normal(0.0, 1.0, size=(N, k_true)))
    X = rng.poisson(1.0, size=(N, D)).astype(np.float32)
    X[:, ::3] = rng.poisson(Z @ V).astype(np.float32)
    return X


def synth_scrna_csr_device(nrows, D, density=0.05, seed=0, device="cuda", sigma_gene=1.5, sigma_cell=0.5,
                           gene_seed=None):
    """scRNA-seq shaped sparse counts on the device: x_bd ~ Poisson(c_b g_d),
    c_b ~ LogNormal(0, sigma_cell^2), g_d ~ LogNormal(m, sigma_gene^2) with m solved so the mean
    P(x>0) equals `density`.  Returns a CsrShard.  Rows are generated in chunks to bound memory.
    `gene_seed` fixes the gene profile g independently of `seed` (row shards of ONE dataset share
    their genes: every rank passes the same gene_seed and its own seed)."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev).manual_seed(int(seed))
    ggen = gen if gene_seed is None else torch.Generator(device=dev).manual_seed(int(gene_seed))
    g = torch.exp(sigma_gene * torch.randn(D, generator=ggen, device=dev, dtype=torch.float64))
    c = torch.exp(sigma_cell * torch.randn(nrows, generator=gen, device=dev, dtype=torch.float64))
    # solve m: mean_{b,d} (1 - exp(-c_b g_d e^m)) = density on a subsample (bisection)
    cs = c[: min(nrows, 2048)]
    lo, hi = -30.0, 10.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        dens = (1.0 - torch.exp(-cs[:, None] * g[None, :] * np.exp(mid))).mean().item()
        lo, hi = (mid, hi) if dens < density else (lo, mid)
    g = g * np.exp(0.5 * (lo + hi))
    rowptrs, cols_l, vals_l = [torch.zeros(1, dtype=torch.int64, device=dev)], [], []
    base = 0
    chunk = max(1, min(nrows, (1 << 27) // max(D, 1)))
    for r0 in range(0, nrows, chunk):
        r1 = min(nrows, r0 + chunk)
        lam = (c[r0:r1, None] * g[None, :]).to(torch.float32)
        x = torch.poisson(lam, generator=gen)
        # every column keeps at least one nonzero (the reference's compute_scales divides by the
        # per-column nonzero count, poisson.py:136-138: an all-zero column turns xi into NaN) ...
        ds = torch.arange(D, device=dev)
        rs = ds % nrows
        m = (rs >= r0) & (rs < r1)
        x[rs[m] - r0, ds[m]] = torch.clamp(x[rs[m] - r0, ds[m]], min=1.0)
        # ... and every row at least one, so that row scaling is defined
        empty = x.sum(1) == 0
        if bool(empty.any()):
            x[empty, int(torch.argmax(g))] = 1.0
        nzmask = x != 0
        counts = nzmask.sum(1)
        rp = torch.cumsum(counts, 0) + base
        idx = nzmask.nonzero(as_tuple=False)
        cols_l.append(idx[:, 1].to(torch.int32))
        vals_l.append(x[nzmask])
        rowptrs.append(rp)
        base = int(rp[-1].item())
        del lam, x, nzmask, idx
    return CsrShard(torch.cat(rowptrs), torch.cat(cols_l), torch.cat(vals_l), D, dev)


# ---------------------------------------------------------------------------
# Host-resident CSR (pinned) -- the streaming path: each step's minibatch is copied H2D, its row
# constants and CSC copy are built on the device, then the step runs.  This is what a reference
# style `data_factory` yielding host batches costs end to end.
# ---------------------------------------------------------------------------
@dataclass
class HostCsrBatch:
    """One minibatch in pinned host memory, as zero-copy views of its `HostCsr`.  `rowptr` holds ABSOLUTE
    offsets (view of the shard's row pointers); `base` is subtracted on the device after the copy."""
    rowptr: torch.Tensor   # int64 [nrows+1], pinned host; offsets relative to `base`
    cols: Optional[torch.Tensor]     # int32 or uint16 (compact) [nnz], pinned host (None in the 2-byte format)
    vals: Optional[torch.Tensor]     # fp32 or uint16 (compact) [nnz], pinned host (None in the 2-byte format)
    D: int
    base: int = 0
    # 2-byte format (spmf_csr_unpack8): column gaps and counts as bytes + the overflow list of large counts
    gaps8: Optional[torch.Tensor] = None
    vals8: Optional[torch.Tensor] = None
    ovf_idx: Optional[torch.Tensor] = None     # int32 / int64 entry indices; `ovf_base` is subtracted on the device
    ovf_val: Optional[torch.Tensor] = None     # fp32
    ovf_base: int = 0

    @property
    def nrows(self):
        return self.rowptr.numel() - 1

    @property
    def nnz(self):
        return int(self.vals8.numel() if self.vals8 is not None else self.vals.numel())

    def nbytes(self):
        n = self.rowptr.numel() * 8
        for t in (self.cols, self.vals, self.gaps8, self.vals8, self.ovf_idx, self.ovf_val):
            if t is not None:
                n += t.numel() * t.element_size()
        return n


def encode_u8(rowptr, cols, vals):
    """CSR -> the 2-byte transfer format (see spmf_csr_unpack8 in include/spmf_b200.h).  Returns (rowptr',
    gaps8, vals8, ovf_idx, ovf_val) as numpy arrays; gaps wider than 256 columns are bridged by explicit
    zero-valued entries (so rowptr' counts those too), counts that are not integers in [0, 254] go to the
    overflow list."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float32)
    n = rowptr.size - 1
    counts = np.diff(rowptr)
    row_of = np.repeat(np.arange(n), counts)
    order = np.lexsort((cols, row_of))                     # columns ascending inside every row
    cols, vals = cols[order], vals[order]
    prev = np.empty_like(cols)
    prev[1:] = cols[:-1]
    prev[rowptr[:-1][counts > 0]] = -1
    gap = cols - prev - 1
    npad = gap // 256
    rep = npad + 1
    ends = np.cumsum(rep)
    total = int(ends[-1]) if ends.size else 0
    gaps8 = np.full(total, 255, dtype=np.uint8)            # padding entries: +256 columns, count 0
    vals8 = np.zeros(total, dtype=np.uint8)
    pos = ends - 1
    gaps8[pos] = (gap - 256 * npad).astype(np.uint8)
    small = (vals >= 0) & (vals <= 254) & (vals == np.round(vals))
    vals8[pos] = np.where(small, vals, 255).astype(np.uint8)
    ovf = ~small
    new_off = np.concatenate([[0], ends]).astype(np.int64)
    return new_off[rowptr], gaps8, vals8, pos[ovf].astype(np.int64), vals[ovf].astype(np.float32)


class HostCsr:
    """A CSR shard kept in pinned host memory; `batch()` slices are zero-copy views.

    compact=True stores column ids as uint16 when D <= 65536 and counts as uint16 when they are
    integers <= 65535 (count data almost always is): 4 B instead of 8 B per nonzero cross PCIe and
    a kernel widens them on the device.  compact="u8" goes to 2 B per nonzero: one byte for the column
    gap inside the row and one for the count (`encode_u8`), expanded by spmf_csr_unpack8 -- with 8 ranks
    streaming 8,192 x 20,000 batches the H2D traffic is what limits the end-to-end step."""

    def __init__(self, rowptr, cols, vals, D, compact=True):
        self.D = int(D)
        self.u8 = compact == "u8"
        if self.u8:
            rp, g8, v8, oi, ov = encode_u8(torch.as_tensor(rowptr).cpu().numpy(), torch.as_tensor(cols).cpu().numpy(),
                                           torch.as_tensor(vals).cpu().numpy())
            self.rowptr = torch.from_numpy(rp).contiguous().pin_memory()
            self.gaps8 = torch.from_numpy(g8).pin_memory()
            self.vals8 = torch.from_numpy(v8).pin_memory()
            self._ovf_pos = oi                                  # sorted (entries are visited in order)
            self.ovf_idx_all = torch.from_numpy(oi.astype(np.int64)).pin_memory()    # absolute positions in the shard
            self.ovf_val_all = torch.from_numpy(ov).pin_memory()
            self.cols = self.vals = None
            self.nrows = self.rowptr.numel() - 1
            return
        self.rowptr = torch.as_tensor(rowptr).to(torch.int64).contiguous().pin_memory()
        cols = torch.as_tensor(cols).to(torch.int32).contiguous()
        vals = torch.as_tensor(vals).to(torch.float32).contiguous()
        self.nrows = self.rowptr.numel() - 1
        if compact and self.D <= 65536:
            cols = cols.to(torch.uint16)
        if compact and vals.numel() and float(vals.max()) <= 65535 and float(vals.min()) >= 0 \
                and bool((vals == vals.round()).all()):
            vals = vals.to(torch.uint16)
        self.cols, self.vals = cols.pin_memory(), vals.pin_memory()

    @classmethod
    def from_shard(cls, shard: CsrShard, compact=True):
        return cls(shard.rowptr.cpu(), shard.cols.cpu(), shard.vals.cpu(), shard.D, compact)

    def batch(self, row0, nrows) -> HostCsrBatch:
        """Zero-copy views (nothing is allocated or pinned per batch)."""
        j0, j1 = int(self.rowptr[row0]), int(self.rowptr[row0 + nrows])
        rp = self.rowptr[row0:row0 + nrows + 1]
        if self.u8:
            a, b = np.searchsorted(self._ovf_pos, [j0, j1])
            oi = self.ovf_idx_all[a:b] if b > a else None       # (slices of the pinned lists: nothing allocated)
            ov = self.ovf_val_all[a:b] if b > a else None
            return HostCsrBatch(rp, None, None, self.D, base=j0, gaps8=self.gaps8[j0:j1], vals8=self.vals8[j0:j1],
                                ovf_idx=oi, ovf_val=ov, ovf_base=j0)
        return HostCsrBatch(rp, self.cols[j0:j1], self.vals[j0:j1], self.D, base=j0)

    def iter_batches(self, batch_rows):
        for r0 in range(0, self.nrows, batch_rows):
            yield self.batch(r0, min(batch_rows, self.nrows - r0))


_DIAG = os.environ.get("SPMF_DIAG_UPLOAD", "")      # timing diagnosis of the streaming path (bench scripts only)
_DENSE_CODE = {torch.uint8: _abi.DENSE_U8, torch.uint16: _abi.DENSE_U16, torch.float32: _abi.DENSE_F32}


@dataclass
class HostDenseBatch:
    """Rows [row0, row0 + nrows) of a `HostDense`: a zero-copy view of the pinned dense matrix."""
    x: torch.Tensor        # [nrows, D] uint8 / uint16 / float32, pinned host, contiguous
    D: int
    nnz: int               # number of nonzeros (sizes the device staging of the uncovered entries)

    @property
    def nrows(self):
        return self.x.shape[0]

    def nbytes(self):
        return self.x.numel() * self.x.element_size()


class HostDense:
    """A dense count matrix kept in pinned host memory in the narrowest integer type that holds it (uint8 /
    uint16, else float32): what a dense-origin workload (BASELINE C2 / C3; the reference's dense tf.data
    batches, tests/spmf_test.py:17-27) streams.  On the device a batch goes straight to the hybrid form
    (spmf_dense_hot_split) -- no CSR round trip."""

    def __init__(self, x):
        x = torch.as_tensor(x)
        xf = x.to(torch.float32)
        self.nrows, self.D = int(x.shape[0]), int(x.shape[1])
        integral = bool((xf == xf.round()).all()) and float(xf.min()) >= 0
        mx = float(xf.max()) if xf.numel() else 0.0
        dt = torch.uint8 if (integral and mx <= 255) else torch.uint16 if (integral and mx <= 65535) else torch.float32
        self.x = xf.to(dt).contiguous().pin_memory()
        nz = (xf != 0).sum(1).to(torch.int64)
        self._nnz_prefix = torch.cat([torch.zeros(1, dtype=torch.int64), torch.cumsum(nz, 0)])

    def batch(self, row0, nrows) -> HostDenseBatch:
        return HostDenseBatch(self.x[row0:row0 + nrows], self.D,
                              int(self._nnz_prefix[row0 + nrows] - self._nnz_prefix[row0]))

    def iter_batches(self, batch_rows):
        for r0 in range(0, self.nrows, batch_rows):
            yield self.batch(r0, min(batch_rows, self.nrows - r0))


class BatchUploader:
    """Reusable device staging for host batches: async H2D of (rowptr, cols, vals), widening of the
    compact format, then the row constants and the CSC copy are built by kernels on the same stream."""

    def __init__(self, device, D, max_rows=0, max_nnz=0, hot=None):
        self.device, self.D = torch.device(device), int(D)
        # hot = (rank, H[, need_hot_csc])
        self.hot = hot if (hot is not None and hot[0] is not None and hot[1] > 0) else None
        self.hot_csc = bool(hot[2]) if (self.hot is not None and len(hot) > 2) else True
        self.hot_version = int(hot[3]) if (self.hot is not None and len(hot) > 3) else 0
        self.graphs = StepGraphs()          # CUDA-graph replays of the step on this staging slot (engine.py)
        self._alloc(max_rows, max_nnz)

    def _alloc(self, rows, nnz):
        dev = self.device
        self.graphs.drop(None)              # captured graphs point into the old staging buffers
        self.cap_rows, self.cap_nnz = int(rows), int(nnz)
        n = max(nnz, 1) + 8
        self.rowptr = torch.empty(rows + 1, dtype=torch.int64, device=dev)
        self.cols = torch.empty(n, dtype=torch.int32, device=dev)
        self.vals = torch.empty(n, dtype=torch.float32, device=dev)
        self.c16 = torch.empty(n, dtype=torch.uint16, device=dev)
        self.v16 = torch.empty(n, dtype=torch.uint16, device=dev)
        self.g8 = torch.empty(n, dtype=torch.uint8, device=dev)       # 2-byte format staging
        self.v8 = torch.empty(n, dtype=torch.uint8, device=dev)
        self.ovf_i = torch.empty(max(n // 64, 1024), dtype=torch.int32, device=dev)
        self.ovf_v = torch.empty(max(n // 64, 1024), dtype=torch.float32, device=dev)
        self.rowsum = torch.empty(max(rows, 1), dtype=torch.float32, device=dev)
        self.lgam = torch.empty(max(rows, 1), dtype=torch.float32, device=dev)
        self.colptr = torch.empty(self.D + 1, dtype=torch.int32, device=dev)
        self.cursor = torch.empty(_abi._lib.spmf_csc_scratch_ints(self.D), dtype=torch.int32, device=dev)
        self.crows = torch.empty(n, dtype=torch.int32, device=dev)
        self.cvals = torch.empty(n, dtype=torch.float32, device=dev)
        self.hot_bufs = None
        if self.hot is not None:
            H = int(self.hot[1])
            na = _abi._lib.spmf_umma_tiled_a_elems(max(rows, 1), _ceil64(H))
            nt = _abi._lib.spmf_umma_tiled_a_elems(H, _ceil64(max(rows, 1)))
            self.hot_bufs = dict(rowptr=torch.empty(rows + 1, dtype=torch.int64, device=dev),
                                 cols=torch.empty(n, dtype=torch.int32, device=dev),
                                 vals=torch.empty(n, dtype=torch.float32, device=dev),
                                 rowmid=torch.empty(max(rows, 1), dtype=torch.int32, device=dev),
                                 xhot=torch.empty(na, dtype=torch.bfloat16, device=dev),
                                 xthot=None,
                                 colptr=self.colptr, crows=self.crows, cvals=self.cvals,
                                 hcolptr=torch.empty(self.D + 1, dtype=torch.int32, device=dev),
                                 hcrows=torch.empty(n, dtype=torch.int32, device=dev),
                                 hcvals=torch.empty(n, dtype=torch.float32, device=dev), scratch=self.cursor)

    def upload_dense(self, hb: HostDenseBatch) -> DeviceBatch:
        """Dense batch: one H2D copy of the slab, then straight to the hybrid form (tile-hybrid engines), or
        to CSR on the device (everything else)."""
        n, nnz = hb.nrows, hb.nnz
        if n > self.cap_rows or nnz > self.cap_nnz:
            self._alloc(max(n, self.cap_rows), max(int(nnz * 1.25), self.cap_nnz))
        dt = hb.x.dtype
        raw = getattr(self, "raw", None)
        if raw is None or raw.dtype != dt or raw.numel() < self.cap_rows * self.D:
            self.raw = raw = torch.empty(max(self.cap_rows, n) * self.D, dtype=dt, device=self.device)
            self.graphs.drop(None)
        rawv = raw[:n * self.D].view(n, self.D)
        rawv.copy_(hb.x, non_blocking=True)
        if self.hot is None or self.hot_csc:
            # no fused tile kernel on the consuming engine: compact to CSR on the device (host sync for nnz)
            sh = CsrShard.from_dense(rawv, self.device)
            return sh.batch(0, n, cache=False)
        H, hb_ = int(self.hot[1]), self.hot_bufs
        _abi.call("spmf_dense_hot_split", _ptr(raw), _DENSE_CODE[dt], n, self.D, _ptr(self.hot[0]), H,
                  _ptr(hb_["rowptr"]), _ptr(hb_["cols"]), _ptr(hb_["vals"]), _ptr(hb_["rowmid"]), _ptr(hb_["xhot"]),
                  _ptr(self.rowsum), _ptr(self.lgam), _stream())
        _abi.call("spmf_csr_to_csc_part", _ptr(hb_["rowptr"]), _ptr(hb_["rowmid"]), 1, _ptr(hb_["cols"]),
                  _ptr(hb_["vals"]), n, self.D, _ptr(hb_["colptr"]), _ptr(hb_["crows"]), _ptr(hb_["cvals"]),
                  _ptr(hb_["scratch"]), _stream())
        db = DeviceBatch(rowptr=None, cols=None, vals=None, rowsum=self.rowsum[:n], lgam=self.lgam[:n], nrows=n,
                         nnz=nnz, D=self.D, dense_raw=raw, dense_dtype=_DENSE_CODE[dt])
        db.hot = HotSplit(H=H, rowptr=hb_["rowptr"], cols=hb_["cols"], vals=hb_["vals"], rowmid=hb_["rowmid"],
                          xhot=hb_["xhot"], xthot=None, colptr=hb_["colptr"], crows=hb_["crows"], cvals=hb_["cvals"],
                          hcolptr=hb_["hcolptr"], hcrows=hb_["hcrows"], hcvals=hb_["hcvals"], has_hot_csc=False,
                          version=self.hot_version)
        db._resident, db._step_graphs, db._nnz_bound = True, self.graphs, self.cap_nnz
        return db

    def upload(self, hb) -> DeviceBatch:
        if isinstance(hb, HostDenseBatch):
            return self.upload_dense(hb)
        n, nnz = hb.nrows, hb.nnz
        if n > self.cap_rows or nnz > self.cap_nnz:
            self._alloc(max(n, self.cap_rows), max(int(nnz * 1.25), self.cap_nnz))
        st = _stream()
        self.rowptr[:n + 1].copy_(hb.rowptr, non_blocking=True)
        if hb.base:
            self.rowptr[:n + 1].sub_(hb.base)              # the host sends a view of its absolute row pointers
        c16 = v16 = packed8 = None
        if hb.gaps8 is not None:
            # 2 bytes per nonzero over PCIe; expanded to int32 / fp32 by one kernel (+ a patch of the few large counts)
            self.g8[:nnz].copy_(hb.gaps8, non_blocking=True)
            self.v8[:nnz].copy_(hb.vals8, non_blocking=True)
            novf = 0 if hb.ovf_idx is None else hb.ovf_idx.numel()
            if novf > self.ovf_i.numel():
                self.ovf_i = torch.empty(2 * novf, dtype=torch.int32, device=self.device)
                self.ovf_v = torch.empty(2 * novf, dtype=torch.float32, device=self.device)
            if novf:
                if hb.ovf_base:
                    # absolute int64 positions: rebased on the device, then narrowed to the kernels' int32
                    if getattr(self, "ovf_i64", None) is None or self.ovf_i64.numel() < novf:
                        self.ovf_i64 = torch.empty(max(2 * novf, 1024), dtype=torch.int64, device=self.device)
                    self.ovf_i64[:novf].copy_(hb.ovf_idx, non_blocking=True)
                    self.ovf_i64[:novf].sub_(int(hb.ovf_base))
                    self.ovf_i[:novf].copy_(self.ovf_i64[:novf])
                else:
                    self.ovf_i[:novf].copy_(hb.ovf_idx, non_blocking=True)
                self.ovf_v[:novf].copy_(hb.ovf_val, non_blocking=True)
            if self.hot is not None and os.environ.get("SPMF_FUSED_UNPACK8", "1") != "0":
                packed8 = (self.g8, self.v8, self.ovf_i, self.ovf_v, novf)       # ... inside the hot split
            else:
                _abi.call("spmf_csr_unpack8", _ptr(self.rowptr), _ptr(self.g8), _ptr(self.v8), n, _ptr(self.ovf_i),
                          _ptr(self.ovf_v), novf, _ptr(self.cols), _ptr(self.vals), st)
        elif hb.cols.dtype == torch.uint16:
            c16 = self.c16[:nnz]
            c16.copy_(hb.cols, non_blocking=True)
        else:
            self.cols[:nnz].copy_(hb.cols, non_blocking=True)
        if hb.gaps8 is not None:
            pass
        elif hb.vals.dtype == torch.uint16:
            v16 = self.v16[:nnz]
            v16.copy_(hb.vals, non_blocking=True)
        else:
            self.vals[:nnz].copy_(hb.vals, non_blocking=True)
        if _DIAG == "skip_prep" and getattr(self, "_diag_db", None) is not None:
            return self._diag_db              # (timing diagnosis only: copies done, device-side build skipped)
        if self.hot is not None:
            # hybrid form: widen, row constants, then the ranked/partitioned CSR + dense bf16 block and
            # its CSC copy -- all on this (copy) stream, into persistent staging
            # (the split reads the compact arrays directly; the wide cols / vals of such a batch are not
            # materialised -- the hybrid step reads the ranked copies only)
            db = DeviceBatch(rowptr=self.rowptr[:n + 1], cols=None if (c16 is not None or packed8) else self.cols,
                             vals=None if (v16 is not None or packed8) else self.vals, rowsum=self.rowsum[:n],
                             lgam=self.lgam[:n], nrows=n, nnz=nnz, D=self.D)
            db.ensure_hot(self.hot[0], int(self.hot[1]), bufs=self.hot_bufs, hot_csc=self.hot_csc,
                          row_consts=True,          # row constants come out of the split's first pass
                          packed=packed8 or (c16, v16), version=self.hot_version)
            # same device arrays for every batch through this slot: the step may be replayed as a graph
            db._resident, db._step_graphs, db._nnz_bound = True, self.graphs, self.cap_nnz
            self._diag_db = db
            return db
        _abi.call("spmf_prepare_batch", _ptr(c16), _ptr(v16), _ptr(self.rowptr), _ptr(self.cols),
                  _ptr(self.vals), n, nnz, self.D, _ptr(self.rowsum), _ptr(self.lgam), _ptr(self.colptr),
                  _ptr(self.crows), _ptr(self.cvals), _ptr(self.cursor), st)
        return DeviceBatch(rowptr=self.rowptr[:n + 1], cols=self.cols, vals=self.vals,
                           rowsum=self.rowsum[:n], lgam=self.lgam[:n], nrows=n, nnz=nnz, D=self.D,
                           colptr=self.colptr, crows=self.crows, cvals=self.cvals)


_PREFETCH_POOL = {}


def prefetch_to_device(host_batches, device, depth=2, hot=None):
    """Generator: uploads HostCsrBatch items `depth` ahead on a copy stream (H2D, widening, row
    constants, CSC build) while the consumer computes on the current stream.  What tf.data's
    `prefetch(AUTOTUNE)` does for the reference's drivers (bin/factorize_csv.py:110-112)."""
    device = torch.device(device)
    it = iter(host_batches)
    # staging buffers and the copy stream persist across calls: no allocation once warmed up
    pool = _PREFETCH_POOL.setdefault(str(device), {"stream": torch.cuda.Stream(device=device), "ups": {}})
    copy, ups = pool["stream"], pool["ups"]
    hot_key = None if hot is None or hot[0] is None or hot[1] <= 0 else (hot[0].data_ptr(), int(hot[1]), tuple(hot[2:]))
    freed = [None] * depth
    queue = []

    def issue(slot):
        try:
            hb = next(it)
        except StopIteration:
            return False
        if not isinstance(hb, (HostCsrBatch, HostDenseBatch)):
            hb = hb["counts"] if isinstance(hb, dict) else hb
        key = (slot, hb.D, hot_key)
        if key not in ups:
            ups[key] = BatchUploader(device, hb.D, hot=hot)
        if freed[slot] is not None:
            copy.wait_event(freed[slot])            # the step that used this staging set is done
        with torch.cuda.stream(copy):
            db = ups[key].upload(hb)
            ev = torch.cuda.Event()
            ev.record(copy)
        queue.append((db, ev, slot))
        return True

    for sidx in range(depth):
        if not issue(sidx):
            break
    while queue:
        db, ev, slot = queue.pop(0)
        main = torch.cuda.current_stream()
        main.wait_event(ev)
        yield db
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())    # consumer has enqueued its step by now
        freed[slot] = done
        issue(slot)
