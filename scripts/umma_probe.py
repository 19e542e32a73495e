#!/usr/bin/env python
"""Bring-up probe of the tcgen05 conventions the fused tile kernel relies on (run on a B200):
MN-major operand descriptors and the accumulator layout for M=64.  Prints which hypothesis matches."""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from spmf_b200 import _abi  # noqa: E402

dev = torch.device("cuda:0")


def tile_img(mat):
    """[R][Kc] -> bytes of the row-major core-matrix tile: 8-row x 16-byte cores, Kc/8 cores per row group."""
    R, Kc = mat.shape
    cpr = Kc // 8
    t = torch.tensor(mat, dtype=torch.float32).to(torch.bfloat16).view(R // 8, 8, cpr, 8)   # rg, r, kc, k
    return t.permute(0, 2, 1, 3).contiguous().view(torch.uint8).flatten().to(dev)            # rg, kc, r, k


def run(a_img, b_img, M, N, a_mn, b_mn, la, sa, sta, lb, sb, stb, nk):
    out = torch.full((128, N), -7.0, dtype=torch.float32, device=dev)
    _abi.call("spmf_umma_probe", a_img.data_ptr(), a_img.numel(), b_img.data_ptr(), b_img.numel(), M, N, a_mn, b_mn,
              la, sa, sta, lb, sb, stb, nk, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return out.cpu().numpy()


rng = np.random.default_rng(0)
W = rng.integers(-3, 4, size=(128, 64)).astype(np.float32)
EV = rng.integers(-3, 4, size=(64, 32)).astype(np.float32)
Z = rng.integers(-3, 4, size=(128, 32)).astype(np.float32)
Bk = rng.integers(-3, 4, size=(32, 64)).astype(np.float32)

# A. both K-major (the configuration the GEMM uses)
out = run(tile_img(W), tile_img(Bk), 128, 32, 0, 0, 128, 1024, 256, 128, 1024, 256, 4)
print("A  K-major x K-major:", np.array_equal(out, W @ Bk.T))

# B. A = W K-major, B = EV^T taken MN-major from the [c][k_lat] tile
ref = W @ EV
for (lb, sb) in ((512, 128),):
    out = run(tile_img(W), tile_img(EV), 128, 32, 0, 1, 128, 1024, 256, lb, sb, 1024, 4)
    print(f"B  B MN-major LBO={lb} SBO={sb}:", np.array_equal(out, ref))

# C. A = W^T MN-major (M=64 columns of W, K=128 rows), B = Z^T MN-major; find the lane layout of M=64
ref = W.T @ Z          # [64][32]
for (la, sa), (lb, sb) in itertools.product(((1024, 128),), ((512, 128),)):
    out = run(tile_img(W), tile_img(Z), 64, 32, 1, 1, la, sa, 2048, lb, sb, 1024, 8)
    lanes = []
    for m in range(64):
        hit = [l for l in range(128) if np.array_equal(out[l], ref[m])]
        lanes.append(hit[0] if len(hit) >= 1 else -1)
    ok = all(l >= 0 for l in lanes)
    print(f"C  A MN (LBO={la},SBO={sa}) B MN (LBO={lb},SBO={sb}): rows found={ok}",
          "lane map:", lanes[:20], "..." if ok else "")
    if ok:
        print("   full lane map:", lanes)

# D. same as C with M=128 (W2: 128 x 128)
W2 = rng.integers(-3, 4, size=(128, 128)).astype(np.float32)
ref = W2.T @ Z
for (la, sa) in ((2048, 128),):
    out = run(tile_img(W2), tile_img(Z), 128, 32, 1, 1, la, sa, 2 * 2048, 512, 128, 1024, 8)
    print(f"D  M=128 A MN (LBO={la},SBO={sa}):", np.array_equal(out, ref))
