// Global-memory layout of the tcgen05 GEMM operands (shared by the GEMM in spmf_umma.cu and by the
// kernels that produce the operands: spmf_hot_split, spmf_split3_transpose).
#pragma once
#include <stdint.h>

namespace spmf {

constexpr int kGemmBM = 128;      // rows of C per CTA (UMMA M)
constexpr int kGemmBK = 64;       // k elements per stage (4 UMMA k-steps of 16)

// ---- operand layouts in global memory ("UMMA-tiled") ----------------------------------------------
// Both operands are stored tile by tile in exactly the byte order the tensor core reads from shared
// memory (K-major, no swizzle: 8-row x 16-byte core matrices, 128 B each; next k-chunk +128 B, next
// 8-row group +1 KiB), so a stage is filled by two contiguous TMA bulk copies and the DRAM side sees
// whole 16 KiB bursts instead of 128-byte slivers of 128 different rows.
//   A (counts, bf16)  : tiles [mt][kc] of 128 rows x 64 k   = 16 KiB each, mt = row / 128, kc = k / 64
//   B3 (3 bf16 terms) : tiles [kc][t]  of N rows x 64 k     = N*128 B each (t = hi, mid, lo adjacent)
constexpr int kTileABytes = kGemmBM * kGemmBK * 2;

// byte offset of the 16-byte chunk (row r, k-chunk kc8 = (k % 64) / 8) inside a [rows][64] tile
__host__ __device__ __forceinline__ uint32_t core_off(int r, int kc8) {
  return (uint32_t)((((r >> 3) * (kGemmBK / 8) + kc8) << 7) + ((r & 7) << 4));
}
// element offset (in bf16 units) of A(row, k) for a matrix with `kchunks` k-chunks
__host__ __device__ __forceinline__ long long tiledA_index(long long row, long long k, long long kchunks) {
  const long long tile = (row >> 7) * kchunks + (k >> 6);
  return tile * (kTileABytes / 2) + (core_off((int)(row & 127), (int)((k & 63) >> 3)) >> 1) + (k & 7);
}
// element offset of B3(term t, channel c, k) with N channels
__host__ __device__ __forceinline__ long long tiledB_index(int t, int c, long long k, int N) {
  const long long tile = (k >> 6) * 3 + t;
  return tile * ((long long)N * kGemmBK) + (core_off(c, (int)((k & 63) >> 3)) >> 1) + (k & 7);
}

}  // namespace spmf
