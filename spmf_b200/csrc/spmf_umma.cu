// tcgen05 (5th-gen tensor core) GEMM for the two count-matrix products of the ADVI step on the
// dense "hot column" block (sm_100a only):
//
//   encode   z_acc[b][c]  = sum_{d<H} X[b][d] * A'[d][c]        (poisson.py:640-643, x @ encoding)
//   grad A'  GA'[d][c]    = sum_b     X[b][d] * dzr[b][c]       (its transpose in the backward)
//
// Both are C[M][N] += A[M][Kd] * B[N][Kd]^T with A = the counts as bf16 (exact for integer counts
// <= 256; anything else stays on the gather path) and B = an fp32 operand split into three bf16
// terms (hi + mid + lo = 24 mantissa bits), so three tcgen05.mma per k-step reproduce the fp32
// product with fp32 accumulation in tensor memory.  N = SV*KP channels (32, 64 or 128).
//
// One CTA = one 128-row tile of C and a range of K (split-K; partial tiles meet through fp32
// atomics in C).  Operands live in global memory already in the canonical K-major no-swizzle UMMA
// byte order (8-row x 16-byte core matrices, LBO = 128 B, SBO = 1 KiB), tile by tile, so a stage is
// two contiguous TMA bulk copies (cp.async.bulk + mbarrier complete_tx), three stages deep.  One
// thread produces, one thread issues the MMAs and commits them to the stage's "empty" mbarrier; the
// accumulator is read back with tcgen05.ld by the four warps (warp w owns TMEM lanes 32w..32w+31).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spmf_b200.h"
#include "spmf_umma_layout.cuh"
#include "spmf_umma_ptx.cuh"

namespace spmf {

constexpr int kGemmStages = 3;
constexpr int kGemmThreads = 128;

template <int N>
struct GemmSmem {
  static constexpr int A_BYTES = kTileABytes;
  static constexpr int B_BYTES = 3 * N * kGemmBK * 2;            // the three terms of one k-chunk
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = kGemmStages * STAGE_BYTES + 128;
};

// C[q][M][N] += A[M][Kd] . (B3[q][0] + B3[q][1] + B3[q][2])[N][Kd]^T, operands UMMA-tiled (see above).
// Warp 0 lane 0: TMA producer.  Warp 1 lane 0: MMA issuer.  All four warps: epilogue.
// N = channels per CTA; the B3 operand may hold n_total = nblocks * N channels (record shapes wider than
// one MMA): blockIdx.z = q * nblocks + nb, and a stage takes the nb-th N-row block of each term's tile.
template <int N>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm3_kernel(const __nv_bfloat16* __restrict__ A, long long a_qstride, int M,
                  const __nv_bfloat16* __restrict__ B, long long b_qstride, float* __restrict__ C,
                  long long ldc, long long c_qstride, int kchunks, int chunks_per_split, int nblocks) {
  using SM = GemmSmem<N>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  __shared__ __align__(8) unsigned long long mbar_store[2 * kGemmStages + 1];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = blockIdx.x;
  const int m0 = mt * kGemmBM;
  const int q = blockIdx.z / nblocks, nb = blockIdx.z - q * nblocks;
  const int c0 = blockIdx.y * chunks_per_split;
  const int c1 = min(kchunks, c0 + chunks_per_split);
  const int nchunks = c1 - c0;
  if (nchunks <= 0) return;
  A += (long long)q * a_qstride;
  B += (long long)q * b_qstride + (long long)nb * (N * kGemmBK);      // N-row block inside every term tile
  C += (long long)q * c_qstride + (long long)nb * N;
  const long long b_term = (long long)nblocks * N * kGemmBK;          // elements between the terms of one k-chunk

  uint32_t full[kGemmStages], empty[kGemmStages];
#pragma unroll
  for (int i = 0; i < kGemmStages; ++i) {
    full[i] = smem_u32(&mbar_store[i]);
    empty[i] = smem_u32(&mbar_store[kGemmStages + i]);
  }
  const uint32_t done = smem_u32(&mbar_store[2 * kGemmStages]);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kGemmStages; ++i) { mbar_init(full[i], 1); mbar_init(empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = tmem_base_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer: two contiguous bulk copies per stage =====
    const __nv_bfloat16* At = A + ((long long)mt * kchunks + c0) * (SM::A_BYTES / 2);
    const __nv_bfloat16* Bt = B + (long long)c0 * 3 * b_term;
    for (int i = 0; i < nchunks; ++i) {
      const int stage = i % kGemmStages;
      if (i >= kGemmStages) mbar_wait(empty[stage], (uint32_t)((i / kGemmStages - 1) & 1));
      const uint32_t sA = sbase + stage * SM::STAGE_BYTES;
      mbar_expect_tx(full[stage], (uint32_t)SM::STAGE_BYTES);
      tma_bulk_g2s(sA, At + (long long)i * (SM::A_BYTES / 2), SM::A_BYTES, full[stage]);
      if (nblocks == 1) {
        tma_bulk_g2s(sA + SM::A_BYTES, Bt + (long long)i * 3 * b_term, SM::B_BYTES, full[stage]);
      } else {
#pragma unroll
        for (int t = 0; t < 3; ++t)
          tma_bulk_g2s(sA + SM::A_BYTES + t * (SM::B_BYTES / 3), Bt + ((long long)i * 3 + t) * b_term, SM::B_BYTES / 3,
                       full[stage]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the whole warp runs the loop, one elected lane issues (spmf_umma_ptx.cuh) =====
    constexpr uint32_t IDESC = umma_idesc_bf16(kGemmBM, N);
    const uint64_t dA = umma_desc(sbase, 128, (kGemmBK / 8) * 128);
    const uint64_t dB = umma_desc(sbase + SM::A_BYTES, 128, (kGemmBK / 8) * 128);
    for (int i = 0; i < nchunks; ++i) {
      const int stage = i % kGemmStages;
      mbar_wait(full[stage], (uint32_t)((i / kGemmStages) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t so = (uint64_t)stage * (uint64_t)(SM::STAGE_BYTES >> 4);
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int j = 0; j < kGemmBK / 16; ++j)
          umma_bf16_elect(tmem_acc, dA + so + (uint64_t)(j * 16), dB + so + (uint64_t)(t * (N * kGemmBK * 2 / 16) + j * 16),
                          IDESC, (i | t | j) ? 1u : 0u);
      umma_commit_elect(empty[stage]);   // frees the stage once these MMAs have read it
    }
    umma_commit_elect(done);             // commits retire in order: covers every MMA above
  }
  __syncwarp();
  mbar_wait(done, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: TMEM lane = row of the tile; warp w reads lanes 32w..32w+31
  const int row = m0 + tid;
  float* crow = C + (long long)row * ldc;
#pragma unroll
  for (int cb = 0; cb < N / 32; ++cb) {
    float v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32), v);
    if (row < M) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        atomicAdd(reinterpret_cast<float4*>(crow + cb * 32 + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)N) : "memory");
  }
}

// fp32 src[R][C] (row stride lds) -> the UMMA-tiled B3 operand (3 bf16 terms, hi + mid + lo = x to
// 24 bits) with k = the source ROW index (i.e. transposed); k in [R, Rpad) is written as zeros.
__global__ void __launch_bounds__(256)
split3_transpose_kernel(const float* __restrict__ src, long long lds, long long src_qstride, int R, int Rpad,
                        int Ccols, __nv_bfloat16* __restrict__ dst, long long dst_qstride) {
  __shared__ float tile[32][33];
  const int q = blockIdx.z;
  src += (long long)q * src_qstride;
  dst += (long long)q * dst_qstride;
  const int r0 = blockIdx.x * 32, cbase = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + ty + 8 * i, c = cbase + tx;
    tile[ty + 8 * i][tx] = (r < R && c < Ccols) ? src[(long long)r * lds + c] : 0.f;
  }
  __syncthreads();
  // each thread: one channel c, 4 consecutive k (8 bytes of a 16-byte core-matrix row)
  const int c = cbase + (threadIdx.x >> 3), kq = (threadIdx.x & 7) * 4;
  if (c < Ccols && r0 + kq < Rpad) {
    __nv_bfloat16 o[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float x = tile[kq + j][threadIdx.x >> 3];        // zero beyond R
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      const float r1 = x - __bfloat162float(h);
      const __nv_bfloat16 m = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(m);
      o[0][j] = h; o[1][j] = m; o[2][j] = __float2bfloat16_rn(r2);
    }
#pragma unroll
    for (int t = 0; t < 3; ++t)
      *reinterpret_cast<uint2*>(dst + tiledB_index(t, c, r0 + kq, Ccols)) = *reinterpret_cast<const uint2*>(o[t]);
  }
}

// Probe (tests / bring-up): one CTA copies raw operand images into shared memory, issues `nk` MMAs with
// the given descriptor fields and dumps the whole 128-lane x N-column accumulator block.  Used to pin
// the MN-major descriptor conventions and the M=64 accumulator layout against numpy.
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const unsigned char* __restrict__ a_img, int a_bytes, const unsigned char* __restrict__ b_img,
                  int b_bytes, int M, int N, int a_mn, int b_mn, int lbo_a, int sbo_a, int step_a, int lbo_b,
                  int sbo_b, int step_b, int nk, float* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* sp = smem_raw + (sbase - smem_u32(smem_raw));
  __shared__ __align__(8) unsigned long long mb;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int a_pad = (a_bytes + 127) & ~127;
  for (int i = tid; i < a_bytes; i += 128) sp[i] = a_img[i];
  for (int i = tid; i < b_bytes; i += 128) sp[a_pad + i] = b_img[i];
  const uint32_t mbar = smem_u32(&mb);
  if (tid == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tacc = tslot;
  // clear the block so untouched lanes read back as zeros: one dummy-free way is tcgen05.st
  {
    const uint32_t taddr = tacc + ((uint32_t)(warp * 32) << 16);
    for (int c = 0; c < N; ++c)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr + c), "r"(0u) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_bf16(M, N, a_mn, b_mn);
    for (int j = 0; j < nk; ++j) {
      const uint64_t ad = umma_desc(sbase + j * step_a, lbo_a, sbo_a);
      const uint64_t bd = umma_desc(sbase + a_pad + j * step_b, lbo_b, sbo_b);
      umma_bf16(tacc, ad, bd, idesc, j ? 1u : 0u);
    }
    umma_commit(mbar);
  }
  __syncwarp();
  mbar_wait(mbar, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int cb = 0; cb < N / 32; ++cb) {
    float v[32];
    tmem_ld32(tacc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32), v);
    for (int j = 0; j < 32; ++j) out[(long long)tid * N + cb * 32 + j] = v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tacc), "r"(256u) : "memory");
}

// GA'[d][c] += sum_b X[b][d] * (B_hi + B_mid + B_lo)[c][b]  with A = X^T taken straight from the X tiles:
// a K-major [128 rows b][64 columns d] tile IS an MN-major operand with M = d (64), K = b, so no
// transposed copy of the counts is ever built.  One CTA = one 64-column chunk (UMMA M = 64) and a
// range of 64-row k-chunks; a stage = half an X tile (8 KiB, contiguous) + one B3 chunk.
template <int N>
struct GemmAtSmem {
  static constexpr int A_BYTES = kTileABytes / 2;                 // 64 rows x 64 columns
  static constexpr int B_BYTES = 3 * N * kGemmBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = kGemmStages * STAGE_BYTES + 128;
};

template <int N>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm3_at_kernel(const __nv_bfloat16* __restrict__ X, int xchunks, int M,
                     const __nv_bfloat16* __restrict__ B, long long b_qstride, float* __restrict__ C,
                     long long ldc, long long c_qstride, int kchunks, int chunks_per_split, int nblocks) {
  using SM = GemmAtSmem<N>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  __shared__ __align__(8) unsigned long long mbar_store[2 * kGemmStages + 1];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kc = blockIdx.x;                  // 64-column chunk of X = 64 rows of C
  const int q = blockIdx.z / nblocks, nb = blockIdx.z - q * nblocks;
  const int c0 = blockIdx.y * chunks_per_split;
  const int c1 = min(kchunks, c0 + chunks_per_split);
  const int nchunks = c1 - c0;
  if (nchunks <= 0) return;
  B += (long long)q * b_qstride + (long long)nb * (N * kGemmBK);
  C += (long long)q * c_qstride + (long long)nb * N;
  const long long b_term = (long long)nblocks * N * kGemmBK;

  uint32_t full[kGemmStages], empty[kGemmStages];
#pragma unroll
  for (int i = 0; i < kGemmStages; ++i) {
    full[i] = smem_u32(&mbar_store[i]);
    empty[i] = smem_u32(&mbar_store[kGemmStages + i]);
  }
  const uint32_t done = smem_u32(&mbar_store[2 * kGemmStages]);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < kGemmStages; ++i) { mbar_init(full[i], 1); mbar_init(empty[i], 1); }
    mbar_init(done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = tmem_base_slot;

  if (warp == 0) {
    if (elect_one()) {
      // k-chunk c (64 rows of X) = half (c & 1) of X tile (c >> 1, kc)
      const __nv_bfloat16* Bt = B + (long long)c0 * 3 * b_term;
      for (int i = 0; i < nchunks; ++i) {
        const int stage = i % kGemmStages, c = c0 + i;
        if (i >= kGemmStages) mbar_wait(empty[stage], (uint32_t)((i / kGemmStages - 1) & 1));
        const uint32_t sA = sbase + stage * SM::STAGE_BYTES;
        const __nv_bfloat16* At = X + ((long long)(c >> 1) * xchunks + kc) * (kTileABytes / 2) + (c & 1) * (SM::A_BYTES / 2);
        mbar_expect_tx(full[stage], (uint32_t)SM::STAGE_BYTES);
        tma_bulk_g2s(sA, At, SM::A_BYTES, full[stage]);
        if (nblocks == 1) {
          tma_bulk_g2s(sA + SM::A_BYTES, Bt + (long long)i * 3 * b_term, SM::B_BYTES, full[stage]);
        } else {
#pragma unroll
          for (int t = 0; t < 3; ++t)
            tma_bulk_g2s(sA + SM::A_BYTES + t * (SM::B_BYTES / 3), Bt + ((long long)i * 3 + t) * b_term,
                         SM::B_BYTES / 3, full[stage]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    constexpr uint32_t IDESC = umma_idesc_bf16(64, N, 1, 0);                  // A MN-major, B K-major
    const uint64_t dA = umma_desc(sbase, 1024, 128);                          // k-group = 8 rows of X, m-group = 8 columns
    const uint64_t dB = umma_desc(sbase + SM::A_BYTES, 128, (kGemmBK / 8) * 128);
    for (int i = 0; i < nchunks; ++i) {
      const int stage = i % kGemmStages;
      mbar_wait(full[stage], (uint32_t)((i / kGemmStages) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t so = (uint64_t)stage * (uint64_t)(SM::STAGE_BYTES >> 4);
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int j = 0; j < kGemmBK / 16; ++j)
          umma_bf16_elect(tmem_acc, dA + so + (uint64_t)(j * (2048 / 16)),
                          dB + so + (uint64_t)(t * (N * kGemmBK * 2 / 16) + j * 16), IDESC, (i | t | j) ? 1u : 0u);
      umma_commit_elect(empty[stage]);
    }
    umma_commit_elect(done);
  }
  __syncwarp();
  mbar_wait(done, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // ---- epilogue: accumulator row m (M = 64) sits in tensor-memory lane (m/16)*32 + m%16
  const int row = kc * 64 + warp * 16 + lane;
  float* crow = C + (long long)row * ldc;
#pragma unroll
  for (int cb = 0; cb < N / 32; ++cb) {
    float v[32];
    tmem_ld32(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32), v);
    if (lane < 16 && row < M) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        atomicAdd(reinterpret_cast<float4*>(crow + cb * 32 + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)N) : "memory");
  }
}

template <int N>
static int launch_gemm3_at(const __nv_bfloat16* X, int xchunks, int M, const __nv_bfloat16* B, long long bq, float* C,
                           long long ldc, long long cq, int Kd, int NQ, int splits, int nblocks, cudaStream_t st) {
  using SM = GemmAtSmem<N>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm3_at_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int kchunks = Kd / kGemmBK;
  const int mchunks = (M + 63) / 64;
  if (splits <= 0) splits = (2 * 148 + mchunks * NQ * nblocks - 1) / (mchunks * NQ * nblocks);
  if (splits < 1) splits = 1;
  if (splits > kchunks) splits = kchunks;
  const int per = (kchunks + splits - 1) / splits;
  splits = (kchunks + per - 1) / per;
  dim3 grid(mchunks, splits, NQ * nblocks);
  umma_gemm3_at_kernel<N><<<grid, kGemmThreads, SM::TOTAL, st>>>(X, xchunks, M, B, bq, C, ldc, cq, kchunks, per, nblocks);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

template <int N>
static int launch_gemm3(const __nv_bfloat16* A, long long aq, int M, const __nv_bfloat16* B, long long bq, float* C,
                        long long ldc, long long cq, int Kd, int NQ, int splits, int nblocks, cudaStream_t st) {
  using SM = GemmSmem<N>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm3_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int kchunks = Kd / kGemmBK;
  const int mtiles = (M + kGemmBM - 1) / kGemmBM;
  if (splits <= 0) splits = (148 + mtiles * NQ * nblocks - 1) / (mtiles * NQ * nblocks);   // one resident CTA per SM
  if (splits < 1) splits = 1;
  if (splits > kchunks) splits = kchunks;
  const int per = (kchunks + splits - 1) / splits;
  splits = (kchunks + per - 1) / per;
  dim3 grid(mtiles, splits, NQ * nblocks);
  umma_gemm3_kernel<N><<<grid, kGemmThreads, SM::TOTAL, st>>>(A, aq, M, B, bq, C, ldc, cq, kchunks, per, nblocks);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

// row-major bf16 [M][ld] -> UMMA-tiled A (test / tooling helper; the product path writes the tiled
// form directly in spmf_hot_split)
__global__ void tile_a_kernel(const __nv_bfloat16* __restrict__ src, long long ld, int M, int Kd,
                              __nv_bfloat16* __restrict__ dst) {
  const long long n = (long long)((M + 127) / 128 * 128) * Kd;
  const long long kchunks = Kd / kGemmBK;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / Kd, k = i - row * Kd;
    dst[tiledA_index(row, k, kchunks)] = row < M ? src[row * ld + k] : __float2bfloat16_rn(0.f);
  }
}

}  // namespace spmf

using namespace spmf;

extern "C" {

long long spmf_umma_tiled_a_elems(long long M, long long Kd) { return (M + 127) / 128 * 128 * ((Kd + 63) / 64 * 64); }
long long spmf_umma_tiled_b_elems(int N, long long Kd) { return 3LL * N * ((Kd + 63) / 64 * 64); }
long long spmf_umma_tiled_a_index(long long row, long long k, long long Kd) { return tiledA_index(row, k, Kd / kGemmBK); }

int spmf_umma_tile_a(const void* src, long long ld, int M, int Kd, void* dst, void* stream) {
  if (!src || !dst || M <= 0 || Kd <= 0 || Kd % kGemmBK) return SPMF_ERR_BAD_ARG;
  tile_a_kernel<<<148 * 4, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)src, ld, M, Kd, (__nv_bfloat16*)dst);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

int spmf_umma_gemm3(const void* A, long long a_qstride, int M, const void* B3, long long b_qstride, float* C,
                    long long ldc, long long c_qstride, int N, int Kd, int NQ, int splits, void* stream) {
  if (!A || !B3 || !C || M <= 0 || Kd <= 0 || NQ <= 0) return SPMF_ERR_BAD_ARG;
  if (Kd % kGemmBK || ldc % 4) return SPMF_ERR_BAD_ARG;          // whole k-chunks; float4 atomics
  if (((uintptr_t)A | (uintptr_t)B3 | (uintptr_t)C) & 15) return SPMF_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* a = (const __nv_bfloat16*)A;
  const __nv_bfloat16* b = (const __nv_bfloat16*)B3;
  switch (N) {
    case 32: return launch_gemm3<32>(a, a_qstride, M, b, b_qstride, C, ldc, c_qstride, Kd, NQ, splits, 1, st);
    case 64: return launch_gemm3<64>(a, a_qstride, M, b, b_qstride, C, ldc, c_qstride, Kd, NQ, splits, 1, st);
    case 128: case 256: case 512:      // wider records: blocks of 128 channels
      return launch_gemm3<128>(a, a_qstride, M, b, b_qstride, C, ldc, c_qstride, Kd, NQ, splits, N / 128, st);
    default: return SPMF_ERR_UNSUPPORTED;
  }
}

int spmf_umma_gemm3_at(const void* X, int x_kd, int x_rows, int M, const void* B3, long long b_qstride, float* C,
                       long long ldc, long long c_qstride, int N, int NQ, int splits, void* stream) {
  if (!X || !B3 || !C || M <= 0 || x_kd <= 0 || x_rows <= 0 || NQ <= 0) return SPMF_ERR_BAD_ARG;
  if (x_kd % kGemmBK || M > x_kd || ldc % 4) return SPMF_ERR_BAD_ARG;
  if (((uintptr_t)X | (uintptr_t)B3 | (uintptr_t)C) & 15) return SPMF_ERR_BAD_ARG;
  const int Kd = (x_rows + 127) / 128 * 128;          // whole X tiles: rows beyond x_rows are zero there
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* a = (const __nv_bfloat16*)X;
  const __nv_bfloat16* b = (const __nv_bfloat16*)B3;
  switch (N) {
    case 32: return launch_gemm3_at<32>(a, x_kd / kGemmBK, M, b, b_qstride, C, ldc, c_qstride, Kd, NQ, splits, 1, st);
    case 64: return launch_gemm3_at<64>(a, x_kd / kGemmBK, M, b, b_qstride, C, ldc, c_qstride, Kd, NQ, splits, 1, st);
    case 128: case 256: case 512:      // wider records: blocks of 128 channels
      return launch_gemm3_at<128>(a, x_kd / kGemmBK, M, b, b_qstride, C, ldc, c_qstride, Kd, NQ, splits, N / 128, st);
    default: return SPMF_ERR_UNSUPPORTED;
  }
}

int spmf_umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes, int M, int N, int a_mn,
                    int b_mn, int lbo_a, int sbo_a, int step_a, int lbo_b, int sbo_b, int step_b, int nk,
                    float* out, void* stream) {
  if (!a_img || !b_img || !out || a_bytes <= 0 || b_bytes <= 0 || nk <= 0) return SPMF_ERR_BAD_ARG;
  if ((M != 64 && M != 128) || N % 32 || N < 32 || N > 256) return SPMF_ERR_BAD_ARG;
  const int smem = ((a_bytes + 127) & ~127) + b_bytes + 256;
  if (smem > 200 * 1024) return SPMF_ERR_BAD_ARG;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const unsigned char*)a_img, a_bytes,
                                                          (const unsigned char*)b_img, b_bytes, M, N, a_mn, b_mn,
                                                          lbo_a, sbo_a, step_a, lbo_b, sbo_b, step_b, nk, out);
  e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

int spmf_split3_transpose(const float* src, long long lds, long long src_qstride, int R, int Rpad, int Ccols,
                          void* dst3, long long dst_qstride, int NQ, void* stream) {
  if (!src || !dst3 || R <= 0 || Rpad < R || Rpad % kGemmBK || Ccols <= 0 || Ccols % 32 || NQ <= 0) return SPMF_ERR_BAD_ARG;
  dim3 grid((Rpad + 31) / 32, Ccols / 32, NQ);
  split3_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, src_qstride, R, Rpad, Ccols,
                                                                  (__nv_bfloat16*)dst3, dst_qstride);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

}  // extern "C"
