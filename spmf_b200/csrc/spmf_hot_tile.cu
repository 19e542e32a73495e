// Fused dense-tile kernel of the hybrid ADVI step (sm_100a, tcgen05): everything the Poisson
// likelihood needs per (row, hot column, draw) -- rate, x log rate, dL/drate = x/rate, and both
// gradient contractions -- for the dense "hot column" block of a minibatch, without ever writing the
// (S,B,D) rate tensor of the reference (poisson.py:174-184) or its gradient.
//
// One CTA = 128 rows x one Monte-Carlo draw; it walks the hot columns in chunks of 64.  Per chunk:
//   P1  S[128x64]   = Z_s . EV_s^T                 tcgen05.mma, accumulator in tensor memory
//   E   lambda = S + phi ; w = x / lambda (x > 0) ; sum x log lambda      CUDA cores, TMEM -> registers,
//       w written to shared memory as two bf16 terms
//   P2  dZ[128xK]  += W . EV_s                     (poisson.py:177 backward to z)
//   P3  GEV[64xK]   = W^T . Z_s , Gphi = W^T . 1   (backward to the decoder / intercept rows)
// fp32 operands are split into two bf16 terms (hi + lo, 16 mantissa bits); each product is three MMAs
// (hi.hi + hi.lo + lo.hi) accumulated in fp32.  The same shared-memory tiles serve as K-major and as
// MN-major operands (a K-major core matrix of T is an MN-major core matrix of T^T), so nothing is
// transposed in software.  Chunk inputs (the bf16 count tile written by spmf_hot_split and a
// pre-split EV/phi block) arrive by TMA bulk copies, double buffered.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spmf_b200.h"
#include "spmf_record.cuh"
#include "spmf_umma_layout.cuh"
#include "spmf_umma_ptx.cuh"

namespace spmf {

// 16-byte chunk (row r, chunk kc) of a core-matrix tile with `cpr` chunks per row
__host__ __device__ __forceinline__ uint32_t core_off_g(int r, int kc, int cpr) {
  return (uint32_t)((((r >> 3) * cpr + kc) << 7) + ((r & 7) << 4));
}

template <int KP>
struct HotTile {
  static constexpr int KK = KP < 16 ? 16 : KP;          // MMA k extent of the latent dimension
  static constexpr int CPR = KK / 8;                    // chunks per row of an EV tile
  static constexpr int ZCPR = CPR + 1;                  // Z rows carry one more chunk: the ones column
  static constexpr int NZ = KK + 8;                     // N of P3: latent dims | 1 | 7 zeros
  static constexpr int EV_TILE = 64 * KK * 2;           // bytes of one bf16 term
  static constexpr int EV_BLOCK = 2 * EV_TILE + 256;    // hi | lo | phi[64] (fp32)
  static constexpr int X_TILE = kTileABytes;            // 128 x 64 bf16 counts
  static constexpr int STAGE = X_TILE + EV_BLOCK;
  static constexpr int Z_TILE = 128 * ZCPR * 16;
  static constexpr int W_TILE = 128 * 64 * 2;
  static constexpr int SMEM = 2 * STAGE + 2 * Z_TILE + 2 * W_TILE + 128;
  static constexpr int TM_S = 0, TM_DZ = 64, TM_GEV = 64 + KK;   // tensor-memory columns
  static constexpr int TM_COLS = 256;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {     // a -> low half
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

template <int NC>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float* v) {
  static_assert(NC == 8 || NC == 16 || NC == 32, "tmem_ld width");
  uint32_t r[NC];
  if constexpr (NC == 32) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
  } else if constexpr (NC == 16) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  } else {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < NC; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ bool rate_ok_t(float lam) { return lam > 0.f && lam <= 3.402823466e38f; }

// ---- per-step operand prep: EV / phi of the hot columns -> per (draw, chunk) TMA blocks -------------
// block = [hi tile | lo tile | phi[64]], tiles are [64 columns][KK latent] bf16 core-matrix tiles.
template <int KP, int SV>
__global__ void __launch_bounds__(256)
hot_ev_tiles_kernel(const float* __restrict__ EV, const float* __restrict__ PH, int D, int H,
                    unsigned char* __restrict__ EVt) {
  using T = HotTile<KP>;
  const int kc = blockIdx.x, s = blockIdx.y, nch = gridDim.x;
  const int q = s / SV, sv = s - q * SV;
  unsigned char* blk = EVt + ((size_t)s * nch + kc) * T::EV_BLOCK;
  for (int t = threadIdx.x; t < 64 * T::CPR; t += blockDim.x) {
    const int c = t / T::CPR, j = t - c * T::CPR;
    const int d = kc * 64 + c;
    float x[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int k = 8 * j + 4 * h;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (d < H && k < KP)
        v = __ldg(reinterpret_cast<const float4*>(EV + ((size_t)q * D + d) * (SV * KP) + rec_pos(KP, SV, sv, k)));
      x[4 * h + 0] = v.x; x[4 * h + 1] = v.y; x[4 * h + 2] = v.z; x[4 * h + 3] = v.w;
    }
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      hi[p] = pack_bf16(x[2 * p], x[2 * p + 1]);
      lo[p] = pack_bf16(x[2 * p] - bf16_lo(hi[p]), x[2 * p + 1] - bf16_hi(hi[p]));
    }
    const uint32_t o = core_off_g(c, j, T::CPR);
    *reinterpret_cast<uint4*>(blk + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(blk + T::EV_TILE + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    if (j == 0)     // padding columns get rate 1 (their counts are 0, so they contribute nothing)
      reinterpret_cast<float*>(blk + 2 * T::EV_TILE)[c] = d < H ? __ldg(PH + ((size_t)q * D + d) * SV + sv) : 1.f;
  }
}

// ---- the fused tile kernel --------------------------------------------------------------------------
// 256 threads: warp w works on tensor-memory lanes 32*(w&3).. (the rows of the tile) and on column
// half (w>>2) of the chunk, so the element-wise phase has eight warps of ILP per CTA; two CTAs per SM
// interleave their MMA and CUDA-core phases.
constexpr int kTileThreads = 256;

template <int KP, int SV>
__global__ void __launch_bounds__(kTileThreads, 2)
hot_tile_kernel(const unsigned char* __restrict__ xhot, const unsigned char* __restrict__ EVt,
                const float* __restrict__ z, int nrows, int D, int H, int nch,
                float* __restrict__ dzacc, float* __restrict__ rowacc, float* __restrict__ GEV,
                float* __restrict__ Gphi) {
  using T = HotTile<KP>;
  constexpr int KK = T::KK, REC = SV * KP;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  unsigned char* sp = smem_raw + (sbase - smem_u32(smem_raw));
  __shared__ __align__(8) unsigned long long mbar_store[4];
  __shared__ uint32_t tmem_slot;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rt = tid & 127;                 // row of the tile = tensor-memory lane
  const int hcol = tid >> 7;                // column half of the chunk this thread works on
  const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
  const int mt = blockIdx.x, s = blockIdx.y;
  const int q = s / SV, sv = s - q * SV;
  const int row = mt * 128 + rt;

  // shared-memory map: [stage 0 | stage 1 | Z hi | Z lo | W hi | W lo]
  const uint32_t sStage0 = sbase;
  const uint32_t sZ0 = sbase + 2u * T::STAGE, sZ1 = sZ0 + (uint32_t)T::Z_TILE;
  const uint32_t sW0 = sZ1 + (uint32_t)T::Z_TILE, sW1 = sW0 + (uint32_t)T::W_TILE;
  unsigned char* const pZ0 = sp + 2 * T::STAGE;
  unsigned char* const pZ1 = pZ0 + T::Z_TILE;
  unsigned char* const pW0 = pZ1 + T::Z_TILE;
  unsigned char* const pW1 = pW0 + T::W_TILE;

  const uint32_t full0 = smem_u32(&mbar_store[0]), full1 = smem_u32(&mbar_store[1]);
  const uint32_t bar_s = smem_u32(&mbar_store[2]), bar_g = smem_u32(&mbar_store[3]);
  if (tid == 0) {
    mbar_init(full0, 1); mbar_init(full1, 1); mbar_init(bar_s, 1);
    mbar_init(bar_g, 2);          // P2 and P3 are issued (and committed) by two different threads
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)T::TM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // ---- Z_s tile: this draw's z of the 128 rows as two bf16 terms + the ones column.  The two
  // thread halves write alternate 16-byte chunks of the row.
  {
    float zr[KK + 8];
#pragma unroll
    for (int k = 0; k < KK + 8; ++k) zr[k] = 0.f;
    if (row < nrows) {
      const float* zp = z + ((size_t)q * nrows + row) * REC;
#pragma unroll
      for (int k = 0; k < KP; k += 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(zp + rec_pos(KP, SV, sv, k)));
        zr[k] = v.x; zr[k + 1] = v.y; zr[k + 2] = v.z; zr[k + 3] = v.w;
      }
    }
    zr[KK] = 1.f;                               // W^T . 1 = column sums of w  (Gphi)
#pragma unroll
    for (int j = 0; j < T::ZCPR; ++j) {
      if ((j & 1) != hcol) continue;
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float a = zr[8 * j + 2 * p], b = zr[8 * j + 2 * p + 1];
        hi[p] = pack_bf16(a, b);
        lo[p] = pack_bf16(a - bf16_lo(hi[p]), b - bf16_hi(hi[p]));
      }
      const uint32_t o = core_off_g(rt, j, T::ZCPR);
      *reinterpret_cast<uint4*>(pZ0 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(pZ1 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tmem_slot;

  const unsigned char* xsrc = xhot + (size_t)mt * nch * T::X_TILE;
  const unsigned char* esrc = EVt + (size_t)s * nch * T::EV_BLOCK;
  auto issue_load = [&](int chunk) {
    const int st = chunk & 1;
    const uint32_t fb = st ? full1 : full0, dst = sStage0 + (uint32_t)(st * T::STAGE);
    mbar_expect_tx(fb, (uint32_t)T::STAGE);
    tma_bulk_g2s(dst, xsrc + (size_t)chunk * T::X_TILE, T::X_TILE, fb);
    tma_bulk_g2s(dst + T::X_TILE, esrc + (size_t)chunk * T::EV_BLOCK, T::EV_BLOCK, fb);
  };
  // Descriptors are loop invariants up to the stage: build them once (the start-address field is
  // the low 14 bits in 16-byte units, so stepping an operand is one 64-bit add), and spread the MMA
  // issue over three threads -- a lone thread retires ~1 instruction per 4 clocks.
  //   warp 0 lane 0: P2        warp 1 lane 0: P3        warp 2 lane 0: P1 of the next chunk
  //   warp 3 lane 0: TMA for chunk i+2
  constexpr uint64_t kStageStep = (uint64_t)(T::STAGE >> 4);
  // P1: S = Z . EV^T   (both K-major), terms hi.hi, hi.lo, lo.hi
  auto issue_p1 = [&](int st) {
    constexpr uint32_t ID = umma_idesc_bf16(128, 64, 0, 0);
    const uint64_t dZk[2] = {umma_desc(sZ0, 128, T::ZCPR * 128), umma_desc(sZ1, 128, T::ZCPR * 128)};
    const uint64_t dEk[2] = {umma_desc(sStage0 + T::X_TILE, 128, T::CPR * 128),
                             umma_desc(sStage0 + T::X_TILE + T::EV_TILE, 128, T::CPR * 128)};
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const uint64_t a = dZk[t == 2 ? 1 : 0], b = dEk[t == 1 ? 1 : 0] + (uint64_t)st * kStageStep;
#pragma unroll
      for (int j = 0; j < KK / 16; ++j)
        umma_bf16(tm + T::TM_S, a + (uint64_t)(j * 16), b + (uint64_t)(j * 16), ID, (t | j) ? 1u : 0u);
    }
    umma_commit(bar_s);
  };
  if (tid == 96) {
    issue_load(0);
    if (nch > 1) issue_load(1);
  }
  if (tid == 64) {
    mbar_wait(full0, 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    issue_p1(0);
  }

  float xlog2 = 0.f, badacc = 0.f;
  for (int i = 0; i < nch; ++i) {
    const int st = i & 1;
    mbar_wait(st ? full1 : full0, (uint32_t)((i >> 1) & 1));   // TMA data visible to this thread
    mbar_wait(bar_s, (uint32_t)(i & 1));                       // S ready in tensor memory
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- E: lambda, w = x / lambda, x log lambda for 32 columns of this thread's row;
    //      W -> shared memory as two bf16 terms.  The rate is clamped into the positive finite range
    //      instead of branching (poisson.py:606-616 guards non-finite rates): a clamped entry with a
    //      nonzero count is flagged through `badacc`.
    {
      const unsigned char* xt = sp + st * T::STAGE;
      const float* ph = reinterpret_cast<const float*>(xt + T::X_TILE + 2 * T::EV_TILE) + 32 * hcol;
      float lam[32];
      tmem_ld<32>(tm + lane_base + T::TM_S + 32 * hcol, lam);
#pragma unroll
      for (int j = 0; j < 4; ++j) {          // 4 chunks of 8 columns
        const uint32_t o = core_off(rt, 4 * hcol + j);
        const uint4 xp = *reinterpret_cast<const uint4*>(xt + o);
        const uint32_t xw[4] = {xp.x, xp.y, xp.z, xp.w};
        const float4 p0 = *reinterpret_cast<const float4*>(ph + 8 * j);
        const float4 p1 = *reinterpret_cast<const float4*>(ph + 8 * j + 4);
        const float phv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        float w[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float x = (e & 1) ? bf16_hi(xw[e >> 1]) : bf16_lo(xw[e >> 1]);
          const float l = lam[8 * j + e] + phv[e];                 // poisson.py:177
          const float lc = fminf(fmaxf(l, 1e-30f), 3.0e38f);
          float lg, rc;
          asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(lc));
          asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(lc));
          w[e] = x * rc;
          xlog2 = fmaf(x, lg, xlog2);
          badacc = fmaf(x, fabsf(l - lc), badacc);                 // 0 unless the rate left (0, inf)
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          hi[p] = pack_bf16(w[2 * p], w[2 * p + 1]);
          lo[p] = pack_bf16(w[2 * p] - bf16_lo(hi[p]), w[2 * p + 1] - bf16_hi(hi[p]));
        }
        *reinterpret_cast<uint4*>(pW0 + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4*>(pW1 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();

    if (tid == 0) {
      // P2: dZ += W . EV      A = W K-major [128 x 64], B = EV MN-major (N = latent, K = 64 columns)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      constexpr uint32_t ID = umma_idesc_bf16(128, KK, 0, 1);
      const uint64_t dWk[2] = {umma_desc(sW0, 128, 1024), umma_desc(sW1, 128, 1024)};
      const uint64_t dEn[2] = {umma_desc(sStage0 + T::X_TILE, T::CPR * 128, 128),
                               umma_desc(sStage0 + T::X_TILE + T::EV_TILE, T::CPR * 128, 128)};
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const uint64_t a = dWk[t == 2 ? 1 : 0], b = dEn[t == 1 ? 1 : 0] + (uint64_t)st * kStageStep;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          umma_bf16(tm + T::TM_DZ, a + (uint64_t)(j * 16), b + (uint64_t)(j * (2 * T::CPR * 128 / 16)), ID,
                    (i | t | j) ? 1u : 0u);
      }
      umma_commit(bar_g);
    } else if (tid == 32) {
      // P3: GEV = W^T . [Z | 1]   A = W MN-major (M = 64 columns, K = 128 rows), B = Z MN-major (N = NZ)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      constexpr uint32_t ID = umma_idesc_bf16(64, T::NZ, 1, 1);
      const uint64_t dWn[2] = {umma_desc(sW0, 1024, 128), umma_desc(sW1, 1024, 128)};
      const uint64_t dZn[2] = {umma_desc(sZ0, T::ZCPR * 128, 128), umma_desc(sZ1, T::ZCPR * 128, 128)};
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const uint64_t a = dWn[t == 2 ? 1 : 0], b = dZn[t == 1 ? 1 : 0];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          umma_bf16(tm + T::TM_GEV, a + (uint64_t)(j * (2048 / 16)), b + (uint64_t)(j * (2 * T::ZCPR * 128 / 16)), ID,
                    (t | j) ? 1u : 0u);
      }
      umma_commit(bar_g);
    } else if (tid == 64 && i + 1 < nch) {
      // P1 of the next chunk keeps the tensor core busy during the flush below
      mbar_wait(((i + 1) & 1) ? full1 : full0, (uint32_t)(((i + 1) >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_p1((i + 1) & 1);
    }
    __syncwarp();
    mbar_wait(bar_g, (uint32_t)(i & 1));
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 96 && i + 2 < nch) issue_load(i + 2);      // this stage's inputs are fully consumed

    // ---- flush GEV / Gphi of the chunk: accumulator row m sits in lane (m/16)*32 + m%16; the two
    //      thread halves take the two halves of the latent range (the second also the ones column)
    {
      const int c = i * 64 + (warp & 3) * 16 + lane;
      const bool act = lane < 16 && c < H;
      float* ge = GEV + ((size_t)q * D + c) * REC;
      const uint32_t ta = tm + lane_base + T::TM_GEV;
      if constexpr (KK == 32) {
        float g[24];
        if (hcol == 0) {
          tmem_ld<16>(ta, g);
          if (act) {
#pragma unroll
            for (int k = 0; k < 16; k += 4)
              if (k < KP) atomicAdd(reinterpret_cast<float4*>(ge + rec_pos(KP, SV, sv, k)), make_float4(g[k], g[k + 1], g[k + 2], g[k + 3]));
          }
        } else {
          tmem_ld<16>(ta + 16, g);
          tmem_ld<8>(ta + 32, g + 16);
          if (act) {
#pragma unroll
            for (int k = 0; k < 16; k += 4)
              if (16 + k < KP) atomicAdd(reinterpret_cast<float4*>(ge + rec_pos(KP, SV, sv, 16 + k)), make_float4(g[k], g[k + 1], g[k + 2], g[k + 3]));
            atomicAdd(Gphi + ((size_t)q * D + c) * SV + sv, g[16]);
          }
        }
      } else {              // KK == 16: latent dims in columns [0,16), the ones column at 16
        float g[16];
        if (hcol == 0) {
          tmem_ld<16>(ta, g);
          if (act) {
#pragma unroll
            for (int k = 0; k < KP; k += 4)
              atomicAdd(reinterpret_cast<float4*>(ge + rec_pos(KP, SV, sv, k)), make_float4(g[k], g[k + 1], g[k + 2], g[k + 3]));
          }
        } else {
          tmem_ld<8>(ta + 16, g);
          if (act) atomicAdd(Gphi + ((size_t)q * D + c) * SV + sv, g[0]);
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  }

  // ---- dZ of the 128 rows (this CTA is the only writer of its (rows, draw) slice; the two thread
  //      halves own the two halves of the latent range) and the row scalars
  {
    constexpr int HK = KK / 2;
    float dzv[HK];
    const uint32_t ta = tm + lane_base + T::TM_DZ + HK * hcol;
    if constexpr (HK == 16) tmem_ld<16>(ta, dzv); else tmem_ld<8>(ta, dzv);
    if (row < nrows) {
      float* dp = dzacc + ((size_t)q * nrows + row) * REC;
#pragma unroll
      for (int k = 0; k < HK; k += 4) {
        const int kk = HK * hcol + k;
        if (kk < KP) {
          float4* p = reinterpret_cast<float4*>(dp + rec_pos(KP, SV, sv, kk));
          float4 v = *p;
          v.x += dzv[k]; v.y += dzv[k + 1]; v.z += dzv[k + 2]; v.w += dzv[k + 3];
          *p = v;
        }
      }
      float* ra = rowacc + ((size_t)q * nrows + row) * 4 * SV;
      atomicAdd(ra + 0 * SV + sv, xlog2 * 0.6931471805599453f);          // two threads per row
      if (badacc != 0.f) atomicAdd(ra + 3 * SV + sv, 1.0f);              // (also true for NaN)
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 3) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"((uint32_t)T::TM_COLS) : "memory");
  }
}

// ---- row finalisation after the cold row pass and the hot tile kernel have both added into dzr / rowacc
template <int KP, int SV>
__global__ void __launch_bounds__(128)
rows_finish_kernel(const float* __restrict__ rowsum, const float* __restrict__ lgam, float inv_xi, int scale_rows,
                   int nrows, const double* __restrict__ vsum, const float* __restrict__ z,
                   float* __restrict__ dzr, float* __restrict__ rowacc) {
  constexpr int REC = SV * KP;
  // one thread per (row, draw): KP latent dims
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = t / SV, sv = t - row * SV, q = blockIdx.y;
  if (row >= nrows) return;
  const float r = scale_rows ? rowsum[row] * inv_xi : 1.f;
  const float* zp = z + ((size_t)q * nrows + row) * REC;
  float* dp = dzr + ((size_t)q * nrows + row) * REC;
  const double* vs = vsum + (size_t)q * REC;
  float zv = 0.f, z2 = 0.f;
  for (int k = 0; k < KP; ++k) {
    const int p = rec_pos(KP, SV, sv, k);
    const float zz = zp[p], v = (float)vs[p];
    zv = fmaf(zz, v, zv);
    z2 = fmaf(zz, zz, z2);
    dp[p] = r * (dp[p] - v - zz);          // dL/dz incl. the HalfNormal(1) z prior (poisson.py:599-604)
  }
  float* ra = rowacc + ((size_t)q * nrows + row) * 4 * SV;
  ra[0 * SV + sv] -= lgam[row];
  ra[1 * SV + sv] = zv;
  ra[2 * SV + sv] = z2;
}

}  // namespace spmf

using namespace spmf;

#define HT_DISPATCH(KP, SV, CALL)                                   \
  do {                                                              \
    if (KP == 32 && SV == 4) { CALL(32, 4); }                       \
    else if (KP == 32 && SV == 2) { CALL(32, 2); }                  \
    else if (KP == 32 && SV == 1) { CALL(32, 1); }                  \
    else if (KP == 16 && SV == 4) { CALL(16, 4); }                  \
    else if (KP == 16 && SV == 2) { CALL(16, 2); }                  \
    else if (KP == 8 && SV == 4) { CALL(8, 4); }                    \
    else return SPMF_ERR_UNSUPPORTED;                               \
  } while (0)

template <int KP, int SV>
static int launch_hot_tile(const void* xhot, const void* EVt, const float* z, int nrows, int D, int H, int S,
                           float* dzacc, float* rowacc, float* GEV, float* Gphi, cudaStream_t st) {
  using T = HotTile<KP>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(hot_tile_kernel<KP, SV>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  dim3 grid((nrows + 127) / 128, S);
  hot_tile_kernel<KP, SV><<<grid, kTileThreads, T::SMEM, st>>>((const unsigned char*)xhot, (const unsigned char*)EVt, z, nrows, D,
                                                     H, (H + 63) / 64, dzacc, rowacc, GEV, Gphi);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

extern "C" {

long long spmf_hot_tile_scratch_bytes(int H, int K, int S) {
  if (H <= 0 || K <= 0 || S <= 0) return 0;
  const int KP = spmf_kpad(K), KK = KP < 16 ? 16 : KP;
  const long long nch = (H + 63) / 64;
  return (long long)S * nch * (2LL * 64 * KK * 2 + 256);
}

int spmf_hot_ev_tiles(const float* EV, const float* PH, int D, int H, int K, int S, void* EVt, void* stream) {
  if (!EV || !PH || !EVt || D <= 0 || H <= 0 || H > D || K <= 0 || K > SPMF_MAX_K || S <= 0) return SPMF_ERR_BAD_ARG;
  const int KP = spmf_kpad(K), SV = spmf_draw_vec(S);
  dim3 grid((H + 63) / 64, S);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL_EVT(KPC, SVC) hot_ev_tiles_kernel<KPC, SVC><<<grid, 256, 0, st>>>(EV, PH, D, H, (unsigned char*)EVt)
  HT_DISPATCH(KP, SV, CALL_EVT);
#undef CALL_EVT
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

int spmf_hot_tile(const void* xhot, const void* EVt, const float* z, int nrows, int D, int H, int K, int S,
                  float* dzacc, float* rowacc, float* GEVnz, float* Gphinz, void* stream) {
  if (!xhot || !EVt || !z || !dzacc || !rowacc || !GEVnz || !Gphinz) return SPMF_ERR_BAD_ARG;
  if (nrows <= 0 || D <= 0 || H <= 0 || H > D || K <= 0 || K > SPMF_MAX_K || S <= 0) return SPMF_ERR_BAD_ARG;
  const int KP = spmf_kpad(K), SV = spmf_draw_vec(S);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = SPMF_OK;
#define CALL_HT(KPC, SVC) rc = launch_hot_tile<KPC, SVC>(xhot, EVt, z, nrows, D, H, S, dzacc, rowacc, GEVnz, Gphinz, st)
  HT_DISPATCH(KP, SV, CALL_HT);
#undef CALL_HT
  return rc;
}

int spmf_rows_finish(const float* rowsum, const float* lgam, float inv_xi, int scale_rows, int nrows, int K, int S,
                     const double* vsum, const float* z, float* dzr, float* rowacc, void* stream) {
  if (!rowsum || !lgam || !vsum || !z || !dzr || !rowacc || nrows <= 0 || K <= 0 || K > SPMF_MAX_K || S <= 0)
    return SPMF_ERR_BAD_ARG;
  const int KP = spmf_kpad(K), SV = spmf_draw_vec(S), NQ = S / SV;
  dim3 grid((unsigned)(((long long)nrows * SV + 127) / 128), NQ);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL_RF(KPC, SVC) rows_finish_kernel<KPC, SVC><<<grid, 128, 0, st>>>(rowsum, lgam, inv_xi, scale_rows, nrows, vsum, z, dzr, rowacc)
  HT_DISPATCH(KP, SV, CALL_RF);
#undef CALL_RF
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

}  // extern "C"
