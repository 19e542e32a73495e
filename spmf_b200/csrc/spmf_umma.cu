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
// atomics in C).  Operands are staged by cp.async into the canonical K-major no-swizzle UMMA layout
// (8-row x 16-byte core matrices, LBO = 128 B, SBO = 1 KiB), three stages deep; one thread issues
// the MMAs and commits them to an mbarrier per stage; the accumulator is read back with tcgen05.ld
// by the four warps (warp w owns TMEM lanes 32w..32w+31).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spmf_b200.h"

namespace spmf {

constexpr int kGemmBM = 128;      // rows of C per CTA (UMMA M)
constexpr int kGemmBK = 64;       // k elements per stage (4 UMMA k-steps of 16)
constexpr int kGemmStages = 3;
constexpr int kGemmThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// 64-bit shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), leading byte offset >> 4 in [16,30) (next 16-byte chunk along K),
// stride byte offset >> 4 in [32,46) (next group of 8 rows), version = 1 in [46,48).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = bf16, both K-major
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar)
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!done);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// One stage: A tile [128][64] bf16 (16 KiB) followed by three B tiles [N][64] bf16 (N*128 B each).
template <int N>
struct GemmSmem {
  static constexpr int A_BYTES = kGemmBM * kGemmBK * 2;
  static constexpr int B_BYTES = N * kGemmBK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + 3 * B_BYTES;
  static constexpr int TOTAL = kGemmStages * STAGE_BYTES + 128;
};

// canonical K-major no-swizzle offset of the 16-byte chunk (row r, k-chunk kc) inside a [rows][64] tile
__device__ __forceinline__ uint32_t core_off(int r, int kc) {
  return (uint32_t)((((r >> 3) * (kGemmBK / 8) + kc) << 7) + ((r & 7) << 4));
}

template <int N>
__global__ void __launch_bounds__(kGemmThreads, 1)
umma_gemm3_kernel(const __nv_bfloat16* __restrict__ A, long long lda, long long a_qstride, int M,
                  const __nv_bfloat16* __restrict__ B, long long ldb, long long b_tstride,
                  long long b_qstride, float* __restrict__ C, long long ldc, long long c_qstride,
                  int kchunks, int chunks_per_split) {
  using SM = GemmSmem<N>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // 128-byte aligned base (descriptor addresses are in 16-byte units)
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  __shared__ __align__(8) unsigned long long mbar_store[kGemmStages + 1];
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * kGemmBM;
  const int q = blockIdx.z;
  const int c0 = blockIdx.y * chunks_per_split;
  const int c1 = min(kchunks, c0 + chunks_per_split);
  const int nchunks = c1 - c0;
  if (nchunks <= 0) return;
  A += (long long)q * a_qstride;
  B += (long long)q * b_qstride;
  C += (long long)q * c_qstride;

  uint32_t mbar[kGemmStages + 1];
#pragma unroll
  for (int i = 0; i <= kGemmStages; ++i) mbar[i] = smem_u32(&mbar_store[i]);
  if (tid == 0) {
#pragma unroll
    for (int i = 0; i <= kGemmStages; ++i) mbar_init(mbar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                 "r"((uint32_t)N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_acc = tmem_base_slot;

  // ---- loader: 16-byte chunks; thread -> (row group, k-chunk) so that 8 consecutive lanes write one
  // contiguous 128-byte core matrix (rows r..r+7 of one k-chunk)
  auto load_stage = [&](int stage, int chunk) {
    const uint32_t sA = sbase + stage * SM::STAGE_BYTES;
    const long long k0 = (long long)chunk * kGemmBK;
    // A: 128 rows x 8 chunks = 1024 chunks / 128 threads
#pragma unroll
    for (int it = 0; it < (kGemmBM * 8) / kGemmThreads; ++it) {
      const int idx = it * kGemmThreads + tid;
      const int r = (idx & 7) | ((idx >> 6) << 3);      // 8 lanes = 8 rows of one core matrix
      const int kc = (idx >> 3) & 7;
      const int row = m0 + r;
      const bool ok = row < M;
      const __nv_bfloat16* src = A + (long long)(ok ? row : 0) * lda + k0 + kc * 8;
      cp_async16(sA + core_off(r, kc), src, ok ? 16u : 0u);
    }
    // B: three [N][64] tiles
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const uint32_t sB = sA + SM::A_BYTES + t * SM::B_BYTES;
      const __nv_bfloat16* Bt = B + (long long)t * b_tstride;
#pragma unroll
      for (int it = 0; it < (N * 8) / kGemmThreads; ++it) {
        const int idx = it * kGemmThreads + tid;
        const int r = (idx & 7) | ((idx >> 6) << 3);
        const int kc = (idx >> 3) & 7;
        cp_async16(sB + core_off(r, kc), Bt + (long long)r * ldb + k0 + kc * 8, 16u);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  constexpr uint32_t IDESC = umma_idesc_bf16(kGemmBM, N);
  // prologue: stages 0 .. S-2
#pragma unroll
  for (int s = 0; s < kGemmStages - 1; ++s) {
    if (s < nchunks) load_stage(s, c0 + s);
    else asm volatile("cp.async.commit_group;" ::: "memory");
  }

  for (int i = 0; i < nchunks; ++i) {
    const int stage = i % kGemmStages;
    asm volatile("cp.async.wait_group %0;" ::"n"(kGemmStages - 2) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> tensor-core reads
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t sA = sbase + stage * SM::STAGE_BYTES;
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const uint32_t sB = sA + SM::A_BYTES + t * SM::B_BYTES;
#pragma unroll
        for (int j = 0; j < kGemmBK / 16; ++j) {
          const uint64_t ad = umma_desc(sA + j * 256, 128, (kGemmBK / 8) * 128);
          const uint64_t bd = umma_desc(sB + j * 256, 128, (kGemmBK / 8) * 128);
          umma_bf16(tmem_acc, ad, bd, IDESC, (i | t | j) ? 1u : 0u);
        }
      }
      umma_commit(mbar[stage]);          // arrives when the MMAs reading this stage are done
    }
    // refill the stage used by iteration i-1 with chunk i+S-1 once its MMAs have retired
    const int nxt = i + kGemmStages - 1;
    if (nxt < nchunks) {
      if (i >= 1) {
        const int pstage = (i - 1) % kGemmStages;
        mbar_wait(mbar[pstage], (uint32_t)(((i - 1) / kGemmStages) & 1));
      }
      load_stage(nxt % kGemmStages, c0 + nxt);
    } else {
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  // all MMAs issued; wait for the last commit (covers every earlier MMA: commits retire in order)
  if (tid == 0) umma_commit(mbar[kGemmStages]);
  mbar_wait(mbar[kGemmStages], 0u);
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
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_acc), "r"((uint32_t)N) : "memory");
  }
}

// fp32 [R][C] (row stride lds) -> three bf16 arrays [C][ldd] (transposed), hi + mid + lo = x to 24 bits.
// Destination columns [R, Rpad) are written as zeros (the GEMM's K padding).
__global__ void __launch_bounds__(256)
split3_transpose_kernel(const float* __restrict__ src, long long lds, long long src_qstride, int R, int Rpad,
                        int Ccols, __nv_bfloat16* __restrict__ dst, long long ldd, long long dst_tstride,
                        long long dst_qstride) {
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
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = cbase + ty + 8 * i, r = r0 + tx;
    if (c < Ccols && r < Rpad) {
      const float x = tile[tx][ty + 8 * i];      // zero beyond R
      const __nv_bfloat16 h = __float2bfloat16_rn(x);
      const float r1 = x - __bfloat162float(h);
      const __nv_bfloat16 m = __float2bfloat16_rn(r1);
      const float r2 = r1 - __bfloat162float(m);
      const __nv_bfloat16 l = __float2bfloat16_rn(r2);
      const long long o = (long long)c * ldd + r;
      dst[o] = h;
      dst[dst_tstride + o] = m;
      dst[2 * dst_tstride + o] = l;
    }
  }
}

template <int N>
static int launch_gemm3(const __nv_bfloat16* A, long long lda, long long aq, int M, const __nv_bfloat16* B,
                        long long ldb, long long bt, long long bq, float* C, long long ldc, long long cq,
                        int Kd, int NQ, int splits, cudaStream_t st) {
  using SM = GemmSmem<N>;
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(umma_gemm3_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::TOTAL);
    if (e != cudaSuccess) return (int)e;
    attr = true;
  }
  const int kchunks = Kd / kGemmBK;
  const int mtiles = (M + kGemmBM - 1) / kGemmBM;
  if (splits <= 0) splits = (2 * 148 + mtiles * NQ - 1) / (mtiles * NQ);   // ~two CTAs' worth of work per SM
  if (splits < 1) splits = 1;
  if (splits > kchunks) splits = kchunks;
  const int per = (kchunks + splits - 1) / splits;
  splits = (kchunks + per - 1) / per;
  dim3 grid((M + kGemmBM - 1) / kGemmBM, splits, NQ);
  umma_gemm3_kernel<N><<<grid, kGemmThreads, SM::TOTAL, st>>>(A, lda, aq, M, B, ldb, bt, bq, C, ldc, cq, kchunks, per);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

}  // namespace spmf

using namespace spmf;

extern "C" {

int spmf_umma_gemm3(const void* A, long long lda, long long a_qstride, int M, const void* B3, long long ldb,
                    long long b_tstride, long long b_qstride, float* C, long long ldc, long long c_qstride,
                    int N, int Kd, int NQ, int splits, void* stream) {
  if (!A || !B3 || !C || M <= 0 || Kd <= 0 || NQ <= 0) return SPMF_ERR_BAD_ARG;
  if (Kd % kGemmBK || lda % 8 || ldb % 8 || ldc % 4) return SPMF_ERR_BAD_ARG;   // 16-byte chunks / float4 atomics
  if (((uintptr_t)A | (uintptr_t)B3 | (uintptr_t)C) & 15) return SPMF_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* a = (const __nv_bfloat16*)A;
  const __nv_bfloat16* b = (const __nv_bfloat16*)B3;
  switch (N) {
    case 32: return launch_gemm3<32>(a, lda, a_qstride, M, b, ldb, b_tstride, b_qstride, C, ldc, c_qstride, Kd, NQ, splits, st);
    case 64: return launch_gemm3<64>(a, lda, a_qstride, M, b, ldb, b_tstride, b_qstride, C, ldc, c_qstride, Kd, NQ, splits, st);
    case 128: return launch_gemm3<128>(a, lda, a_qstride, M, b, ldb, b_tstride, b_qstride, C, ldc, c_qstride, Kd, NQ, splits, st);
    default: return SPMF_ERR_UNSUPPORTED;
  }
}

int spmf_split3_transpose(const float* src, long long lds, long long src_qstride, int R, int Rpad, int Ccols,
                          void* dst3, long long ldd, long long dst_tstride, long long dst_qstride, int NQ,
                          void* stream) {
  if (!src || !dst3 || R <= 0 || Rpad < R || Rpad > ldd || Ccols <= 0 || NQ <= 0) return SPMF_ERR_BAD_ARG;
  dim3 grid((Rpad + 31) / 32, (Ccols + 31) / 32, NQ);
  split3_transpose_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, lds, src_qstride, R, Rpad, Ccols,
                                                                  (__nv_bfloat16*)dst3, ldd, dst_tstride, dst_qstride);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? SPMF_OK : (int)e;
}

}  // extern "C"
