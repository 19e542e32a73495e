// Device-side state of the exact non-finite guard (poisson.py:606-616) -- shared by the kernels that
// raise it (sparse / tensor-core row passes), the dense re-evaluation (spmf_dense.cu) and the backward /
// loss-part kernels that must then drop the closed-form sum(rate) terms (spmf_params.cu).
#pragma once
#include <stdint.h>
#include <string.h>

namespace spmf {

// `flag` bit 0 = some kernel of this step met a non-finite log-likelihood (triggers the conditional
// launches), bit 1 = the step runs a dense-only link; any bit set = the data term is the dense one (no
// closed-form sum(rate) terms downstream).  `nbad` = number of non-finite entries over all draws
// (statistics pass).  `minkey` = (order-preserving bits of the smallest finite log-likelihood << 32) |
// low 32 bits of that entry's linear index (s*B + b)*D + d.
struct GuardState {
  int flag;
  int nbad;
  unsigned long long minkey;
  unsigned bar;        // arrival counter of the slow path's grid barrier (0 between launches)
  unsigned done;       // blocks that have left the slow-path kernel
};

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned ordered_bits(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
#endif
#if defined(__CUDACC__)
__host__ __device__
#endif
inline float ordered_float(unsigned o) {
  const unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// the replacement value of poisson.py:609: min(finite portion, with zeros for the others) - 10
#if defined(__CUDACC__)
__host__ __device__
#endif
inline float guard_min_val(unsigned long long minkey) {
  const float m = minkey == ~0ull ? 0.f : ordered_float((unsigned)(minkey >> 32));
  return (m < 0.f ? m : 0.f) - 10.f;
}

}  // namespace spmf
