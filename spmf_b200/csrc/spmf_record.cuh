// Layout of one gather record (the SV*KP floats of draw group q for one feature d or one row b).
//
// A record is stored k-vector-major so that a slot of LPN = SV*RG lanes reads it with VPL fully
// coalesced vector loads while every lane ends up owning KP/RG latent dims of ONE draw:
//     k = (i*RG + kg)*VW + w,   position = ((i*SV + s)*RG + kg)*VW + w
// (i < VPL vectors per lane, kg < RG lanes sharing draw s, w < VW floats per vector).  The
// k-contraction is then KP/RG in-lane FMAs plus log2(RG) shuffle steps -- two lanes per draw at
// K=32 -- instead of a wide butterfly.  Shared by the CUDA kernels, the host check and (through
// spmf_rec_pos) the Python packing code.
#pragma once
#include "spmf_math.cuh"

namespace spmf {

struct RecMap { int VW, NV, VPL, RG; };

SPMF_HD RecMap rec_map(int KP) {
  RecMap m;
  m.VW = KP < 4 ? KP : 4;
  m.NV = KP / m.VW;
  m.VPL = m.NV < 4 ? m.NV : 4;
  m.RG = m.NV / m.VPL;
  return m;
}

SPMF_HD int rec_pos(int KP, int SV, int sv, int k) {
  const RecMap m = rec_map(KP);
  const int nv = k / m.VW, w = k - nv * m.VW;
  const int i = nv / m.RG, kg = nv - i * m.RG;
  return ((i * SV + sv) * m.RG + kg) * m.VW + w;
}

}  // namespace spmf
