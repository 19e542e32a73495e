// TEST INFRASTRUCTURE: runs the __host__ __device__ bodies of spmf_b200/csrc/spmf_model.cuh on the
// CPU (serial loops standing in for warps / lanes) so the fp32 math of the draw, prior, entropy and
// backward kernels can be checked against the float64 oracle without a GPU.  Built by
// tests/hostcheck/__init__.py with g++; never shipped, never used by the product path.
#include <cstring>
#include <vector>

#include "../../spmf_b200/csrc/spmf_model.cuh"

using namespace spmf;

extern "C" {

void hc_gamma_grad(const float* a, const float* x, float* out, int n) {
  for (int i = 0; i < n; ++i) out[i] = gamma_sample_der_alpha(a[i], x[i]);
}
// 4-draw variant used by gamma_kernel: every group of 4 consecutive x share a[4*g]
void hc_gamma_grad4(const float* a, const float* x, float* out, int n) {
  for (int i = 0; i + 4 <= n; i += 4) {
    float xs[4] = {x[i], x[i + 1], x[i + 2], x[i + 3]}, o[4];
    gamma_sample_der_alpha4(a[i], digammaf_pos(a[i]), xs, 4, o);
    for (int j = 0; j < 4; ++j) out[i + j] = o[j];
  }
}
void hc_digamma(const float* x, float* out, int n) {
  for (int i = 0; i < n; ++i) out[i] = digammaf_pos(x[i]);
}
void hc_softplus(const float* x, float* out, float* sg, int n) {
  for (int i = 0; i < n; ++i) { out[i] = softplusf(x[i]); sg[i] = sigmoidf(x[i]); }
}
void hc_normals(float* out, long long n, unsigned stream, unsigned step, unsigned long long seed) {
  uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
  for (long long i = 0; i * 4 < n; ++i) {
    U4 ctr = {(uint32_t)(i & 0xffffffffu), (uint32_t)(i >> 32), stream, step};
    U4 r = philox4x32_10(ctr, k0, k1);
    float v[4];
    box_muller(r.x, r.y, &v[0], &v[1]);
    box_muller(r.z, r.w, &v[2], &v[3]);
    for (int j = 0; j < 4; ++j)
      if (i * 4 + j < n) out[i * 4 + j] = v[j];
  }
}
void hc_gammas(float* out, long long n, float alpha, unsigned stream, unsigned long long seed) {
  uint32_t k0 = (uint32_t)(seed & 0xffffffffu), k1 = (uint32_t)(seed >> 32);
  for (long long i = 0; i < n; ++i)
    out[i] = gamma_draw(alpha, (uint32_t)(i & 0xffffffffu), (uint32_t)(i >> 32), stream, k0, k1);
}
void hc_philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, unsigned k0, unsigned k1,
               unsigned* out) {
  U4 r = philox4x32_10(U4{c0, c1, c2, c3}, k0, k1);
  out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

void hc_layout(int D, int K, int S, long long* toff, long long* noff) {
  Layout L = make_layout(D, K, S);
  for (int i = 0; i <= NUM_TENSORS; ++i) toff[i] = L.toff[i];
  for (int i = 0; i <= NUM_VARS; ++i) noff[i] = L.noff[i];
}

// operands for every draw: Ap/EV [S][D][K], PH [S][D]  (plain layout, SV = 1, KP = K)
void hc_draw_operands(const float* P, const float* N, const float* eta, int D, int K, int S,
                      float* Ap, float* EV, float* PH) {
  Layout L = make_layout(D, K, S);
  for (int d = 0; d < D; ++d) {
    LaneState<4> st[32];
    FeatState f;
    for (int lane = 0; lane < 32; ++lane) lane_init<4>(st[lane], L, P, d, lane);
    feat_init(f, L, P, d);
    for (int s = 0; s < S; ++s) {
      FeatDraw fd = feat_draw(f, L, N, d, s);
      for (int lane = 0; lane < 32; ++lane)
        for (int i = 0; i < 4; ++i) {
          int k = lane + 32 * i;
          if (k < K) {
            long long idx = ((long long)s * D + d) * K + k;
            lane_operands<4>(st[lane], L, N, eta, d, lane, i, s, fd.a, &Ap[idx], &EV[idx], nullptr, nullptr);
          }
        }
      PH[(long long)s * D + d] = eta[d] * fd.b * fd.w.y;
    }
  }
}

// full backward given the data-term upstream gradients (plain [S][D][K] / [S][D] layouts,
// GEV already includes the closed-form -zcolsum, Gphi_nz does NOT include -B).
void hc_backward_params(const float* P, const float* N, const float* eta, int D, int K, int S,
                        const float* GAp, const float* GEV, const float* Gphinz, float batch_rows,
                        float u_tau_scale, float s_tau_scale, float decay, float w_entropy,
                        float w_prior, int world, float* grads, double* parts /*[S][16]*/) {
  Layout L = make_layout(D, K, S);
  Hyper h;
  h.vw_identity = 0;
  h.u_tau_b = 1.f / (u_tau_scale * u_tau_scale);
  h.s_tau_b = 1.f / (s_tau_scale * s_tau_scale);
  h.decay = decay; h.w_entropy = w_entropy; h.w_prior = w_prior;
  h.rep_scale = 1.f / (float)world; h.batch_rows = batch_rows;
  std::vector<double> dutau((size_t)S * K, 0.0);
  // dg/dalpha of every Gamma draw, as gamma_grad_kernel computes it
  std::vector<float> Gv((size_t)L.noff[NUM_VARS], 0.f);
  for (int v = VAR_UETA; v < NUM_VARS; ++v)
    for (long long i = 0; i < L.vsize[v] * S; ++i) {
      long long e = i % L.vsize[v];
      Gv[L.noff[v] + i] = gamma_sample_der_alpha(softplusf(P[L.toff[2 * v] + e]), N[L.noff[v] + i]);
    }
  const float* G = Gv.data();
  std::memset(parts, 0, sizeof(double) * S * NUM_PARTS);
  const float invS = 1.f / (float)S;
  const float wer = h.w_entropy * h.rep_scale;
  for (int d = 0; d < D; ++d) {
    LaneState<4> st[32];
    FeatState f;
    for (int lane = 0; lane < 32; ++lane) lane_init<4>(st[lane], L, P, d, lane, h.decay);
    feat_init(f, L, P, d);
    for (int s = 0; s < S; ++s) {
      FeatDraw fd = feat_draw(f, L, N, d, s);
      float da = 0.f, pp[5] = {0, 0, 0, 0, 0};
      for (int lane = 0; lane < 32; ++lane)
        for (int i = 0; i < 4; ++i) {
          int k = lane + 32 * i;
          if (k < K) {
            long long idx = ((long long)s * D + d) * K + k;
            DkUp up{GAp[idx], GEV[idx]};
            DkOut o = lane_step<4>(st[lane], L, h, N, G, eta, d, lane, i, s, fd.a, up);
            da += o.da;
            dutau[(size_t)s * K + k] += o.dutau;
            for (int j = 0; j < 5; ++j) pp[j] += o.parts[j];
          }
        }
      float fp[7];
      feat_step(f, fd, L, h, N, G, eta, d, s, da, Gphinz[(long long)s * D + d], fp);
      double* o = parts + (size_t)s * NUM_PARTS;
      o[P_U] += pp[0]; o[P_V] += pp[1]; o[P_UETA] += pp[2]; o[P_UETAA] += pp[3];
      o[P_W] += fp[0]; o[P_S] += fp[1]; o[P_SETA] += fp[2]; o[P_STAU] += fp[3];
      o[P_SETAA] += fp[4]; o[P_STAUA] += fp[5]; o[P_LOGQ] += (double)pp[4] + fp[6];
    }
    for (int lane = 0; lane < 32; ++lane)
      for (int i = 0; i < 4; ++i) {
        int k = lane + 32 * i;
        if (k < K) {
          long long e = (long long)d * K + k;
          nparam_finish(st[lane].u[i], P[L.toff[U_RHO] + e], invS, wer, &grads[L.toff[U_LOC] + e], &grads[L.toff[U_RHO] + e]);
          nparam_finish(st[lane].v[i], P[L.toff[V_RHO] + e], invS, wer, &grads[L.toff[V_LOC] + e], &grads[L.toff[V_RHO] + e]);
          gparam_finish(st[lane].ue[i], P[L.toff[UETA_C] + e], P[L.toff[UETA_B] + e], invS, &grads[L.toff[UETA_C] + e], &grads[L.toff[UETA_B] + e]);
          gparam_finish(st[lane].ua[i], P[L.toff[UETAA_C] + e], P[L.toff[UETAA_B] + e], invS, &grads[L.toff[UETAA_C] + e], &grads[L.toff[UETAA_B] + e]);
        }
      }
    nparam_finish(f.w, P[L.toff[W_RHO] + d], invS, wer, &grads[L.toff[W_LOC] + d], &grads[L.toff[W_RHO] + d]);
    nparam_finish(f.s0, P[L.toff[S_RHO] + d], invS, wer, &grads[L.toff[S_LOC] + d], &grads[L.toff[S_RHO] + d]);
    nparam_finish(f.s1, P[L.toff[S_RHO] + D + d], invS, wer, &grads[L.toff[S_LOC] + D + d], &grads[L.toff[S_RHO] + D + d]);
    gparam_finish(f.se0, P[L.toff[SETA_C] + d], P[L.toff[SETA_B] + d], invS, &grads[L.toff[SETA_C] + d], &grads[L.toff[SETA_B] + d]);
    gparam_finish(f.se1, P[L.toff[SETA_C] + D + d], P[L.toff[SETA_B] + D + d], invS, &grads[L.toff[SETA_C] + D + d], &grads[L.toff[SETA_B] + D + d]);
    gparam_finish(f.st, P[L.toff[STAU_C] + d], P[L.toff[STAU_B] + d], invS, &grads[L.toff[STAU_C] + d], &grads[L.toff[STAU_B] + d]);
    gparam_finish(f.sea0, P[L.toff[SETAA_C] + d], P[L.toff[SETAA_B] + d], invS, &grads[L.toff[SETAA_C] + d], &grads[L.toff[SETAA_B] + d]);
    gparam_finish(f.sea1, P[L.toff[SETAA_C] + D + d], P[L.toff[SETAA_B] + D + d], invS, &grads[L.toff[SETAA_C] + D + d], &grads[L.toff[SETAA_B] + D + d]);
    gparam_finish(f.sta, P[L.toff[STAUA_C] + d], P[L.toff[STAUA_B] + d], invS, &grads[L.toff[STAUA_C] + d], &grads[L.toff[STAUA_B] + d]);
  }
  for (int k = 0; k < K; ++k) {
    LatState t;
    lat_init(t, L, P, k);
    for (int s = 0; s < S; ++s) {
      float pp[3];
      lat_step(t, L, h, N, G, k, s, (float)dutau[(size_t)s * K + k], pp);
      double* o = parts + (size_t)s * NUM_PARTS;
      o[P_UTAU] += pp[0]; o[P_UTAUA] += pp[1]; o[P_LOGQ] += pp[2];
    }
    gparam_finish(t.ut, P[L.toff[UTAU_C] + k], P[L.toff[UTAU_B] + k], invS, &grads[L.toff[UTAU_C] + k], &grads[L.toff[UTAU_B] + k]);
    gparam_finish(t.uta, P[L.toff[UTAUA_C] + k], P[L.toff[UTAUA_B] + k], invS, &grads[L.toff[UTAUA_C] + k], &grads[L.toff[UTAUA_B] + k]);
  }
}

}  // extern "C"
