// Training side of the LEAF frontend: the training forward and the backward (parameter gradients of
// sum(out * grad_out), optionally the gradient w.r.t. the waveform).
//
// Replaces autograd through reference frontend.py:78-89 (what loss.backward() does in train.py:258;
// the reference graph has ~7*N_frames nodes because of the Python EMA loop).  Formulas: SURVEY A.2.
//
// Tensor-core path (geometries the training kernel covers, k1_tc_train_filters_per_group() > 0):
//   leafk_forward_train   T0 k0_banks_train_kernel   banks h, tau*h, (tau^2/sigma^3 - 1/sigma)*h      (k0_banks.cu)
//                         T1 k1_tc_kernel<.,.,1,.>   y, z, v correlations; pools e and the three bilinear forms
//                                                    Q_mu, Q_sigma, Q_poolw with the pooling windows     (k1_tc_kernel.cuh)
//                         T2 k2_pcen_kernel          p -> floor -> PCEN -> out; assembles p and the Q's  (k2_pcen.cu)
//   leafk_backward_saved  B1 bwd_pcen_kernel         per (clip, filter) row: PCEN + smoother backward (forward scan to
//                                                    rebuild M, reverse affine scan for the smoother adjoint), floor
//                                                    mask -> dp; row sums for alpha, delta, root, ema_w, bias and the
//                                                    contractions sum_n dp[n] Q[n]
//                         B3 bwd_finish_kernel       fixed-order sums over clips, clamp masks, chain factors -> 7 grads
//   The backward runs NO correlation: sum_t de[t] q[t] with de[t] = sum_n dp[n] g[t - t_n] equals sum_n dp[n] Q[n]
//   where Q is q pooled with the forward's own windows, so the forward pools the q's next to the energy.
// Generic path (any geometry; also the waveform gradient of every geometry):
//                         G1 bwd_generic_kernel      FP32 CUDA-core correlations y, z, v per (clip, tile, filter), the
//                                                    per-filter sums, and dx partials per tile
//                         G2 bwd_dx_sum_kernel       adds the overlapping dx partials of neighbouring tiles in tile order
// Everything is deterministic (no floating-point atomics).
#include "../../include/leafk.h"
#include "leafk_common.cuh"
#include "k1_tc_layout.cuh"
#include "k2_pcen_args.cuh"

#include <cstring>

namespace leafk {

int fail(int code, const char* fmt, ...);
void count_launch(int n);
void prof_mark(int which, cudaStream_t stream);      // leafk_profile_begin/end: events around K0 / K1 / K2
void launch_k0(const float* kernel, const float* pool_w, int F, int K, int Kp, int C2p, float* cprm,
               float* w32, float* g32, uint8_t* w16, int tc_cg, int tc_groups, int* tc_perm, int* tc_zones,
               float prune_c, float prune_c3, int* done, int n_done, cudaStream_t stream);
void launch_k0_train(const float* kernel, const float* pool_w, int F, int K, int Kp, int FB, int n_groups,
                     float* tprm, uint8_t* w16t, int* done, int n_done, cudaStream_t stream);
void geom_apply_prep(const leafk_config* cfg, Geom* g);
void bank_bounds(int K, float* mu_hi, float* sigma_lo, float* sigma_hi, float* pool_lo);
int k1_tc_train_filters_per_group(int K, int H);
cudaError_t launch_k1_tc_train(const Geom& g, const float* x, const uint8_t* w16t, int FB, int n_groups,
                               const float* tprm, float* ppart, int* done, cudaStream_t stream);
cudaError_t launch_k2(const Geom& g, const float* ppart, const PcenArgs& a, cudaStream_t stream);

constexpr int B1_FPB = 8;      // filters per block (= warps)
constexpr int B1_SEG = 128;    // frames per scan step

struct PcenBwdArgs {
  const float* p;        // (B,F,N) floored pooled energies saved by the forward
  const float* gout;     // (B,F,N)
  const float* q;        // (3,B,F,N) pooled bilinear forms saved by the training forward, or null
  const float* alpha; const float* delta; const float* root; const float* ema_w;
  float* scratch;        // (B,F,N,3): t1 = G*D^-alpha, dM, p_n - M_{n-1}
  float* dp;             // (B,F,N) gradient w.r.t. the pooled energies (after the floor mask), or null
  float* rpart;          // (B,F,8): d_delta, d_alpha_hat, d_root_hat, d_w_hat, d_bias, S_mu, S_sigma, S_poolw
  float pcen_floor, clamp_min;
  int compression;
};

__global__ void __launch_bounds__(B1_FPB * 32)
bwd_pcen_kernel(int B, int F, int N, const PcenBwdArgs a) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fgroups = (F + B1_FPB - 1) / B1_FPB;
  const int b = blockIdx.x / fgroups;
  const int f = (blockIdx.x % fgroups) * B1_FPB + warp;
  if (f >= F) return;                                   // warps are independent (no block barrier below)
  const size_t row = ((size_t)b * F + f) * N;
  const size_t qstride = (size_t)B * F * N;
  const int nseg = (N + B1_SEG - 1) / B1_SEG;

  float s_delta = 0.f, s_alpha = 0.f, s_root = 0.f, s_w = 0.f, s_bias = 0.f, s_mu = 0.f, s_sg = 0.f, s_pw = 0.f;
  float w = 0.f, om = 1.f;
  // rows of <= 128 frames (1 s clips) keep the adjoint seeds in registers between the two sweeps; longer rows park
  // them in the scratch buffer
  const bool in_regs = nseg == 1;
  float k_t1[4] = {0.f, 0.f, 0.f, 0.f}, k_dM[4] = {0.f, 0.f, 0.f, 0.f}, k_pm[4] = {0.f, 0.f, 0.f, 0.f};

  if (a.compression) {
    w = clamp_nan(a.ema_w[f], 0.f, 1.f);
    const float alpha = min_nan(a.alpha[f], 1.0f);
    const float q = 1.0f / max_nan(a.root[f], 1.0f);
    const float delta = a.delta[f];
    const float dq = powf(delta, q);
    const float dq1 = powf(delta, q - 1.0f);
    const float ldelta = logf(delta);
    om = 1.0f - w;
    // ---------------- forward sweep: rebuild M, per-element adjoint seeds -----------------------
    float carry = a.p[row];                               // M_{-1} = p_0   (postprocessing.py:15)
    for (int sg = 0; sg < nseg; ++sg) {
      const int seg0 = sg * B1_SEG, seg_n = min(B1_SEG, N - seg0);
      float p[4], go[4];
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nl = lane * 4 + j;
        const bool ok = nl < seg_n;
        p[j] = ok ? a.p[row + seg0 + nl] : 0.f;
        go[j] = ok ? a.gout[row + seg0 + nl] : 0.f;
        cnt += ok;
      }
      float A = 1.f, C = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < cnt) { C = fmaf(om, C, w * p[j]); A *= om; }
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float Ap = __shfl_up_sync(0xffffffffu, A, o), Cp = __shfl_up_sync(0xffffffffu, C, o);
        if (lane >= o) { C = fmaf(A, Cp, C); A *= Ap; }
      }
      float Aex = __shfl_up_sync(0xffffffffu, A, 1), Cex = __shfl_up_sync(0xffffffffu, C, 1);
      if (lane == 0) { Aex = 1.f; Cex = 0.f; }
      float state = fmaf(Aex, carry, Cex);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nl = lane * 4 + j;
        const float prev = state;
        state = fmaf(om, state, w * p[j]);
        if (nl < seg_n) {
          // one log2 + one exp2 per power (D, u > 0 wherever the forward is finite); ln = log2 * ln 2
          const float D = a.pcen_floor + state;
          const float l2D = log2f(D);
          const float Dma = exp2f(-alpha * l2D);
          const float u = p[j] * Dma + delta;
          const float l2u = log2f(u);
          const float uq1 = exp2f((q - 1.0f) * l2u);
          const float G = go[j] * q * uq1;
          s_delta += G - go[j] * q * dq1;
          s_alpha -= G * p[j] * Dma * (l2D * 0.69314718056f);
          s_root -= go[j] * (u * uq1 * (l2u * 0.69314718056f) - dq * ldelta) * (q * q);
          const float v_t1 = G * Dma, v_dM = -G * alpha * p[j] * Dma / D, v_pm = p[j] - prev;
          if (in_regs) {
            k_t1[j] = v_t1; k_dM[j] = v_dM; k_pm[j] = v_pm;
          } else {
            float* sc = a.scratch + (row + seg0 + nl) * 3;
            sc[0] = v_t1; sc[1] = v_dM; sc[2] = v_pm;
          }
        }
      }
      const float Al = __shfl_sync(0xffffffffu, A, 31), Cl = __shfl_sync(0xffffffffu, C, 31);
      carry = fmaf(Al, carry, Cl);
    }
  }
  // ---------------- reverse sweep: lambda_n = dM_n + (1-w) lambda_{n+1};  dp ------------------------
  float lam_carry = 0.f;
  for (int sg = nseg - 1; sg >= 0; --sg) {
    const int seg0 = sg * B1_SEG, seg_n = min(B1_SEG, N - seg0);
    float dpv[4];
    if (a.compression) {
      float t1[4], dM[4], pm[4];
      int cnt = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nl = lane * 4 + j;
        const bool ok = nl < seg_n;
        if (in_regs) {
          t1[j] = k_t1[j]; dM[j] = k_dM[j]; pm[j] = k_pm[j];      // zero beyond the row's last frame
        } else {
          const float* sc = a.scratch + (row + seg0 + (ok ? nl : 0)) * 3;
          t1[j] = ok ? sc[0] : 0.f; dM[j] = ok ? sc[1] : 0.f; pm[j] = ok ? sc[2] : 0.f;
        }
        cnt += ok;
      }
      // composite of this lane's frames, applied from the last frame backwards:  L -> om*L + dM_j
      float A = 1.f, C = 0.f;
#pragma unroll
      for (int j = 3; j >= 0; --j)
        if (j < cnt) { C = fmaf(om, C, dM[j]); A *= om; }
      // inclusive scan from the high lanes down
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float An = __shfl_down_sync(0xffffffffu, A, o), Cn = __shfl_down_sync(0xffffffffu, C, o);
        if (lane + o < 32) { C = fmaf(A, Cn, C); A *= An; }
      }
      float Aex = __shfl_down_sync(0xffffffffu, A, 1), Cex = __shfl_down_sync(0xffffffffu, C, 1);
      if (lane == 31) { Aex = 1.f; Cex = 0.f; }
      float lam = fmaf(Aex, lam_carry, Cex);          // lambda of the frame just after this lane's last one
#pragma unroll
      for (int j = 3; j >= 0; --j) {
        const int nl = lane * 4 + j;
        if (j < cnt) {
          lam = fmaf(om, lam, dM[j]);
          s_w = fmaf(lam, pm[j], s_w);
          // frame 0 also seeds the state: dM_0/dp_0 = w + (1-w)
          const float coef = (seg0 + nl == 0) ? (w + om) : w;
          dpv[j] = fmaf(coef, lam, t1[j]);
        } else {
          dpv[j] = 0.f;
        }
      }
      const float A0 = __shfl_sync(0xffffffffu, A, 0), C0 = __shfl_sync(0xffffffffu, C, 0);
      lam_carry = fmaf(A0, lam_carry, C0);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nl = lane * 4 + j;
        dpv[j] = (nl < seg_n) ? a.gout[row + seg0 + nl] : 0.f;
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nl = lane * 4 + j;
      if (nl < seg_n) {
        const size_t at = row + seg0 + nl;
        const float pv = a.p[at];
        const float d = (pv > a.clamp_min) ? dpv[j] : ((pv != pv) ? pv : 0.f);   // torch.maximum(., 1e-5)  frontend.py:84
        s_bias += d;
        if (a.dp != nullptr) a.dp[at] = d;
        if (a.q != nullptr) {
          s_mu = fmaf(d, a.q[at], s_mu);
          s_sg = fmaf(d, a.q[qstride + at], s_sg);
          s_pw = fmaf(d, a.q[2 * qstride + at], s_pw);
        }
      }
    }
  }
  // ---------------- per-row sums ------------------------------------------------------------------
  float v[8] = {s_delta, s_alpha, s_root, s_w, s_bias, s_mu, s_sg, s_pw};
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  if (lane == 0) {
    float* r = a.rpart + ((size_t)b * F + f) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i];
  }
}

// ------------------------------------------------------------------------------------------------
// G1: generic FP32 backward.  One block = 128 consecutive e-samples of one clip at a time (persistent over
// (clip, tile) units), one thread per sample.  Per filter the thread correlates its sample window with the bank
// (three running sums per component: sum x w, sum x tau w, sum x tau^2 w give y, z and v), forms the pooling adjoint
// de[t] = sum_n dp[n] g[k_n], accumulates the three per-filter sums and, when the waveform gradient is wanted, spreads
// dy = 2 y de back over the window: dx[j] += sum_t dy[t] W[j - t + padL]  (last bullet).
constexpr int G_TL = 128;

struct GenericArgs {
  const float* x;        // waveform window (fp32 or int16, Geom::x_fmt)
  const float* w32;      // (Kp, C2p) fp32 bank, tap-major (k0_banks_kernel)
  const float* g32;      // (K, F) pooling windows, tap-major
  const float* cprm;     // (F, 8) constrained parameters (CP_SIGMA)
  const float* dp;       // (B, F, N)
  float* gpart;          // (gridDim.x, F, 4) per-block sums {S_mu, S_sigma, S_poolw, 0}, or null
  float* xpart;          // (B, n_tiles, G_TL + K - 1) dx partials per tile, or null
};

template <bool PARAMS, bool DX>
__global__ void __launch_bounds__(G_TL)
bwd_generic_kernel(const Geom g, const GenericArgs a) {
  extern __shared__ __align__(16) float gs[];
  const int WL = G_TL + g.K - 1;                      // samples a tile's outputs depend on / dx entries it touches
  float* xs = gs;                                     // [WL] x~[ts - padL + i]
  float* wre = xs + WL;                               // [K]
  float* wim = wre + g.K;                             // [K]
  float* dyr = wim + g.K;                             // [G_TL]
  float* dyi = dyr + G_TL;                            // [G_TL]
  float* dxs = dyi + G_TL;                            // [WL]   (DX)
  float* sacc = dxs + (DX ? WL : 0);                  // [F*3]  (PARAMS)
  float* red = sacc + (PARAMS ? 3 * g.F : 0);         // [4*3]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float centre = 0.5f * (float)(g.K - 1);
  if (PARAMS)
    for (int i = tid; i < 3 * g.F; i += G_TL) sacc[i] = 0.f;
  const long long n_units = (long long)g.B * g.n_tiles;
  for (long long u = blockIdx.x; u < n_units; u += gridDim.x) {
    const int b = (int)(u / g.n_tiles), tile = (int)(u % g.n_tiles);
    const long long ts = (long long)tile * G_TL;
    const long long t = ts + tid;
    const bool tok = t < g.T_total;
    __syncthreads();                                  // previous unit done with xs / dxs
    const ClipView cv = clip_view(g, b);
    for (int i = tid; i < WL; i += G_TL) {
      const long long s = ts - g.padL + i;
      xs[i] = (s >= 0 && s < g.T_total) ? clip_sample(g, a.x, cv, s) : 0.f;
      if (DX) dxs[i] = 0.f;
    }
    // frames whose pooling window holds sample t: k = t + padL - n H in [0, K)
    int n_lo = (int)ceildiv_ll(t + g.padL - g.K + 1, g.H), n_hi = (int)floordiv_ll(t + g.padL, g.H);
    if (n_lo < 0) n_lo = 0;
    if (n_hi > g.N_total - 1) n_hi = g.N_total - 1;
    for (int f = 0; f < g.F; ++f) {
      __syncthreads();                                // xs ready; previous filter done with wre / wim / dy
      for (int k = tid; k < g.K; k += G_TL) {
        wre[k] = a.w32[(size_t)k * g.C2p + 2 * f];
        wim[k] = a.w32[(size_t)k * g.C2p + 2 * f + 1];
      }
      __syncthreads();
      float y0r = 0.f, y0i = 0.f, y1r = 0.f, y1i = 0.f, y2r = 0.f, y2i = 0.f;
      const float* xp = xs + tid;
      const float tau0 = (float)(-(g.K / 2));
      if (PARAMS) {
        for (int k = 0; k < g.K; ++k) {
          const float tau = tau0 + (float)k;
          const float x0 = xp[k], x1 = x0 * tau, x2 = x1 * tau;
          const float wr = wre[k], wi = wim[k];
          y0r = fmaf(x0, wr, y0r); y0i = fmaf(x0, wi, y0i);
          y1r = fmaf(x1, wr, y1r); y1i = fmaf(x1, wi, y1i);
          y2r = fmaf(x2, wr, y2r); y2i = fmaf(x2, wi, y2i);
        }
      } else {
        for (int k = 0; k < g.K; ++k) {
          const float x0 = xp[k];
          y0r = fmaf(x0, wre[k], y0r); y0i = fmaf(x0, wim[k], y0i);
        }
      }
      // pooling adjoint of this sample
      float de = 0.f, dg2 = 0.f;
      if (tok) {
        const float* dprow = a.dp + ((size_t)b * g.F + f) * g.N_total;
        for (int n = n_lo; n <= n_hi; ++n) {
          const int k = (int)(t + g.padL - (long long)n * g.H);
          const float wgt = __ldg(a.g32 + (size_t)k * g.F + f) * dprow[n];
          const float kc = (float)k - centre;
          de += wgt;
          dg2 = fmaf(wgt, kc * kc, dg2);
        }
      }
      if (PARAMS) {
        const float sg = __ldg(a.cprm + (size_t)f * 8 + CP_SIGMA);
        const float inv_s = 1.0f / sg, inv_s3 = inv_s * inv_s * inv_s;
        const float vr = fmaf(inv_s3, y2r, -inv_s * y0r), vi = fmaf(inv_s3, y2i, -inv_s * y0i);
        float v0 = tok ? de * (y0i * y1r - y0r * y1i) : 0.f;
        float v1 = tok ? de * (y0r * vr + y0i * vi) : 0.f;
        float v2 = tok ? dg2 * fmaf(y0r, y0r, y0i * y0i) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          v0 += __shfl_xor_sync(0xffffffffu, v0, o);
          v1 += __shfl_xor_sync(0xffffffffu, v1, o);
          v2 += __shfl_xor_sync(0xffffffffu, v2, o);
        }
        if (lane == 0) { red[warp * 3] = v0; red[warp * 3 + 1] = v1; red[warp * 3 + 2] = v2; }
      }
      if (DX) {
        dyr[tid] = tok ? 2.0f * de * y0r : 0.f;
        dyi[tid] = tok ? 2.0f * de * y0i : 0.f;
      }
      __syncthreads();
      if (PARAMS && tid < 3)
        sacc[3 * f + tid] += (red[tid] + red[3 + tid]) + (red[6 + tid] + red[9 + tid]);
      if (DX) {
        // dx entry i of the tile (sample ts - padL + i) receives dy[t'] W[i - t'] for t' in the tile, 0 <= i - t' < K
        for (int i = tid; i < WL; i += G_TL) {
          const int t0 = i - (g.K - 1) > 0 ? i - (g.K - 1) : 0, t1 = i < G_TL - 1 ? i : G_TL - 1;
          float s = 0.f;
          for (int tt = t0; tt <= t1; ++tt) s = fmaf(dyr[tt], wre[i - tt], fmaf(dyi[tt], wim[i - tt], s));
          dxs[i] += s;
        }
      }
    }
    if (DX) {
      __syncthreads();
      float* dst = a.xpart + (size_t)u * WL;
      for (int i = tid; i < WL; i += G_TL) dst[i] = dxs[i];
    }
  }
  if (PARAMS) {
    __syncthreads();
    for (int i = tid; i < g.F; i += G_TL) {
      float* dst = a.gpart + ((size_t)blockIdx.x * g.F + i) * 4;
      dst[0] = sacc[3 * i]; dst[1] = sacc[3 * i + 1]; dst[2] = sacc[3 * i + 2]; dst[3] = 0.f;
    }
  }
}

// G2: dx[b][j] = sum over the tiles whose window holds j of xpart[b][tile][j - (tile*G_TL - padL)], in tile order
__global__ void bwd_dx_sum_kernel(const Geom g, const float* __restrict__ xpart, float* __restrict__ dx) {
  const int WL = G_TL + g.K - 1;
  const long long total = (long long)g.B * g.T_total;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(idx / g.T_total);
    const long long j = idx % g.T_total;
    // tiles with 0 <= j - tile*G_TL + padL < WL
    long long lo = ceildiv_ll(j + g.padL - WL + 1, G_TL), hi = floordiv_ll(j + g.padL, G_TL);
    if (lo < 0) lo = 0;
    if (hi > g.n_tiles - 1) hi = g.n_tiles - 1;
    float s = 0.f;
    for (long long tl = lo; tl <= hi; ++tl)
      s += xpart[((size_t)b * g.n_tiles + tl) * WL + (size_t)(j - tl * G_TL + g.padL)];
    dx[idx] = s;
  }
}

// ------------------------------------------------------------------------------------------------
struct FinishArgs {
  const float* rpart;    // (B,F,8)
  const float* gpart;    // (n_gpart,F,4) sums of the generic kernel, or null (then S_* come from rpart[5..7])
  int B, F, n_gpart, K, compression;
  float mu_hi, sigma_lo, sigma_hi, pool_lo;
  leafk_params prm;
  leafk_grads g;
};

// one block per filter: fixed-order (thread-strided, then tree) sums over clips / partial rows
__global__ void __launch_bounds__(128) bwd_finish_kernel(const FinishArgs a) {
  __shared__ float sh[8][128];
  const int f = blockIdx.x, tid = threadIdx.x;
  float r[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int b = tid; b < a.B; b += 128) {
    const float4* p = reinterpret_cast<const float4*>(a.rpart + ((size_t)b * a.F + f) * 8);
    const float4 u = p[0], v = p[1];
    r[0] += u.x; r[1] += u.y; r[2] += u.z; r[3] += u.w; r[4] += v.x; r[5] += v.y; r[6] += v.z; r[7] += v.w;
  }
  if (a.gpart != nullptr) {
    r[5] = r[6] = r[7] = 0.f;
    for (int c = tid; c < a.n_gpart; c += 128) {
      const float4 u = *reinterpret_cast<const float4*>(a.gpart + ((size_t)c * a.F + f) * 4);
      r[5] += u.x; r[6] += u.y; r[7] += u.z;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[i][tid] = r[i];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o)
#pragma unroll
      for (int i = 0; i < 8; ++i) sh[i][tid] += sh[i][tid + o];
    __syncthreads();
  }
  if (tid != 0) return;
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = sh[i][0];
  // Gabor parameters: d/dmu = 2*S_mu, d/dsigma = 2*S_sigma, gated by the clamps (convolution.py:20-21); a NaN
  // parameter yields a NaN gradient (the comparisons below would silently zero it)
  const float th0 = a.prm.kernel[2 * f], th1 = a.prm.kernel[2 * f + 1];
  a.g.kernel[2 * f] = (th0 != th0) ? th0 : ((th0 >= 0.f && th0 <= a.mu_hi) ? 2.0f * r[5] : 0.f);
  a.g.kernel[2 * f + 1] = (th1 != th1) ? th1 : ((th1 >= a.sigma_lo && th1 <= a.sigma_hi) ? 2.0f * r[6] : 0.f);
  // pooling width: dg/ds = g * (k-c)^2 / (s^3 c^2)   (impulse_responses.py:75-80)
  const float sraw = a.prm.pool_w[f];
  const float ps = clamp_nan(sraw, a.pool_lo, 0.5f);
  const float c = 0.5f * (float)(a.K - 1);
  a.g.pool_w[f] = (sraw != sraw) ? sraw : ((sraw >= a.pool_lo && sraw <= 0.5f) ? r[7] / (ps * ps * ps * c * c) : 0.f);
  if (a.g.pool_b) a.g.pool_b[f] = r[4];
  if (a.compression) {
    const float al = a.prm.alpha[f], ro = a.prm.root[f], w = a.prm.ema_w[f];
    a.g.delta[f] = r[0];
    a.g.alpha[f] = (al != al) ? al : ((al < 1.0f) ? r[1] : (al == 1.0f ? 0.5f * r[1] : 0.f));      // torch.min tie -> 1/2
    a.g.root[f] = (ro != ro) ? ro : ((ro > 1.0f) ? r[2] : (ro == 1.0f ? 0.5f * r[2] : 0.f));       // torch.max tie -> 1/2
    a.g.ema_w[f] = (w != w) ? w : ((w >= 0.f && w <= 1.0f) ? r[3] : 0.f);                          // clamp passes at the bounds
  }
}

// ------------------------------------------------------------------------------------------------
// Host side
struct TrainPlan {
  int FB, n_groups, Fpad, Kp, N, n_tiles, SL;
  size_t off_tprm, off_w16t, off_done, off_ppart, total;
};

static bool train_plan(const leafk_config* cfg, int B, int T, TrainPlan* pl) {
  const int K = cfg->K, H = cfg->H, F = cfg->F;
  pl->Kp = (K + 15) / 16 * 16;
  pl->FB = ((cfg->algo & 15) == LEAFK_ALGO_FP32) ? 0 : k1_tc_train_filters_per_group(K, H);
  if (pl->FB == 0) return false;
  pl->n_groups = (F + pl->FB - 1) / pl->FB;
  pl->Fpad = pl->n_groups * pl->FB;
  const int padL = K / 2 + (K % 2) - 1, padR = K / 2;
  pl->N = (T + padL + padR - K) / H + 1;
  pl->n_tiles = (T + tc::TILE - 1) / tc::TILE;
  pl->SL = (tc::TILE + K - 2) / H + 1;
  size_t off = 0;
  pl->off_done = off;    off += align256(sizeof(int) * ((size_t)B + 16));     // error word (int 0), then the counters
  pl->off_tprm = off;    off += align256(sizeof(float) * 8 * pl->Fpad);
  pl->off_w16t = off;    off += align256(tc::t_group_bytes(6 * pl->FB, pl->Kp) * (size_t)pl->n_groups);
  pl->off_ppart = off;   off += align256(sizeof(float) * (size_t)B * pl->n_tiles * 4 * F * pl->SL);
  pl->total = off;
  return true;
}

static void whole_clip_geom(const leafk_config* cfg, int B, int T, int tile_len, Geom* out) {
  Geom g;
  memset(&g, 0, sizeof(g));
  const int K = cfg->K, H = cfg->H;
  g.B = B; g.F = cfg->F; g.K = K; g.H = H;
  g.padL = K / 2 + (K % 2) - 1; g.padR = K / 2;
  g.C2 = 2 * cfg->F; g.C2p = (g.C2 + 7) / 8 * 8; g.Kp = (K + 15) / 16 * 16;
  g.T_total = T; g.t_off = 0; g.T_win = T; g.ldx = T;
  g.N_total = (T + g.padL + g.padR - K) / H + 1; g.n_begin = 0; g.n_count = g.N_total;
  g.te_lo = 0; g.te_hi = T; g.TL = tile_len; g.n_tiles = (T + tile_len - 1) / tile_len;
  g.SL = (tile_len + K - 2) / H + 1;
  g.x_fmt = cfg->input_format == LEAFK_INPUT_S16 ? 1 : 0;
  geom_apply_prep(cfg, &g);
  *out = g;
}

static int check_common(const leafk_config* cfg, const leafk_params* prm, int B, int T) {
  if (!cfg || !prm) return fail(LEAFK_EINVAL, "null pointer argument");
  if (cfg->F < 1 || cfg->K < 2 || cfg->H < 1 || B < 1 || T < 1) return fail(LEAFK_EINVAL, "bad shape");
  if (T > (1 << 30)) return fail(LEAFK_EINVAL, "clip too long (%d samples)", T);
  if (!prm->kernel || !prm->pool_w) return fail(LEAFK_EINVAL, "null Gabor / pooling parameter");
  if (cfg->compression && (!prm->alpha || !prm->delta || !prm->root || !prm->ema_w))
    return fail(LEAFK_EINVAL, "compression=1 needs alpha, delta, root, ema_w");
  return LEAFK_OK;
}

int train_supported(int F, int K, int H) {
  if (F < 1 || K < 2 || H < 1) return 0;
  return k1_tc_train_filters_per_group(K, H) > 0 ? 1 : 0;
}

size_t train_workspace_bytes(const leafk_config* cfg, int B, int T) {
  if (!cfg || cfg->F < 1 || cfg->K < 2 || cfg->H < 1 || B < 1 || T < 1) return 0;
  TrainPlan pl;
  if (!train_plan(cfg, B, T, &pl)) return 0;
  return pl.total;
}

int forward_train_run(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T, float* out,
                      float* saved, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  int rc = check_common(cfg, prm, B, T);
  if (rc) return rc;
  if (!x || !out || !saved || !workspace) return fail(LEAFK_EINVAL, "null pointer argument");
  TrainPlan pl;
  if (!train_plan(cfg, B, T, &pl))
    return fail(LEAFK_EINVAL, "training forward: geometry (K=%d,H=%d) not covered by the tensor-core kernel "
                              "(use leafk_forward + leafk_backward)", cfg->K, cfg->H);
  if (pl.total > workspace_bytes)
    return fail(LEAFK_EWORKSPACE, "training workspace %zu bytes < %zu needed", workspace_bytes, pl.total);
  uint8_t* base = (uint8_t*)workspace;
  float* tprm = (float*)(base + pl.off_tprm);
  uint8_t* w16t = base + pl.off_w16t;
  int* err_word = (int*)(base + pl.off_done);
  int* done = err_word + 16;
  float* ppart = (float*)(base + pl.off_ppart);
  Geom g;
  whole_clip_geom(cfg, B, T, tc::TILE, &g);
  const size_t bfn = (size_t)B * cfg->F * pl.N;

  prof_mark(0, stream);
  launch_k0_train(prm->kernel, prm->pool_w, cfg->F, cfg->K, pl.Kp, pl.FB, pl.n_groups, tprm, w16t, err_word, B + 16, stream);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "k0_train launch: %s", cudaGetErrorString(err));
  prof_mark(1, stream);
  err = launch_k1_tc_train(g, x, w16t, pl.FB, pl.n_groups, tprm, ppart, done, stream);
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "k1_tc_train launch: %s", cudaGetErrorString(err));
  prof_mark(2, stream);
  PcenArgs a;
  memset(&a, 0, sizeof(a));
  a.pool_b = prm->pool_b; a.alpha = prm->alpha; a.delta = prm->delta; a.root = prm->root; a.ema_w = prm->ema_w;
  a.out = out; a.saved_p = saved; a.ldo_b = (long long)cfg->F * pl.N; a.ldo_f = pl.N;
  a.pcen_floor = cfg->pcen_floor; a.clamp_min = cfg->clamp_min; a.compression = cfg->compression;
  a.done = done; a.done_target = g.n_tiles * pl.n_groups * 8; a.err = err_word;
  a.q_out = saved + bfn;
  err = launch_k2(g, ppart, a, stream);
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "k2 launch: %s", cudaGetErrorString(err));
  prof_mark(3, stream);
  count_launch(4);
  return LEAFK_OK;
}

// ---- backward from saved tensors (+ generic pieces) -----------------------------------------------------------
struct BwdPlan {
  int N, n_tiles_g, n_blocks_g;
  size_t off_scratch, off_rpart, off_dp, off_cprm, off_w32, off_g32, off_gpart, off_xpart, total;
};

static int generic_blocks(int B, int n_tiles) {
  long long units = (long long)B * n_tiles;
  return (int)(units < 1184 ? units : 1184);            // 8 blocks of 128 threads per SM on a 148-SM part
}

// need_generic: the per-filter sums come from the FP32 kernel (no saved Q's); want_dx: waveform gradient
static void bwd_plan(const leafk_config* cfg, int B, int T, bool need_generic, bool want_dx, BwdPlan* pl) {
  const int K = cfg->K, H = cfg->H, F = cfg->F;
  const int padL = K / 2 + (K % 2) - 1, padR = K / 2;
  const int Kp = (K + 15) / 16 * 16, C2p = (2 * F + 7) / 8 * 8;
  pl->N = (T + padL + padR - K) / H + 1;
  pl->n_tiles_g = (T + G_TL - 1) / G_TL;
  pl->n_blocks_g = generic_blocks(B, pl->n_tiles_g);
  const size_t bfn = (size_t)B * F * pl->N;
  size_t off = 0;
  pl->off_scratch = off; off += align256(sizeof(float) * bfn * 3);
  pl->off_rpart = off;   off += align256(sizeof(float) * (size_t)B * F * 8);
  pl->off_dp = off;      off += (need_generic || want_dx) ? align256(sizeof(float) * bfn) : 0;
  pl->off_cprm = off;    off += (need_generic || want_dx) ? align256(sizeof(float) * 8 * F) : 0;
  pl->off_w32 = off;     off += (need_generic || want_dx) ? align256(sizeof(float) * (size_t)Kp * C2p) : 0;
  pl->off_g32 = off;     off += (need_generic || want_dx) ? align256(sizeof(float) * (size_t)K * F) : 0;
  pl->off_gpart = off;   off += need_generic ? align256(sizeof(float) * (size_t)pl->n_blocks_g * F * 4) : 0;
  pl->off_xpart = off;   off += want_dx ? align256(sizeof(float) * (size_t)B * pl->n_tiles_g * (G_TL + K - 1)) : 0;
  pl->total = off;
}

size_t backward_saved_workspace_bytes(const leafk_config* cfg, int B, int T, int want_grad_x) {
  if (!cfg || cfg->F < 1 || cfg->K < 2 || cfg->H < 1 || B < 1 || T < 1) return 0;
  BwdPlan pl;
  bwd_plan(cfg, B, T, false, want_grad_x != 0, &pl);
  return pl.total;
}

// saved_q == nullptr: generic path (per-filter sums from the FP32 kernel); x may be null when neither that nor grad_x
// is needed.
static int backward_core(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
                         const float* grad_out, const float* saved_p, const float* saved_q, const leafk_grads* grads,
                         float* grad_x, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  int rc = check_common(cfg, prm, B, T);
  if (rc) return rc;
  if (!grad_out || !saved_p || !grads || !workspace) return fail(LEAFK_EINVAL, "null pointer argument");
  if (!grads->kernel || !grads->pool_w) return fail(LEAFK_EINVAL, "null Gabor / pooling gradient");
  if (cfg->compression && (!grads->alpha || !grads->delta || !grads->root || !grads->ema_w))
    return fail(LEAFK_EINVAL, "compression=1 needs the gradients of alpha, delta, root, ema_w");
  const bool need_generic = saved_q == nullptr, want_dx = grad_x != nullptr;
  if ((need_generic || want_dx) && !x) return fail(LEAFK_EINVAL, "the waveform is needed for this backward");
  if (want_dx && cfg->input_format == LEAFK_INPUT_S16)
    return fail(LEAFK_EINVAL, "no gradient w.r.t. an int16 waveform");
  if (want_dx && cfg->prep)
    return fail(LEAFK_EINVAL, "no gradient w.r.t. a waveform that is cropped / normalised on the fly (cfg->prep)");
  BwdPlan pl;
  bwd_plan(cfg, B, T, need_generic, want_dx, &pl);
  if (pl.total > workspace_bytes)
    return fail(LEAFK_EWORKSPACE, "backward workspace %zu bytes < %zu needed", workspace_bytes, pl.total);
  const int F = cfg->F, K = cfg->K;
  uint8_t* base = (uint8_t*)workspace;
  float* scratch = (float*)(base + pl.off_scratch);
  float* rpart = (float*)(base + pl.off_rpart);
  float* dp = (need_generic || want_dx) ? (float*)(base + pl.off_dp) : nullptr;
  float* gpart = need_generic ? (float*)(base + pl.off_gpart) : nullptr;

  PcenBwdArgs pa;
  pa.p = saved_p; pa.gout = grad_out; pa.q = saved_q; pa.alpha = prm->alpha; pa.delta = prm->delta; pa.root = prm->root;
  pa.ema_w = prm->ema_w; pa.scratch = scratch; pa.dp = dp; pa.rpart = rpart; pa.pcen_floor = cfg->pcen_floor;
  pa.clamp_min = cfg->clamp_min; pa.compression = cfg->compression;
  const int fgroups = (F + B1_FPB - 1) / B1_FPB;
  bwd_pcen_kernel<<<(unsigned)((long long)B * fgroups), B1_FPB * 32, 0, stream>>>(B, F, pl.N, pa);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "bwd_pcen launch: %s", cudaGetErrorString(err));
  int launches = 1;

  if (need_generic || want_dx) {
    Geom g;
    whole_clip_geom(cfg, B, T, G_TL, &g);
    float* cprm = (float*)(base + pl.off_cprm);
    float* w32 = (float*)(base + pl.off_w32);
    float* g32 = (float*)(base + pl.off_g32);
    launch_k0(prm->kernel, prm->pool_w, F, K, g.Kp, g.C2p, cprm, w32, g32, nullptr, 16, 1, nullptr, nullptr, 0.f, 0.f,
              nullptr, 0, stream);
    err = cudaGetLastError();
    if (err != cudaSuccess) return fail(LEAFK_ECUDA, "k0 launch: %s", cudaGetErrorString(err));
    GenericArgs ga;
    ga.x = x; ga.w32 = w32; ga.g32 = g32; ga.cprm = cprm; ga.dp = dp; ga.gpart = gpart;
    ga.xpart = want_dx ? (float*)(base + pl.off_xpart) : nullptr;
    const int WL = G_TL + K - 1;
    const size_t smem = sizeof(float) * ((size_t)WL + 2 * K + 2 * G_TL + (want_dx ? WL : 0) + (need_generic ? 3 * F : 0) + 16);
    auto launch = [&](auto kern) -> cudaError_t {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      kern<<<pl.n_blocks_g, G_TL, smem, stream>>>(g, ga);
      return cudaGetLastError();
    };
    if (smem > 220 * 1024) return fail(LEAFK_EINVAL, "window too long for the generic backward (%d taps)", K);
    if (need_generic && want_dx) err = launch(bwd_generic_kernel<true, true>);
    else if (need_generic) err = launch(bwd_generic_kernel<true, false>);
    else err = launch(bwd_generic_kernel<false, true>);
    if (err != cudaSuccess) return fail(LEAFK_ECUDA, "bwd_generic launch: %s", cudaGetErrorString(err));
    launches += 2;
    if (want_dx) {
      const long long total = (long long)B * T;
      const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
      bwd_dx_sum_kernel<<<blocks, 256, 0, stream>>>(g, ga.xpart, grad_x);
      err = cudaGetLastError();
      if (err != cudaSuccess) return fail(LEAFK_ECUDA, "bwd_dx_sum launch: %s", cudaGetErrorString(err));
      ++launches;
    }
  }

  FinishArgs fa;
  fa.rpart = rpart; fa.gpart = gpart; fa.B = B; fa.F = F; fa.n_gpart = pl.n_blocks_g; fa.K = K;
  fa.compression = cfg->compression;
  bank_bounds(K, &fa.mu_hi, &fa.sigma_lo, &fa.sigma_hi, &fa.pool_lo);
  fa.prm = *prm; fa.g = *grads;
  bwd_finish_kernel<<<F, 128, 0, stream>>>(fa);
  err = cudaGetLastError();
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "bwd_finish launch: %s", cudaGetErrorString(err));
  count_launch(launches + 1);
  return LEAFK_OK;
}

int backward_saved_run(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
                       const float* grad_out, const float* saved, const leafk_grads* grads, float* grad_x,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!cfg || !saved) return fail(LEAFK_EINVAL, "null pointer argument");
  const int K = cfg->K, H = cfg->H;
  if (cfg->F < 1 || K < 2 || H < 1 || B < 1 || T < 1) return fail(LEAFK_EINVAL, "bad shape");
  const int N = (T + (K / 2 + (K % 2) - 1) + K / 2 - K) / H + 1;
  const size_t bfn = (size_t)B * cfg->F * N;
  return backward_core(cfg, prm, x, B, T, grad_out, saved, saved + bfn, grads, grad_x, workspace, workspace_bytes, stream);
}

// leafk_backward: from the waveform and the saved pooled energies only.  Tensor-core geometries re-run the training
// forward into the workspace to obtain the Q's; every other geometry takes the generic FP32 kernel.
size_t bwd_workspace_bytes(const leafk_config* cfg, int B, int T) {
  if (!cfg || cfg->F < 1 || cfg->K < 2 || cfg->H < 1 || B < 1 || T < 1) return 0;
  TrainPlan tp;
  BwdPlan pl;
  const int K = cfg->K, H = cfg->H;
  const int N = (T + (K / 2 + (K % 2) - 1) + K / 2 - K) / H + 1;
  const size_t bfn = (size_t)B * cfg->F * N;
  if (train_plan(cfg, B, T, &tp)) {
    bwd_plan(cfg, B, T, false, true, &pl);
    return pl.total + align256(sizeof(float) * bfn * 5) + tp.total;      // + out, saved(4) of the re-run, its workspace
  }
  bwd_plan(cfg, B, T, true, true, &pl);
  return pl.total;
}

int bwd_run(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
            const float* grad_out, const float* saved_p, const leafk_grads* grads, float* grad_x,
            void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  int rc = check_common(cfg, prm, B, T);
  if (rc) return rc;
  if (!x || !grad_out || !saved_p || !grads || !workspace) return fail(LEAFK_EINVAL, "null pointer argument");
  if (bwd_workspace_bytes(cfg, B, T) > workspace_bytes)
    return fail(LEAFK_EWORKSPACE, "backward workspace %zu bytes < %zu needed", workspace_bytes, bwd_workspace_bytes(cfg, B, T));
  TrainPlan tp;
  if (train_plan(cfg, B, T, &tp)) {
    BwdPlan pl;
    bwd_plan(cfg, B, T, false, true, &pl);
    const size_t bfn = (size_t)B * cfg->F * tp.N;
    uint8_t* base = (uint8_t*)workspace;
    float* re_out = (float*)(base + pl.total);
    float* re_saved = re_out + bfn;
    void* re_ws = base + pl.total + align256(sizeof(float) * bfn * 5);
    rc = forward_train_run(cfg, prm, x, B, T, re_out, re_saved, re_ws, tp.total, stream);
    if (rc) return rc;
    // the caller's saved_p is the same tensor the re-run produced; the Q's come from the re-run
    return backward_core(cfg, prm, x, B, T, grad_out, saved_p, re_saved + bfn, grads, grad_x, workspace, pl.total, stream);
  }
  return backward_core(cfg, prm, x, B, T, grad_out, saved_p, nullptr, grads, grad_x, workspace, workspace_bytes, stream);
}

}  // namespace leafk
