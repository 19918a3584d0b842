// Backward of the LEAF frontend: parameter gradients of sum(out * grad_out).
//
// Replaces autograd through reference frontend.py:78-89 (what loss.backward() does in train.py:258;
// the reference graph has ~7*N_frames nodes because of the Python EMA loop).  Formulas: SURVEY A.2.
//
//   B1  bwd_pcen_kernel     per (clip, filter) row: PCEN + smoother backward (forward scan to rebuild M,
//                           reverse affine scan for the smoother adjoint), floor mask; emits dp in
//                           frame-major layout (B,N,F) and per-row sums for alpha, delta, root, ema_w, bias.
//   K0b k0_banks_bwd_kernel derivative banks (k0_banks.cu)
//   B2  k1_tc_kernel<.,.,1> the three correlations y, z, v on tensor cores with the backward epilogue
//                           (k1_tc.cu): per-CTA partial sums S_mu, S_sigma, S_poolw per filter
//   B3  bwd_finish_kernel   fixed-order sums over clips / CTAs, clamp masks, chain-rule factors -> the 7 grads
// Everything is deterministic (no atomics).
#include "../../include/leafk.h"
#include "leafk_common.cuh"
#include "k1_tc_layout.cuh"

#include <cstring>

namespace leafk {

int fail(int code, const char* fmt, ...);
void count_launch(int n);
void launch_k0_bwd(const float* kernel, const float* pool_w, int F, int K, int Kp, int FB, int n_groups,
                   float* bprm, uint8_t* w16b, cudaStream_t stream);
void bank_bounds(int K, float* mu_hi, float* sigma_lo, float* sigma_hi, float* pool_lo);
cudaError_t launch_k1_tc_bwd(const Geom& g, const float* x, const uint8_t* w16b, int FB, int n_groups,
                             const float* dpT, const float* bprm, float* bpart, int* ctas_per_group,
                             int skip_xlo, cudaStream_t stream);

constexpr int B1_FPB = 8;      // filters per block (= warps)
constexpr int B1_SEG = 128;    // frames per scan step

struct PcenBwdArgs {
  const float* p;        // (B,F,N) floored pooled energies saved by the forward
  const float* gout;     // (B,F,N)
  const float* alpha; const float* delta; const float* root; const float* ema_w;
  float* scratch;        // (B,F,N,3): t1 = G*D^-alpha, dM, p_n - M_{n-1}
  float* dpT;            // (B,N,F)
  float* rpart;          // (B,F,8): d_delta, d_alpha_hat, d_root_hat, d_w_hat, d_bias
  float pcen_floor, clamp_min;
  int compression;
};

__global__ void __launch_bounds__(B1_FPB * 32)
bwd_pcen_kernel(int B, int F, int N, const PcenBwdArgs a) {
  __shared__ float tile[B1_FPB][B1_SEG + 1];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fgroups = (F + B1_FPB - 1) / B1_FPB;
  const int b = blockIdx.x / fgroups;
  const int f0 = (blockIdx.x % fgroups) * B1_FPB;
  const int f = f0 + warp;
  const bool fok = f < F;
  const size_t row = ((size_t)b * F + (fok ? f : 0)) * N;
  const int nseg = (N + B1_SEG - 1) / B1_SEG;

  float s_delta = 0.f, s_alpha = 0.f, s_root = 0.f, s_w = 0.f, s_bias = 0.f;
  float w = 0.f, om = 1.f;

  if (a.compression) {
    float alpha = 1.f, delta = 0.f, q = 1.f, dq = 0.f, dq1 = 0.f, ldelta = 0.f;
    if (fok) {
      w = fminf(fmaxf(a.ema_w[f], 0.f), 1.f);
      alpha = fminf(a.alpha[f], 1.0f);
      q = 1.0f / fmaxf(a.root[f], 1.0f);
      delta = a.delta[f];
      dq = powf(delta, q);
      dq1 = powf(delta, q - 1.0f);
      ldelta = logf(delta);
    }
    om = 1.0f - w;
    // ---------------- forward sweep: rebuild M, per-element adjoint seeds -----------------------
    float carry = fok ? a.p[row] : 0.f;                 // M_{-1} = p_0   (postprocessing.py:15)
    for (int sg = 0; sg < nseg && fok; ++sg) {
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
          const float D = a.pcen_floor + state;
          const float Dma = 1.0f / powf(D, alpha);
          const float u = p[j] * Dma + delta;
          const float uq1 = powf(u, q - 1.0f);
          const float G = go[j] * q * uq1;
          s_delta += G - go[j] * q * dq1;
          s_alpha -= G * p[j] * Dma * logf(D);
          s_root -= go[j] * (u * uq1 * logf(u) - dq * ldelta) * (q * q);
          float* sc = a.scratch + (row + seg0 + nl) * 3;
          sc[0] = G * Dma;
          sc[1] = -G * alpha * p[j] * Dma / D;
          sc[2] = p[j] - prev;
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
    __syncthreads();
    if (fok) {
      float dpv[4];
      if (a.compression) {
        float t1[4], dM[4], pm[4];
        int cnt = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int nl = lane * 4 + j;
          const bool ok = nl < seg_n;
          const float* sc = a.scratch + (row + seg0 + (ok ? nl : 0)) * 3;
          t1[j] = ok ? sc[0] : 0.f; dM[j] = ok ? sc[1] : 0.f; pm[j] = ok ? sc[2] : 0.f;
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
          const float pv = a.p[row + seg0 + nl];
          const float d = (pv > a.clamp_min) ? dpv[j] : 0.f;      // torch.maximum(., 1e-5)  frontend.py:84
          s_bias += d;
          tile[warp][nl] = d;
        }
      }
    } else {
      for (int nl = lane; nl < seg_n; nl += 32) tile[warp][nl] = 0.f;
    }
    __syncthreads();
    for (int idx = tid; idx < seg_n * B1_FPB; idx += blockDim.x) {
      const int fl = idx % B1_FPB, nl = idx / B1_FPB;
      if (f0 + fl < F) a.dpT[((size_t)b * N + seg0 + nl) * F + f0 + fl] = tile[fl][nl];
    }
  }
  // ---------------- per-row sums ------------------------------------------------------------------
  float v[5] = {s_delta, s_alpha, s_root, s_w, s_bias};
#pragma unroll
  for (int i = 0; i < 5; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  if (fok && lane == 0) {
    float* r = a.rpart + ((size_t)b * F + f) * 8;
#pragma unroll
    for (int i = 0; i < 5; ++i) r[i] = v[i];
  }
}

struct FinishArgs {
  const float* rpart;    // (B,F,8)
  const float* bpart;    // (ctas_per_group, Fpad, 4)
  const float* bprm;     // (Fpad, 8)
  int B, F, Fpad, ctas_per_group, K, compression;
  float mu_hi, sigma_lo, sigma_hi, pool_lo;
  leafk_params prm;
  leafk_grads g;
};

__global__ void bwd_finish_kernel(const FinishArgs a) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.F) return;
  float r[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int b = 0; b < a.B; ++b) {
    const float* p = a.rpart + ((size_t)b * a.F + f) * 8;
#pragma unroll
    for (int i = 0; i < 5; ++i) r[i] += p[i];
  }
  float s[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < a.ctas_per_group; ++c) {
    const float* p = a.bpart + ((size_t)c * a.Fpad + f) * 4;
#pragma unroll
    for (int i = 0; i < 3; ++i) s[i] += p[i];
  }
  // Gabor parameters: d/dmu = 2*S_mu, d/dsigma = 2*S_sigma, gated by the clamps (convolution.py:20-21)
  const float th0 = a.prm.kernel[2 * f], th1 = a.prm.kernel[2 * f + 1];
  a.g.kernel[2 * f] = (th0 >= 0.f && th0 <= a.mu_hi) ? 2.0f * s[0] : 0.f;
  a.g.kernel[2 * f + 1] = (th1 >= a.sigma_lo && th1 <= a.sigma_hi) ? 2.0f * s[1] : 0.f;
  // pooling width: dg/ds = g * (k-c)^2 / (s^3 c^2)   (impulse_responses.py:75-80)
  const float sraw = a.prm.pool_w[f];
  const float ps = a.bprm[(size_t)f * 8 + 5];
  const float c = 0.5f * (float)(a.K - 1);
  a.g.pool_w[f] = (sraw >= a.pool_lo && sraw <= 0.5f) ? s[2] / (ps * ps * ps * c * c) : 0.f;
  if (a.g.pool_b) a.g.pool_b[f] = r[4];
  if (a.compression) {
    const float al = a.prm.alpha[f], ro = a.prm.root[f], w = a.prm.ema_w[f];
    a.g.delta[f] = r[0];
    a.g.alpha[f] = (al < 1.0f) ? r[1] : (al == 1.0f ? 0.5f * r[1] : 0.f);      // torch.min tie -> 1/2
    a.g.root[f] = (ro > 1.0f) ? r[2] : (ro == 1.0f ? 0.5f * r[2] : 0.f);       // torch.max tie -> 1/2
    a.g.ema_w[f] = (w >= 0.f && w <= 1.0f) ? r[3] : 0.f;                       // clamp passes at the bounds
  }
}

// ------------------------------------------------------------------------------------------------
struct BwdPlan {
  int FB, n_groups, Fpad, Kp, N, n_tiles, SL, max_ctas;
  size_t off_bprm, off_w16b, off_dpT, off_scratch, off_rpart, off_bpart, total;
};

static bool bwd_plan(const leafk_config* cfg, int B, int T, BwdPlan* pl) {
  const int K = cfg->K, H = cfg->H, F = cfg->F;
  pl->Kp = (K + 15) / 16 * 16;
  const int nslot = tc::slots_per_thread(K, H);
  pl->FB = tc::bwd_filters_per_group(pl->Kp, nslot);
  if (pl->FB == 0) return false;
  pl->n_groups = (F + pl->FB - 1) / pl->FB;
  pl->Fpad = pl->n_groups * pl->FB;
  const int padL = K / 2 + (K % 2) - 1, padR = K / 2;
  pl->N = (T + padL + padR - K) / H + 1;
  pl->n_tiles = (T + tc::TILE - 1) / tc::TILE;
  pl->SL = (tc::TILE + K - 2) / H + 1;
  pl->max_ctas = 256;                                   // >= SMs per group on any part
  size_t off = 0;
  pl->off_bprm = off;    off += align256(sizeof(float) * 8 * pl->Fpad);
  pl->off_w16b = off;    off += align256(tc::b_group_bytes(6 * pl->FB, pl->Kp) * (size_t)pl->n_groups);
  pl->off_dpT = off;     off += align256(sizeof(float) * (size_t)B * pl->N * F);
  pl->off_scratch = off; off += align256(sizeof(float) * (size_t)B * pl->N * F * 3);
  pl->off_rpart = off;   off += align256(sizeof(float) * (size_t)B * F * 8);
  pl->off_bpart = off;   off += align256(sizeof(float) * (size_t)pl->max_ctas * pl->Fpad * 4);
  pl->total = off;
  return true;
}

size_t bwd_workspace_bytes(const leafk_config* cfg, int B, int T) {
  if (!cfg || cfg->F < 1 || cfg->K < 2 || cfg->H < 1 || B < 1 || T < 1) return 0;
  BwdPlan pl;
  if (!bwd_plan(cfg, B, T, &pl)) return 0;
  return pl.total;
}

int bwd_run(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
            const float* grad_out, const float* saved_p, const leafk_grads* grads, float* grad_x,
            void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  if (!cfg || !prm || !x || !grad_out || !saved_p || !grads || !workspace)
    return fail(LEAFK_EINVAL, "null pointer argument");
  if (grad_x != nullptr)
    return fail(LEAFK_EINVAL, "gradient w.r.t. the waveform is not implemented (train.py never needs it)");
  if (!prm->kernel || !prm->pool_w || !grads->kernel || !grads->pool_w)
    return fail(LEAFK_EINVAL, "null Gabor / pooling parameter or gradient");
  if (cfg->compression && (!prm->alpha || !prm->delta || !prm->root || !prm->ema_w || !grads->alpha ||
                           !grads->delta || !grads->root || !grads->ema_w))
    return fail(LEAFK_EINVAL, "compression=1 needs alpha, delta, root, ema_w and their gradients");
  if (cfg->F < 1 || cfg->K < 2 || cfg->H < 1 || B < 1 || T < 1) return fail(LEAFK_EINVAL, "bad shape");
  BwdPlan pl;
  if (!bwd_plan(cfg, B, T, &pl))
    return fail(LEAFK_EINVAL, "backward: geometry (K=%d,H=%d) not covered by the tensor-core kernel", cfg->K, cfg->H);
  if (pl.total > workspace_bytes)
    return fail(LEAFK_EWORKSPACE, "backward workspace %zu bytes < %zu needed", workspace_bytes, pl.total);
  const int F = cfg->F, K = cfg->K, H = cfg->H;
  uint8_t* base = (uint8_t*)workspace;
  float* bprm = (float*)(base + pl.off_bprm);
  uint8_t* w16b = base + pl.off_w16b;
  float* dpT = (float*)(base + pl.off_dpT);
  float* scratch = (float*)(base + pl.off_scratch);
  float* rpart = (float*)(base + pl.off_rpart);
  float* bpart = (float*)(base + pl.off_bpart);

  Geom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.F = F; g.K = K; g.H = H;
  g.padL = K / 2 + (K % 2) - 1; g.padR = K / 2;
  g.C2 = 2 * F; g.C2p = (g.C2 + 7) / 8 * 8; g.Kp = pl.Kp;
  g.T_total = T; g.t_off = 0; g.T_win = T; g.ldx = T;
  g.N_total = pl.N; g.n_begin = 0; g.n_count = pl.N;
  g.te_lo = 0; g.te_hi = T; g.TL = tc::TILE; g.n_tiles = pl.n_tiles; g.SL = pl.SL;
  g.x_fmt = cfg->input_format == LEAFK_INPUT_S16 ? 1 : 0;

  PcenBwdArgs pa;
  pa.p = saved_p; pa.gout = grad_out; pa.alpha = prm->alpha; pa.delta = prm->delta; pa.root = prm->root;
  pa.ema_w = prm->ema_w; pa.scratch = scratch; pa.dpT = dpT; pa.rpart = rpart; pa.pcen_floor = cfg->pcen_floor;
  pa.clamp_min = cfg->clamp_min; pa.compression = cfg->compression;
  const int fgroups = (F + B1_FPB - 1) / B1_FPB;
  bwd_pcen_kernel<<<(unsigned)((long long)B * fgroups), B1_FPB * 32, 0, stream>>>(B, F, pl.N, pa);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "bwd_pcen launch: %s", cudaGetErrorString(err));

  launch_k0_bwd(prm->kernel, prm->pool_w, F, K, pl.Kp, pl.FB, pl.n_groups, bprm, w16b, stream);
  err = cudaGetLastError();
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "k0_bwd launch: %s", cudaGetErrorString(err));

  // rows of bpart that no CTA owns (uneven pair split, fewer unit pairs than SM pairs) must read as zero
  err = cudaMemsetAsync(bpart, 0, sizeof(float) * (size_t)pl.max_ctas * pl.Fpad * 4, stream);
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "bpart reset: %s", cudaGetErrorString(err));
  int ctas_per_group = 0;
  const int skip_xlo = (cfg->algo & LEAFK_BWD_2PRODUCT) ? 1 : 0;
  err = launch_k1_tc_bwd(g, x, w16b, pl.FB, pl.n_groups, dpT, bprm, bpart, &ctas_per_group, skip_xlo, stream);
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "k1_tc_bwd launch: %s", cudaGetErrorString(err));
  if (ctas_per_group > pl.max_ctas) return fail(LEAFK_EINVAL, "internal: partial buffer too small");

  FinishArgs fa;
  fa.rpart = rpart; fa.bpart = bpart; fa.bprm = bprm; fa.B = B; fa.F = F; fa.Fpad = pl.Fpad;
  fa.ctas_per_group = ctas_per_group; fa.K = K; fa.compression = cfg->compression;
  bank_bounds(K, &fa.mu_hi, &fa.sigma_lo, &fa.sigma_hi, &fa.pool_lo);
  fa.prm = *prm; fa.g = *grads;
  bwd_finish_kernel<<<(F + 127) / 128, 128, 0, stream>>>(fa);
  err = cudaGetLastError();
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "bwd_finish launch: %s", cudaGetErrorString(err));
  count_launch(4);
  return LEAFK_OK;
}

}  // namespace leafk
