// K2: pooled-energy assembly + floor + PCEN.
//
// Replaces   the bias add of GaussianLowPass          reference pooling.py:41
//            torch.maximum(outputs, 1e-5)             reference frontend.py:84
//            ExponentialMovingAverage.forward         reference postprocessing.py:13-28
//            PCENLayer.forward                        reference postprocessing.py:62-69
// The reference runs the smoother as a Python loop over frames (5 tiny ops per frame); here it is
// a warp scan over affine maps M -> (1-w) M + w p, 128 frames per step, with the state carried in
// a register between steps (and in/out of the kernel for chunked long-form audio).
//
// One CTA = (clip b, 8 filters); warp = one filter; lane = 4 consecutive frames.  HBM traffic is
// the partial sums in (~6 B/element incl. overlap), `out` and optionally `p` out: this kernel is
// the only HBM-bound stage of the path, and it is <1 % of the forward time.
#include "leafk_common.cuh"
#include "k2_pcen_args.cuh"

namespace leafk {

constexpr int K2_FPB = 8;      // filters per block (= warps)
constexpr int K2_SEG = 128;    // frames per scan step


__global__ void __launch_bounds__(K2_FPB * 32)
k2_pcen_kernel(const Geom g, const float* __restrict__ ppart, const PcenArgs a) {
  __shared__ float ps[K2_FPB][K2_SEG + 4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int fgroups = (g.F + K2_FPB - 1) / K2_FPB;
  const int b = blockIdx.x / fgroups;
  const int f0 = (blockIdx.x % fgroups) * K2_FPB;
  const int f = f0 + warp;
  const bool fok = f < g.F;

  float w = 0.f, alpha = 1.f, delta = 0.f, q = 1.f, dq = 0.f;
  if (fok && a.compression) {
    w = fminf(fmaxf(a.ema_w[f], 0.f), 1.f);                  // postprocessing.py:14
    alpha = fminf(a.alpha[f], 1.0f);                         // postprocessing.py:63
    const float r = fmaxf(a.root[f], 1.0f);                  // postprocessing.py:64
    q = 1.0f / r;
    delta = a.delta[f];
    dq = powf(delta, q);
  }
  const float om = 1.0f - w;
  float carry = 0.f;
  bool have_carry = false;
  if (fok && a.compression && a.ema_in != nullptr) {
    carry = a.ema_in[(size_t)b * g.F + f];
    have_carry = true;
  }

  for (int seg0 = 0; seg0 < g.n_count; seg0 += K2_SEG) {
    const int seg_n = min(K2_SEG, g.n_count - seg0);
    // ---- phase A: assemble p for (8 filters) x (seg_n frames) --------------------------------
    __syncthreads();
    for (int idx = tid; idx < seg_n * K2_FPB; idx += blockDim.x) {
      const int fl = idx % K2_FPB, nl = idx / K2_FPB;
      const int ff = f0 + fl;
      float v = 0.f;
      if (ff < g.F) {
        const int n = g.n_begin + seg0 + nl;
        long long wlo = (long long)n * g.H - g.padL, whi = wlo + g.K - 1;
        if (wlo < g.te_lo) wlo = g.te_lo;
        if (whi > g.te_hi - 1) whi = g.te_hi - 1;
        const int i0 = (int)((wlo - g.te_lo) / g.TL), i1 = (int)((whi - g.te_lo) / g.TL);
        float s = 0.f;
        for (int i = i0; i <= i1; ++i) {
          const int nf = first_frame_of(g, g.te_lo + (long long)i * g.TL);
          s += __ldg(ppart + (((size_t)b * g.n_tiles + i) * g.SL + (n - nf)) * g.F + ff);
        }
        if (a.pool_b != nullptr) s += __ldg(a.pool_b + ff);
        v = fmaxf(s, a.clamp_min);                             // frontend.py:84
      }
      ps[fl][nl] = v;
    }
    __syncthreads();
    if (!fok) continue;

    // ---- phase B: smoother scan + compression ------------------------------------------------
    float p[4];
    int cnt = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nl = lane * 4 + j;
      p[j] = (nl < seg_n) ? ps[warp][nl] : 0.f;
      cnt += (nl < seg_n);
    }
    float m[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.compression) {
      if (!have_carry) {                                       // state starts at the first frame
        carry = ps[warp][0];                                   // postprocessing.py:15
        have_carry = true;
      }
      // lane-local composite map  M -> A*M + C  over this lane's frames
      float A = 1.f, C = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (j < cnt) { C = fmaf(om, C, w * p[j]); A *= om; }
      // inclusive scan of the maps across lanes
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float Ap = __shfl_up_sync(0xffffffffu, A, o);
        const float Cp = __shfl_up_sync(0xffffffffu, C, o);
        if (lane >= o) { C = fmaf(A, Cp, C); A *= Ap; }
      }
      float Aex = __shfl_up_sync(0xffffffffu, A, 1), Cex = __shfl_up_sync(0xffffffffu, C, 1);
      if (lane == 0) { Aex = 1.f; Cex = 0.f; }
      float state = fmaf(Aex, carry, Cex);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        state = fmaf(om, state, w * p[j]);                     // postprocessing.py:22
        m[j] = state;
      }
      const float Al = __shfl_sync(0xffffffffu, A, 31), Cl = __shfl_sync(0xffffffffu, C, 31);
      carry = fmaf(Al, carry, Cl);
    }
    float* orow = a.out + (size_t)b * a.ldo_b + (size_t)f * a.ldo_f + seg0;
    float* prow = a.saved_p ? a.saved_p + (size_t)b * a.ldo_b + (size_t)f * a.ldo_f + seg0 : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nl = lane * 4 + j;
      if (nl < seg_n) {
        float o = p[j];
        if (a.compression) {
          const float d = a.pcen_floor + m[j];
          const float u = p[j] / powf(d, alpha) + delta;       // postprocessing.py:66
          o = powf(u, q) - dq;
        }
        orow[nl] = o;
        if (prow) prow[nl] = p[j];
      }
    }
  }
  if (fok && a.compression && a.ema_out != nullptr && lane == 0) a.ema_out[(size_t)b * g.F + f] = carry;
}

cudaError_t launch_k2(const Geom& g, const float* ppart, const PcenArgs& a, cudaStream_t stream) {
  const int fgroups = (g.F + K2_FPB - 1) / K2_FPB;
  k2_pcen_kernel<<<(unsigned)((long long)g.B * fgroups), K2_FPB * 32, 0, stream>>>(g, ppart, a);
  return cudaGetLastError();
}

}  // namespace leafk
