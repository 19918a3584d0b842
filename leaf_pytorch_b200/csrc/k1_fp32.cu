// K1 (CUDA-core variant): Gabor correlation + squared modulus + Gaussian pooling partials, FP32 FMA.
//
// Replaces   F.conv1d(pad(x), bank)        reference convolution.py:91-98
//            SquaredModulus.forward        reference frontend.py:15-19
//            GaussianLowPass.forward       reference pooling.py:31-42   (bias is added in K2)
// without ever materialising the (B,2F,T) activation the reference writes to memory.
//
// This is the general-geometry path (any F, K, H) and the arithmetic reference for the tensor-core
// kernel in k1_tc_kernel.cuh; it is bound by FP32-FMA issue, not HBM (~12.9 kFLOP per byte).
//
// One CTA = one (clip, tile of TL=512 e-samples).  The clip window (TL + Kp samples) sits in
// shared memory; the bank is streamed through shared memory in slices of 16 taps (cp.async, double
// buffered).  A warp owns 8 channels (4 filters); a lane owns 4 consecutive samples, so per tap a
// lane issues 32 FMAs against 2 broadcast LDS.128 (taps) and 1/4 LDS.128 (sliding sample window).
// After the 401 taps the lane holds e for 4 samples x 4 filters; the pooling partials of the
// frames overlapping the chunk are reduced with warp shuffles and parked per (chunk,slot,filter) so
// the final per-tile sum has a fixed order (bit-reproducible run to run).
#include "leafk_common.cuh"

namespace leafk {

constexpr int F32_TL = 512;   // e-samples per tile
constexpr int F32_TC = 128;   // samples per warp chunk (32 lanes x 4)
constexpr int F32_KS = 16;    // taps per staged slice
constexpr int F32_NCH = F32_TL / F32_TC;

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__global__ void __launch_bounds__(512)
k1_fp32_kernel(const Geom g, const float* __restrict__ x, const float* __restrict__ w32,
               const float* __restrict__ g32, float* __restrict__ ppart, int rounds) {
  extern __shared__ __align__(16) float smem[];
  const int XS = F32_TL + g.Kp + 8;                    // window length (floats, multiple of 4)
  float* xs = smem;                                    // [XS]
  float* wsl = xs + XS;                                // [2][KS][C2p]
  float* pitem = wsl + 2 * F32_KS * g.C2p;             // [NCH][SL][F]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int tile = blockIdx.x % g.n_tiles;
  const int b = blockIdx.x / g.n_tiles;
  const long long ts = g.te_lo + (long long)tile * F32_TL;        // first e-sample of the tile
  const long long te = (ts + F32_TL < g.te_hi) ? ts + F32_TL : g.te_hi;
  const int n_first = first_frame_of(g, ts);

  // ---- stage the sample window: xs[i] = x~[ts - padL + i], zero outside the clip -------------
  const ClipView cv = clip_view(g, b);
  for (int i = tid; i < XS; i += blockDim.x) {
    const long long a = ts - g.padL + i;
    const long long wi = a - g.t_off;
    float v = 0.f;
    if (a >= 0 && a < g.T_total && wi >= 0 && wi < g.T_win) v = clip_sample(g, x, cv, wi);
    xs[i] = v;
  }
  for (int i = tid; i < F32_NCH * g.SL * g.F; i += blockDim.x) pitem[i] = 0.f;

  const int n_slices = g.Kp / F32_KS;
  const int slice_floats = F32_KS * g.C2p;             // multiple of 4
  const int n_chunks = (int)((te - ts + F32_TC - 1) / F32_TC);

  for (int r = 0; r < rounds; ++r) {
    const int grp = r * nwarps + warp;                 // channel group of 8
    const bool active = grp * 8 < g.C2p;
    for (int ch = 0; ch < n_chunks; ++ch) {
      float acc[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;

      const float* xp = xs + ch * F32_TC + lane * 4;
      // prologue: slice 0 -> buffer 0
      for (int i = tid * 4; i < slice_floats; i += blockDim.x * 4) cp_async16(wsl + i, w32 + i);
      cp_async_commit();
      for (int s = 0; s < n_slices; ++s) {
        if (s + 1 < n_slices) {
          float* dst = wsl + ((s + 1) & 1) * slice_floats;
          const float* src = w32 + (size_t)(s + 1) * slice_floats;
          for (int i = tid * 4; i < slice_floats; i += blockDim.x * 4) cp_async16(dst + i, src + i);
          cp_async_commit();
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncthreads();                               // slice s (and, first time, xs) visible
        if (active) {
          const float* wb = wsl + (s & 1) * slice_floats + grp * 8;
          const int k0 = s * F32_KS;
          float4 cur = *reinterpret_cast<const float4*>(xp + k0);
          float xr[8] = {cur.x, cur.y, cur.z, cur.w, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int kk = 0; kk < F32_KS; kk += 4) {
            const float4 nx = *reinterpret_cast<const float4*>(xp + k0 + kk + 4);
            xr[4] = nx.x; xr[5] = nx.y; xr[6] = nx.z; xr[7] = nx.w;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4 wa = *reinterpret_cast<const float4*>(wb + (kk + u) * g.C2p);
              const float4 wc = *reinterpret_cast<const float4*>(wb + (kk + u) * g.C2p + 4);
              const float w[8] = {wa.x, wa.y, wa.z, wa.w, wc.x, wc.y, wc.z, wc.w};
#pragma unroll
              for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(xr[u + j], w[c], acc[j][c]);
            }
            xr[0] = xr[4]; xr[1] = xr[5]; xr[2] = xr[6]; xr[3] = xr[7];
          }
        }
        __syncthreads();                               // everyone done with buffer s&1 before refill
      }

      // ---- modulus + pooling partials of this chunk ------------------------------------------
      if (active) {
        const long long tc0 = ts + (long long)ch * F32_TC;            // chunk's first sample
        const long long tc1 = (tc0 + F32_TC < te) ? tc0 + F32_TC : te; // one past its last
        const int nA = first_frame_of(g, tc0);
        const int nB = last_frame_of(g, tc1 - 1);
        float e[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int q = 0; q < 4; ++q) e[j][q] = acc[j][2 * q] * acc[j][2 * q] + acc[j][2 * q + 1] * acc[j][2 * q + 1];
        const long long t0 = tc0 + lane * 4;
        for (int n = nA; n <= nB; ++n) {
          const long long kbase = t0 + g.padL - (long long)n * g.H;   // tap index of sample t0 in frame n
          float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const long long k = kbase + j;
            if (k >= 0 && k < g.K && t0 + j < te) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int f = grp * 4 + q;
                if (f < g.F) part[q] = fmaf(__ldg(g32 + (size_t)k * g.F + f), e[j][q], part[q]);
              }
            }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part[q] += __shfl_xor_sync(0xffffffffu, part[q], o);
          }
          if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int f = grp * 4 + q;
              if (f < g.F) pitem[(ch * g.SL + (n - n_first)) * g.F + f] = part[q];
            }
          }
        }
      }
    }
  }
  __syncthreads();
  // ---- fixed-order sum over chunks -> partial pooled sums of this tile -----------------------
  float* dst = ppart + ((size_t)b * g.F * g.n_tiles + tile) * g.SL;      // layout [clip][filter][tile][slot]
  for (int i = tid; i < g.SL * g.F; i += blockDim.x) {       // i = f*SL + slot
    const int f = i / g.SL, slot = i % g.SL;
    float s = 0.f;
    for (int ch = 0; ch < F32_NCH; ++ch) s += pitem[(ch * g.SL + slot) * g.F + f];
    dst[(size_t)f * g.n_tiles * g.SL + slot] = s;
  }
}

size_t k1_fp32_smem_bytes(const Geom& g) {
  return sizeof(float) * ((size_t)(F32_TL + g.Kp + 8) + 2 * F32_KS * g.C2p + (size_t)F32_NCH * g.SL * g.F);
}

cudaError_t launch_k1_fp32(const Geom& g, const float* x, const float* w32, const float* g32,
                           float* ppart, cudaStream_t stream) {
  const int ngroups = g.C2p / 8;
  const int rounds = (ngroups + 15) / 16;
  const int nwarps = (ngroups + rounds - 1) / rounds;
  const size_t smem = k1_fp32_smem_bytes(g);
  cudaError_t err = cudaFuncSetAttribute(k1_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) return err;
  const long long nblk = (long long)g.B * g.n_tiles;
  k1_fp32_kernel<<<(unsigned)nblk, nwarps * 32, smem, stream>>>(g, x, w32, g32, ppart, rounds);
  return cudaGetLastError();
}

}  // namespace leafk
