// K2: pooled-energy assembly + floor + PCEN.
//
// Replaces   the bias add of GaussianLowPass          reference pooling.py:41
//            torch.maximum(outputs, 1e-5)             reference frontend.py:84
//            ExponentialMovingAverage.forward         reference postprocessing.py:13-28
//            PCENLayer.forward                        reference postprocessing.py:62-69
// The reference runs the smoother as a Python loop over frames (5 tiny ops per frame); here it is
// a warp scan over affine maps M -> (1-w) M + w p, 4 x 32 frames per step, with the state carried in
// a register between steps (and in/out of the kernel for chunked long-form audio).
//
// One warp = one (clip, filter) row; lane = frame within a group of 32 consecutive frames, so the
// partial sums (layout [clip][filter][tile][slot]: the partials of a row are contiguous) are read and `out` is written with
// unit stride across lanes.  Grid-stride over rows, no shared memory, no block barrier.  This is the
// only HBM-bound stage of the path (~14 B per output element).
#include "leafk_common.cuh"
#include "k2_pcen_args.cuh"
#include <cuda_bf16.h>
#include <cstring>

namespace leafk {

constexpr int K2_WARPS = 8;

// x^y for x > 0 as exp2(y*log2 x) with the accurate log2f/exp2f (no fast-math): ~1e-7 * max(1,|y log2 x|)
// relative, a third of the instructions of powf (whose special-case handling is not needed: x = floor + M > 0;
// a negative delta gives NaN exactly like powf / the reference).
__device__ __forceinline__ float pow_pos(float x, float y) { return exp2f(y * log2f(x)); }

// ROWBLOCK = false: one warp per (clip, filter) row (many short rows: 1 s clips).  ROWBLOCK = true: one block of 8 warps
// per row, for long rows (10 s / 60 s clips, N >= 512 frames): the smoother is a chain along the frames, so a single
// warp walks a 1000-frame row in 8 dependent steps of 128 frames while most of the GPU idles (512 rows at F=64, B=8).
// Here warp w takes the frames [n0 + 128 w, n0 + 128 w + 128) of a 1024-frame super-step, publishes the composite
// affine map of its segment, and picks up its carry-in by composing the maps of the warps before it (<= 7 FMAs).
template <bool ROWBLOCK>
__global__ void __launch_bounds__(K2_WARPS * 32, 6)
k2_pcen_kernel(const Geom g, const float* __restrict__ ppart, const PcenArgs a, int tl_shift, unsigned long long hop_magic) {
  __shared__ float s_A[K2_WARPS], s_C[K2_WARPS], s_first;
  // Programmatic dependent launch: this grid is scheduled under the tail of the kernel that writes the partial sums.
  // After the tensor-core K1 it does not wait for that whole grid: K1 counts, per clip, the tiles whose partial sums are
  // stored (release), and a row starts as soon as its clip is complete (acquire) -- clips finish in index order and
  // the SMs whose CTAs have run out of tiles (a third of them idle through K1's last tile) take the early clips, so
  // most of this kernel hides under K1's tail.  After any other producer: wait for the grid.
  if (a.done == nullptr) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = g.B * g.F;
  const int te_lo = (int)g.te_lo, te_hi = (int)g.te_hi;       // clip length <= 2^30 (checked on the host)
  const int n_end = g.n_begin + g.n_count;
  const int FV = a.q_out != nullptr ? 4 * g.F : g.F;            // virtual filters per (clip, tile) block
  const size_t tile_stride = (size_t)g.SL;                      // layout [clip][virtual filter][tile][slot]: a row's partials are contiguous

  for (int row = ROWBLOCK ? blockIdx.x : blockIdx.x * K2_WARPS + warp; row < rows;
       row += ROWBLOCK ? gridDim.x : gridDim.x * K2_WARPS) {
    const int b = row / g.F, f = row - b * g.F;
    if (a.done != nullptr) {
      if (lane == 0) {
        const int* flag = a.done + b;
        int v;
        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v < a.done_target) {
          const long long t0 = global_timer_ns();
          do {
            __nanosleep(100);
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            // bounded (K1 itself may be waiting up to H2D_TIMEOUT_NS for a host copy): report instead of hanging or
            // killing the context; the host API returns the error word after its next synchronisation
            if (v < a.done_target && global_timer_ns() - t0 > 2 * H2D_TIMEOUT_NS) {
              if (a.err != nullptr) atomicExch(a.err, LEAFK_ASYNC_K1_TIMEOUT);
              break;
            }
          } while (v < a.done_target);
        }
      }
      __syncwarp();
    }
    const float* pbase = ppart + ((size_t)b * FV + f) * g.n_tiles * g.SL;
    const size_t ooff = (size_t)b * a.ldo_b + (size_t)f * a.ldo_f;
    float* orow = a.out + ooff;
    __nv_bfloat16* orow16 = reinterpret_cast<__nv_bfloat16*>(a.out) + ooff;
    float* prow = a.saved_p ? a.saved_p + (size_t)b * a.ldo_b + (size_t)f * a.ldo_f : nullptr;

    float w = 0.f, alpha = 1.f, delta = 0.f, q = 1.f, dq = 0.f, om = 1.f, bias = 0.f;
    float carry = 0.f;
    bool have_carry = false;
    bool have_prm = false;

    // 128 frames per step: 4 independent groups of 32 consecutive frames (lane = frame within group), so the
    // loads, the 4 local scans and the 4 PCEN evaluations of a step overlap; only 4 FMAs chain the carry.
    for (int nb0 = g.n_begin; nb0 < n_end; nb0 += ROWBLOCK ? 128 * K2_WARPS : 128) {
      const int n0 = nb0 + (ROWBLOCK ? 128 * warp : 0);
      float p[4];
      bool ok[4];
      // partial pooled sums first (the longest latency of the row), parameters while they are in flight
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = n0 + 32 * u + lane;
        ok[u] = n < n_end;
        p[u] = 0.f;
        if (ok[u]) {
          int wlo = n * g.H - g.padL, whi = wlo + g.K - 1;
          if (wlo < te_lo) wlo = te_lo;
          if (whi > te_hi - 1) whi = te_hi - 1;
          const int i0 = (wlo - te_lo) >> tl_shift, i1 = (whi - te_lo) >> tl_shift;
          float s = 0.f;
          for (int i = i0; i <= i1; ++i) {                    // <= ceil(K/TL)+1 tiles, in tile order
            // first frame whose window reaches the tile start ts: max(n_begin, ceil((ts + padL - K + 1) / H))
            const int num = te_lo + (i << tl_shift) + g.padL - g.K + 1;
            int nf = num <= 0 ? 0 : (int)div_magic((unsigned)(num + g.H - 1), hop_magic);
            if (nf < g.n_begin) nf = g.n_begin;
            s += __ldcg(pbase + (size_t)i * tile_stride + (n - nf));   // written by a grid that may still be running: L2, not the read-only path
          }
          p[u] = s;
        }
      }
      if (!have_prm) {
        have_prm = true;
        bias = a.pool_b ? __ldg(a.pool_b + f) : 0.f;
        if (a.compression) {
          w = clamp_nan(__ldg(a.ema_w + f), 0.f, 1.f);             // postprocessing.py:14 (NaN propagates like torch.clamp)
          alpha = min_nan(__ldg(a.alpha + f), 1.0f);               // postprocessing.py:63
          q = 1.0f / max_nan(__ldg(a.root + f), 1.0f);             // postprocessing.py:64-65
          delta = __ldg(a.delta + f);
          dq = pow_pos(delta, q);
          om = 1.0f - w;
          if (a.ema_in != nullptr) {
            carry = a.ema_in[(size_t)b * g.F + f];
            have_carry = true;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) p[u] = max_nan(p[u] + bias, a.clamp_min);  // pooling.py:41, frontend.py:84 (torch.maximum keeps NaN)
      float o[4] = {p[0], p[1], p[2], p[3]};
      if (a.compression) {
        if (!have_carry) {                                    // smoother starts at the first frame
          if constexpr (ROWBLOCK) {
            if (warp == 0 && lane == 0) s_first = p[0];
            __syncthreads();
            carry = s_first;
          } else {
            carry = __shfl_sync(0xffffffffu, p[0], 0);        // postprocessing.py:15
          }
          have_carry = true;
        }
        float A[4], C[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { A[u] = ok[u] ? om : 1.f; C[u] = ok[u] ? w * p[u] : 0.f; }   // M -> A*M + C
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float Ap = __shfl_up_sync(0xffffffffu, A[u], d), Cp = __shfl_up_sync(0xffffffffu, C[u], d);
            if (lane >= d) { C[u] = fmaf(A[u], Cp, C[u]); A[u] *= Ap; }
          }
        }
        float next_carry = 0.f;
        if constexpr (ROWBLOCK) {
          // composite map of this warp's 128 frames, then the carry-in from the warps before it
          float Aw = 1.f, Cw = 0.f;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float At = __shfl_sync(0xffffffffu, A[u], 31), Ct = __shfl_sync(0xffffffffu, C[u], 31);
            Cw = fmaf(At, Cw, Ct); Aw *= At;
          }
          __syncthreads();                                    // previous super-step's maps consumed
          if (lane == 0) { s_A[warp] = Aw; s_C[warp] = Cw; }
          __syncthreads();
          float c_in = carry, c_all = carry;
#pragma unroll
          for (int k = 0; k < K2_WARPS; ++k) {
            c_all = fmaf(s_A[k], c_all, s_C[k]);
            if (k + 1 == warp) c_in = c_all;
          }
          next_carry = c_all;
          carry = c_in;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float m = fmaf(A[u], carry, C[u]);            // smoother state after this lane's frame
          carry = __shfl_sync(0xffffffffu, m, 31);
          const float dd = a.pcen_floor + m;
          const float uu = p[u] / pow_pos(dd, alpha) + delta; // postprocessing.py:66
          o[u] = pow_pos(uu, q) - dq;
        }
        if constexpr (ROWBLOCK) carry = next_carry;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int n = n0 + 32 * u + lane;
        if (ok[u]) {
          if (a.out_bf16) orow16[n - g.n_begin] = __float2bfloat16_rn(o[u]);
          else orow[n - g.n_begin] = o[u];
          if (prow) prow[n - g.n_begin] = p[u];
        }
      }
    }
    if (a.compression && a.ema_out != nullptr && lane == 0 && (!ROWBLOCK || warp == 0)) a.ema_out[(size_t)b * g.F + f] = carry;
  }
}

// Training forward: Q[kind-1][b][f][n] = sum over the tiles overlapping frame n of the partial sums of pooled quantity
// `kind` (1..3: Q_mu, Q_sigma, Q_poolw; k1_tc_kernel.cuh), in tile order.  One thread per output element, frames fastest:
// nothing but index arithmetic and <= ceil(K/TL)+1 loads, so the kernel lives on memory-level parallelism (full
// occupancy) -- as rows of the PCEN kernel (one warp per 100 frames, 59 registers) it took 0.44 ms at 1024 x 80 x 100.
__global__ void __launch_bounds__(256)
q_assemble_kernel(const Geom g, const float* __restrict__ ppart, float* __restrict__ q_out, int tl_shift,
                  unsigned long long hop_magic) {
  const int te_lo = (int)g.te_lo, te_hi = (int)g.te_hi;
  const int n4 = (g.n_count + 3) / 4;                         // groups of 4 consecutive frames per thread
  const long long rows = 3LL * g.B * g.F, total = rows * n4;
  const size_t per_kind = (size_t)g.B * g.F;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / n4;                            // (kind-1, b, f)
    const int q4 = (int)(idx - row * n4);
    const int kind = (int)(row / per_kind) + 1;
    const long long bf = row - (long long)(kind - 1) * per_kind;
    const int b = (int)(bf / g.F), f = (int)(bf - (long long)b * g.F);
    const float* pbase = ppart + ((size_t)b * 4 * g.F + (size_t)kind * g.F + f) * g.n_tiles * g.SL;
    float v[4][2];
    int extra_lo[4], extra_hi[4];                              // tiles beyond the first two of a frame (long windows only)
#pragma unroll
    for (int u = 0; u < 4; ++u) {                              // all loads of the 4 frames are issued before any is used
      const int n = g.n_begin + 4 * q4 + u;
      v[u][0] = v[u][1] = 0.f;
      extra_lo[u] = 1; extra_hi[u] = 0;
      if (n < g.n_begin + g.n_count) {
        int wlo = n * g.H - g.padL, whi = wlo + g.K - 1;
        if (wlo < te_lo) wlo = te_lo;
        if (whi > te_hi - 1) whi = te_hi - 1;
        const int i0 = (wlo - te_lo) >> tl_shift, i1 = (whi - te_lo) >> tl_shift;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int i = i0 + k;
          if (i <= i1) {
            const int num = te_lo + (i << tl_shift) + g.padL - g.K + 1;
            int nf = num <= 0 ? 0 : (int)div_magic((unsigned)(num + g.H - 1), hop_magic);
            if (nf < g.n_begin) nf = g.n_begin;
            v[u][k] = __ldg(pbase + (size_t)i * g.SL + (n - nf));
          }
        }
        extra_lo[u] = i0 + 2; extra_hi[u] = i1;
      }
    }
    float* qrow = q_out + (size_t)row * g.n_count + 4 * q4;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int n = g.n_begin + 4 * q4 + u;
      if (n >= g.n_begin + g.n_count) break;
      float sum = v[u][0] + v[u][1];                          // tile order
      for (int i = extra_lo[u]; i <= extra_hi[u]; ++i) {
        const int num = te_lo + (i << tl_shift) + g.padL - g.K + 1;
        int nf = num <= 0 ? 0 : (int)div_magic((unsigned)(num + g.H - 1), hop_magic);
        if (nf < g.n_begin) nf = g.n_begin;
        sum += __ldg(pbase + (size_t)i * g.SL + (n - nf));
      }
      qrow[u] = sum;
    }
  }
}

cudaError_t launch_k2(const Geom& g, const float* ppart, const PcenArgs& a, cudaStream_t stream) {
  int tl_shift = 0;
  while ((1 << tl_shift) < g.TL) ++tl_shift;
  if ((1 << tl_shift) != g.TL) return cudaErrorInvalidValue;   // tile lengths are powers of two
  const long long rows = (long long)g.B * g.F;
  // long rows: one block per row (the choice depends on the frame count only, so a clip's features do not depend on
  // the batch it is in); else one warp per row.  Blocks beyond the resident wave are scheduled as earlier ones
  // retire (rows past the grid limit loop inside the kernel)
  const bool rowblock = a.q_out == nullptr && g.n_count >= 512;
  long long blocks = rowblock ? rows : (rows + K2_WARPS - 1) / K2_WARPS;
  if (blocks > (1LL << 20)) blocks = 1LL << 20;
  const unsigned long long hop_magic = div_magic_of((unsigned)g.H);
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)blocks); lc.blockDim = dim3(K2_WARPS * 32); lc.dynamicSmemBytes = 0; lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at; lc.numAttrs = 1;
  cudaError_t err = rowblock ? cudaLaunchKernelEx(&lc, k2_pcen_kernel<true>, g, ppart, a, tl_shift, hop_magic)
                             : cudaLaunchKernelEx(&lc, k2_pcen_kernel<false>, g, ppart, a, tl_shift, hop_magic);
  if (err != cudaSuccess || a.q_out == nullptr) return err;
  // training forward: the three pooled bilinear forms, one thread per output element (after K2 in stream order, i.e.
  // after the whole Gabor grid)
  const long long total = 3LL * g.B * g.F * ((g.n_count + 3) / 4);    // 4 consecutive frames per thread
  long long qblocks = (total + 255) / 256;
  if (qblocks > (1LL << 22)) qblocks = 1LL << 22;
  q_assemble_kernel<<<(unsigned)qblocks, 256, 0, stream>>>(g, ppart, a.q_out, tl_shift, hop_magic);
  return cudaGetLastError();
}

}  // namespace leafk
