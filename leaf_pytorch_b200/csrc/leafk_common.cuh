// Shared geometry / workspace layout for the LEAF frontend kernels (sm_100a).
//
// Path being replaced: leaf_pytorch.frontend.Leaf.forward, reference frontend.py:78-89.
// Index conventions follow SURVEY.md Appendix A.1:
//   y[c,t] = sum_k W[c,k] * x~[t - padL + k]          (convolution.py:92,97; cross-correlation)
//   e[f,t] = y[2f,t]^2 + y[2f+1,t]^2                  (frontend.py:15-19)
//   p[f,n] = bias_f + sum_k g_f[k] * e~[f, n*H - padL + k]   (pooling.py:33-41), e~ = 0 outside [0,T)
//
// Work decomposition ("tiles"): the e-sample axis of every clip is cut into disjoint tiles of TL
// samples.  A tile owns no frame: it emits the partial pooled sums of every frame whose window
// overlaps it (<= SL of them, "slots"), and the PCEN kernel adds the <= ceil(K/TL)+1 partials of a
// frame in tile order (deterministic, no atomics, no recomputed halo).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace leafk {

struct Geom {
  int B, F, K, H;
  int padL, padR;
  int C2;             // 2F
  int C2p;            // 2F rounded up to 8 (fp32 path channel padding)
  int Kp;             // K rounded up to 16 (zero taps)
  long long T_total;  // clip length in samples
  long long t_off;    // first sample held by the x window
  int T_win;          // samples in the window
  long long ldx;      // row stride of the x window
  int n_begin, n_count, N_total;
  long long te_lo, te_hi;  // e-samples [te_lo, te_hi) are needed by frames [n_begin, n_begin+n_count)
  int TL;                  // tile length in e-samples
  int n_tiles;             // tiles per clip
  int SL;                  // slots (frames) a tile can touch
  int x_fmt;               // 0: float32 samples, 1: int16 PCM (value = s / 32768)
  // Optional on-the-fly clip preparation (reference data pipeline: PadToSize / CenterCrop / RandomCrop / the zero
  // padding of the collate function / PeakNormalization, utilities/data/raw_transforms.py:121-140,143-160,334-344 and
  // utilities/data/utils.py:8-28).  Sample i of prepared clip b is raw[(i + clip_start[b])] of its row, where indices
  // outside [0, clip_len[b]) wrap around (clip_wrap) or read as zero, divided by clip_div[b].  All null: rows are
  // used as they are.
  const int* clip_start;
  const int* clip_len;
  const float* clip_div;
  const float* clip_padval;
  int clip_pad;            // LEAFK_PAD_*: 0 zero, 1 wrap, 2 edge (replicate), 3 per-clip value
};

// What a kernel needs to read clip b: row offset, crop start, raw length, divisor.
struct ClipView {
  size_t row;
  int start, len;          // clip lengths are below 2^30 (checked on the host), crop offsets within +-2^29
  float div, padval;
  bool prep;
};

// sample i of a clip row (row = first sample of the clip in the window buffer)
__device__ __forceinline__ float load_sample(const float* base, size_t row_elems, long long i, int fmt) {
  if (fmt == 0) return base[row_elems + i];
  return (float)reinterpret_cast<const short*>(base)[row_elems + i] * (1.0f / 32768.0f);
}

#ifdef __CUDACC__
__device__ __forceinline__ ClipView clip_view(const Geom& g, int b) {
  ClipView v;
  v.row = (size_t)b * g.ldx;
  v.prep = g.clip_start != nullptr || g.clip_len != nullptr || g.clip_div != nullptr;
  v.start = g.clip_start ? g.clip_start[b] : 0;
  v.len = g.clip_len ? g.clip_len[b] : (int)g.T_total;
  v.div = g.clip_div ? g.clip_div[b] : 1.0f;
  v.padval = g.clip_padval ? g.clip_padval[b] : 0.0f;
  return v;
}
// sample wi of the window (= sample t_off + wi of the prepared clip); the caller has checked 0 <= wi < T_win and that
// the sample lies inside [0, T_total)
__device__ __forceinline__ float clip_sample(const Geom& g, const float* x, const ClipView& v, long long wi) {
  if (!v.prep) return load_sample(x, v.row, wi, g.x_fmt);
  int j = (int)(g.t_off + wi) + v.start;
  if (j < 0 || j >= v.len) {
    if (g.clip_pad == 3) return v.padval / v.div;
    if (g.clip_pad == 0 || v.len <= 0) return 0.f;
    if (g.clip_pad == 2) {
      j = j < 0 ? 0 : v.len - 1;
    } else {
      j %= v.len;
      if (j < 0) j += v.len;
    }
  }
  return load_sample(x, v.row, j, g.x_fmt) / v.div;
}
#endif

__host__ __device__ inline long long floordiv_ll(long long a, long long b) {
  return (a >= 0) ? a / b : -((-a + b - 1) / b);
}
__host__ __device__ inline long long ceildiv_ll(long long a, long long b) {
  return floordiv_ll(a + b - 1, b);
}

// first frame whose pooling window reaches sample ts or later (clamped to n_begin)
__host__ __device__ inline int first_frame_of(const Geom& g, long long ts) {
  long long n = ceildiv_ll(ts + g.padL - g.K + 1, g.H);
  return (int)(n < g.n_begin ? g.n_begin : n);
}
// last frame whose pooling window starts at or before sample tl (clamped to the produced range)
__host__ __device__ inline int last_frame_of(const Geom& g, long long tl) {
  long long n = floordiv_ll(tl + g.padL, g.H);
  long long hi = (long long)g.n_begin + g.n_count - 1;
  return (int)(n > hi ? hi : n);
}

// q = a / d for 0 <= a < 2^32 with the round-up magic number m = floor(2^64 / d) + 1 (exact for every 32-bit a;
// m = 0 encodes d = 1).  A hardware-less integer division by a runtime value costs ~25 instructions.
__host__ __device__ inline unsigned long long div_magic_of(unsigned d) { return d <= 1u ? 0ULL : ~0ULL / (unsigned long long)d + 1ULL; }
#ifdef __CUDACC__
__device__ __forceinline__ unsigned div_magic(unsigned a, unsigned long long m) {
  return m == 0ULL ? a : (unsigned)__umul64hi((unsigned long long)a, m);
}
#endif

// ---- workspace layout (all offsets in bytes, 256-aligned) -------------------------------------
struct Workspace {
  size_t off_cprm;   // float[F*8]   constrained parameters + derived per-filter constants
  size_t off_w32;    // float[Kp*C2p] Gabor bank, tap-major ([k][c]), zero padded
  size_t off_g32;    // float[K*F]   Gaussian pooling windows, tap-major ([k][f])
  size_t off_w16;    // uint8[...]   fp16 hi/lo bank in UMMA smem layout (tensor-core path)
  size_t off_tcmap;  // int[Fp] sorted position -> filter, then int[groups][16] zone table (k1_tc_layout.cuh)
  size_t off_ppart;  // float[B][F][n_tiles][SL] partial pooled sums (slot fastest; a (clip, filter) row's partials contiguous)
  size_t off_flags;  // int[64]      slice-ready flags of the host-pipelined forward
  size_t off_done;   // int[B]       per clip: epilogue warps of K1 that have stored a tile of it (K2 starts a clip at
                     //              n_tiles * n_groups * 8, see k2_pcen.cu)
  size_t total;
};

// per-filter constants produced by k0 (float[8] per filter)
enum { CP_MU = 0, CP_SIGMA = 1, CP_NORM = 2, CP_INV2S2 = 3, CP_POOLS = 4, CP_POOLA = 5, CP_WSCALE = 6, CP_PAD = 7 };

// asynchronous error codes (error word in the workspace, read back by leafk_async_status)
constexpr int LEAFK_ASYNC_H2D_TIMEOUT = 1;      // a slice-ready flag of the host-pipelined forward never arrived
constexpr int LEAFK_ASYNC_K1_TIMEOUT = 2;       // K2 gave up waiting for K1's per-clip completion counter
constexpr long long H2D_TIMEOUT_NS = 10LL * 1000 * 1000 * 1000;
#ifdef __CUDACC__
// torch.clamp / torch.minimum / torch.maximum propagate NaN (fminf / fmaxf drop it): a NaN parameter or waveform must
// come out as NaN features like in the reference, not as a silently clamped value
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) { return (v != v) ? v : fminf(fmaxf(v, lo), hi); }
__device__ __forceinline__ float max_nan(float v, float c) { return (v != v) ? v : fmaxf(v, c); }
__device__ __forceinline__ float min_nan(float v, float c) { return (v != v) ? v : fminf(v, c); }
__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace leafk
