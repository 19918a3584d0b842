// Input-side helpers of the frontend: what the reference does per clip on the CPU before the batch
// reaches Leaf.forward -- crop / pad to a fixed length and peak normalisation -- expressed as per-clip (start, length,
// divisor) triples that the Gabor kernels apply while they stage the waveform (leafk_common.cuh: clip_sample), so the
// prepared batch is never written to memory.
//
// Replaces   PadToSize('wrap') / CenterCrop / RandomCrop      reference utilities/data/raw_transforms.py:121-160
//            zero padding to the longest clip (collate)        reference utilities/data/utils.py:8-28
//            PeakNormalization(apply_to="only_too_loud_sounds") reference utilities/data/raw_transforms.py:334-344
//            (torch_audiomentations: divide a clip by max|x| when max|x| > 1)
#include "../../include/leafk.h"
#include "leafk_common.cuh"

#include <cstring>

namespace leafk {

int fail(int code, const char* fmt, ...);
void count_launch(int n);

// one block per clip: divisor[b] = max_i |prepared sample i| if it exceeds 1 (only_too_loud) / if it is > 0, else 1.
// A clip holding a NaN is left alone (torch.max gives NaN, and NaN > 1 is false).
__global__ void __launch_bounds__(256)
peak_divisor_kernel(const Geom g, const float* __restrict__ x, int only_too_loud, float* __restrict__ div_out) {
  __shared__ float red[8];
  __shared__ int red_nan[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ClipView cv = clip_view(g, b);
  float mx = 0.f;
  int has_nan = 0;
  for (long long i = tid; i < g.T_total; i += blockDim.x) {
    const float v = clip_sample(g, x, cv, i);
    has_nan |= (v != v);
    mx = fmaxf(mx, fabsf(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o);
  }
  if (lane == 0) { red[warp] = mx; red_nan[warp] = has_nan; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) { mx = fmaxf(mx, red[w]); has_nan |= red_nan[w]; }
    float d = 1.0f;
    if (!has_nan && (only_too_loud ? mx > 1.0f : mx > 0.f)) d = mx;
    div_out[b] = d;
  }
}

// one block per clip: smallest raw sample over [0, length); NaN if the clip holds one (torch.min propagates it)
__global__ void __launch_bounds__(256)
clip_minimum_kernel(const Geom g, const float* __restrict__ x, float* __restrict__ min_out) {
  __shared__ float red[8];
  __shared__ int red_nan[8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ClipView cv = clip_view(g, b);
  float mn = 3.0e38f;
  int has_nan = 0;
  for (int i = tid; i < cv.len; i += blockDim.x) {
    const float v = load_sample(x, cv.row, i, g.x_fmt);
    has_nan |= (v != v);
    mn = fminf(mn, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    has_nan |= __shfl_xor_sync(0xffffffffu, has_nan, o);
  }
  if (lane == 0) { red[warp] = mn; red_nan[warp] = has_nan; }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) { mn = fminf(mn, red[w]); has_nan |= red_nan[w]; }
    min_out[b] = has_nan ? __int_as_float(0x7fc00000) : (cv.len > 0 ? mn : 0.f);
  }
}

void geom_apply_prep(const leafk_config* cfg, Geom* g) {
  const leafk_clip_prep* p = cfg->prep;
  if (p == nullptr) return;
  g->clip_start = p->start; g->clip_len = p->length; g->clip_div = p->divisor;
  g->clip_padval = p->pad_value; g->clip_pad = p->pad_mode;
  if (p->ld > 0) g->ldx = p->ld;
}

}  // namespace leafk

using namespace leafk;

extern "C" int leafk_clip_minimum(const leafk_config* cfg, const float* x, int B, int T, float* minimum_out, void* stream) {
  if (!cfg || !x || !minimum_out) return fail(LEAFK_EINVAL, "null pointer argument");
  if (B < 1 || T < 1) return fail(LEAFK_EINVAL, "bad B/T (%d,%d)", B, T);
  Geom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.T_total = T; g.T_win = T; g.t_off = 0; g.ldx = T;
  g.x_fmt = cfg->input_format == LEAFK_INPUT_S16 ? 1 : 0;
  geom_apply_prep(cfg, &g);
  clip_minimum_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(g, x, minimum_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "clip_minimum launch: %s", cudaGetErrorString(e));
  count_launch(1);
  return LEAFK_OK;
}

extern "C" int leafk_peak_divisors(const leafk_config* cfg, const float* x, int B, int T, int only_too_loud,
                                   float* divisor_out, void* stream) {
  if (!cfg || !x || !divisor_out) return fail(LEAFK_EINVAL, "null pointer argument");
  if (B < 1 || T < 1) return fail(LEAFK_EINVAL, "bad B/T (%d,%d)", B, T);
  if (cfg->input_format != LEAFK_INPUT_F32 && cfg->input_format != LEAFK_INPUT_S16)
    return fail(LEAFK_EINVAL, "unknown input_format %d", cfg->input_format);
  Geom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.T_total = T; g.T_win = T; g.t_off = 0; g.ldx = T;
  g.x_fmt = cfg->input_format == LEAFK_INPUT_S16 ? 1 : 0;
  geom_apply_prep(cfg, &g);
  g.clip_div = nullptr;                                  // the divisors are what is being computed
  peak_divisor_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(g, x, only_too_loud, divisor_out);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "peak_divisor launch: %s", cudaGetErrorString(e));
  count_launch(1);
  return LEAFK_OK;
}
