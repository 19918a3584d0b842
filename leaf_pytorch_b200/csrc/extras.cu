// The two optional stages the reference's constructor declares but never implemented:
//   preemp=True          reference frontend.py:40-41   ("Pre-emp functionality not implemented yet..")
//   mean_var_norm=True   reference frontend.py:62-63   ("Instance Norm functionality not added yet..")
// Semantics follow the original LEAF (google-research/leaf-audio, leaf_audio/frontend.py): the pre-emphasis is a
// learnable 2-tap 'same' correlation in front of the Gabor bank, initialised to (-0.97, 1):
//      xp[t] = w0 x[t] + w1 x[t+1]          (x[T] = 0; 'same' padding of a 2-tap kernel is (0, 1), utils.py:5-10)
// and the mean/variance normalisation is an instance norm over the frames of every (clip, filter) row without
// affine parameters:  o[n] = (v[n] - mean_n v) / sqrt(var_n v + 1e-5)   (biased variance, torch.nn.InstanceNorm1d).
// Both run as small stand-alone kernels around the fused frontend (forward and backward each); the fused kernels
// supply the gradient w.r.t. their input waveform for the pre-emphasis.
#include "../../include/leafk.h"
#include "leafk_common.cuh"

namespace leafk {
int fail(int code, const char* fmt, ...);
void count_launch(int n);

__global__ void preemp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, long long B, int T,
                                  float* __restrict__ y) {
  const float w0 = w[0], w1 = w[1];
  const long long total = B * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const float nx = (t + 1 < T) ? x[i + 1] : 0.f;
    y[i] = fmaf(w1, nx, w0 * x[i]);
  }
}

// dx[t] = w0 g[t] + w1 g[t-1];  per-block partial sums of dw0 = sum g[t] x[t], dw1 = sum g[t] x[t+1]
__global__ void __launch_bounds__(256)
preemp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ g, long long B,
                  int T, float* __restrict__ dx, float* __restrict__ part) {
  __shared__ float red[2][8];
  const float w0 = w[0], w1 = w[1];
  const long long total = B * T;
  float s0 = 0.f, s1 = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    const float gi = g[i], xi = x[i];
    const float nx = (t + 1 < T) ? x[i + 1] : 0.f;
    const float pg = (t > 0) ? g[i - 1] : 0.f;
    if (dx != nullptr) dx[i] = fmaf(w1, pg, w0 * gi);
    s0 = fmaf(gi, xi, s0);
    s1 = fmaf(gi, nx, s1);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  if (lane == 0) { red[0][warp] = s0; red[1][warp] = s1; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += red[threadIdx.x][k];
    part[2 * blockIdx.x + threadIdx.x] = s;
  }
}
__global__ void preemp_bwd_finish_kernel(const float* __restrict__ part, int n, float* __restrict__ dw) {
  if (threadIdx.x < 2) {
    float s = 0.f;
    for (int k = 0; k < n; ++k) s += part[2 * k + threadIdx.x];       // fixed order: deterministic
    dw[threadIdx.x] = s;
  }
}

// one warp per (clip, filter) row of N frames
__global__ void __launch_bounds__(256)
instnorm_fwd_kernel(const float* __restrict__ v, long long rows, int N, float eps, float* __restrict__ o,
                    float* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = v + row * N;
  float s = 0.f;
  for (int n = lane; n < N; n += 32) s += p[n];
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) s += __shfl_xor_sync(0xffffffffu, s, k);
  const float mean = s / (float)N;
  float q = 0.f;
  for (int n = lane; n < N; n += 32) { const float d = p[n] - mean; q = fmaf(d, d, q); }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) q += __shfl_xor_sync(0xffffffffu, q, k);
  const float rstd = rsqrtf(q / (float)N + eps);
  for (int n = lane; n < N; n += 32) o[row * N + n] = (p[n] - mean) * rstd;
  if (lane == 0) { stats[2 * row] = mean; stats[2 * row + 1] = rstd; }
}
// dv = rstd * (g - mean(g) - o * mean(g o)),  o = (v - mean) * rstd
__global__ void __launch_bounds__(256)
instnorm_bwd_kernel(const float* __restrict__ v, const float* __restrict__ stats, const float* __restrict__ g,
                    long long rows, int N, float* __restrict__ dv) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float mean = stats[2 * row], rstd = stats[2 * row + 1];
  const float* p = v + row * N;
  const float* gr = g + row * N;
  float sg = 0.f, sgo = 0.f;
  for (int n = lane; n < N; n += 32) { const float o = (p[n] - mean) * rstd; sg += gr[n]; sgo = fmaf(gr[n], o, sgo); }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) { sg += __shfl_xor_sync(0xffffffffu, sg, k); sgo += __shfl_xor_sync(0xffffffffu, sgo, k); }
  const float mg = sg / (float)N, mgo = sgo / (float)N;
  for (int n = lane; n < N; n += 32) {
    const float o = (p[n] - mean) * rstd;
    dv[row * N + n] = rstd * (gr[n] - mg - o * mgo);
  }
}
}  // namespace leafk

using namespace leafk;

extern "C" {

int leafk_preemp_forward(const float* x, const float* w2, int B, int T, float* y, void* stream) {
  if (!x || !w2 || !y) return fail(LEAFK_EINVAL, "null pointer argument");
  if (B < 1 || T < 1) return fail(LEAFK_EINVAL, "bad B/T (%d,%d)", B, T);
  const long long total = (long long)B * T;
  const int blocks = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
  preemp_fwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, w2, B, T, y);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "preemp_fwd launch: %s", cudaGetErrorString(e));
  count_launch(1);
  return LEAFK_OK;
}

size_t leafk_preemp_backward_workspace_bytes(void) { return sizeof(float) * 2 * 1184; }

int leafk_preemp_backward(const float* x, const float* w2, const float* grad_y, int B, int T, float* grad_x,
                          float* grad_w2, void* workspace, size_t workspace_bytes, void* stream) {
  if (!x || !w2 || !grad_y || !grad_w2 || !workspace) return fail(LEAFK_EINVAL, "null pointer argument");
  if (B < 1 || T < 1) return fail(LEAFK_EINVAL, "bad B/T (%d,%d)", B, T);
  if (workspace_bytes < leafk_preemp_backward_workspace_bytes()) return fail(LEAFK_EWORKSPACE, "workspace too small");
  const long long total = (long long)B * T;
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  preemp_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, w2, grad_y, B, T, grad_x, (float*)workspace);
  preemp_bwd_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const float*)workspace, blocks, grad_w2);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "preemp_bwd launch: %s", cudaGetErrorString(e));
  count_launch(2);
  return LEAFK_OK;
}

int leafk_instnorm_forward(const float* v, long long rows, int N, float eps, float* out, float* stats, void* stream) {
  if (!v || !out || !stats) return fail(LEAFK_EINVAL, "null pointer argument");
  if (rows < 1 || N < 1) return fail(LEAFK_EINVAL, "bad shape");
  instnorm_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(v, rows, N, eps, out, stats);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "instnorm_fwd launch: %s", cudaGetErrorString(e));
  count_launch(1);
  return LEAFK_OK;
}

int leafk_instnorm_backward(const float* v, const float* stats, const float* grad_out, long long rows, int N,
                            float* grad_v, void* stream) {
  if (!v || !stats || !grad_out || !grad_v) return fail(LEAFK_EINVAL, "null pointer argument");
  if (rows < 1 || N < 1) return fail(LEAFK_EINVAL, "bad shape");
  instnorm_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(v, stats, grad_out, rows, N, grad_v);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "instnorm_bwd launch: %s", cudaGetErrorString(e));
  count_launch(1);
  return LEAFK_OK;
}

}  // extern "C"
