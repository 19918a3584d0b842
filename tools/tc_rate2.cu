// tc_rate2: tight (fully unrolled, uniform-datapath) tcgen05.mma issue loop, as the real kernel
// will use, to separate tensor-pipe execution time from issue overhead.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../leaf_pytorch_b200/csrc/tc_ptx.cuh"
using namespace leafk::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ int g_timeout = 0;

// MODE 0: N1 (A0) + N2 (A1) same accumulator    MODE 1: two phases interleaved k-step by k-step
template <int N1, int N2, int MODE>
__global__ void __launch_bounds__(128) rate_kernel(int phases, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  constexpr int NB = N1;
  constexpr int SMEM = 16384 + 26 * NB * 32;
  for (int i = tid; i < SMEM / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 16384);
  if (warp == 0) {
    const bool leader = elect_one();
    constexpr uint32_t i1 = idesc_f16(128, N1);
    constexpr uint32_t i2 = idesc_f16(128, N2 ? N2 : 16);
    const uint64_t a_hi = smem_desc(sA, 16, 128), a_lo = smem_desc(sA + 2944, 16, 128);
    const uint64_t a_hi2 = smem_desc(sA + 2 * 2944, 16, 128), a_lo2 = smem_desc(sA + 3 * 2944, 16, 128);
    const uint64_t b0 = smem_desc(sB, NB * 16, 128);
    uint32_t phase[2] = {0, 0};
    const long long t0 = clock64();
    for (int ph = 0; ph < phases; ++ph) {
      const int s = ph & 1;
      if (ph >= 2) {
        bool ok = false;
        for (long long i = 0; i < 20000000LL && !ok; ++i) ok = mbar_try_wait(&bar[s], phase[s]);
        if (!ok) { g_timeout = 1; break; }
        phase[s] ^= 1;
      }
      const uint32_t d = tmem + s * 256;
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < 26; ++ks) {
          const uint64_t b = b0 + (uint64_t)(ks * (NB * 32 / 16));
          if (MODE == 0) {
            mma_f16_ss(d, a_hi + 2 * ks, b, i1, ks > 0);
            if (N2) mma_f16_ss(d, a_lo + 2 * ks, b, i2, 1);
          } else {           // two independent accumulators (two phases in flight), 128 columns apart
            mma_f16_ss(d, a_hi + 2 * ks, b, i1, ks > 0);
            mma_f16_ss(d + 128, a_hi2 + 2 * ks, b, i1, ks > 0);
            if (N2) { mma_f16_ss(d, a_lo + 2 * ks, b, i2, 1); mma_f16_ss(d + 128, a_lo2 + 2 * ks, b, i2, 1); }
          }
        }
        mma_commit(&bar[s]);
      }
      __syncwarp();
    }
    for (int s = 0; s < 2; ++s) {
      bool ok = false;
      for (long long i = 0; i < 20000000LL && !ok; ++i) ok = mbar_try_wait(&bar[s], phase[s]);
      if (!ok) g_timeout = 1;
    }
    if (leader) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

template <int N1, int N2, int MODE>
void run(const char* name, int nsm, long long* dcyc) {
  constexpr int SMEM = 16384 + 26 * N1 * 32;
  CK(cudaFuncSetAttribute(rate_kernel<N1, N2, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  const int phases = 64;
  rate_kernel<N1, N2, MODE><<<nsm, 128, SMEM>>>(phases, dcyc);
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate_kernel<N1, N2, MODE><<<nsm, 128, SMEM>>>(phases, dcyc);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> cyc(nsm);
  CK(cudaMemcpy(cyc.data(), dcyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
  long long mx = 0;
  for (auto v : cyc) mx = v > mx ? v : mx;
  const double ksteps = 26.0 * phases;
  const double cols = (N1 + N2) * (MODE == 1 ? 2 : 1);
  printf("%-44s %7.1f cyc/k-step  math floor %5.1f  -> %5.1f%% of tensor peak   %.3f ms  clk %.0f MHz\n", name, mx / ksteps,
         cols / 2.0, 100.0 * (cols / 2.0) / (mx / ksteps), ms, mx / (ms * 1e3));
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  long long* dcyc;
  CK(cudaMalloc(&dcyc, sizeof(long long) * nsm));
  run<16, 0, 0>("N=16 (issue floor, 1 MMA/k-step)", nsm, dcyc);
  run<16, 16, 0>("N=16+16 (issue floor, 2 MMA/k-step)", nsm, dcyc);
  run<80, 0, 0>("N=80", nsm, dcyc);
  run<160, 0, 0>("N=160", nsm, dcyc);
  run<256, 0, 0>("N=256", nsm, dcyc);
  run<160, 80, 0>("N=160(Ahi)+80(Alo)  [real k-step]", nsm, dcyc);
  run<128, 64, 0>("N=128(Ahi)+64(Alo)  [F=32 / F=64 split]", nsm, dcyc);
  run<256, 128, 0>("N=256(Ahi)+128(Alo) [F=64]", nsm, dcyc);
  run<80, 48, 1>("2 phases interleaved: N=80+48 each (toy)", nsm, dcyc);
  run<128, 0, 1>("2 independent chains N=128", nsm, dcyc);
  run<96, 48, 1>("2 phases interleaved: N=96+48 each", nsm, dcyc);
  int tf = 0;
  CK(cudaMemcpyFromSymbol(&tf, g_timeout, sizeof(int)));
  printf("tc_rate2: %s\n", tf ? "TIMEOUT" : "done");
  return 0;
}
