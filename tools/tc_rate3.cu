// tc_rate3: issue/execute rate of cta_group::2 (CTA-pair, M = 256) tcgen05.mma pairs as a function of N:
// how much a support-pruned k-step (N = 2*na main + na corr) really costs.  Operands are zeros; only timing.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include "../leaf_pytorch_b200/csrc/tc_ptx.cuh"
using namespace leafk::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int KSTEPS = 26, KP = 416, XS_LEN = 1440;
__device__ int g_timeout = 0;

// per k-step: MMA(N1, A = copy 0) [+ MMA(N2, A = copy 1)] [+ MMA(N3, A = copy 0)]
template <int N1, int N2, int N3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128) pair_rate(int reps, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  constexpr int R1 = 128 * KP * 2;                     // up to 128 B rows per CTA
  for (int i = tid; i < (8192 + R1) / 4; i += 128) ((uint32_t*)smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc_pair<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  long long t0 = 0;
  if (warp == 0 && rank == 0) {
    const bool leader = elect_one();
    const uint64_t a_hi = smem_desc(smem_u32(smem), 16, 128), a_lo = smem_desc(smem_u32(smem) + XS_LEN * 2, 16, 128);
    const uint64_t b1 = smem_desc(smem_u32(smem + 8192), 128 * 16, 128);
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
          mma_f16_ss_pair(tmem, a_hi + 2 * ks, b1 + (uint64_t)(ks * (128 * 32 / 16)), idesc_f16(256, N1), ks > 0);
          if (N2) mma_f16_ss_pair(tmem, a_lo + 2 * ks, b1 + (uint64_t)(ks * (128 * 32 / 16)), idesc_f16(256, N2 ? N2 : 16), 1);
          if (N3) mma_f16_ss_pair(tmem + 256, a_hi + 2 * ks, b1 + (uint64_t)(ks * (128 * 32 / 16)), idesc_f16(256, N3 ? N3 : 16), ks > 0);
        }
      }
      __syncwarp();
    }
    if (leader) mma_commit_pair(&bar);
    __syncwarp();
  }
  bool ok = false;
  for (long long i = 0; i < 20000000LL && !ok; ++i) ok = mbar_try_wait(&bar, 0);
  if (!ok) g_timeout = 1;
  if (warp == 0 && rank == 0 && tid == 0) cycles[blockIdx.x / 2] = clock64() - t0;
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair<512>(tmem);
}

template <int N1, int N2, int N3>
void run(int grid, long long* dc) {
  constexpr int smem = 8192 + 128 * KP * 2 + 1024;
  CK(cudaFuncSetAttribute(pair_rate<N1, N2, N3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int reps = 64;
  pair_rate<N1, N2, N3><<<grid, 128, smem>>>(reps, dc);
  CK(cudaDeviceSynchronize());
  pair_rate<N1, N2, N3><<<grid, 128, smem>>>(reps, dc);
  CK(cudaDeviceSynchronize());
  long long cyc;
  CK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
  const double per = cyc / (double)(reps * KSTEPS);
  printf("N = %3d + %3d + %3d : %6.1f cycles per k-step   (math floor %5.1f)\n", N1, N2, N3, per, (N1 + N2 + N3) / 2.0);
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int grid = prop.multiProcessorCount / 2 * 2;
  long long* dc;
  CK(cudaMalloc(&dc, 8 * 128));
  printf("single MMA per k-step (pair mode, M = 256)\n");
  run<16, 0, 0>(grid, dc); run<32, 0, 0>(grid, dc); run<64, 0, 0>(grid, dc); run<96, 0, 0>(grid, dc);
  run<128, 0, 0>(grid, dc); run<160, 0, 0>(grid, dc); run<192, 0, 0>(grid, dc); run<240, 0, 0>(grid, dc); run<256, 0, 0>(grid, dc);
  printf("k-step of the kernel at na active channels: N = 2 na (A hi) + na (A lo)\n");
  run<32, 16, 0>(grid, dc); run<64, 32, 0>(grid, dc); run<96, 48, 0>(grid, dc); run<128, 64, 0>(grid, dc); run<160, 80, 0>(grid, dc);
  printf("same with an independent accumulator chain interleaved (N3, A hi)\n");
  run<64, 32, 64>(grid, dc); run<96, 48, 96>(grid, dc);
  int tf;
  CK(cudaMemcpyFromSymbol(&tf, g_timeout, sizeof(int)));
  printf("tc_rate3: %s\n", tf ? "TIMEOUT" : "done");
  return 0;
}
