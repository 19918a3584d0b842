// tc_rate: issue / shared-memory-operand rate of tcgen05.mma kind::f16 (M=128) for the k-step
// patterns the Gabor kernel could use.  One CTA per SM, operands resident in shared memory, real
// operand sizes (Toeplitz A copies of 2880 B, B bank 26 k-steps x NB rows).  Values are zero.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../leaf_pytorch_b200/csrc/tc_ptx.cuh"
using namespace leafk::ptx;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

struct Pat {
  int n1, n2, n3;     // N of up to three MMAs per k-step (0 = absent)
  int a1, a2, a3;     // which A copy each uses (0..3)
  int d1, d2, d3;     // accumulator column offset of each
  int canonical;      // 1: A from a canonical (non-overlapping) 4 KB tile per k-step parity
  int phases;         // phases (of 26 k-steps) per launch
  int acc_alt;        // 1: odd k-steps use accumulator +256 columns... (independent chains)
  int b_fixed;        // 1: same B slice every k-step
  int a_fixed;        // 1: same A slice every k-step
  int tmem_half;      // 1: allocate/use 256 columns only (2 CTAs per SM)
};

__device__ int g_timeout = 0;

__global__ void __launch_bounds__(128) rate_kernel(Pat p, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 8 * 256 * 32) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init_fence(); }
  if (warp == 0) { if (p.tmem_half) tmem_alloc<256>(&tmem_base_s); else tmem_alloc<512>(&tmem_base_s); }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t sA = smem_u32(smem), sB = smem_u32(smem + 16384);
  if (warp == 0) {
    const bool leader = elect_one();
    const int NB = p.n1;                                  // B rows (layout stride)
    const uint32_t i1 = idesc_f16(128, p.n1), i2 = p.n2 ? idesc_f16(128, p.n2) : 0, i3 = p.n3 ? idesc_f16(128, p.n3) : 0;
    const uint64_t a_step = p.canonical ? 0 : 2;          // +32 B per k-step in the start-address field
    uint64_t A[4];
    for (int c = 0; c < 4; ++c)
      A[c] = p.canonical ? smem_desc(sA + c * 4096, 2048, 128) : smem_desc(sA + c * 2944, 16, 128);
    const uint64_t b0 = smem_desc(sB, NB * 16, 128);
    const uint64_t b_step = (uint64_t)(NB * 32) >> 4;
    uint32_t phase[2] = {0, 0};
    const long long t0 = clock64();
    for (int ph = 0; ph < p.phases; ++ph) {
      const int s = ph & 1;
      if (ph >= 2) {
        bool ok = false;
        for (long long i = 0; i < 20000000LL && !ok; ++i) ok = mbar_try_wait(&bar[s], phase[s]);
        if (!ok) { g_timeout = 1; break; }
        phase[s] ^= 1;
      }
      const uint32_t d = tmem + (p.tmem_half ? 0 : s * 256);
      uint64_t a1 = A[p.a1], a2 = A[p.a2], a3 = A[p.a3], b = b0;
#pragma unroll 2
      for (int ks = 0; ks < 26; ++ks) {
        const uint32_t dd = d + ((p.acc_alt && (ks & 1)) ? 128 : 0);
        const uint64_t bb = p.b_fixed ? b0 : b0 + (uint64_t)(ks & 7) * b_step;
        if (leader) {
          mma_f16_ss(dd + p.d1, a1, bb, i1, ks > (p.acc_alt ? 1 : 0));
          if (p.n2) mma_f16_ss(dd + p.d2, a2, bb, i2, (ks > (p.acc_alt ? 1 : 0)) | (p.d2 == p.d1));
          if (p.n3) mma_f16_ss(dd + p.d3, a3, bb, i3, 1);
        }
        if (!p.a_fixed) { a1 += a_step; a2 += a_step; a3 += a_step; }
      }
      if (leader) mma_commit(&bar[s]);
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
  if (warp == 0) { if (p.tmem_half) tmem_dealloc<256>(tmem); else tmem_dealloc<512>(tmem); }
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  long long* dcyc;
  CK(cudaMalloc(&dcyc, sizeof(long long) * nsm * 2));
  const int smem = 16384 + 8 * 256 * 32;
  CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  struct Named { const char* name; Pat p; };
  const Named tests[] = {
      //                                            n1  n2  n3 a1 a2 a3 d1 d2 d3 can ph alt bfx afx half
      {"N=16 (issue)                          ", {16, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 0, 0, 0, 0}},
      {"N=80                                  ", {80, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 0, 0, 0, 0}},
      {"N=160                                 ", {160, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 0, 0, 0, 0}},
      {"N=256                                 ", {256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 0, 0, 0, 0}},
      {"N=160+80 (real)                       ", {160, 80, 0, 0, 1, 0, 0, 0, 0, 0, 64, 0, 0, 0, 0}},
      {"N=160+80 B fixed                      ", {160, 80, 0, 0, 1, 0, 0, 0, 0, 0, 64, 0, 1, 0, 0}},
      {"N=160+80 A fixed                      ", {160, 80, 0, 0, 1, 0, 0, 0, 0, 0, 64, 0, 0, 1, 0}},
      {"N=160+80 A,B fixed                    ", {160, 80, 0, 0, 1, 0, 0, 0, 0, 0, 64, 0, 1, 1, 0}},
      {"N=128 acc alternating                 ", {128, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 1, 0, 0, 0}},
      {"N=128 acc same                        ", {128, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 0, 0, 0, 0}},
      {"N=80+48 acc alternating               ", {80, 48, 0, 0, 1, 0, 0, 0, 0, 0, 64, 1, 0, 0, 0}},
      {"N=160+80 half TMEM, 1 CTA/SM          ", {160, 80, 0, 0, 1, 0, 0, 0, 0, 0, 64, 0, 0, 0, 1}},
      {"N=160+80 half TMEM, 2 CTA/SM          ", {160, 80, 0, 0, 1, 0, 0, 0, 0, 0, 64, 0, 0, 0, 2}},
      {"N=256 half TMEM, 2 CTA/SM             ", {256, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 0, 0, 0, 2}},
      {"N=80 half TMEM, 2 CTA/SM              ", {80, 0, 0, 0, 0, 0, 0, 0, 0, 0, 64, 0, 0, 0, 2}},
  };

  for (const Named& t : tests) {
    for (int grid : {t.p.tmem_half == 2 ? 2 * nsm : nsm}) {
      rate_kernel<<<grid, 128, smem>>>(t.p, dcyc);
      CK(cudaDeviceSynchronize());
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      rate_kernel<<<grid, 128, smem>>>(t.p, dcyc);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> cyc(grid);
      CK(cudaMemcpy(cyc.data(), dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
      long long mx = 0;
      for (auto v : cyc) mx = v > mx ? v : mx;
      const double ksteps = 26.0 * t.p.phases;
      const double cols = t.p.n1 + t.p.n2 + t.p.n3;
      const double bytes = 4096.0 * (1 + (t.p.n2 > 0) + (t.p.n3 > 0)) + 32.0 * cols;
      printf("%s grid=%3d: %7.1f cyc/k-step (math floor %5.1f)  %6.1f smemB/cyc  %.3f ms\n", t.name, grid, mx / ksteps,
             cols / 2.0, bytes / (mx / ksteps), ms);
    }
  }
  int tf = 0;
  CK(cudaMemcpyFromSymbol(&tf, g_timeout, sizeof(int)));
  printf("tc_rate: %s\n", tf ? "TIMEOUT" : "done");
  return 0;
}
