// tc_probe: hardware probe for the two assumptions the tensor-core Gabor kernel rests on.
//   (1) tcgen05.mma kind::f16 with hand-built no-swizzle K-major descriptors gives exact results
//       (canonical, non-overlapping operand tiles);
//   (2) the same instruction accepts an A descriptor whose core matrices OVERLAP
//       (LBO = 16 B, SBO = 128 B): row m of the operand is then xs[8m .. 8m+15], i.e. a Toeplitz
//       (sliding-window) matrix read straight out of a linear sample buffer;
//   (3) issue-rate of the N=160 + N=80 MMA pair the kernel uses (is the A read from shared memory
//       exposed at small N?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/tc_probe tools/tc_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../leaf_pytorch_b200/csrc/tc_ptx.cuh"

using namespace leafk::ptx;

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);      \
      exit(2);                                                                             \
    }                                                                                      \
  } while (0)

constexpr int M = 128;
constexpr int KSTEPS = 26;         // 416 taps
constexpr int XS_LEN = 1440;       // fp16 samples in the linear buffer (8*127 + 415 < 1440)

struct ProbeArgs {
  const __half* a;   // mode 0: canonical A tile (128x16) bytes ; mode 1: linear xs[XS_LEN]
  const __half* b;   // B operand bytes in kernel layout (NB rows, ksteps)
  float* d;          // [128][NB] out
  int NB;            // rows of B (multiple of 16)
  int mode;          // 0 canonical single k-step, 1 Toeplitz 26 k-steps (+ second MMA N=NB/2 from lo rows)
  int pair;          // mode 1: also issue the N=NB/2 MMA against a second xs copy (hi/lo scheme)
};

__global__ void __launch_bounds__(128) probe_kernel(ProbeArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int NB = p.NB;
  uint8_t* sA = smem;                        // up to 2*XS_LEN*2 bytes (hi copy, lo copy) or 4096
  uint8_t* sB = smem + 8192;
  const int a_bytes = (p.mode == 0) ? M * 16 * 2 : XS_LEN * 2 * (p.pair ? 2 : 1);
  const int b_bytes = (p.mode == 0) ? NB * 16 * 2 : NB * 16 * 2 * KSTEPS;
  for (int i = tid; i < a_bytes / 4; i += blockDim.x) ((uint32_t*)sA)[i] = ((const uint32_t*)p.a)[i];
  for (int i = tid; i < b_bytes / 4; i += blockDim.x) ((uint32_t*)sB)[i] = ((const uint32_t*)p.b)[i];
  if (tid == 0) { mbar_init(&bar, 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc<256>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    if (p.mode == 0) {
      const uint64_t ad = smem_desc(smem_u32(sA), 2048, 128);
      const uint64_t bd = smem_desc(smem_u32(sB), NB * 16, 128);
      mma_f16_ss(tmem, ad, bd, idesc_f16(M, NB), 0);
    } else {
      for (int ks = 0; ks < KSTEPS; ++ks) {
        const uint64_t ad = smem_desc(smem_u32(sA) + ks * 32, 16, 128);          // overlapping rows
        const uint64_t bd = smem_desc(smem_u32(sB) + ks * NB * 32, NB * 16, 128);
        mma_f16_ss(tmem, ad, bd, idesc_f16(M, NB), ks > 0);
        if (p.pair) {
          const uint64_t ad2 = smem_desc(smem_u32(sA) + XS_LEN * 2 + ks * 32, 16, 128);
          mma_f16_ss(tmem, ad2, bd, idesc_f16(M, NB / 2), 1);                     // lo(x) * hi(W)
        }
      }
    }
    mma_commit(&bar);
  }
  for (long long i = 0; i < 20000000LL; ++i)
    if (mbar_try_wait(&bar, 0)) break;
  tc_fence_after();
  for (int c0 = 0; c0 < NB; c0 += 16) {
    float v[16];
    tmem_ld16_sync(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) p.d[(size_t)tid * NB + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

// ---------------------------------------------------------------- throughput
struct RateArgs {
  int NB;        // N of the first MMA (0 = skip)
  int N2;        // N of the second MMA (0 = skip)
  int toeplitz;  // A descriptor: 1 overlapping (LBO 16), 0 canonical (LBO 2048)
  int batches;   // commits
  int per_batch; // k-steps per commit
  long long* cycles;
};

__device__ int g_timeout_flag = 0;
__device__ __forceinline__ bool bounded_wait(uint64_t* bar, uint32_t parity) {
  for (long long i = 0; i < 20000000LL; ++i)
    if (mbar_try_wait(bar, parity)) return true;
  g_timeout_flag = 1;
  return false;
}

__global__ void __launch_bounds__(128) rate_kernel(RateArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (64 * 1024) / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;   // zeros: values irrelevant
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init_fence(); }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 16384;
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    // two barriers, one per accumulator buffer: a batch may complete before the issuing thread
    // gets to wait on it, so consecutive batches must not share a barrier (phase ABA)
    uint32_t phase[2] = {0, 0};
    for (int bt = 0; bt < p.batches; ++bt) {
      if (bt >= 2) { if (!bounded_wait(&bar[bt & 1], phase[bt & 1])) break; phase[bt & 1] ^= 1; }
      for (int ks = 0; ks < p.per_batch; ++ks) {
        const int kk = ks % 26;
        const uint64_t ad = p.toeplitz ? smem_desc(smem_u32(sA) + kk * 32, 16, 128)
                                       : smem_desc(smem_u32(sA) + (kk & 1) * 4096, 2048, 128);
        const uint64_t bd = smem_desc(smem_u32(sB) + (kk % 8) * 5120, 2560, 128);
        if (p.NB) mma_f16_ss(tmem + (bt & 1) * 256, ad, bd, idesc_f16(M, p.NB), ks > 0);
        if (p.N2) mma_f16_ss(tmem + (bt & 1) * 256, ad, bd, idesc_f16(M, p.N2), 1);
      }
      mma_commit(&bar[bt & 1]);
    }
    bounded_wait(&bar[0], phase[0]);
    bounded_wait(&bar[1], phase[1]);
    t1 = clock64();
    p.cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

static size_t b_off(int NB, int n, int k) {   // same formula as k1_tc_layout.cuh
  return (size_t)(k / 16) * NB * 32 + (size_t)((k % 16) / 8) * NB * 16 + (size_t)(n / 8) * 128 + (size_t)(n % 8) * 16 +
         (size_t)(k % 8) * 2;
}

int main() {
  setvbuf(stdout, NULL, _IONBF, 0);
  int dev = 0;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  printf("device %s sm_%d%d SMs=%d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  srand(1234);
  int failures = 0;

  // ---------------- test 1: canonical layout, single k-step, N = 160 and 80
  for (int NB : {160, 80, 128, 256}) {
    std::vector<__half> A(M * 16), Bm((size_t)NB * 16);
    std::vector<float> Af(M * 16), Bf((size_t)NB * 16);
    for (int m = 0; m < M; ++m)
      for (int k = 0; k < 16; ++k) {
        float v = (float)(rand() % 9 - 4);
        Af[m * 16 + k] = v;
        A[(m / 8) * 64 + (k / 8) * 1024 + (m % 8) * 8 + (k % 8)] = __float2half(v);   // in halves
      }
    std::vector<uint8_t> Bbytes((size_t)NB * 32, 0);
    for (int n = 0; n < NB; ++n)
      for (int k = 0; k < 16; ++k) {
        float v = (float)(rand() % 9 - 4) * 0.5f;
        Bf[n * 16 + k] = v;
        *(__half*)&Bbytes[b_off(NB, n, k)] = __float2half(v);
      }
    __half *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, 8192)); CK(cudaMalloc(&dB, Bbytes.size())); CK(cudaMalloc(&dD, sizeof(float) * M * NB));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bbytes.data(), Bbytes.size(), cudaMemcpyHostToDevice));
    ProbeArgs pa{dA, dB, dD, NB, 0, 0};
    const int smem = 8192 + NB * 32 + 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    probe_kernel<<<1, 128, smem>>>(pa);
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)M * NB);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < NB; ++n) {
        float r = 0;
        for (int k = 0; k < 16; ++k) r += Af[m * 16 + k] * Bf[n * 16 + k];
        if (r != D[(size_t)m * NB + n]) { if (bad < 4) printf("  mismatch m=%d n=%d got %f want %f\n", m, n, D[(size_t)m * NB + n], r); ++bad; }
      }
    printf("test1 canonical N=%d: %s (%d mismatches)\n", NB, bad ? "FAIL" : "ok", bad);
    failures += bad != 0;
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }

  // ---------------- test 2: Toeplitz (overlapping) A descriptor, 26 k-steps, hi/lo pair
  for (int pair = 0; pair < 2; ++pair) {
    const int NB = 160, CG = 80, KP = KSTEPS * 16;
    std::vector<__half> xs(2 * XS_LEN);
    std::vector<float> xh(XS_LEN), xl(XS_LEN);
    for (int i = 0; i < XS_LEN; ++i) {
      xh[i] = (float)(rand() % 9 - 4);
      xl[i] = (float)(rand() % 5 - 2) * 0.25f;
      xs[i] = __float2half(xh[i]);
      xs[XS_LEN + i] = __float2half(xl[i]);
    }
    std::vector<uint8_t> Bbytes((size_t)NB * KP * 2, 0);
    std::vector<float> Wf((size_t)NB * KP);
    for (int n = 0; n < NB; ++n)
      for (int k = 0; k < KP; ++k) {
        float v = (k < 401) ? (float)(rand() % 7 - 3) * 0.5f : 0.f;
        Wf[(size_t)n * KP + k] = v;
        *(__half*)&Bbytes[b_off(NB, n, k)] = __float2half(v);
      }
    __half *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, xs.size() * 2)); CK(cudaMalloc(&dB, Bbytes.size())); CK(cudaMalloc(&dD, sizeof(float) * M * NB));
    CK(cudaMemcpy(dA, xs.data(), xs.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, Bbytes.data(), Bbytes.size(), cudaMemcpyHostToDevice));
    ProbeArgs pa{dA, dB, dD, NB, 1, pair};
    const int smem = 8192 + (int)Bbytes.size() + 1024;
    probe_kernel<<<1, 128, smem>>>(pa);
    CK(cudaDeviceSynchronize());
    std::vector<float> D((size_t)M * NB);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < M; ++m)
      for (int n = 0; n < NB; ++n) {
        double r = 0;
        for (int k = 0; k < KP; ++k) {
          r += (double)xh[8 * m + k] * Wf[(size_t)n * KP + k];
          if (pair && n < CG) r += (double)xl[8 * m + k] * Wf[(size_t)n * KP + k];
        }
        if ((float)r != D[(size_t)m * NB + n]) { if (bad < 4) printf("  mismatch m=%d n=%d got %f want %f\n", m, n, D[(size_t)m * NB + n], (float)r); ++bad; }
      }
    printf("test2 toeplitz pair=%d: %s (%d mismatches of %d)\n", pair, bad ? "FAIL" : "ok", bad, M * NB);
    failures += bad != 0;
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }

  // ---------------- test 3: issue rate
  {
    long long* dcyc;
    const int nsm = prop.multiProcessorCount;
    CK(cudaMalloc(&dcyc, sizeof(long long) * nsm));
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    struct Cfg { int NB, N2, toe; const char* name; };
    const Cfg cfgs[] = {{160, 0, 1, "N=160 toeplitz"}, {160, 0, 0, "N=160 canonical"}, {80, 0, 1, "N=80 toeplitz"},
                        {80, 0, 0, "N=80 canonical"},  {160, 80, 1, "N=160+80 toeplitz"}, {160, 80, 0, "N=160+80 canonical"},
                        {256, 0, 1, "N=256 toeplitz"}, {256, 0, 0, "N=256 canonical"}, {240, 0, 1, "N=240 toeplitz"},
                        {128, 64, 1, "N=128+64 toeplitz"}, {256, 128, 1, "N=256+128 toeplitz"}};
    for (const Cfg& c : cfgs) {
      for (int grid : {1, nsm}) {
        RateArgs ra{c.NB, c.N2, c.toe, 64, 52, dcyc};
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        rate_kernel<<<grid, 128, 64 * 1024>>>(ra);   // warm
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        rate_kernel<<<grid, 128, 64 * 1024>>>(ra);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> cyc(grid);
        CK(cudaMemcpy(cyc.data(), dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
        long long mx = 0;
        for (auto v : cyc) mx = v > mx ? v : mx;
        const double ksteps = 64.0 * 52.0;
        const double macs = ksteps * 128.0 * 16.0 * (c.NB + c.N2);
        printf("test3 %-22s grid=%3d: %8.1f cyc/k-step  %7.1f MAC/cyc/SM  kernel %.3f ms  (%.1f dense TFLOP/s chip)\n", c.name,
               grid, mx / ksteps, macs / mx, ms, 2.0 * macs * grid / (ms * 1e-3) / 1e12);
      }
    }
    cudaFree(dcyc);
  }
  int tf = 0;
  CK(cudaMemcpyFromSymbol(&tf, g_timeout_flag, sizeof(int)));
  if (tf) { printf("a bounded mbarrier wait TIMED OUT\n"); failures++; }
  printf("tc_probe: %s\n", failures ? "FAILURES" : "ALL OK");
  return failures ? 1 : 0;
}
