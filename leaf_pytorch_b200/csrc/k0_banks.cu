// K0: filter-bank prologue.  One tiny launch per forward/backward.
//
// Replaces, for the learnable parameters of one Leaf module:
//   GaborConstraint.forward                 reference convolution.py:15-22
//   gabor_filters / gabor_impulse_response  reference impulse_responses.py:5-16, 66-71
//   the re/im interleave                    reference convolution.py:77-90
//   gaussian_lowpass                        reference impulse_responses.py:74-80
// The fp32 operation order of those lines is kept (mu*tau is rounded to fp32 BEFORE sin/cos, the
// envelope argument is (1/(2 sigma^2)) * (-(tau^2)), the result is (norm*carrier)*envelope), so
// the banks agree with the reference's to an ulp or two of expf/sincosf.
//
// Outputs (all in the caller's workspace):
//   cprm[f][8]      constrained / derived per-filter constants (see CP_* in leafk_common.cuh)
//   w32[k][c]       fp32 bank, tap-major, channels padded to C2p, taps padded to Kp with zeros
//   g32[k][f]       fp32 Gaussian pooling windows, tap-major
//   w16             fp16 hi/lo bank in the tcgen05 shared-memory layout (see k1_tc_kernel.cuh), optional
#include "leafk_common.cuh"
#include "k1_tc_layout.cuh"
#include <cuda_fp16.h>
#include <math.h>

namespace leafk {

struct BankConsts {
  float mu_hi;      // (float)pi
  float sigma_lo;   // 4*sqrt(2 ln 2)/pi      convolution.py:18
  float sigma_hi;   // K*sqrt(2 ln 2)/pi      convolution.py:19
  float sqrt_2pi;   // sqrt(2*pi) as the reference rounds it (impulse_responses.py:6)
  float pool_lo;    // 2/K                    impulse_responses.py:75
};

// Width-sorted channel order and per-k-step active counts of the pruned tensor-core layout
// (k1_tc_layout.cuh, "SUPPORT PRUNING").  Every block recomputes the (tiny) sort redundantly so that one
// launch suffices: sm_key[f'] = clamped sigma (or -1 for padding filters), sm_pos[f'] = rank of f' in
// ascending (key, index) order, sm_na[s] = active channels of THIS filter's group at k-step s.
struct TcPrune {
  int* perm;          // [groups][FG] (group, slot) -> filter index (>= F: padding)
  int* zones;         // [groups][tc::ZONE_INTS]: ints [0,16) {lo_L, hi_L} = k-steps with >= 16 L channels running,
                      // L = 1..CG/16; ints [16,32) {na3 on the rising zone of level L, on its falling zone}
  int* kcodes;        // [groups][Kp/16] (na1/16) | (na3/16) << 4 per k-step (profiling read-back)
  float c;            // support radius in sigmas; <= 0: no pruning
  float c3;           // radius beyond which only the main product runs; <= 0: everywhere all products
  int* done;          // [n_done] per-clip completion counters of K1 -> K2, zeroed here
  int n_done;
};

__global__ void __launch_bounds__(128)
k0_banks_kernel(const float* __restrict__ kernel, const float* __restrict__ pool_w, BankConsts bc,
                int F, int K, int Kp, int C2p, float* __restrict__ cprm, float* __restrict__ w32,
                float* __restrict__ g32, uint8_t* __restrict__ w16, int tc_cg, int tc_groups, TcPrune pr) {
  extern __shared__ float k0_smem[];
  // programmatic dependent launch: the consumer (K1) may be scheduled while this grid runs; it waits for this grid's
  // completion (griddepcontrol.wait) before it touches the banks
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int f = blockIdx.x;                 // filter (or a zero-padding channel pair when f >= F)
  const int k_of_thread = blockIdx.y * blockDim.x + threadIdx.x;   // one tap per thread, gridDim.y chunks of taps
  if (pr.done != nullptr)
    for (int i = (blockIdx.y * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x; i < pr.n_done;
         i += gridDim.x * gridDim.y * blockDim.x)
      pr.done[i] = 0;
  const size_t grp_bytes = tc::b_group_bytes(tc_cg, Kp);
  const int FG = tc_cg / 2, Fp = FG * tc_groups, ks = Kp / tc::KSTEP;
  float* sm_key = k0_smem;
  int* sm_pos = reinterpret_cast<int*>(k0_smem + Fp);       // rank | first k-step inside 5.5 sigma << 16 | last << 24 (ks <= 128)
  int* sm_r3 = sm_pos + Fp;                                  // first k-step inside 3.7 sigma | last << 8
  int* sm_na1 = sm_r3 + Fp;                                  // per k-step: channels of this filter's group that run
  int* sm_na3 = sm_na1 + ks;                                 //             channels that run all three products
  int my_pos = 0;
  if (w16 != nullptr) {
    for (int i = threadIdx.x; i < Fp; i += blockDim.x)
    {
      float key = -1.0f;                                     // padding filter
      if (i < F) {
        const float raw = kernel[2 * i + 1];
        key = (raw != raw) ? bc.sigma_hi : fminf(fmaxf(raw, bc.sigma_lo), bc.sigma_hi);   // NaN: widest (never pruned)
      }
      sm_key[i] = key;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < Fp; i += blockDim.x) {
      const float ki = sm_key[i];
      int r = 0;
      for (int j = 0; j < Fp; ++j) {
        const float kj = sm_key[j];
        r += (kj < ki || (kj == ki && j < i)) ? 1 : 0;
      }
      int lo, hi, lo3, hi3;
      tc::kstep_range(ki, pr.c, K, Kp, &lo, &hi);
      tc::kstep_range(ki, pr.c3, K, Kp, &lo3, &hi3);
      if (lo3 < lo) lo3 = lo;                                // (c3 <= 0 or > c: never more than the channels that run)
      if (hi3 > hi) hi3 = hi;
      sm_pos[i] = r | (lo << 16) | (hi << 24);
      sm_r3[i] = lo3 | (hi3 << 8);
    }
    __syncthreads();
    my_pos = sm_pos[f < Fp ? f : 0] & 0xffff;
    const int my_grp = tc::group_of(my_pos, tc_groups);
    for (int s = threadIdx.x; s < ks; s += blockDim.x) {
      int cnt = 0, cnt3 = 0;
      for (int j = 0; j < Fp; ++j) {
        const int v = sm_pos[j], v3 = sm_r3[j];
        const bool mine = tc::group_of(v & 0xffff, tc_groups) == my_grp;
        cnt += (mine && s >= ((v >> 16) & 0xff) && s <= ((v >> 24) & 0xff)) ? 1 : 0;
        cnt3 += (mine && s >= (v3 & 0xff) && s <= ((v3 >> 8) & 0xff)) ? 1 : 0;
      }
      int nf = (cnt + 7) / 8 * 8, nf3 = (cnt3 + 7) / 8 * 8;
      if (nf > FG || s == tc::first_kstep(Kp)) nf = FG;
      if (nf3 > nf || s == tc::first_kstep(Kp)) nf3 = nf;
      sm_na1[s] = 2 * nf;
      sm_na3[s] = 2 * nf3;
    }
    __syncthreads();
    // Make na3 constant on every zone of constant na1 (a zone = one side of the middle k-step): the largest value
    // inside the zone; where every channel runs (na1 = CG) every channel also runs all three products.  The issue
    // loop of k1 then has one (na1, na3) pair per zone and no more zones than with one pruning level.
    int na3_aligned = 0;
    const int sf = tc::first_kstep(Kp);
    if ((int)threadIdx.x < ks) {
      const int s = threadIdx.x, n1 = sm_na1[s];
      na3_aligned = n1;
      if (n1 < tc_cg) {
        na3_aligned = 0;
        for (int t = (s < sf ? 0 : sf + 1); t < (s < sf ? sf : ks); ++t)
          if (sm_na1[t] == n1 && sm_na3[t] > na3_aligned) na3_aligned = sm_na3[t];
      }
    }
    __syncthreads();
    if ((int)threadIdx.x < ks) sm_na3[threadIdx.x] = na3_aligned;
    __syncthreads();
    if (f < Fp && blockIdx.y == 0) {
      if (threadIdx.x == 0) pr.perm[my_grp * FG + tc::slot_of(my_pos, tc_groups)] = f;
      if (tc::slot_of(my_pos, tc_groups) == 0 && threadIdx.x < tc_cg / 16) {   // the group's first filter publishes the zone tables
        const int L = threadIdx.x + 1;
        int lo = ks, hi = -1, n3r = 0, n3f = 0;
        for (int s = 0; s < ks; ++s) {
          if (sm_na1[s] >= 16 * L) { lo = s < lo ? s : lo; hi = s; }
          if (sm_na1[s] == 16 * L) { if (s < sf) n3r = sm_na3[s]; else n3f = sm_na3[s]; }
        }
        pr.zones[my_grp * tc::ZONE_INTS + 2 * (L - 1)] = lo;
        pr.zones[my_grp * tc::ZONE_INTS + 2 * (L - 1) + 1] = hi;
        pr.zones[my_grp * tc::ZONE_INTS + 16 + 2 * (L - 1)] = n3r;
        pr.zones[my_grp * tc::ZONE_INTS + 16 + 2 * (L - 1) + 1] = n3f;
      }
      // per-k-step schedule for the profiling read-back: (na1/16) | (na3/16) << 4
      if (tc::slot_of(my_pos, tc_groups) == 0)
        for (int s = threadIdx.x; s < ks; s += blockDim.x)
          pr.kcodes[my_grp * ks + s] = (sm_na1[s] >> 4) | ((sm_na3[s] >> 4) << 4);
    }
  }
  // tensor-core image: taps of sorted channel c = 2*j + q of group grp, only where the k-step keeps the channel
  auto store_tc = [&](int k, int q, float scaled) {
    const int grp = tc::group_of(my_pos, tc_groups), j = tc::slot_of(my_pos, tc_groups), c = 2 * j + q;
    const int na1 = sm_na1[k / tc::KSTEP], na3 = sm_na3[k / tc::KSTEP];
    if (c < tc_cg - na1) return;                       // outside the active suffix: never read by the MMAs
    uint8_t* gb = w16 + (size_t)grp * grp_bytes;
    const __half hi = __float2half_rn(scaled);
    *reinterpret_cast<__half*>(gb + tc::p_hi_main(tc_cg, Kp, c, na1, na3, k)) = hi;
    if (c < tc_cg - na3) return;                       // main product only
    const __half lo = __float2half_rn(scaled - __half2float(hi));
    *reinterpret_cast<__half*>(gb + tc::p_lo_main(tc_cg, Kp, c, na1, na3, k)) = lo;
    *reinterpret_cast<__half*>(gb + tc::p_hi_corr(tc_cg, Kp, c, na3, k)) = hi;
  };
  if (f >= F) {                             // padded channels: zero taps in both layouts
    for (int k = k_of_thread; k < Kp; k += gridDim.y * blockDim.x) {
      for (int q = 0; q < 2; ++q) {
        const int c = 2 * f + q;
        if (c < C2p) w32[(size_t)k * C2p + c] = 0.f;
        if (w16 != nullptr && f < Fp) store_tc(k, q, 0.f);
      }
    }
    return;
  }
  const float mu = clamp_nan(kernel[2 * f], 0.f, bc.mu_hi);
  const float sg = clamp_nan(kernel[2 * f + 1], bc.sigma_lo, bc.sigma_hi);
  const float norm = 1.0f / (bc.sqrt_2pi * sg);
  const float inv2s2 = 1.0f / (2.0f * (sg * sg));
  const float ps = clamp_nan(pool_w[f], bc.pool_lo, 0.5f);
  const float den = (ps * 0.5f) * (float)(K - 1);
  const float centre = (float)(0.5 * (double)(K - 1));

  // power-of-two scale that puts the filter's peak (norm, at tau = 0) in [2^13, 2^14) for the
  // fp16 hi/lo split; exact, undone in the epilogue.
  int wexp;
  (void)frexpf(norm, &wexp);                // norm = m * 2^wexp, m in [0.5,1)
  const int wshift = 14 - wexp;             // norm * 2^wshift in [2^13, 2^14)

  if (threadIdx.x == 0 && blockIdx.y == 0) {
    float* c = cprm + (size_t)f * 8;
    c[CP_MU] = mu;
    c[CP_SIGMA] = sg;
    c[CP_NORM] = norm;
    c[CP_INV2S2] = inv2s2;
    c[CP_POOLS] = ps;
    // g_f[k] = exp2(pool_a * (k - centre)^2)
    c[CP_POOLA] = (float)(-0.5 * 1.4426950408889634 / ((double)den * (double)den));
    c[CP_WSCALE] = (float)wshift;
    c[CP_PAD] = 0.f;
  }

  for (int k = k_of_thread; k < Kp; k += gridDim.y * blockDim.x) {
    float wr = 0.f, wi = 0.f;
    if (k < K) {
      const float tau = (float)(k - K / 2);
      const float env = expf(inv2s2 * (-(tau * tau)));
      const float ph = mu * tau;
      float sn, cs;
      sincosf(ph, &sn, &cs);
      wr = (norm * cs) * env;
      wi = (norm * sn) * env;
      const float r = ((float)k - centre) / den;
      g32[(size_t)k * F + f] = expf(-0.5f * (r * r));
    }
    w32[(size_t)k * C2p + 2 * f] = wr;
    w32[(size_t)k * C2p + 2 * f + 1] = wi;
    if (w16 != nullptr) {
      // hi/lo fp16 split of the scaled taps, stored in the CTA-pair regions of the channel's group
      store_tc(k, 0, ldexpf(wr, wshift));
      store_tc(k, 1, ldexpf(wi, wshift));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Training banks.  For every filter three complex banks in the tcgen05 TRAINING layout (k1_tc_layout.cuh; fp16
// hi/lo, own power-of-two scale each):  h (kind 0),  tau*h (kind 1, d/dmu up to the factor i),
// (tau^2/sigma^3 - 1/sigma)*h (kind 2, d/dsigma).  Channel order: tc::train_channel().
// tprm[f] = {pool exp2 coefficient, shift_y, shift_z, shift_v, sigma, pool_s, 0, 0}.
__global__ void __launch_bounds__(128)
k0_banks_train_kernel(const float* __restrict__ kernel, const float* __restrict__ pool_w, BankConsts bc, int F,
                      int K, int Kp, int FB, float* __restrict__ tprm, uint8_t* __restrict__ w16t, int* done, int n_done) {
  __shared__ float smax[3][4];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");     // K1 waits for this grid's completion itself
  const int f = blockIdx.x;                          // padded filter index (f >= F: zero banks)
  const int CG = 6 * FB;
  const int grp = f / FB, fl = f % FB;
  uint8_t* gb = w16t + (size_t)grp * tc::t_group_bytes(CG, Kp);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (done != nullptr)
    for (int i = blockIdx.x * blockDim.x + tid; i < n_done; i += gridDim.x * blockDim.x) done[i] = 0;
  if (f >= F) {
    for (int k = tid; k < Kp; k += blockDim.x)
      for (int kind = 0; kind < 3; ++kind)
        for (int ri = 0; ri < 2; ++ri) {
          const int c = tc::train_channel(FB, fl, kind, ri);
          const __half z = __float2half_rn(0.f);
          *reinterpret_cast<__half*>(gb + tc::t_hi(CG, Kp, c, k)) = z;
          *reinterpret_cast<__half*>(gb + tc::t_lo(CG, Kp, c, k)) = z;
        }
    if (tid == 0) {
      float* bp = tprm + (size_t)f * 8;
      bp[0] = -1.0f;
      for (int i = 1; i < 8; ++i) bp[i] = 0.f;
    }
    return;
  }
  const float mu = clamp_nan(kernel[2 * f], 0.f, bc.mu_hi);
  const float sg = clamp_nan(kernel[2 * f + 1], bc.sigma_lo, bc.sigma_hi);
  const float norm = 1.0f / (bc.sqrt_2pi * sg);
  const float inv2s2 = 1.0f / (2.0f * (sg * sg));
  const float ps = clamp_nan(pool_w[f], bc.pool_lo, 0.5f);
  const float den = (ps * 0.5f) * (float)(K - 1);
  const float inv_s3 = 1.0f / (sg * sg * sg), inv_s = 1.0f / sg;

  // pass 1: peak magnitude of each kind
  float mx[3] = {0.f, 0.f, 0.f};
  for (int k = tid; k < K; k += blockDim.x) {
    const float tau = (float)(k - K / 2);
    const float a = norm * expf(inv2s2 * (-(tau * tau)));     // |h| (carrier has modulus 1)
    mx[0] = fmaxf(mx[0], a);
    mx[1] = fmaxf(mx[1], fabsf(tau) * a);
    mx[2] = fmaxf(mx[2], fabsf(tau * tau * inv_s3 - inv_s) * a);
  }
#pragma unroll
  for (int q = 0; q < 3; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx[q] = fmaxf(mx[q], __shfl_xor_sync(0xffffffffu, mx[q], o));
    if (lane == 0) smax[q][warp] = mx[q];
  }
  __syncthreads();
  int shift[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const float m = fmaxf(fmaxf(smax[q][0], smax[q][1]), fmaxf(smax[q][2], smax[q][3]));
    int ex = 0;
    if (m > 0.f && m < 3.0e38f) (void)frexpf(m, &ex);
    shift[q] = (m > 0.f && m < 3.0e38f) ? 14 - ex : 0;
  }
  if (tid == 0) {
    float* bp = tprm + (size_t)f * 8;
    bp[0] = (float)(-0.5 * 1.4426950408889634 / ((double)den * (double)den));
    bp[1] = (float)shift[0]; bp[2] = (float)shift[1]; bp[3] = (float)shift[2];
    bp[4] = sg; bp[5] = ps; bp[6] = 0.f; bp[7] = 0.f;
  }
  // pass 2: write the banks
  for (int k = tid; k < Kp; k += blockDim.x) {
    float v[3][2] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
    if (k < K) {
      const float tau = (float)(k - K / 2);
      const float env = expf(inv2s2 * (-(tau * tau)));
      float sn, cs;
      sincosf(mu * tau, &sn, &cs);
      const float wr = (norm * cs) * env, wi = (norm * sn) * env;
      const float cs_ = tau * tau * inv_s3 - inv_s;
      v[0][0] = wr; v[0][1] = wi;
      v[1][0] = tau * wr; v[1][1] = tau * wi;
      v[2][0] = cs_ * wr; v[2][1] = cs_ * wi;
    }
#pragma unroll
    for (int kind = 0; kind < 3; ++kind)
#pragma unroll
      for (int ri = 0; ri < 2; ++ri) {
        const float sc = ldexpf(v[kind][ri], shift[kind]);
        const __half hi = __float2half_rn(sc);
        const __half lo = __float2half_rn(sc - __half2float(hi));
        const int c = tc::train_channel(FB, fl, kind, ri);
        *reinterpret_cast<__half*>(gb + tc::t_hi(CG, Kp, c, k)) = hi;
        *reinterpret_cast<__half*>(gb + tc::t_lo(CG, Kp, c, k)) = lo;
      }
  }
}

static BankConsts make_consts(int K) {
  BankConsts bc;
  const float root_2ln2 = sqrtf(2.0f * logf(2.0f));
  bc.mu_hi = (float)M_PI;
  bc.sigma_lo = (4.0f * root_2ln2) / (float)M_PI;
  bc.sigma_hi = ((float)K * root_2ln2) / (float)M_PI;
  bc.sqrt_2pi = sqrtf(2.0f * (float)M_PI);
  bc.pool_lo = (float)(2.0 / (double)K);
  return bc;
}

void bank_bounds(int K, float* mu_hi, float* sigma_lo, float* sigma_hi, float* pool_lo) {
  const BankConsts bc = make_consts(K);
  *mu_hi = bc.mu_hi; *sigma_lo = bc.sigma_lo; *sigma_hi = bc.sigma_hi; *pool_lo = bc.pool_lo;
}

void launch_k0_train(const float* kernel, const float* pool_w, int F, int K, int Kp, int FB, int n_groups,
                     float* tprm, uint8_t* w16t, int* done, int n_done, cudaStream_t stream) {
  k0_banks_train_kernel<<<n_groups * FB, 128, 0, stream>>>(kernel, pool_w, make_consts(K), F, K, Kp, FB, tprm, w16t,
                                                         done, n_done);
}

void launch_k0(const float* kernel, const float* pool_w, int F, int K, int Kp, int C2p, float* cprm,
               float* w32, float* g32, uint8_t* w16, int tc_cg, int tc_groups, int* tc_perm, int* tc_zones,
               float prune_c, float prune_c3, int* done, int n_done, cudaStream_t stream) {
  const BankConsts bc = make_consts(K);
  int nblk = C2p / 2;
  const int Fp = tc_cg * tc_groups / 2;
  if (w16 != nullptr && Fp > nblk) nblk = Fp;
  const size_t smem = (w16 != nullptr) ? sizeof(float) * (3 * (size_t)Fp + 2 * (Kp / tc::KSTEP)) : 0;
  // one tap per thread: gridDim.y chunks of 128 taps (the kernel is latency-bound: sincosf + expf + scattered
  // 2-byte stores per tap; 40 blocks looping over 416 taps took 14 us)
  const dim3 grid((unsigned)nblk, (unsigned)((Kp + 127) / 128));
  k0_banks_kernel<<<grid, 128, smem, stream>>>(kernel, pool_w, bc, F, K, Kp, C2p, cprm, w32, g32, w16, tc_cg,
                                               tc_groups, TcPrune{tc_perm, tc_zones, tc_zones + (size_t)tc_groups * tc::ZONE_INTS, prune_c, prune_c3, done, n_done});
}

}  // namespace leafk
