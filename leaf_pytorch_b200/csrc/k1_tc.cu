// K1 (tensor-core variant): Gabor correlation as a Toeplitz GEMM on tcgen05, fused with the squared
// modulus and the Gaussian pooling partials.
//
// Replaces   F.conv1d(pad(x), bank)        reference convolution.py:91-98   (1.03 GFLOP per audio-second)
//            SquaredModulus.forward        reference frontend.py:15-19
//            GaussianLowPass.forward       reference pooling.py:31-42       (bias is added in K2)
//
// Arithmetic.  y[t,c] = sum_k x~[t+k] W[c,k] is computed as D[128 x NB] += A[128 x 16] * B[16 x NB] with
// fp16 operands and fp32 accumulation in tensor memory.  Plain fp16 (or TF32) inputs miss the 1e-4
// parity target (3.4e-4), so both operands are split x = xh + xl, W = Wh + Wl after an
// exact power-of-two scaling (per tile for x, per filter for W) that puts |xh|,|Wh| < 2^14, and the
// three significant products are formed per k-step with two MMAs:
//      D[:, 0:CG)   += xh * Wh       \_ one MMA, N = 2*CG (B rows = [Wh ; Wl])
//      D[:, CG:2CG) += xh * Wl       /
//      D[:, 0:CG)   += xl * Wh          one MMA, N = CG
// (xl*Wl ~ 2^-22 is dropped.)  The epilogue adds the two column halves: ~2^-21 relative, fp32 class.
// Forward: per k-step only the channels whose filter is still inside its support run (N = na1 + na3 / na3, two
// pruning levels, channels in ascending-width order, lo columns mirrored): see k1_tc_layout.cuh and issue_zone below.
//
// The A operand is never materialised: see k1_tc_layout.cuh (overlapping-core-matrix descriptor on
// 8 shifted linear copies of the sample window; row m of phase p = output sample ts + 8m + p).
//
// One persistent CTA per SM (CTA pairs, tcgen05 cta_group::2), 12 warps:
//   warps 0-7   epilogue: tcgen05.ld of a finished phase (128 rows x NB columns), hi+lo add, re^2+im^2,
//               Gaussian window weight by ex2.approx of a per-filter coefficient times (k-centre)^2, FMA
//               into <= NSLOT frame accumulators per (row, filter) kept in registers for the whole tile;
//               at the end of the tile ONE recursive-halving shuffle reduction over the rows for all (<= 5)
//               frames at once, a fixed-order sum over the four row quadrants and one store of the tile's
//               partial pooled sums; a phase later the tile is published to K2 (per-clip counters).
//   warp  8     allocates tensor memory and issues every MMA (one elected lane, uniform control flow),
//               zone by zone of constant active channel counts.
//   warps 9-11  producers: load the sample window into registers, find its max, scale, split to fp16
//               hi/lo, and write the 8 shifted copies; copy p of the next tile is rebuilt as soon as phase p
//               of the current tile has been consumed (per-phase full/empty mbarriers).
// Accumulators rotate through NST = 512/NB tensor-memory stages so the epilogue of phase p overlaps
// the MMAs of phases p+1.. .  The bank of the CTA's channel group stays resident in shared memory.
// The kernel is chained to k0 (before) and k2 (after) by programmatic dependent launch.
#include "leafk_common.cuh"
#include "k1_tc_layout.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <cstdio>
#include <cstring>

#ifndef LEAFK_EXP
#define LEAFK_EXP 0      // timing experiments only (results wrong): 1 = no epilogue loads + arithmetic, 2 = producers skip the
                         // copies, 4 = epilogue loads only, 8 = epilogue arithmetic only,
                         // 16 = no epilogue work on the MMA warp's scheduler (quadrant 0), 32 = only there
#endif

namespace leafk {

using namespace ptx;

namespace tc {
constexpr int EPI_WARPS = 8;
constexpr int MMA_WARP = 8;
constexpr int PROD_WARP0 = 9;
constexpr int PROD_WARPS = 3;
constexpr int PROD_THREADS = PROD_WARPS * 32;
constexpr int NTHREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32;   // 384
constexpr int BAR_PROD = 1, BAR_EPI = 2;                       // named barrier ids

struct Misc {                 // small shared state behind the big regions
  uint64_t a_full[NPHASE];
  uint64_t a_empty[NPHASE];
  uint64_t acc_full[4];
  uint64_t acc_empty[4];
  uint32_t tmem_base;
  int sx_ring[4];
  float red[4];
};
static_assert(sizeof(Misc) <= 512, "Misc must fit the reserved tail");
}  // namespace tc


// Producer step for phase P: wait until the MMAs of phase P of the previous tile have drained, then write copy_P (hi
// and lo): chunk jj of the copy = staged halves [8jj+P, 8jj+P+8).
// Each producer thread owns the chunks jj = ptid + c * PROD_THREADS (c < NC) of every
// copy and keeps the 16 scaled hi / lo halves [8 jj, 8 jj + 16) they are cut from in registers (wh / wl, 8 words per
// chunk), so a tile costs the shared-memory pipe only the 16 copy stores: the staging round trips (fp32 store + load,
// half store, two 16-byte loads per 16-byte chunk) were ~850 of the ~1200 wavefronts the producers added per tile to
// a pipe that the tensor-core operand fetches keep ~90 % busy.
template <int P, int NC>              // NC = chunks per thread: ceil((127 + Kp/8) / 96), 2 for the 401-tap window, <= 4
__device__ __forceinline__ void build_copy_regs(tc::Misc* misc, const uint32_t (&wh)[NC][8],
                                                const uint32_t (&wl)[NC][8], uint8_t* s_acopy, int acb, int nchunk,
                                                int ptid, int lane, int it) {
  mbar_wait(&misc->a_empty[P], (uint32_t)((it & 1) ^ 1));
  constexpr int s = P >> 1;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int jj = ptid + c * tc::PROD_THREADS;
    if (jj < (((LEAFK_EXP & 2) != 0) ? 0 : nchunk)) {
      uint4 oh, ol;
      if ((P & 1) == 0) {
        oh = make_uint4(wh[c][s], wh[c][s + 1], wh[c][s + 2], wh[c][s + 3]);
        ol = make_uint4(wl[c][s], wl[c][s + 1], wl[c][s + 2], wl[c][s + 3]);
      } else {
        oh = make_uint4(__funnelshift_r(wh[c][s], wh[c][s + 1], 16), __funnelshift_r(wh[c][s + 1], wh[c][s + 2], 16),
                        __funnelshift_r(wh[c][s + 2], wh[c][s + 3], 16), __funnelshift_r(wh[c][s + 3], wh[c][s + 4], 16));
        ol = make_uint4(__funnelshift_r(wl[c][s], wl[c][s + 1], 16), __funnelshift_r(wl[c][s + 1], wl[c][s + 2], 16),
                        __funnelshift_r(wl[c][s + 2], wl[c][s + 3], 16), __funnelshift_r(wl[c][s + 3], wl[c][s + 4], 16));
      }
      *reinterpret_cast<uint4*>(s_acopy + (size_t)(2 * P) * acb + (size_t)jj * 16) = oh;
      *reinterpret_cast<uint4*>(s_acopy + (size_t)(2 * P + 1) * acb + (size_t)jj * 16) = ol;
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) mbar_arrive_rank0(&misc->a_full[P]);      // the MMA issuer lives in rank 0 of the pair
}

// Extra arguments of the backward variant (MODE 1): the epilogue turns the three correlations
// y, z = x*(tau h), v = x*((tau^2/sigma^3 - 1/sigma) h) into the per-filter sums that give the gradients of
// centre, width and pooling width ("equivalent without forming dW").
// Optional host-pipelining hook: when `ready` is non-null the producers wait, before touching clip b, until
// ready[b / clips_per_flag] != 0.  The flags are set by stream-ordered 32-bit writes that follow each slice of
// the H2D copy on another stream, so ONE persistent launch overlaps the whole PCIe transfer (leafk_forward_host).
struct TcReady {
  const int* ready;
  int clips_per_flag;
  long long* perf;      // optional: CTA 0 stores {SM cycles, nanoseconds} of its lifetime (effective SM clock)
};

// Forward only: the width-sorted channel order and the per-k-step active channel counts written by k0
// (k1_tc_layout.cuh, "SUPPORT PRUNING").
struct TcMap {
  int* done;            // [B] per-clip completion counters for K2 (one increment per epilogue warp and stored tile)
  const int* perm;      // [n_groups * CG/2] sorted position -> filter index (>= F: padding)
  const int* zones;     // [n_groups][tc::ZONE_INTS]: ints [0,16) {lo_L, hi_L}, L = 1..CG/16: k-steps with >= 16 L channels
                        // running; ints [16,32) {na3 of level L's rising zone, of its falling zone}
};

struct TcBwdArgs {
  const float* dpT;     // (B, N, F) gradient w.r.t. the floored pooled energies, frame-major
  const float* bprm;    // (Fpad, 8): [0] pooling exp2 coefficient, [1..3] power-of-two shifts of the y,z,v banks
  float* bpart;         // (ctas_per_group, Fpad, 4) per-CTA partial sums: {S_mu, S_sigma, S_poolw, 0}
  int Fpad;             // n_groups * FB
  int skip_xlo;         // 1: drop the x_lo*W_hi product (2-product mode): the waveform enters with fp16 rounding
                        //    (zero-mean, averages out over the B*T terms of a gradient), the banks keep hi+lo
};

// ---- pruned MMA issue (forward) ------------------------------------------------------------------------------
// The active channel count na1(s) is unimodal in the k-step s (nested, centred supports), so the k-steps split into
// at most 2*CG/16 - 1 ZONES of constant na1: level L (na1 >= 16 L) is active on the k-step interval [lo_L, hi_L],
// intervals nested; k0 makes na3 (channels that run all three products) constant on every zone.  k0 publishes the bounds; the issuing warp makes them warp-uniform registers (redux) and runs
// one short loop per zone with NA a compile-time constant: running descriptors advanced by immediates, i.e. two
// 64-bit uniform adds per MMA like the unpruned loop.  (Measured alternatives: a per-k-step table in shared memory
// cost ~18 R2UR and 190 cycles per k-step, a per-k-step switch with immediate offsets -- jump tables -- 340.)
template <int CG, int NA>
__device__ __forceinline__ void issue_zone(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b1, uint64_t b2, int s0,
                                           int s1, int na3, uint32_t accumulate_first) {
  // NA channels run (compile time), the last na3 <= NA of them all three products (zone-uniform run-time value):
  // main MMA N = NA + na3 into columns [CG-NA, CG+na3), corr MMA N = na3 into [CG-na3, CG)
  // running descriptors: only the low words (start address field, 16-byte units) move, and never carry
  uint32_t ah = (uint32_t)a_hi + (uint32_t)(2 * s0), al = (uint32_t)a_lo + (uint32_t)(2 * s0);
  uint32_t bb1 = (uint32_t)b1 + (uint32_t)(s0 * 2 * CG + (CG - NA)), bb2 = (uint32_t)b2 + (uint32_t)(s0 * CG + (CG - na3) / 2);
  const uint32_t ahh = (uint32_t)(a_hi >> 32), alh = (uint32_t)(a_lo >> 32), b1h = (uint32_t)(b1 >> 32), b2h = (uint32_t)(b2 >> 32);
  const uint32_t d1 = d + (uint32_t)(CG - NA), d2 = d + (uint32_t)(CG - na3);
  const uint32_t id1 = idesc_f16(256, 0) | ((uint32_t)((NA + na3) >> 3) << 17);
  const uint32_t id2 = idesc_f16(256, 0) | ((uint32_t)(na3 >> 3) << 17);
  if (na3 > 0) {
#pragma unroll 1
    for (int s = s0; s < s1; ++s) {
      mma_f16_ss_pair_w(d1, ah, ahh, bb1, b1h, id1, (s > s0) ? 1u : accumulate_first);   // x_hi * [W_hi | W_lo]
      mma_f16_ss_pair_w(d2, al, alh, bb2, b2h, id2, 1);                                  // x_lo * W_hi
      ah += 2; al += 2; bb1 += (uint32_t)(2 * CG); bb2 += (uint32_t)CG;
    }
  } else {
#pragma unroll 1
    for (int s = s0; s < s1; ++s) {
      mma_f16_ss_pair_w(d1, ah, ahh, bb1, b1h, id1, (s > s0) ? 1u : accumulate_first);   // x_hi * W_hi only
      ah += 2; bb1 += (uint32_t)(2 * CG);
    }
  }
}
// zones outside the centre one, levels L = LV .. 1 (rising side [lo_L, lo_{L+1}), falling side (hi_{L+1}, hi_L]);
// z3r / z3f: channels of the level's rising / falling zone that run all three products
template <int CG, int LV>
__device__ __forceinline__ void issue_outer_zones(uint32_t d, uint64_t a_hi, uint64_t a_lo, uint64_t b1, uint64_t b2,
                                                  const int* zlo, const int* zhi, const int* z3r, const int* z3f) {
  if constexpr (LV >= 1) {
    issue_zone<CG, 16 * LV>(d, a_hi, a_lo, b1, b2, zlo[LV - 1], zlo[LV], z3r[LV - 1], 1);
    issue_zone<CG, 16 * LV>(d, a_hi, a_lo, b1, b2, zhi[LV] + 1, zhi[LV - 1] + 1, z3f[LV - 1], 1);
    issue_outer_zones<CG, LV - 1>(d, a_hi, a_lo, b1, b2, zlo, zhi, z3r, z3f);
  }
}

// ---- tile-end reduction over the 32 rows of a warp ----------------------------------------------------------
// Sum v[f][0..N) over the lanes by recursive halving, NFR independent arrays at once: at exchange distance D the
// lanes with bit D clear keep the lower half of the indices and receive the partner's partial sums of it, the
// others the upper half; after the last step lane L holds the complete sums of ONE index,
// halving_index<N0,16>(L) (-1: a padding slot), for every f.  ~N shuffles per array instead of 5 N, and the NFR
// arrays advance together, so a tile costs 5 dependent shuffle levels in all.
template <int NFR, int N, int D>
__device__ __forceinline__ void halving_multi(const float (&v)[NFR][N], int lane, float (&out)[NFR]) {
  if constexpr (N == 1) {
#pragma unroll
    for (int f = 0; f < NFR; ++f) out[f] = v[f][0];
#pragma unroll
    for (int o = D; o > 0; o >>= 1) {
#pragma unroll
      for (int f = 0; f < NFR; ++f) out[f] += __shfl_xor_sync(0xffffffffu, out[f], o);
    }
  } else {
    static_assert(D >= 1, "more values than lanes");
    constexpr int H = (N + 1) / 2;
    const bool up = (lane & D) != 0;
    float k[NFR][H];
#pragma unroll
    for (int f = 0; f < NFR; ++f) {
#pragma unroll
      for (int i = 0; i < H; ++i) {
        const float lo_v = v[f][i], hi_v = (i + H < N) ? v[f][i + H < N ? i + H : 0] : 0.f;
        k[f][i] = (up ? hi_v : lo_v) + __shfl_xor_sync(0xffffffffu, up ? lo_v : hi_v, D);
      }
    }
    halving_multi<NFR, H, D / 2>(k, lane, out);
  }
}
template <int N, int D>
__device__ __forceinline__ int halving_index(int lane) {
  if constexpr (N == 1 || D == 0) {
    return 0;
  } else {
    constexpr int H = (N + 1) / 2;
    const int inner = halving_index<H, D / 2>(lane);
    const int idx = ((lane & D) ? H : 0) + inner;
    return (inner < 0 || idx >= N) ? -1 : idx;
  }
}

// Producer warps (3 warps of each CTA): per tile load the sample window, find its max, scale by a power of two,
// split to fp16 hi/lo in registers and write the 8 shifted copies, copy p as soon as phase p of the previous tile has
// been consumed.
template <int NC>
__device__ __forceinline__ void producer_loop(const Geom& g, const float* __restrict__ x, const TcReady& rdy,
                                              const tc::SmemPlan& sp, tc::Misc* misc, uint8_t* s_acopy, int tid, int lane,
                                              int warp, uint32_t rank, int pair_in_grp, int pairs_in_grp,
                                              long long n_units, long long n_pair_units) {
  using namespace tc;
  const int ptid = tid - PROD_WARP0 * 32;
  const int nchunk = sp.CL / 8;            // 16-byte chunks per copy
  int it = 0;
  for (long long pu = pair_in_grp; pu < n_pair_units; pu += pairs_in_grp, ++it) {
    const long long u = 2 * pu + rank;
    const bool valid = u < n_units;                // odd unit count: the last pair's rank 1 runs on zeros
    const int b = valid ? (int)(u / g.n_tiles) : 0, tile = valid ? (int)(u % g.n_tiles) : 0;
    const long long ts = g.te_lo + (long long)tile * TILE;
    const size_t xrow = (size_t)b * g.ldx;
    if (valid && rdy.ready != nullptr && ptid == 0) {       // clip b still in flight over PCIe?
      const int* flag = rdy.ready + b / rdy.clips_per_flag;
      int v;
      unsigned spins = 0;
      do {
        asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (v == 0) { __nanosleep(200); if (++spins > (1u << 24)) __trap(); }
      } while (v == 0);
    }
    named_bar_sync(BAR_PROD, PROD_THREADS);        // clip b resident; max scratch of the previous tile consumed
    // this thread's samples: 16 per chunk (staged index 8 jj .. 8 jj + 15, sample ts - padL + index)
    float v[NC][16];
    float mx = 0.f;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const int jj = ptid + c * PROD_THREADS;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int idx = 8 * jj + i;
        const long long a = ts - g.padL + idx, wi = a - g.t_off;
        float val = 0.f;
        if (valid && jj < nchunk && idx < sp.LX && a >= 0 && a < g.T_total && wi >= 0 && wi < g.T_win)
          val = load_sample(x, xrow, wi, g.x_fmt);   // coherent load: may have just landed
        v[c][i] = val;
        mx = fmaxf(mx, fabsf(val));
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) misc->red[warp - PROD_WARP0] = mx;
    named_bar_sync(BAR_PROD, PROD_THREADS);
    mx = fmaxf(misc->red[0], fmaxf(misc->red[1], misc->red[2]));
    int sx = 0;
    if (mx > 0.f && mx < 3.0e38f) {
      int ex;
      (void)frexpf(mx, &ex);                       // mx = m * 2^ex, m in [0.5,1)
      sx = 14 - ex;                                // mx * 2^sx in [2^13, 2^14)
      sx = sx < -100 ? -100 : (sx > 100 ? 100 : sx);
    }
    if (ptid == 0) misc->sx_ring[it & 3] = sx;
    const float scale = ldexpf(1.0f, sx);
    uint32_t wh[NC][8], wl[NC][8];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float a0 = v[c][2 * i] * scale, a1 = v[c][2 * i + 1] * scale;
        const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
        const __half l0 = __float2half_rn(a0 - __half2float(h0)), l1 = __float2half_rn(a1 - __half2float(h1));
        wh[c][i] = (uint32_t)__half_as_ushort(h0) | ((uint32_t)__half_as_ushort(h1) << 16);
        wl[c][i] = (uint32_t)__half_as_ushort(l0) | ((uint32_t)__half_as_ushort(l1) << 16);
      }
    }
    build_copy_regs<0, NC>(misc, wh, wl, s_acopy, sp.acb, nchunk, ptid, lane, it);
    build_copy_regs<1, NC>(misc, wh, wl, s_acopy, sp.acb, nchunk, ptid, lane, it);
    build_copy_regs<2, NC>(misc, wh, wl, s_acopy, sp.acb, nchunk, ptid, lane, it);
    build_copy_regs<3, NC>(misc, wh, wl, s_acopy, sp.acb, nchunk, ptid, lane, it);
    build_copy_regs<4, NC>(misc, wh, wl, s_acopy, sp.acb, nchunk, ptid, lane, it);
    build_copy_regs<5, NC>(misc, wh, wl, s_acopy, sp.acb, nchunk, ptid, lane, it);
    build_copy_regs<6, NC>(misc, wh, wl, s_acopy, sp.acb, nchunk, ptid, lane, it);
    build_copy_regs<7, NC>(misc, wh, wl, s_acopy, sp.acb, nchunk, ptid, lane, it);
  }
}

// KS > 0: number of k-steps known at compile time (26 for the default 401-tap window): the MMA issue loop is
// fully unrolled with immediate descriptor offsets -- with a runtime trip count the per-iteration descriptor
// arithmetic made the single issuing lane the bottleneck (149 cycles per k-step measured vs 124 issued tight).
template <int CG, int NSLOT, int MODE, int KS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc::NTHREADS, 1)
k1_tc_kernel(const Geom g, const float* __restrict__ x, const uint8_t* __restrict__ w16,
             const float* __restrict__ cprm, float* __restrict__ ppart, int n_groups, const TcBwdArgs ba,
             const TcReady rdy, const TcMap tm) {
  using namespace tc;
  constexpr int NB = 2 * CG;                 // accumulator columns per stage (hi | lo products)
  constexpr int NST = (512 / NB) > 4 ? 4 : (512 / NB);
  constexpr int FPT = CG / 4;                // filters per epilogue thread
  constexpr uint32_t IDESC_MAIN = idesc_f16(256, NB);   // M = 256: 128 rows from each CTA of the pair
  constexpr uint32_t IDESC_CORR = idesc_f16(256, CG);

  extern __shared__ __align__(1024) uint8_t smem[];
  const SmemPlan sp = smem_plan(CG, g.Kp, g.SL, MODE, NSLOT);
  uint8_t* s_w = smem + sp.off_w;
  uint8_t* s_acopy = smem + sp.off_acopy;
  float* s_pw = reinterpret_cast<float*>(smem + sp.off_pw);
  float* s_red = reinterpret_cast<float*>(smem + sp.off_red);
  int4* s_out = reinterpret_cast<int4*>(smem + sp.off_out);
  Misc* misc = reinterpret_cast<Misc*>(smem + sp.off_misc);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long perf_c0 = 0, perf_t0 = 0;
  if (rdy.perf != nullptr && blockIdx.x == 0 && tid == 0) {
    perf_c0 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(perf_t0));
  }
  // CTA pair = cluster of 2 (same TPC).  Both CTAs serve the same channel group; the pair takes two units
  // (tiles) per iteration, rank r the unit 2*pair_unit + r.  Rank 0 issues the MMAs for both.
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int grp = pair % n_groups;
  const int pair_in_grp = pair / n_groups;
  const int pairs_in_grp = (n_pairs - grp + n_groups - 1) / n_groups;
  const long long n_units = (long long)g.B * g.n_tiles;
  const long long n_pair_units = (n_units + 1) / 2;
  const int cta_in_grp = pair_in_grp * 2 + (int)rank;    // row of the backward partial buffer
  const int ksteps = g.Kp / KSTEP;

  // ---- one-time setup ---------------------------------------------------------------------------
  // Programmatic dependent launch (forward): this grid may have been scheduled while the bank prologue k0 was still
  // running -- wait for its completion before reading anything it wrote; and let the PCEN kernel be scheduled as soon
  // as SMs free up at the tail of this grid (it waits for our completion itself).
  if constexpr (MODE == 0) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  }
  {
    // this CTA's bank regions: R1 (main MMA) then R2 (corr MMA), from the group's global image
    const uint8_t* gimg = w16 + (size_t)grp * b_group_bytes(CG, g.Kp);
    const size_t r1 = r1_bytes(CG, g.Kp), r2 = r2_bytes(CG, g.Kp);
    const uint4* src1 = reinterpret_cast<const uint4*>(gimg + rank * r1);
    const uint4* src2 = reinterpret_cast<const uint4*>(gimg + 2 * r1 + rank * r2);
    uint4* dst = reinterpret_cast<uint4*>(s_w);
    for (int i = tid; i < (int)(r1 / 16); i += NTHREADS) dst[i] = __ldg(src1 + i);
    for (int i = tid; i < (int)(r2 / 16); i += NTHREADS) dst[r1 / 16 + i] = __ldg(src2 + i);
  }
  if (tid == 0) {
    // a_full / acc_empty live on rank 0 and collect arrivals from BOTH CTAs; a_empty / acc_full are local
    // and are signalled in both CTAs by the multicast tcgen05.commit of rank 0.
    for (int p = 0; p < NPHASE; ++p) { mbar_init(&misc->a_full[p], 2 * PROD_WARPS); mbar_init(&misc->a_empty[p], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&misc->acc_full[s], 1); mbar_init(&misc->acc_empty[s], 2 * EPI_WARPS); }
    mbar_init_fence();
  }
  if (warp == MMA_WARP) tmem_alloc_pair<512>(&misc->tmem_base);
  fence_proxy_async_smem();                  // bank written with generic stores, read by the MMA (async proxy)
  tc_fence_before();
  cluster_sync_all();                        // barriers + TMEM of both CTAs ready before any remote arrive / MMA
  tc_fence_after();
  const uint32_t tmem = misc->tmem_base;

  if (warp >= PROD_WARP0) {
    // =========================================== PRODUCERS ======================================
    const int cpt = (sp.CL / 8 + PROD_THREADS - 1) / PROD_THREADS;     // 16-byte chunks of a copy per producer thread
    if (cpt <= 2)
      producer_loop<2>(g, x, rdy, sp, misc, s_acopy, tid, lane, warp, rank, pair_in_grp, pairs_in_grp, n_units, n_pair_units);
    else if (cpt == 3)
      producer_loop<3>(g, x, rdy, sp, misc, s_acopy, tid, lane, warp, rank, pair_in_grp, pairs_in_grp, n_units, n_pair_units);
    else
      producer_loop<4>(g, x, rdy, sp, misc, s_acopy, tid, lane, warp, rank, pair_in_grp, pairs_in_grp, n_units, n_pair_units);
  } else if (warp == MMA_WARP) {
    // =========================================== MMA ISSUER (rank 0 only) =======================
    if (rank == 0) {
      const bool leader = elect_one();
      const uint32_t a_base = smem_u32(s_acopy), w_base = smem_u32(s_w);
      const uint64_t b1_desc0 = smem_desc(w_base, CG * 16, 128);                               // R1: CG rows per CTA
      const uint64_t b2_desc0 = smem_desc(w_base + (uint32_t)r1_bytes(CG, g.Kp), (CG / 2) * 16, 128);  // R2: CG/2 rows
      const uint64_t b1_step = (uint64_t)((CG * 32) >> 4), b2_step = (uint64_t)(((CG / 2) * 32) >> 4);
      // Forward: zone bounds of this channel group (k0's support pruning), made warp-uniform with a redux so
      // that the zone loops run on uniform registers.
#if (LEAFK_EXP & 64)
      long long dbg_t[3] = {0, 0, 0};
#endif
      constexpr int LMAX = CG / 16;
      int zlo[LMAX], zhi[LMAX], z3r[LMAX], z3f[LMAX];
      if constexpr (MODE == 0) {
        const int* z = tm.zones + (size_t)grp * tc::ZONE_INTS;
#pragma unroll
        for (int L = 0; L < LMAX; ++L) {
          zlo[L] = __reduce_max_sync(0xffffffffu, __ldg(z + 2 * L));
          zhi[L] = __reduce_max_sync(0xffffffffu, __ldg(z + 2 * L + 1));
          z3r[L] = __reduce_max_sync(0xffffffffu, __ldg(z + 16 + 2 * L));
          z3f[L] = __reduce_max_sync(0xffffffffu, __ldg(z + 16 + 2 * L + 1));
        }
      }
      int it = 0;
      for (long long pu = pair_in_grp; pu < n_pair_units; pu += pairs_in_grp, ++it) {
#pragma unroll 1
        for (int p = 0; p < NPHASE; ++p) {
          const int gp = it * NPHASE + p;
          const int st = gp % NST;
#if (LEAFK_EXP & 64)
          const long long tq0 = clock64();
#endif
          mbar_wait_cluster(&misc->a_full[p], (uint32_t)(it & 1));
#if (LEAFK_EXP & 64)
          const long long tq1 = clock64();
#endif
          mbar_wait_cluster(&misc->acc_empty[st], (uint32_t)(((gp / NST) & 1) ^ 1));
#if (LEAFK_EXP & 64)
          const long long tq2 = clock64();
          dbg_t[0] += tq1 - tq0; dbg_t[1] += tq2 - tq1;
#endif
          tc_fence_after();
          const uint32_t d = tmem + (uint32_t)(st * NB);
          const uint64_t a_hi = smem_desc(a_base + (uint32_t)((2 * p) * sp.acb), 16, 128);
          const uint64_t a_lo = smem_desc(a_base + (uint32_t)((2 * p + 1) * sp.acb), 16, 128);
          if constexpr (MODE == 0) {
            if (leader) {
              // centre zone first (every channel; its first MMA initialises all 2*CG accumulator columns)
              issue_zone<CG, CG>(d, a_hi, a_lo, b1_desc0, b2_desc0, zlo[LMAX - 1], zhi[LMAX - 1] + 1, CG, 0);
              issue_outer_zones<CG, LMAX - 1>(d, a_hi, a_lo, b1_desc0, b2_desc0, zlo, zhi, z3r, z3f);
              mma_commit_pair(&misc->a_empty[p]);
              mma_commit_pair(&misc->acc_full[st]);
            }
          } else if (leader) {
            if (MODE == 1 && ba.skip_xlo) {
#pragma unroll 4
              for (int ks = 0; ks < ksteps; ++ks)
                mma_f16_ss_pair(d, a_hi + (uint64_t)(2 * ks), b1_desc0 + (uint64_t)ks * b1_step, IDESC_MAIN, ks > 0);
            } else if constexpr (KS > 0) {
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                mma_f16_ss_pair(d, a_hi + (uint64_t)(2 * ks), b1_desc0 + (uint64_t)(ks * ((CG * 32) >> 4)), IDESC_MAIN, ks > 0);
                mma_f16_ss_pair(d, a_lo + (uint64_t)(2 * ks), b2_desc0 + (uint64_t)(ks * (((CG / 2) * 32) >> 4)), IDESC_CORR, 1);
              }
            } else {
#pragma unroll 2
              for (int ks = 0; ks < ksteps; ++ks) {
                mma_f16_ss_pair(d, a_hi + (uint64_t)(2 * ks), b1_desc0 + (uint64_t)ks * b1_step, IDESC_MAIN, ks > 0);
                mma_f16_ss_pair(d, a_lo + (uint64_t)(2 * ks), b2_desc0 + (uint64_t)ks * b2_step, IDESC_CORR, 1);
              }
            }
            mma_commit_pair(&misc->a_empty[p]);
            mma_commit_pair(&misc->acc_full[st]);
          }
          __syncwarp();
#if (LEAFK_EXP & 64)
          dbg_t[2] += clock64() - tq2;
#endif
        }
      }
#if (LEAFK_EXP & 64)
      if (blockIdx.x == 0 && lane == 0)
        printf("MMA warp: wait a_full %lld, wait acc_empty %lld, issue %lld cycles (tiles %d)\n", dbg_t[0], dbg_t[1], dbg_t[2], it);
#endif
    }
  } else if constexpr (MODE == 0) {
    // =========================================== EPILOGUE (forward) =============================
    const int e = warp, q = e & 3, hh = e >> 2;
    const int etid = tid;                               // 0..255
    const int m = 32 * q + lane;                        // accumulator row
    // this thread's filters: sorted positions hh*FPT + i of the group; perm gives the filter they belong to
    const int* gperm = tm.perm + (size_t)grp * (CG / 2);
#if (LEAFK_EXP & 64)
    long long dbg_e[5] = {0, 0, 0, 0, 0};
#endif
    float pa[FPT];
#pragma unroll
    for (int i = 0; i < FPT; ++i) {
      const int f = __ldg(gperm + hh * FPT + i);
      pa[i] = (f < g.F) ? __ldg(cprm + (size_t)f * 8 + CP_POOLA) : -1.0f;
    }
    const float centre = 0.5f * (float)(g.K - 1);
    const int n_last = g.n_begin + g.n_count - 1;
    float* red = s_red + (size_t)e * FPT * 33;          // this warp's transpose buffer (generic tile-end row sums)
    int sig_b = -1;                                     // clip whose last stored tile K2 has not been told about yet
    const int red_fi = halving_index<FPT, 16>(lane);    // filter whose row sum the halving reduction leaves in this lane
    // output table of the tile-end store (slot fastest, the layout K2 reads); read back by the same threads only
    for (int idx = etid; idx < g.SL * (CG / 2); idx += EPI_WARPS * 32) {
      const int fl = idx / g.SL, slot = idx - fl * g.SL;
      const int f = __ldg(gperm + fl);                      // sorted position -> filter
      int4 o = make_int4(((fl / FPT) * 4 * g.SL + slot) * FPT + fl % FPT, -1, 0, 0);
      if (f < g.F) { o.y = f * g.SL + slot; o.z = (int)__ldg(cprm + (size_t)f * 8 + CP_WSCALE); }
      s_out[idx] = o;
    }
    int it = 0;
    for (long long pu = pair_in_grp; pu < n_pair_units; pu += pairs_in_grp, ++it) {
      const long long u = 2 * pu + rank;
      const bool valid = u < n_units;
      const int b = valid ? (int)(u / g.n_tiles) : 0, tile = valid ? (int)(u % g.n_tiles) : 0;
      const long long ts = g.te_lo + (long long)tile * TILE;
      const long long te = !valid ? ts : ((ts + TILE < g.te_hi) ? ts + TILE : g.te_hi);   // invalid unit: all rows masked
      const int n_first = first_frame_of(g, ts);
      const long long tb = ts + 8 * m;                  // this row's 8 samples: tb .. tb+7
      const int nb = first_frame_of(g, tb);
      float acc[FPT][NSLOT];
#pragma unroll
      for (int i = 0; i < FPT; ++i)
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) acc[i][j] = 0.f;

#pragma unroll 1
      for (int p = 0; p < NPHASE; ++p) {
        const int gp = it * NPHASE + p;
        const int st = gp % NST;
        const long long t = tb + p;
        float dj[NSLOT];
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const int n = nb + j;
          const long long k = t + g.padL - (long long)n * g.H;
          const bool ok = (k >= 0) && (k < g.K) && (t < te) && (n <= n_last);
          const float kc = (float)k - centre;
          dj[j] = ok ? kc * kc : 1.0e30f;               // ex2(pa * 1e30) = 0: outside the window
        }
#if (LEAFK_EXP & 64)
        const long long te0 = clock64();
#endif
        mbar_wait(&misc->acc_full[st], (uint32_t)((gp / NST) & 1));
#if (LEAFK_EXP & 64)
        const long long te1 = clock64();
        dbg_e[0] += te1 - te0;
#endif
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * NB + hh * (CG / 2));
        const uint32_t tlo = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * NB + 2 * CG - 8 - hh * (CG / 2));
#pragma unroll
        for (int c = 0; c < (((LEAFK_EXP & 1) || ((LEAFK_EXP & 16) && q == 0) || ((LEAFK_EXP & 32) && q != 0)) ? 0 : FPT / 4); ++c) {
          // hi products of the 4 filters at columns hh*CG/2 + 8c ..; their lo products sit in the mirrored
          // 8-column block of the lo half, filters in reverse order (k1_tc_layout.cuh)
          float ym[8], yc[8];
          if constexpr ((LEAFK_EXP & 8) != 0) {            // experiment: the arithmetic without the TMEM loads
#pragma unroll
            for (int i = 0; i < 8; ++i) { ym[i] = dj[0] * (float)(i + c); yc[i] = dj[1] + (float)i; }
          } else {
            tmem_ld8x2_sync(taddr + 8 * c, tlo - 8 * c, ym, yc);
          }
          if constexpr ((LEAFK_EXP & 4) != 0) {            // experiment: the TMEM loads without the arithmetic
            acc[0][0] += ym[0] + yc[7];
            continue;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float re = ym[2 * i] + yc[2 * (3 - i)], im = ym[2 * i + 1] + yc[2 * (3 - i) + 1];
            const float en = fmaf(re, re, im * im);
#pragma unroll
            for (int j = 0; j < NSLOT; ++j)
              acc[4 * c + i][j] = fmaf(ex2_approx(pa[4 * c + i] * dj[j]), en, acc[4 * c + i][j]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_rank0(&misc->acc_empty[st]);
        if (p == 1 && sig_b >= 0) {          // the previous tile's partial sums were stored a phase ago: publish them to K2
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(tm.done + sig_b, 1);
          sig_b = -1;
        }
#if (LEAFK_EXP & 64)
        dbg_e[1] += clock64() - te1;
#endif
      }
#if (LEAFK_EXP & 64)
      const long long te2 = clock64();
#endif

      // ---- reduce over the rows of this warp: per-frame sums -> s_pw[buf][e][slot][fi] ----------
      // This section is on the critical path: while the epilogue warps are in it nobody drains the accumulator
      // stages, and the three stages cover only ~8 k cycles of MMAs.  The shared-memory pipe is ~86 % busy with
      // tensor-core operand fetches, so every dependent trip through it (shuffle level, store/load pair) costs
      // 150-350 cycles: the row sums therefore run as ONE recursive-halving reduction over all (<= 5) frames at once
      // -- 5 dependent shuffle levels per tile instead of 5 per frame or per (frame, filter).
      float* pw_buf = s_pw + (size_t)(it & 1) * (EPI_WARPS * g.SL * FPT);
      float* pw = pw_buf + (size_t)e * g.SL * FPT;
      for (int i = lane; i < g.SL * FPT; i += 32) pw[i] = 0.f;
      __syncwarp();
      const int nb_lo = __shfl_sync(0xffffffffu, nb, 0);
      int nb_hi = __shfl_sync(0xffffffffu, nb, 31) + NSLOT - 1;
      constexpr int NFR = NSLOT + 2;                        // frames the fast path covers
      if (NSLOT == 3 && nb_hi - nb_lo < NFR) {
        if (nb_hi > n_last) nb_hi = n_last;
        const int eo = nb - nb_lo;                          // 0..2: this lane's first frame relative to the warp's
        constexpr int H1 = (FPT + 1) / 2;
        const bool up = (lane & 16) != 0;
        float k1[NFR][H1];
#pragma unroll
        for (int d = 0; d < NFR; ++d) {
#pragma unroll
          for (int i = 0; i < H1; ++i) {
            float lo_v = 0.f, hi_v = 0.f;
#pragma unroll
            for (int jj = 0; jj < NSLOT; ++jj) {
              if (d - jj >= 0 && d - jj <= 2) {             // frame d is slot jj of the lanes with eo == d - jj
                lo_v = (eo == d - jj) ? acc[i][jj] : lo_v;
                if (i + H1 < FPT) hi_v = (eo == d - jj) ? acc[i + H1 < FPT ? i + H1 : 0][jj] : hi_v;
              }
            }
            k1[d][i] = (up ? hi_v : lo_v) + __shfl_xor_sync(0xffffffffu, up ? lo_v : hi_v, 16);
          }
        }
        float tot[NFR];
        halving_multi<NFR, H1, 8>(k1, lane, tot);
#pragma unroll
        for (int d = 0; d < NFR; ++d) {
          const int slot = nb_lo + d - n_first;
          if (red_fi >= 0 && nb_lo + d <= nb_hi && slot < g.SL) pw[slot * FPT + red_fi] = tot[d];
        }
      } else {
        // generic geometry (more frames per warp): row sums through the transpose buffer, one frame at a time
        if (nb_hi > n_last) nb_hi = n_last;
        for (int n = nb_lo; n <= nb_hi; ++n) {
          const int slot = n - n_first;
          if (slot >= g.SL) break;
          const int j = n - nb;
#pragma unroll
          for (int i = 0; i < FPT; ++i) {
            float v = 0.f;
#pragma unroll
            for (int jj = 0; jj < NSLOT; ++jj) v = (j == jj) ? acc[i][jj] : v;
            red[i * 33 + lane] = v;
          }
          __syncwarp();
          if (lane < FPT) {
            const float* rr = red + lane * 33;
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int r = 0; r < 32; r += 4) { s0 += rr[r]; s1 += rr[r + 1]; s2 += rr[r + 2]; s3 += rr[r + 3]; }
            pw[slot * FPT + lane] = (s0 + s1) + (s2 + s3);
          }
          __syncwarp();
        }
      }
#if (LEAFK_EXP & 64)
      const long long te3 = clock64();
      dbg_e[3] += te3 - te2;
#endif
      named_bar_sync(BAR_EPI, EPI_WARPS * 32);
      // ---- fixed-order sum over the four row quadrants, undo the scaling, store -----------------
      const int sx = misc->sx_ring[it & 3];
      float* dst = ppart + ((size_t)b * g.n_tiles + tile) * g.SL * g.F;
      for (int idx = etid; idx < g.SL * (CG / 2); idx += EPI_WARPS * 32) {
        const int4 o = s_out[idx];
        if (valid && o.y >= 0) {
          const float* src = pw_buf + o.x;
          float s = 0.f;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) s += src[(size_t)qq * g.SL * FPT];
          dst[o.y] = scalbnf(s, -2 * (sx + o.z));
        }
      }
      sig_b = (valid && tm.done != nullptr) ? b : -1;   // published off the critical path (phase 1 of the next tile)
#if (LEAFK_EXP & 64)
      dbg_e[2] += clock64() - te2;
#endif
    }
    if (sig_b >= 0) {
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(tm.done + sig_b, 1);
    }
#if (LEAFK_EXP & 64)
    if (blockIdx.x == 0 && tid == 0)
      printf("epilogue warp 0: wait acc_full %lld, phase work %lld, tile end %lld (reduction %lld) cycles\n", dbg_e[0], dbg_e[1], dbg_e[2], dbg_e[3]);
#endif
  } else {
    // =========================================== EPILOGUE (backward) ============================
    constexpr int FB = CG / 6;                          // filters per group
    constexpr int FPB = FB / 2;                         // filters per epilogue thread
    const int e = warp, q = e & 3, hh = e >> 2;
    const int m = 32 * q + lane;
    const int fbase = grp * FB + hh * FPB;              // first filter of this thread
    float pa[FPB];
    int shy[FPB], shz[FPB], shv[FPB];
#pragma unroll
    for (int i = 0; i < FPB; ++i) {
      const float* bp = ba.bprm + (size_t)(fbase + i) * 8;
      pa[i] = __ldg(bp + 0);
      shy[i] = (int)__ldg(bp + 1); shz[i] = (int)__ldg(bp + 2); shv[i] = (int)__ldg(bp + 3);
    }
    float a_mu[FPB], a_sg[FPB], a_pw[FPB];
#pragma unroll
    for (int i = 0; i < FPB; ++i) { a_mu[i] = 0.f; a_sg[i] = 0.f; a_pw[i] = 0.f; }
    const float centre = 0.5f * (float)(g.K - 1);
    const int n_last = g.n_begin + g.n_count - 1;
    int it = 0;
    for (long long pu = pair_in_grp; pu < n_pair_units; pu += pairs_in_grp, ++it) {
      const long long u = 2 * pu + rank;
      const bool valid = u < n_units;
      const int b = valid ? (int)(u / g.n_tiles) : 0, tile = valid ? (int)(u % g.n_tiles) : 0;
      const long long ts = g.te_lo + (long long)tile * TILE;
      const long long te = !valid ? ts : ((ts + TILE < g.te_hi) ? ts + TILE : g.te_hi);   // invalid unit: all rows masked
      const long long tb = ts + 8 * m;
      const int nb = first_frame_of(g, tb);
      float dpv[NSLOT][FPB];
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) {
        const int n = nb + j;
        const bool nok = (n >= g.n_begin) && (n <= n_last);
        const float* row = ba.dpT + ((size_t)b * g.N_total + (nok ? n : 0)) * g.F;
#pragma unroll
        for (int i = 0; i < FPB; ++i) dpv[j][i] = (nok && fbase + i < g.F) ? __ldg(row + fbase + i) : 0.f;
      }
      float sy[FPB], sz[FPB], sv[FPB];
      bool have_scale = false;
#pragma unroll 1
      for (int p = 0; p < NPHASE; ++p) {
        const int gp = it * NPHASE + p;
        const int st = gp % NST;
        const long long t = tb + p;
        float dj[NSLOT];
#pragma unroll
        for (int j = 0; j < NSLOT; ++j) {
          const int n = nb + j;
          const long long k = t + g.padL - (long long)n * g.H;
          const bool ok = (k >= 0) && (k < g.K) && (t < te) && (n <= n_last);
          const float kc = (float)k - centre;
          dj[j] = ok ? kc * kc : 1.0e30f;
        }
        mbar_wait(&misc->acc_full[st], (uint32_t)((gp / NST) & 1));
        tc_fence_after();
        if (!have_scale) {                                 // sx of this tile is published before its first phase
          const int sx = misc->sx_ring[it & 3];
#pragma unroll
          for (int i = 0; i < FPB; ++i) {
            sy[i] = scalbnf(1.0f, -(sx + shy[i])); sz[i] = scalbnf(1.0f, -(sx + shz[i])); sv[i] = scalbnf(1.0f, -(sx + shv[i]));
          }
          have_scale = true;
        }
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * NB + hh * (CG / 2));
#pragma unroll
        for (int c = 0; c < FPB / 4; ++c) {
          float ym[8], yc[8], zm[8], zc[8], vm[8], vc[8];
          tmem_ld8x2_sync(taddr + 8 * c, taddr + CG + 8 * c, ym, yc);
          tmem_ld8x2_sync(taddr + FB + 8 * c, taddr + CG + FB + 8 * c, zm, zc);
          tmem_ld8x2_sync(taddr + 2 * FB + 8 * c, taddr + CG + 2 * FB + 8 * c, vm, vc);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int fi = 4 * c + i;
            const float yre = (ym[2 * i] + yc[2 * i]) * sy[fi], yim = (ym[2 * i + 1] + yc[2 * i + 1]) * sy[fi];
            const float zre = (zm[2 * i] + zc[2 * i]) * sz[fi], zim = (zm[2 * i + 1] + zc[2 * i + 1]) * sz[fi];
            const float vre = (vm[2 * i] + vc[2 * i]) * sv[fi], vim = (vm[2 * i + 1] + vc[2 * i + 1]) * sv[fi];
            float de = 0.f, dgs = 0.f;
#pragma unroll
            for (int j = 0; j < NSLOT; ++j) {
              const float wgt = ex2_approx(pa[fi] * dj[j]) * dpv[j][fi];
              de += wgt;
              dgs = fmaf(wgt, dj[j] < 1.0e29f ? dj[j] : 0.f, dgs);
            }
            const float en = fmaf(yre, yre, yim * yim);
            a_mu[fi] = fmaf(de, yim * zre - yre * zim, a_mu[fi]);
            a_sg[fi] = fmaf(de, yre * vre + yim * vim, a_sg[fi]);
            a_pw[fi] = fmaf(en, dgs, a_pw[fi]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_rank0(&misc->acc_empty[st]);
      }
    }
    // ---- CTA reduction of the 3*FPB sums per thread -> one partial row per filter ----------------
    float* red = s_pw;                                  // [8 warps][32]
#pragma unroll
    for (int i = 0; i < FPB; ++i) {
      float v0 = a_mu[i], v1 = a_sg[i], v2 = a_pw[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        v0 += __shfl_xor_sync(0xffffffffu, v0, o);
        v1 += __shfl_xor_sync(0xffffffffu, v1, o);
        v2 += __shfl_xor_sync(0xffffffffu, v2, o);
      }
      if (lane == 0) { red[e * 32 + 3 * i] = v0; red[e * 32 + 3 * i + 1] = v1; red[e * 32 + 3 * i + 2] = v2; }
    }
    named_bar_sync(BAR_EPI, EPI_WARPS * 32);
    if (tid < 2 * FPB * 3) {
      const int h2 = tid / (FPB * 3), r = tid % (FPB * 3);
      float sacc = 0.f;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) sacc += red[(h2 * 4 + qq) * 32 + r];
      const int f = grp * FB + h2 * FPB + r / 3;
      ba.bpart[((size_t)cta_in_grp * ba.Fpad + f) * 4 + (r % 3)] = sacc;
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  cluster_sync_all();                        // nobody may still signal a peer barrier / use TMEM
  if (warp == tc::MMA_WARP) tmem_dealloc_pair<512>(tmem);
  if (rdy.perf != nullptr && blockIdx.x == 0 && tid == 0) {
    long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    rdy.perf[0] = clock64() - perf_c0;
    rdy.perf[1] = t1 - perf_t0;
  }
}

// ------------------------------------------------------------------------------------------------
bool k1_tc_supported(const Geom& g, const char** why) {
  const int nslot = tc::slots_per_thread(g.K, g.H);
  if (nslot > 5) { if (why) *why = "hop too small relative to the window (more than 5 frames per 8 samples)"; return false; }
  if (g.Kp > 2048) { if (why) *why = "window longer than 2048 taps"; return false; }
  int ng, cg;
  if (!tc::channel_groups(g.C2, g.Kp, g.SL, nslot, &ng, &cg)) {
    if (why) *why = "shared-memory plan does not fit for any channel grouping";
    return false;
  }
  return true;
}

template <int CG, int NSLOT, int KS>
static cudaError_t launch_inst_ks(const Geom& g, const float* x, const uint8_t* w16, const float* cprm, float* ppart,
                                  int n_groups, int grid, int smem, cudaStream_t stream, const TcReady& rdy, const TcMap& tm) {
  // the opt-in to > 48 KB of dynamic shared memory is per function and device: raise it only when this launch needs
  // more than any earlier one asked for (a driver call per forward otherwise)
  static thread_local int smem_set[64] = {0};
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  if (dev < 0 || dev >= 64 || smem_set[dev] < smem) {
    err = cudaFuncSetAttribute(k1_tc_kernel<CG, NSLOT, 0, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    if (dev >= 0 && dev < 64) smem_set[dev] = smem;
  }
  TcBwdArgs none{nullptr, nullptr, nullptr, 0, 0};
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(tc::NTHREADS); lc.dynamicSmemBytes = (size_t)smem; lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // may start under the tail of k0 (see kernel)
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at; lc.numAttrs = 1;
  return cudaLaunchKernelEx(&lc, k1_tc_kernel<CG, NSLOT, 0, KS>, g, x, w16, cprm, ppart, n_groups, none, rdy, tm);
}
template <int CG, int NSLOT>
static cudaError_t launch_inst(const Geom& g, const float* x, const uint8_t* w16, const float* cprm, float* ppart,
                               int n_groups, int grid, int smem, cudaStream_t stream, const TcReady& rdy, const TcMap& tm) {
  return launch_inst_ks<CG, NSLOT, 0>(g, x, w16, cprm, ppart, n_groups, grid, smem, stream, rdy, tm);
}

template <int CG, int NSLOT, int KS>
static cudaError_t launch_bwd_inst_ks(const Geom& g, const float* x, const uint8_t* w16, int n_groups, int grid,
                                      const TcBwdArgs& ba, cudaStream_t stream) {
  const int smem = tc::smem_plan(CG, g.Kp, g.SL, 1).total;
  cudaError_t err = cudaFuncSetAttribute(k1_tc_kernel<CG, NSLOT, 1, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (err != cudaSuccess) return err;
  k1_tc_kernel<CG, NSLOT, 1, KS><<<grid, tc::NTHREADS, smem, stream>>>(g, x, w16, nullptr, nullptr, n_groups, ba,
                                                                       TcReady{nullptr, 1, nullptr}, TcMap{nullptr, nullptr, nullptr});
  return cudaGetLastError();
}
template <int CG, int NSLOT>
static cudaError_t launch_bwd_inst(const Geom& g, const float* x, const uint8_t* w16, int n_groups, int grid,
                                   const TcBwdArgs& ba, cudaStream_t stream) {
  if constexpr (CG == 96) {
    if (g.Kp == 26 * tc::KSTEP) return launch_bwd_inst_ks<CG, NSLOT, 26>(g, x, w16, n_groups, grid, ba, stream);
  }
  return launch_bwd_inst_ks<CG, NSLOT, 0>(g, x, w16, n_groups, grid, ba, stream);
}

static int sm_count(cudaError_t* err) {
  static int n_sm_cached[64] = {0};
  int dev = 0;
  *err = cudaGetDevice(&dev);
  if (*err != cudaSuccess) return 0;
  if (dev < 64 && n_sm_cached[dev] == 0) {
    int v = 0;
    *err = cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
    if (*err != cudaSuccess) return 0;
    n_sm_cached[dev] = v;
  }
  return dev < 64 ? n_sm_cached[dev] : 148;
}

// Backward correlation pass: FB filters per group (16 -> CG 96, 8 -> CG 48).  Returns the number of CTAs per
// group through *ctas_per_group (rows of bpart the final reduction must add).
cudaError_t launch_k1_tc_bwd(const Geom& g, const float* x, const uint8_t* w16b, int FB, int n_groups,
                             const float* dpT, const float* bprm, float* bpart, int* ctas_per_group,
                             int skip_xlo, cudaStream_t stream) {
  cudaError_t err;
  const int n_sm = sm_count(&err);
  if (err != cudaSuccess) return err;
  const int grid = tc::pair_grid(n_sm, n_groups, (long long)g.B * g.n_tiles);
  *ctas_per_group = 2 * (((grid / 2) + n_groups - 1) / n_groups);   // rows of bpart (zero-filled by the caller)
  TcBwdArgs ba{dpT, bprm, bpart, n_groups * FB, skip_xlo};
  const int nslot = tc::slots_per_thread(g.K, g.H);
  if (FB == 16 && nslot <= 3) return launch_bwd_inst<96, 3>(g, x, w16b, n_groups, grid, ba, stream);
  if (FB == 8 && nslot <= 3) return launch_bwd_inst<48, 3>(g, x, w16b, n_groups, grid, ba, stream);
  if (FB == 8 && nslot <= 5) return launch_bwd_inst<48, 5>(g, x, w16b, n_groups, grid, ba, stream);
  return cudaErrorNotSupported;
}

template <int CG>
static cudaError_t launch_cg(int nslot, const Geom& g, const float* x, const uint8_t* w16, const float* cprm,
                             float* ppart, int n_groups, int grid, int smem, cudaStream_t stream, const TcReady& rdy, const TcMap& tm) {
  if (nslot <= 3) return launch_inst<CG, 3>(g, x, w16, cprm, ppart, n_groups, grid, smem, stream, rdy, tm);
  if constexpr (CG <= 64) return launch_inst<CG, 5>(g, x, w16, cprm, ppart, n_groups, grid, smem, stream, rdy, tm);
  return cudaErrorNotSupported;
}

cudaError_t launch_k1_tc(const Geom& g, const float* x, const uint8_t* w16, const float* cprm, float* ppart,
                         int tc_cg, int tc_groups, const int* tc_perm, const int* tc_zones, int* done, cudaStream_t stream,
                         const int* ready, int clips_per_flag, long long* perf) {
  const TcReady rdy{ready, clips_per_flag < 1 ? 1 : clips_per_flag, perf};
  const TcMap tm{done, tc_perm, tc_zones};
  cudaError_t err;
  const int n_sm = sm_count(&err);
  if (err != cudaSuccess) return err;
  const int grid = tc::pair_grid(n_sm, tc_groups, (long long)g.B * g.n_tiles);
  const int nslot = tc::slots_per_thread(g.K, g.H);
  const int smem = tc::smem_plan(tc_cg, g.Kp, g.SL, 0, nslot > 3 ? 5 : 3).total;
  switch (tc_cg) {
    case 16: return launch_cg<16>(nslot, g, x, w16, cprm, ppart, tc_groups, grid, smem, stream, rdy, tm);
    case 32: return launch_cg<32>(nslot, g, x, w16, cprm, ppart, tc_groups, grid, smem, stream, rdy, tm);
    case 48: return launch_cg<48>(nslot, g, x, w16, cprm, ppart, tc_groups, grid, smem, stream, rdy, tm);
    case 64: return launch_cg<64>(nslot, g, x, w16, cprm, ppart, tc_groups, grid, smem, stream, rdy, tm);
    case 80: return launch_cg<80>(nslot, g, x, w16, cprm, ppart, tc_groups, grid, smem, stream, rdy, tm);
    case 96: return launch_cg<96>(nslot, g, x, w16, cprm, ppart, tc_groups, grid, smem, stream, rdy, tm);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace leafk
