// K1 (tensor-core variant), kernel template -- included by k1_tc.cu (host side) and k1_tc_inst_*.cu (instantiations).
// Gabor correlation as a Toeplitz GEMM on tcgen05, fused with the squared
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
// Registers: the CTA launches with 168 per thread; warps 8-11 (MMA issuer + producers) give some back
// (setmaxnreg.dec to 136) and the two epilogue warpgroups take them (setmaxnreg.inc to 184), so the accumulator-heavy
// epilogues (up to 96 accumulators per thread in training) do not spill.
//
// MODE 1 = TRAINING forward (Leaf.forward when parameters require grad).  Besides y it runs the two derivative
// banks z = x*(tau h), v = x*((tau^2/sigma^3 - 1/sigma) h) ("equivalent without forming dW") and pools,
// with the SAME Gaussian windows as the energy, the three bilinear forms the parameter gradients need:
//      Q_mu[n]    = sum_t g[k] (y_im z_re - y_re z_im)       dL/dmu    = 2 sum_n dp[n] Q_mu[n]
//      Q_sigma[n] = sum_t g[k] (y_re v_re + y_im v_im)       dL/dsigma = 2 sum_n dp[n] Q_sigma[n]
//      Q_pw[n]    = sum_t g[k] (k - c)^2 e[t]                dL/ds     = sum_n dp[n] Q_pw[n] / (s^3 c^2)
// because sum_t de[t] q[t] with de[t] = sum_n dp[n] g[t - t_n] is sum_n dp[n] (pooled q)[n].  The backward pass then
// needs no correlation at all (bwd.cu): forward + backward cost 3 correlations instead of 1 + 3.
#pragma once
#include "leafk_common.cuh"
#include "k1_tc_layout.cuh"
#include "tc_ptx.cuh"

#include <cuda_fp16.h>
#include <cstdio>
#include <cstring>


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
  uint64_t bank_full;         // this CTA's bank has landed (bulk copy, complete_tx)
  uint64_t bank_pair;         // rank 0: both CTAs' banks have landed
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
    if (jj < nchunk) {
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

// Optional host-pipelining hook: when `ready` is non-null the producers wait, before touching clip b, until
// ready[b / clips_per_flag] != 0.  The flags are set by stream-ordered 32-bit writes that follow each slice of
// the H2D copy on another stream, so ONE persistent launch overlaps the whole PCIe transfer (leafk_forward_host).
// A flag that does not arrive within ~10 s (copy stalled, stream torn down) does not kill the context: the producer
// records LEAFK_ASYNC_H2D_TIMEOUT in *err, goes on with whatever is in the buffer, and the host API reports the
// error word after its next synchronisation (leafk_async_status).
struct TcReady {
  const int* ready;
  int clips_per_flag;
  long long* perf;      // optional: CTA 0 stores {SM cycles, nanoseconds} of its lifetime (effective SM clock)
  int* err;             // asynchronous error word in the workspace (0 = none), may be null
};

// Forward: the width-sorted channel order and the per-k-step active channel counts written by k0
// (k1_tc_layout.cuh, "SUPPORT PRUNING").  Both modes: the per-clip completion counters read by K2.
struct TcMap {
  int* done;            // [B] per-clip completion counters for K2 (one increment per epilogue warp and stored tile)
  const int* perm;      // [n_groups * CG/2] sorted position -> filter index (>= F: padding)
  const int* zones;     // [n_groups][tc::ZONE_INTS]: ints [0,16) {lo_L, hi_L}, L = 1..CG/16: k-steps with >= 16 L channels
                        // running; ints [16,32) {na3 of level L's rising zone, of its falling zone}
};

// Training forward (MODE 1): per-filter constants written by k0_banks_train_kernel.
struct TcTrainArgs {
  const float* tprm;    // (Fpad, 8): [0] pooling exp2 coefficient, [1..3] power-of-two shifts of the y,z,v banks
  int Fpad;             // n_groups * FB
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
  bool gave_up = false;                    // a ready flag timed out: stop waiting for the others too
  int it = 0;
  for (long long pu = pair_in_grp; pu < n_pair_units; pu += pairs_in_grp, ++it) {
    const long long u = 2 * pu + rank;
    const bool valid = u < n_units;                // odd unit count: the last pair's rank 1 runs on zeros
    const int b = valid ? (int)(u / g.n_tiles) : 0, tile = valid ? (int)(u % g.n_tiles) : 0;
    const long long ts = g.te_lo + (long long)tile * TILE;
    const ClipView cv = clip_view(g, b);
    if (valid && rdy.ready != nullptr && ptid == 0 && !gave_up) {       // clip b still in flight over PCIe?
      const int* flag = rdy.ready + b / rdy.clips_per_flag;
      int v;
      asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (v == 0) {
        const long long t0 = global_timer_ns();
        do {
          __nanosleep(200);
          asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
          if (v == 0 && global_timer_ns() - t0 > H2D_TIMEOUT_NS) {      // stalled copy: report, do not trap
            if (rdy.err != nullptr) atomicExch(rdy.err, LEAFK_ASYNC_H2D_TIMEOUT);
            gave_up = true;
            break;
          }
        } while (v == 0);
      }
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
          val = clip_sample(g, x, cv, wi);           // coherent load: may have just landed
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


// ---- tile end, part 1: row sums over the 32 rows of a warp ----------------------------------------------------
// acc[vi][j]: this row's partial pooled sums of NV "virtual filters" (forward: CG/4 filters; training: FB/2 filters
// x 4 pooled quantities) for its NSLOT frame slots (frames nb .. nb+NSLOT-1).  Writes the warp's sums to
// pw[slot * NV + vi], slot = frame - n_first.
// This section is on the critical path: while the epilogue warps are in it nobody drains the accumulator stages.
// The shared-memory pipe is ~86 % busy with tensor-core operand fetches, so every dependent trip through it (shuffle
// level, store/load pair) costs 150-350 cycles: the row sums therefore run as ONE recursive-halving reduction over
// all (<= 5) frames at once -- 5 dependent shuffle levels per tile instead of 5 per frame or per (frame, filter).
template <int NV, int NSLOT>
__device__ __forceinline__ void tile_end_rowsums(const float (&acc)[NV][NSLOT], float* pw, float* red, int lane, int nb,
                                                 int n_first, int n_last, int SL) {
  for (int i = lane; i < SL * NV; i += 32) pw[i] = 0.f;
  __syncwarp();
  const int nb_lo = __shfl_sync(0xffffffffu, nb, 0);
  int nb_hi = __shfl_sync(0xffffffffu, nb, 31) + NSLOT - 1;
  constexpr int NFR = NSLOT + 2;                        // frames the fast path covers
  if (NSLOT == 3 && nb_hi - nb_lo < NFR) {
    if (nb_hi > n_last) nb_hi = n_last;
    const int red_vi = halving_index<NV, 16>(lane);     // virtual filter whose row sum the reduction leaves in this lane
    const int eo = nb - nb_lo;                          // 0..2: this lane's first frame relative to the warp's
    constexpr int H1 = (NV + 1) / 2;
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
            if (i + H1 < NV) hi_v = (eo == d - jj) ? acc[i + H1 < NV ? i + H1 : 0][jj] : hi_v;
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
      if (red_vi >= 0 && nb_lo + d <= nb_hi && slot < SL) pw[slot * NV + red_vi] = tot[d];
    }
  } else {
    // generic geometry (more frames per warp): row sums through the transpose buffer, one frame at a time
    if (red == nullptr) __trap();                       // lean plan: the host admits only geometries that never get here
    if (nb_hi > n_last) nb_hi = n_last;
    for (int n = nb_lo; n <= nb_hi; ++n) {
      const int slot = n - n_first;
      if (slot >= SL) break;
      const int j = n - nb;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v = 0.f;
#pragma unroll
        for (int jj = 0; jj < NSLOT; ++jj) v = (j == jj) ? acc[i][jj] : v;
        red[i * 33 + lane] = v;
      }
      __syncwarp();
      for (int vi = lane; vi < NV; vi += 32) {
        const float* rr = red + vi * 33;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int r = 0; r < 32; r += 4) { s0 += rr[r]; s1 += rr[r + 1]; s2 += rr[r + 2]; s3 += rr[r + 3]; }
        pw[slot * NV + vi] = (s0 + s1) + (s2 + s3);
      }
      __syncwarp();
    }
  }
}

// ---- tile end, part 2: fixed-order sum over the four row quadrants, undo the power-of-two scaling, store ------
// s_out[idx] = {offset of quadrant 0's sum in the reduction buffer, offset in the tile's partial-sum block or -1,
// bank exponent, 0}; the stored value is sum * 2^-(2 sx + exponent).
__device__ __forceinline__ void tile_end_store(const float* pw_buf, const int4* s_out, int n_entries, int etid,
                                               bool valid, float* dst, int sx, int qstride) {
  for (int idx = etid; idx < n_entries; idx += tc::EPI_WARPS * 32) {
    const int4 o = s_out[idx];
    if (valid && o.y >= 0) {
      const float* src = pw_buf + o.x;
      float s = 0.f;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) s += src[(size_t)qq * qstride];
      dst[o.y] = scalbnf(s, -(2 * sx + o.z));
    }
  }
}

// KS > 0: number of k-steps known at compile time (26 for the default 401-tap window): the training MMA issue loop is
// fully unrolled with immediate descriptor offsets -- with a runtime trip count the per-iteration descriptor
// arithmetic made the single issuing lane the bottleneck (149 cycles per k-step measured vs 124 issued tight).
// BIGP: producers keep 3-4 chunks per thread (windows longer than ~520 taps) and need more registers.
template <int CG, int NSLOT, int MODE, int KS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(tc::NTHREADS, 1)
k1_tc_kernel(const Geom g, const float* __restrict__ x, const uint8_t* __restrict__ w16,
             const float* __restrict__ cprm, float* __restrict__ ppart, int n_groups, const TcTrainArgs ta,
             const TcReady rdy, const TcMap tm) {
  using namespace tc;
  constexpr int NB = (MODE == 0) ? 2 * CG : CG;   // accumulator columns per stage (forward: hi | lo products)
  constexpr int NST = (512 / NB) > 4 ? 4 : (512 / NB);
  constexpr int NV = virt_per_thread(CG, MODE);   // virtual filters per epilogue thread

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
    perf_t0 = global_timer_ns();
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
  const int ksteps = g.Kp / KSTEP;
  const int cpt = (sp.CL / 8 + PROD_THREADS - 1) / PROD_THREADS;     // 16-byte chunks of a copy per producer thread

  // ---- one-time setup ---------------------------------------------------------------------------
  if (tid == 0) {
    // a_full / acc_empty / bank_pair live on rank 0 and collect arrivals from BOTH CTAs; a_empty / acc_full are local
    // and are signalled in both CTAs by the multicast tcgen05.commit of rank 0.
    for (int p = 0; p < NPHASE; ++p) { mbar_init(&misc->a_full[p], 2 * PROD_WARPS); mbar_init(&misc->a_empty[p], 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(&misc->acc_full[s], 1); mbar_init(&misc->acc_empty[s], 2 * EPI_WARPS); }
    mbar_init(&misc->bank_full, 1);
    mbar_init(&misc->bank_pair, 2);
    mbar_init_fence();
  }
  if (warp == MMA_WARP) tmem_alloc_pair<512>(&misc->tmem_base);
  tc_fence_before();
  cluster_sync_all();                        // barriers + TMEM of both CTAs ready before any remote arrive / MMA
  tc_fence_after();
  const uint32_t tmem = misc->tmem_base;
  // Programmatic dependent launch: this grid may have been scheduled while the bank prologue k0 was still running --
  // wait for its completion before reading anything it wrote; and let the PCEN kernel be scheduled as soon as SMs
  // free up at the tail of this grid (it waits for our completion itself).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp >= MMA_WARP) {
    // ---- warpgroup 2: MMA issuer (warp 8) + producers (warps 9-11): give registers to the epilogue warpgroups
    if (cpt <= 2) setmaxnreg_dec<136>(); else setmaxnreg_dec<152>();
    if (warp >= PROD_WARP0) {
      // =========================================== PRODUCERS ======================================
      if (cpt <= 2)
        producer_loop<2>(g, x, rdy, sp, misc, s_acopy, tid, lane, warp, rank, pair_in_grp, pairs_in_grp, n_units, n_pair_units);
      else if (cpt == 3)
        producer_loop<3>(g, x, rdy, sp, misc, s_acopy, tid, lane, warp, rank, pair_in_grp, pairs_in_grp, n_units, n_pair_units);
      else
        producer_loop<4>(g, x, rdy, sp, misc, s_acopy, tid, lane, warp, rank, pair_in_grp, pairs_in_grp, n_units, n_pair_units);
    } else {
      // =========================================== MMA ISSUER ====================================
      const bool leader = elect_one();
      // this CTA's bank image: one bulk copy (TMA) per region, straight into the operand layout k0 wrote; the
      // generic proxy never touches the bank
      {
        size_t off1, len1, off2 = 0, len2 = 0;
        if constexpr (MODE == 0) {
          // R1 (main MMA) then R2 (corr MMA), from the group's global image [R1 cta0 | R1 cta1 | R2 cta0 | R2 cta1]
          const size_t r1 = r1_bytes(CG, g.Kp), r2 = r2_bytes(CG, g.Kp);
          off1 = (size_t)grp * b_group_bytes(CG, g.Kp) + rank * r1; len1 = r1;
          off2 = (size_t)grp * b_group_bytes(CG, g.Kp) + 2 * r1 + rank * r2; len2 = r2;
        } else {
          off1 = (size_t)grp * t_group_bytes(CG, g.Kp) + rank * t_cta_bytes(CG, g.Kp); len1 = t_cta_bytes(CG, g.Kp);
        }
        if (leader) {
          mbar_expect_tx(&misc->bank_full, (uint32_t)(len1 + len2));
          // <= 64 KB per copy
          for (size_t o = 0; o < len1; o += 65536)
            bulk_copy_g2s(s_w + o, w16 + off1 + o, (uint32_t)(len1 - o < 65536 ? len1 - o : 65536), &misc->bank_full);
          for (size_t o = 0; o < len2; o += 65536)
            bulk_copy_g2s(s_w + len1 + o, w16 + off2 + o, (uint32_t)(len2 - o < 65536 ? len2 - o : 65536), &misc->bank_full);
        }
        __syncwarp();
        mbar_wait(&misc->bank_full, 0);
        if (lane == 0) mbar_arrive_rank0(&misc->bank_pair);
      }
      if (rank == 0) {
        mbar_wait_cluster(&misc->bank_pair, 0);          // the peer's half of the bank rows is resident too
        const uint32_t a_base = smem_u32(s_acopy), w_base = smem_u32(s_w);
        // Forward: zone bounds of this channel group (k0's support pruning), made warp-uniform with a redux so
        // that the zone loops run on uniform registers.
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
        // forward: R1 = CG rows per CTA, R2 = CG/2 rows; training: HI and LO regions of CG/2 rows each
        const uint64_t b1_desc0 = (MODE == 0) ? smem_desc(w_base, CG * 16, 128) : smem_desc(w_base, (CG / 2) * 16, 128);
        const uint64_t b2_desc0 = (MODE == 0)
            ? smem_desc(w_base + (uint32_t)r1_bytes(CG, g.Kp), (CG / 2) * 16, 128)
            : smem_desc(w_base + (uint32_t)t_region_bytes(CG, g.Kp), (CG / 2) * 16, 128);
        constexpr uint32_t IDESC_TRAIN = idesc_f16(256, CG);   // M = 256: 128 rows from each CTA of the pair
        constexpr uint32_t T_STEP = (uint32_t)(((CG / 2) * 32) >> 4);  // descriptor units per k-step slab (training)
        int it = 0;
        for (long long pu = pair_in_grp; pu < n_pair_units; pu += pairs_in_grp, ++it) {
#pragma unroll 1
          for (int p = 0; p < NPHASE; ++p) {
            const int gp = it * NPHASE + p;
            const int st = gp % NST;
            mbar_wait_cluster(&misc->a_full[p], (uint32_t)(it & 1));
            mbar_wait_cluster(&misc->acc_empty[st], (uint32_t)(((gp / NST) & 1) ^ 1));
            tc_fence_after();
            const uint32_t d = tmem + (uint32_t)(st * NB);
            const uint64_t a_hi = smem_desc(a_base + (uint32_t)((2 * p) * sp.acb), 16, 128);
            const uint64_t a_lo = smem_desc(a_base + (uint32_t)((2 * p + 1) * sp.acb), 16, 128);
            if (leader) {
              if constexpr (MODE == 0) {
                // centre zone first (every channel; its first MMA initialises all 2*CG accumulator columns)
                issue_zone<CG, CG>(d, a_hi, a_lo, b1_desc0, b2_desc0, zlo[LMAX - 1], zhi[LMAX - 1] + 1, CG, 0);
                issue_outer_zones<CG, LMAX - 1>(d, a_hi, a_lo, b1_desc0, b2_desc0, zlo, zhi, z3r, z3f);
              } else if constexpr (KS > 0) {
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                  mma_f16_ss_pair(d, a_hi + (uint64_t)(2 * ks), b1_desc0 + (uint64_t)(ks * T_STEP), IDESC_TRAIN, ks > 0);   // x_hi * W_hi
                  mma_f16_ss_pair(d, a_hi + (uint64_t)(2 * ks), b2_desc0 + (uint64_t)(ks * T_STEP), IDESC_TRAIN, 1);        // x_hi * W_lo
                  mma_f16_ss_pair(d, a_lo + (uint64_t)(2 * ks), b1_desc0 + (uint64_t)(ks * T_STEP), IDESC_TRAIN, 1);        // x_lo * W_hi
                }
              } else {
#pragma unroll 2
                for (int ks = 0; ks < ksteps; ++ks) {
                  mma_f16_ss_pair(d, a_hi + (uint64_t)(2 * ks), b1_desc0 + (uint64_t)ks * T_STEP, IDESC_TRAIN, ks > 0);
                  mma_f16_ss_pair(d, a_hi + (uint64_t)(2 * ks), b2_desc0 + (uint64_t)ks * T_STEP, IDESC_TRAIN, 1);
                  mma_f16_ss_pair(d, a_lo + (uint64_t)(2 * ks), b1_desc0 + (uint64_t)ks * T_STEP, IDESC_TRAIN, 1);
                }
              }
              mma_commit_pair(&misc->a_empty[p]);
              mma_commit_pair(&misc->acc_full[st]);
            }
            __syncwarp();
          }
        }
      }
    }
  } else {
    // =========================================== EPILOGUE =========================================
    if (cpt <= 2) setmaxnreg_inc<184>(); else setmaxnreg_inc<176>();
    const int e = warp, q = e & 3, hh = e >> 2;
    const int etid = tid;                               // 0..255
    const int m = 32 * q + lane;                        // accumulator row
    const float centre = 0.5f * (float)(g.K - 1);
    const int n_last = g.n_begin + g.n_count - 1;
    constexpr bool LEAN = lean_plan(CG, MODE);
    float* red = LEAN ? nullptr : s_red + (size_t)e * NV * 33;   // this warp's transpose buffer (generic tile-end row sums)
    int sig_b = -1;                                     // clip whose last stored tile K2 has not been told about yet
    const int FV = (MODE == 0) ? g.F : 4 * g.F;         // virtual filters per (clip, tile) block of partial sums

    // per-thread filter constants and the output table of the tile-end store (slot fastest, the layout K2 reads);
    // the table is read back by the same threads only
    constexpr int NF = (MODE == 0) ? NV : NV / 4;       // real filters per thread
    float pa[NF];
    if constexpr (MODE == 0) {
      // this thread's filters: sorted positions hh*NV + i of the group; perm gives the filter they belong to
      const int* gperm = tm.perm + (size_t)grp * (CG / 2);
#pragma unroll
      for (int i = 0; i < NF; ++i) {
        const int f = __ldg(gperm + hh * NV + i);
        pa[i] = (f < g.F) ? __ldg(cprm + (size_t)f * 8 + CP_POOLA) : -1.0f;
      }
      for (int idx = etid; idx < g.SL * (CG / 2); idx += EPI_WARPS * 32) {
        const int fl = idx / g.SL, slot = idx - fl * g.SL;
        const int f = __ldg(gperm + fl);                      // sorted position -> filter
        int4 o = make_int4(((fl / NV) * 4 * g.SL + slot) * NV + fl % NV, -1, 0, 0);
        if (f < g.F) { o.y = f * g.n_tiles * g.SL + slot; o.z = 2 * (int)__ldg(cprm + (size_t)f * 8 + CP_WSCALE); }
        s_out[idx] = o;
      }
    } else {
      constexpr int FB = CG / 6;
      const int fbase = grp * FB + hh * NF;             // first filter of this thread
#pragma unroll
      for (int i = 0; i < NF; ++i) pa[i] = __ldg(ta.tprm + (size_t)(fbase + i) * 8);
      // entries: (column half h2, virtual filter vi = kind*NF + i, slot)
      for (int idx = etid; idx < g.SL * 2 * NV; idx += EPI_WARPS * 32) {
        const int vl = idx / g.SL, slot = idx - vl * g.SL;
        const int h2 = vl / NV, vi = vl - h2 * NV, kind = vi / NF, i = vi - kind * NF;
        const int f = grp * FB + h2 * NF + i;
        int4 o = make_int4(((h2 * 4) * g.SL + slot) * NV + vi, -1, 0, 0);
        if (f < g.F) {
          const float* tp = ta.tprm + (size_t)f * 8;
          const int shy = (int)__ldg(tp + 1), shz = (int)__ldg(tp + 2), shv = (int)__ldg(tp + 3);
          o.y = (kind * g.F + f) * g.n_tiles * g.SL + slot;
          o.z = shy + (kind == 1 ? shz : (kind == 2 ? shv : shy));
        }
        s_out[idx] = o;
      }
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
      float acc[NV][NSLOT];
#pragma unroll
      for (int i = 0; i < NV; ++i)
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
        mbar_wait(&misc->acc_full[st], (uint32_t)((gp / NST) & 1));
        tc_fence_after();
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * NB + hh * (CG / 2));
        if constexpr (MODE == 0) {
          const uint32_t tlo = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(st * NB + 2 * CG - 8 - hh * (CG / 2));
          // software-pipelined tensor-memory loads: chunk c+1 is in flight while chunk c is processed
          uint32_t bm[2][8], bc[2][8];
          tmem_ld8x2_issue(taddr, tlo, bm[0], bc[0]);
#pragma unroll
          for (int c = 0; c < NV / 4; ++c) {
            // hi products of the 4 filters at columns hh*CG/2 + 8c ..; their lo products sit in the mirrored
            // 8-column block of the lo half, filters in reverse order (k1_tc_layout.cuh)
            tmem_ld_wait(bm[c & 1], bc[c & 1]);
            if (c + 1 < NV / 4) tmem_ld8x2_issue(taddr + 8 * (c + 1), tlo - 8 * (c + 1), bm[(c + 1) & 1], bc[(c + 1) & 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float re = __uint_as_float(bm[c & 1][2 * i]) + __uint_as_float(bc[c & 1][2 * (3 - i)]);
              const float im = __uint_as_float(bm[c & 1][2 * i + 1]) + __uint_as_float(bc[c & 1][2 * (3 - i) + 1]);
              const float en = fmaf(re, re, im * im);
#pragma unroll
              for (int j = 0; j < NSLOT; ++j)
                acc[4 * c + i][j] = fmaf(ex2_approx(pa[4 * c + i] * dj[j]), en, acc[4 * c + i][j]);
            }
          }
        } else {
          constexpr int FB = CG / 6;
          float dv[NSLOT];                                // (k - c)^2 inside the window, 0 outside
#pragma unroll
          for (int j = 0; j < NSLOT; ++j) dv[j] = dj[j] < 1.0e29f ? dj[j] : 0.f;
          // y, z, v accumulators (all three split products already summed in tensor memory) of 4 filters per chunk;
          // chunk c+1 is in flight while chunk c is processed
          uint32_t by[2][8], bz[2][8], bv[2][8];
          tmem_ld8_issue(taddr, by[0]); tmem_ld8_issue(taddr + FB, bz[0]); tmem_ld8_issue(taddr + 2 * FB, bv[0]);
#pragma unroll
          for (int c = 0; c < NF / 4; ++c) {
            tmem_ld_wait(by[c & 1], bz[c & 1], bv[c & 1]);
            if (c + 1 < NF / 4) {
              tmem_ld8_issue(taddr + 8 * (c + 1), by[(c + 1) & 1]);
              tmem_ld8_issue(taddr + FB + 8 * (c + 1), bz[(c + 1) & 1]);
              tmem_ld8_issue(taddr + 2 * FB + 8 * (c + 1), bv[(c + 1) & 1]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int fi = 4 * c + i;
              const float yre = __uint_as_float(by[c & 1][2 * i]), yim = __uint_as_float(by[c & 1][2 * i + 1]);
              const float zre = __uint_as_float(bz[c & 1][2 * i]), zim = __uint_as_float(bz[c & 1][2 * i + 1]);
              const float vre = __uint_as_float(bv[c & 1][2 * i]), vim = __uint_as_float(bv[c & 1][2 * i + 1]);
              const float en = fmaf(yre, yre, yim * yim);
              const float qm = fmaf(yim, zre, -(yre * zim));
              const float qs = fmaf(yre, vre, yim * vim);
#pragma unroll
              for (int j = 0; j < NSLOT; ++j) {
                const float wgt = ex2_approx(pa[fi] * dj[j]);
                acc[fi][j] = fmaf(wgt, en, acc[fi][j]);
                acc[NF + fi][j] = fmaf(wgt, qm, acc[NF + fi][j]);
                acc[2 * NF + fi][j] = fmaf(wgt, qs, acc[2 * NF + fi][j]);
                acc[3 * NF + fi][j] = fmaf(wgt * dv[j], en, acc[3 * NF + fi][j]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_rank0_relaxed(&misc->acc_empty[st]);   // no memory published: TMEM reads are complete
        if (p == 1 && sig_b >= 0) {          // the previous tile's partial sums were stored a phase ago: publish them to K2
          __threadfence();
          __syncwarp();
          if (lane == 0) atomicAdd(tm.done + sig_b, 1);
          sig_b = -1;
        }
      }

      // ---- tile end: row sums of this warp -> s_pw[buf][e][slot][vi], then the quadrant sum and the store --------
      float* pw_buf = s_pw + (LEAN ? (size_t)0 : (size_t)(it & 1) * (EPI_WARPS * g.SL * NV));
      tile_end_rowsums<NV, NSLOT>(acc, pw_buf + (size_t)e * g.SL * NV, red, lane, nb, n_first, n_last, g.SL);
      named_bar_sync(BAR_EPI, EPI_WARPS * 32);
      const int sx = misc->sx_ring[it & 3];
      float* dst = ppart + ((size_t)b * FV * g.n_tiles + tile) * g.SL;     // layout [clip][virtual filter][tile][slot]
      tile_end_store(pw_buf, s_out, g.SL * 2 * NV, etid, valid, dst, sx, g.SL * NV);
      sig_b = (valid && tm.done != nullptr) ? b : -1;   // published off the critical path (phase 1 of the next tile)
    }
    if (sig_b >= 0) {
      __threadfence();
      __syncwarp();
      if (lane == 0) atomicAdd(tm.done + sig_b, 1);
    }
  }

  // ---- teardown -----------------------------------------------------------------------------------
  tc_fence_before();
  cluster_sync_all();                        // nobody may still signal a peer barrier / use TMEM
  if (warp == tc::MMA_WARP) tmem_dealloc_pair<512>(tmem);
  if (rdy.perf != nullptr && blockIdx.x == 0 && tid == 0) {
    rdy.perf[0] = clock64() - perf_c0;
    rdy.perf[1] = global_timer_ns() - perf_t0;
  }
}


// ---- launch wrappers (instantiated in k1_tc_inst_*.cu, one group of channel-group sizes per translation unit so that the
// 17 instantiations of the kernel compile in parallel) ----------------------------------------------------------
template <int CG, int NSLOT, int MODE, int KS>
cudaError_t launch_inst(const Geom& g, const float* x, const uint8_t* w16, const float* cprm, float* ppart,
                               int n_groups, int grid, int smem, cudaStream_t stream, const TcTrainArgs& ta,
                               const TcReady& rdy, const TcMap& tm) {
  // the opt-in to > 48 KB of dynamic shared memory is per function and device: raise it only when this launch needs
  // more than any earlier one asked for (a driver call per forward otherwise)
  static thread_local int smem_set[64] = {0};
  int dev = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  if (dev < 0 || dev >= 64 || smem_set[dev] < smem) {
    err = cudaFuncSetAttribute(k1_tc_kernel<CG, NSLOT, MODE, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (err != cudaSuccess) return err;
    if (dev >= 0 && dev < 64) smem_set[dev] = smem;
  }
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid); lc.blockDim = dim3(tc::NTHREADS); lc.dynamicSmemBytes = (size_t)smem; lc.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // may start under the tail of k0 (see kernel)
  at[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = at; lc.numAttrs = 1;
  return cudaLaunchKernelEx(&lc, k1_tc_kernel<CG, NSLOT, MODE, KS>, g, x, w16, cprm, ppart, n_groups, ta, rdy, tm);
}


template <int CG>
cudaError_t launch_cg(int nslot, const Geom& g, const float* x, const uint8_t* w16, const float* cprm,
                             float* ppart, int n_groups, int grid, int smem, cudaStream_t stream, const TcReady& rdy, const TcMap& tm) {
  const TcTrainArgs none{nullptr, 0};
  if (nslot <= 3) return launch_inst<CG, 3, 0, 0>(g, x, w16, cprm, ppart, n_groups, grid, smem, stream, none, rdy, tm);
  if constexpr (CG <= 64) return launch_inst<CG, 5, 0, 0>(g, x, w16, cprm, ppart, n_groups, grid, smem, stream, none, rdy, tm);
  return cudaErrorNotSupported;
}


}  // namespace leafk
