// Shared-memory operand layouts and the shared-memory budget of the tensor-core (tcgen05) Gabor
// kernel.  Used by the bank prologue (k0 writes the B operand into global memory already in this
// layout), by k1_tc_kernel.cuh (which copies it into shared memory unchanged) and by the host planner.
//
// B operand = Gabor bank of one channel group, fp16, "N x K, K-major, no swizzle" canonical
// layout: 8x8 core matrices of 128 contiguous bytes (8 rows x 16 B).  The kernel runs on CTA PAIRS
// (tcgen05 cta_group::2, M = 256): each MMA takes its 128 A rows per CTA from that CTA's own shared
// memory and HALF of its B rows from each CTA, so per CTA the bank is stored as two regions
//   R1 (main MMA, N = 2*CG):  CG rows  -- CTA0: hi halves of the CG channels, CTA1: lo halves
//   R2 (corr MMA, N = CG):    CG/2 rows -- CTA0: hi halves of channels [0,CG/2), CTA1: of [CG/2,CG)
// A region with R rows stores, per k-step of 16 taps, the two 8-tap core-matrix columns one after
// the other:   byte(n,k) = (k/16)*R*32 + ((k%16)/8)*R*16 + (n/8)*128 + (n%8)*16 + (k%8)*2
// so for k-step s the descriptor is {start = base + s*R*32, LBO = R*16 (K direction), SBO = 128}.
// Global image of a group (what k0 writes, what each CTA copies): [R1 cta0 | R1 cta1 | R2 cta0 | R2 cta1].
// (Halving the B rows each CTA feeds its tensor core is what takes the shared-memory pipe off the
// critical path: 94 operand wavefronts per k-step instead of 124, ncu r01.)
//
//
// SUPPORT PRUNING (forward bank only), two levels.  Taps of filter f beyond |tau| > ceil(5.5 sigma_f) are below
// exp(-5.5^2/2) = 2.7e-7 of its peak (the level of the fp16 hi/lo split itself): k-steps whose 16 taps all lie
// outside contribute nothing for f.  And the two CORRECTION products (x_hi*W_lo, x_lo*W_hi; 2^-11 of the main one)
// are only needed where the taps themselves are not small: beyond ceil(3.7 sigma_f) the envelope is below 1.1e-3, the
// dropped terms below 5e-7 of peak*|x| (a 7e-8 relative error of y for broadband input).  k0 sorts the filters by width
// (ascending; padding filters first), deals the sorted list round-robin to the channel groups, and for every
// (group, k-step) finds
//   na1[g][s] = 2 * (filters of the group, rounded up to 8, inside 5.5 sigma)   channels that run at all
//   na3[g][s] = 2 * (filters of the group, rounded up to 8, inside 3.7 sigma)   channels that run all 3 products
// -- always the LAST na1 / na3 columns of the group's sorted order, na3 <= na1, both unimodal in s.  na1 is published
// as nested k-step intervals (level L, na1 >= 16 L, is active on [lo_L, hi_L], L = 1..CG/16); na3 is rounded up to
// its largest value inside every zone of constant na1 (and to CG where na1 = CG), so that the issue loop of k1 has
// one (na1, na3) pair per zone, and published per level and side.  The MMAs of k-step s run
// with N = na1 + na3 (main, A = x_hi) into accumulator columns [CG-na1, CG+na3) and N = na3 (corr, A = x_lo) into
// [CG-na3, CG).  With h = (na1+na3)/2 rows of B from each CTA of the pair and ONE descriptor start offset for both:
//   main, CTA0 rows [CG-na1, CG-na1+h)   hi of sorted channels CG-na1 .. CG-na1+h-1            (row = channel)
//   main, CTA1 rows [CG-na1, CG-h)       hi of the remaining channels CG-na1+h .. CG-1         (row = channel - h)
//         CTA1 rows [CG-h, CG-h+na3)     lo halves in REVERSED filter order: D column CG+i holds the lo product of
//                                        channel lo_channel(i) = 2*(FG-1 - i/2) + i%2, so the lo product of a
//                                        channel lands in the same column for every (na1, na3)
//   corr, CTA0 rows [CG/2-na3/2, CG/2)   hi of channels CG-na3 .. CG-na3/2-1            (row = c - CG/2 + na3/2)
//   corr, CTA1 rows [CG/2-na3/2, CG/2)   hi of channels CG-na3/2 .. CG-1                (row = c - CG/2)
// The middle k-step first_kstep(Kp) (|tau| <= 16 for some tap: inside every real filter's support) always runs with
// na1 = na3 = CG and is issued first with accumulate = 0, so every accumulator column is initialised.  With
// na1 = na3 = CG everywhere this is the unpruned layout up to the order of the lo columns.
//
// A operand = NOT materialised.  For phase p (0..7) a linear fp16 copy of the scaled sample window,
// shifted by p samples, sits in shared memory: copy_p[i] = x~[ts - padL + p + i].  The descriptor
// {start = copy_p + s*32 B, LBO = 16 B, SBO = 128 B} makes row m of the 128x16 operand read
// copy_p[8m + 16s .. 8m + 16s + 15]: a Toeplitz matrix whose core matrices overlap in memory
// (verified exact on B200 by tools/tc_probe.cu).  Row m of phase p is output sample ts + 8m + p.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace leafk {
namespace tc {

constexpr int KSTEP = 16;            // taps per MMA (kind::f16, K = 16)
constexpr int TILE = 1024;           // e-samples per tile: 128 MMA rows x 8 phases
constexpr int NPHASE = 8;
constexpr int MAX_CG = 128;          // largest channel group instantiated (NB = 256 accumulator columns, 2 stages)
constexpr int MAX_CG_FULL = 96;      // largest group with the full shared-memory plan (see lean_plan)
constexpr int SMEM_LIMIT = 227 * 1024;

__host__ __device__ inline size_t region_offset(int R, int n, int k) {
  return (size_t)(k / 16) * R * 32 + (size_t)((k % 16) / 8) * R * 16 + (size_t)(n / 8) * 128 +
         (size_t)(n % 8) * 16 + (size_t)(k % 8) * 2;
}
__host__ __device__ inline size_t r1_bytes(int CG, int Kp) { return (size_t)Kp * CG * 2; }
__host__ __device__ inline size_t r2_bytes(int CG, int Kp) { return (size_t)Kp * (CG / 2) * 2; }
__host__ __device__ inline size_t b_cta_bytes(int CG, int Kp) { return r1_bytes(CG, Kp) + r2_bytes(CG, Kp); }
__host__ __device__ inline size_t b_group_bytes(int CG, int Kp) { return 2 * b_cta_bytes(CG, Kp); }

// ---- pruned forward layout: byte offsets inside a group's global image for sorted channel c = 2*j + ri -------
// (FG = CG/2 filters per group, j = sorted position in the group, na1 / na3 = active channels of the k-step, k = tap)
__host__ __device__ inline size_t p_hi_main(int CG, int Kp, int c, int na1, int na3, int k) {
  const int h = (na1 + na3) / 2;
  if (c < CG - na1 + h) return region_offset(CG, c, k);                       // CTA0
  return r1_bytes(CG, Kp) + region_offset(CG, c - h, k);                       // CTA1, ahead of its lo rows
}
__host__ __device__ inline int lo_column(int CG, int c) {                      // lo column (relative to CG) of channel c
  const int FG = CG / 2, j = c >> 1, ri = c & 1;
  return 2 * (FG - 1 - j) + ri;
}
__host__ __device__ inline size_t p_lo_main(int CG, int Kp, int c, int na1, int na3, int k) {
  const int h = (na1 + na3) / 2;
  return r1_bytes(CG, Kp) + region_offset(CG, CG - h + lo_column(CG, c), k);
}
__host__ __device__ inline size_t p_hi_corr(int CG, int Kp, int c, int na3, int k) {
  const int h = CG / 2;
  if (c < CG - na3 / 2) return 2 * r1_bytes(CG, Kp) + region_offset(h, c - h + na3 / 2, k);
  return 2 * r1_bytes(CG, Kp) + r2_bytes(CG, Kp) + region_offset(h, c - h, k);
}
// k-steps [lo, hi] that hold a tap inside the support |tau| <= ceil(c * sigma) of a filter of width sigma
// (K-tap window, tau = k - K/2).  c <= 0: every k-step; sigma < 0 (padding filter): none (lo > hi).
__host__ __device__ inline void kstep_range(float sigma, float c, int K, int Kp, int* lo, int* hi) {
  if (!(c > 0.f)) { *lo = 0; *hi = Kp / KSTEP - 1; return; }
  if (sigma < 0.f) { *lo = 1; *hi = 0; return; }
  const float Rf = ceilf(c * sigma);
  const int R = Rf > (float)K ? K : (int)Rf, kc = K / 2;
  const int k0 = kc - R < 0 ? 0 : kc - R, k1 = kc + R > K - 1 ? K - 1 : kc + R;
  *lo = k0 / KSTEP; *hi = k1 / KSTEP;
}
constexpr float PRUNE_C = 5.5f;      // support radius in units of sigma: beyond it a filter's taps are skipped
constexpr float PRUNE_C3 = 3.7f;     // beyond it only the main product x_hi*W_hi runs (no lo / correction products)
// k-step the forward kernel issues first (accumulate = 0): k0 keeps every channel of every group active there
__host__ __device__ constexpr int first_kstep(int Kp) { return (Kp / KSTEP - 1) / 2; }
// width rank (ascending) -> channel group and slot inside it: the ranks are DEALT to the groups round-robin, so
// every group holds the same mix of narrow and wide filters (same pruning profile, same cost per tile) and the
// slots of a group are still in ascending width
__host__ __device__ inline int group_of(int rank, int n_groups) { return rank % n_groups; }
__host__ __device__ inline int slot_of(int rank, int n_groups) { return rank / n_groups; }
constexpr int ZONE_INTS = 32;       // ints per channel group in the zone table: [0,16) {lo_L, hi_L} of na1, L = 1..CG/16 <= 6;
                                    // [16,32) {na3 on level L's rising zone, on its falling zone}

// TRAINING LAYOUT (mode 1).  A group holds FB filters x 3 kinds (h, tau*h, (tau^2/sigma^3 - 1/sigma)*h) x (re, im)
// = CG = 6*FB channels (train_channel()).  All three split products of a k-step accumulate into the SAME CG
// accumulator columns with three MMAs of N = CG:  x_hi*W_hi (accumulate = k-step > 0),  x_hi*W_lo,  x_lo*W_hi.
// Each CTA of the pair supplies the bank rows of its own half of the columns (CTA r: columns [r*CG/2, (r+1)*CG/2)),
// so per CTA the bank is two regions of CG/2 rows -- HI then LO -- in the k-step-slab layout of region_offset();
// the third MMA reads the HI region again.  Global image of a group: [HI cta0 | LO cta0 | HI cta1 | LO cta1].
__host__ __device__ inline size_t t_region_bytes(int CG, int Kp) { return (size_t)Kp * (CG / 2) * 2; }
__host__ __device__ inline size_t t_cta_bytes(int CG, int Kp) { return 2 * t_region_bytes(CG, Kp); }
__host__ __device__ inline size_t t_group_bytes(int CG, int Kp) { return 2 * t_cta_bytes(CG, Kp); }
__host__ __device__ inline size_t t_hi(int CG, int Kp, int c, int k) {
  const int h = CG / 2;
  return (size_t)(c / h) * t_cta_bytes(CG, Kp) + region_offset(h, c % h, k);
}
__host__ __device__ inline size_t t_lo(int CG, int Kp, int c, int k) { return t_hi(CG, Kp, c, k) + t_region_bytes(CG, Kp); }

// Byte offsets of the kernel's dynamic shared memory regions.
struct SmemPlan {
  int CL;          // halves per shifted copy: 8*127 + Kp
  int LX;          // samples a tile's copies are cut from: CL + 8 (held in the producers' registers)
  int acb;         // bytes per copy
  int off_w, off_acopy, off_pw, off_red, off_out, off_misc, total;
};

// Virtual filters per epilogue thread: forward CG/4 filters; training (mode 1) FB/2 filters x 4 pooled quantities
// (e, q_mu, q_sigma, q_poolw), FB = CG/6 filters per group.
__host__ __device__ constexpr int virt_per_thread(int CG, int mode) { return mode == 0 ? CG / 4 : 4 * (CG / 12); }

// Channel groups above 96 channels (F = 49..64 as ONE group: the A copies are built once per tile instead of twice and
// the pruned MMAs stay above the 39-cycle A-fetch floor) only fit the 227 KB with a LEAN plan: a single buffer of
// per-warp row sums (safe: with <= 4 accumulator stages no epilogue warp can run a whole tile ahead of another) and no
// transpose buffer for the generic tile-end path -- so they need the fast tile-end path on every tile:
// ceil(248 / H) + NSLOT - 1 < NSLOT + 2, i.e. H >= 124, and NSLOT = 3.
__host__ __device__ constexpr bool lean_plan(int CG, int mode) { return mode == 0 && CG > MAX_CG_FULL; }
__host__ __device__ inline bool lean_geometry_ok(int K, int H) { return H >= 124 && (K + 6) / H + 1 <= 3; }

// mode 0: forward, mode 1: training forward (banks h, tau*h, (tau^2/sigma^3 - 1/sigma)*h; see "TRAINING LAYOUT")
__host__ __device__ inline SmemPlan smem_plan(int CG, int Kp, int SL, int mode = 0, int nslot = 3) {
  SmemPlan s;
  s.CL = 8 * 127 + Kp;
  s.LX = s.CL + 8;
  s.acb = s.CL * 2;
  const int NV = virt_per_thread(CG, mode);
  int off = 0;
  s.off_w = off;      off += (mode == 0) ? (int)b_cta_bytes(CG, Kp) : (int)t_cta_bytes(CG, Kp);
  s.off_acopy = off;  off += 16 * s.acb;
  const bool lean = lean_plan(CG, mode);
  s.off_pw = off;     off += (lean ? 1 : 2) * 8 * SL * NV * 4;   // buffers of per-warp row sums [8 warps][SL][NV]
  // per epilogue warp an NV x 33 float transpose buffer (generic tile-end row sums)
  s.off_red = (off + 15) / 16 * 16;
  // per (virtual filter, slot) output of a tile {offset in the reduction buffer, offset in the tile's partial-sum
  // block or -1, power-of-two exponent of the bank scaling, 0}: tile independent, built once per CTA
  s.off_out = s.off_red + (lean ? 0 : 8 * NV * 33 * 4);
  s.off_out = (s.off_out + 15) / 16 * 16;
  s.off_misc = s.off_out + SL * (2 * NV) * 16;
  s.off_misc = (s.off_misc + 15) / 16 * 16;
  s.total = s.off_misc + 1024;
  return s;
}

// CTAs to launch: pairs (clusters of 2); at most one pair per two SMs, at least one pair per channel group,
// no more pairs than there are unit pairs per group.
inline int pair_grid(int n_sm, int n_groups, long long n_units) {
  long long pairs = n_sm / 2;
  const long long want = ((n_units + 1) / 2) * n_groups;
  if (pairs > want) pairs = want;
  if (pairs < n_groups) pairs = n_groups;
  return (int)(2 * pairs);
}

// slots (frames) one thread's 8 consecutive samples can touch
__host__ __device__ inline int slots_per_thread(int K, int H) { return (K + 6) / H + 1; }

// Split C2 channels into the fewest groups of CG channels (CG a multiple of 16, <= MAX_CG) whose
// shared-memory plan fits.  Returns false when no group size fits.
__host__ __device__ inline bool channel_groups(int C2, int Kp, int SL, int nslot, int* n_groups, int* CG, int K = 0, int H = 0) {
  const int max_cg = nslot > 3 ? 64 : MAX_CG;
  for (int g = 1; g <= 64; ++g) {
    int cg = (C2 + g - 1) / g;
    cg = (cg + 15) / 16 * 16;
    if (cg > max_cg) continue;
    if (cg > MAX_CG_FULL && !(K > 0 && lean_geometry_ok(K, H))) continue;
    if (smem_plan(cg, Kp, SL, 0, nslot > 3 ? 5 : 3).total > SMEM_LIMIT) continue;
    *n_groups = (C2 + cg - 1) / cg;
    *CG = cg;
    return true;
  }
  *n_groups = 0;
  *CG = 16;
  return false;
}

// Training forward: a group holds FB filters x 3 kinds (y, dy/dmu, dy/dsigma) x (re, im) = 6*FB channels.
// Channel of (filter fl in group, kind, ri):  half = fl / (FB/2);  c = half*(3*FB) + kind*FB + (fl % (FB/2))*2 + ri,
// so each epilogue half (columns [half*3FB, (half+1)*3FB), one CTA's bank rows) sees kind-major blocks of FB columns.
__host__ __device__ inline int train_channel(int FB, int fl, int kind, int ri) {
  const int hf = FB / 2;
  return (fl / hf) * (3 * FB) + kind * FB + (fl % hf) * 2 + ri;
}
// filters per training group: 16 (CG = 96) when the plan fits, else 8 (CG = 48); 0 = unsupported
__host__ __device__ inline int train_filters_per_group(int Kp, int SL, int nslot) {
  if (nslot <= 3 && smem_plan(96, Kp, SL, 1, 3).total <= SMEM_LIMIT) return 16;
  if (nslot <= 5 && smem_plan(48, Kp, SL, 1, 5).total <= SMEM_LIMIT) return 8;
  return 0;
}

}  // namespace tc
}  // namespace leafk
