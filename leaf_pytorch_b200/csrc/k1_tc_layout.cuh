// Shared-memory operand layouts of the tensor-core (tcgen05) Gabor kernel, shared between the
// bank prologue (k0, which writes the B operand into global memory already in this layout) and
// k1_tc.cu (which bulk-copies it into shared memory unchanged).
//
// B operand = Gabor bank of one channel group, fp16, "N x K, K-major, no swizzle" canonical
// layout: 8x8 core matrices of 128 contiguous bytes (8 rows x 16 B).  Rows n: [0,CG) are the
// hi halves of the group's CG channels, [CG,2CG) the lo halves (NB = 2*CG rows).  The taps are cut
// into k-steps of 16 (one tcgen05.mma.kind::f16 each); per k-step the two 8-tap core-matrix
// columns are stored one after the other:
//   byte(n,k) = (k/16)*NB*32 + ((k%16)/8)*NB*16 + (n/8)*128 + (n%8)*16 + (k%8)*2
// so for k-step s the descriptor is {start = base + s*NB*32, LBO = NB*16 (K direction),
// SBO = 128 (N direction)}.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace leafk {
namespace tc {

constexpr int KSTEP = 16;          // taps per MMA (kind::f16, K = 16)
constexpr int MAX_CG = 80;         // channels per CTA group (hi+lo = 160 accumulator columns)

__host__ __device__ inline size_t b_group_bytes(int CG, int Kp) { return (size_t)Kp * (2 * CG) * 2; }

__host__ __device__ inline size_t b_offset(int CG, int n, int k) {
  const int NB = 2 * CG;
  return (size_t)(k / 16) * NB * 32 + (size_t)((k % 16) / 8) * NB * 16 + (size_t)(n / 8) * 128 +
         (size_t)(n % 8) * 16 + (size_t)(k % 8) * 2;
}

// channels per group: split C2 channels into the fewest groups of <= MAX_CG channels, each a
// multiple of 16 (tcgen05 M=128 needs N % 16 == 0)
__host__ __device__ inline void channel_groups(int C2, int* n_groups, int* CG) {
  int g = (C2 + MAX_CG - 1) / MAX_CG;
  int cg = (C2 + g - 1) / g;
  cg = (cg + 15) / 16 * 16;
  *n_groups = (C2 + cg - 1) / cg;
  *CG = cg;
}

}  // namespace tc
}  // namespace leafk
