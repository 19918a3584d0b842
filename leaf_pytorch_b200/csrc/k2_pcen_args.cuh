// Argument block of the PCEN kernel (k2_pcen.cu), shared with the host API.
#pragma once
namespace leafk {
struct PcenArgs {
  const float* pool_b;   // may be null
  const float* alpha;
  const float* delta;
  const float* root;
  const float* ema_w;
  const float* ema_in;   // may be null
  float* ema_out;        // may be null
  float* out;
  float* saved_p;        // may be null
  long long ldo_b, ldo_f;
  float pcen_floor, clamp_min;
  int compression;
  const int* done;       // per-clip completion counters written by the tensor-core K1, or null (then the kernel waits
  int done_target;       // for the whole preceding grid); a clip is ready at done[b] >= done_target
  int* err;              // asynchronous error word (LEAFK_ASYNC_K1_TIMEOUT when a counter never completes), may be null
  // training forward: the partial sums hold 4 pooled quantities per filter ([b][kind*F + f][tile][slot]); kinds 1..3
  // (Q_mu, Q_sigma, Q_poolw, see k1_tc_kernel.cuh) are only assembled (no bias / floor / PCEN) into q_out[(kind-1)][b][f][n]
  float* q_out;          // (3,B,F,N) contiguous or null (plain forward: one quantity per filter)
  int out_bf16;          // 1: `out` points to bfloat16 elements (same strides, in elements); features for a bf16 backbone
};
}  // namespace leafk
