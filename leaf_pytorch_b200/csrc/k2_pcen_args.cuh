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
};
}  // namespace leafk
