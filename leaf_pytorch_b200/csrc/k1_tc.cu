// K1 (tensor-core variant), host side: geometry coverage and launch dispatch.  The kernel template lives in
// k1_tc_kernel.cuh; its instantiations are compiled in k1_tc_inst_{a,b,c,t}.cu.
#include "k1_tc_kernel.cuh"

namespace leafk {

// explicit instantiations elsewhere
#define LEAFK_EXTERN_CG(CG) extern template cudaError_t launch_cg<CG>(int, const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcReady&, const TcMap&);
LEAFK_EXTERN_CG(16) LEAFK_EXTERN_CG(32) LEAFK_EXTERN_CG(48) LEAFK_EXTERN_CG(64) LEAFK_EXTERN_CG(80) LEAFK_EXTERN_CG(96) LEAFK_EXTERN_CG(112) LEAFK_EXTERN_CG(128)
#undef LEAFK_EXTERN_CG
#define LEAFK_EXTERN_TRAIN(CG, NSLOT, KS) extern template cudaError_t launch_inst<CG, NSLOT, 1, KS>(const Geom&, const float*, const uint8_t*, const float*, float*, int, int, int, cudaStream_t, const TcTrainArgs&, const TcReady&, const TcMap&);
LEAFK_EXTERN_TRAIN(96, 3, 26) LEAFK_EXTERN_TRAIN(96, 3, 0) LEAFK_EXTERN_TRAIN(48, 5, 0)
#undef LEAFK_EXTERN_TRAIN

// ------------------------------------------------------------------------------------------------
bool k1_tc_supported(const Geom& g, const char** why) {
  const int nslot = tc::slots_per_thread(g.K, g.H);
  if (nslot > 5) { if (why) *why = "hop too small relative to the window (more than 5 frames per 8 samples)"; return false; }
  if (g.Kp > 2048) { if (why) *why = "window longer than 2048 taps"; return false; }
  int ng, cg;
  if (!tc::channel_groups(g.C2, g.Kp, g.SL, nslot, &ng, &cg, g.K, g.H)) {
    if (why) *why = "shared-memory plan does not fit for any channel grouping";
    return false;
  }
  return true;
}

// filters per group of the training kernel for this geometry (0: not covered -> generic fp32 backward)
int k1_tc_train_filters_per_group(int K, int H) {
  const int Kp = (K + 15) / 16 * 16;
  if (Kp > 2048) return 0;
  const int SL = (tc::TILE + K - 2) / H + 1;
  return tc::train_filters_per_group(Kp, SL, tc::slots_per_thread(K, H));
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

// Training forward: FB filters per group (16 -> CG 96, 8 -> CG 48); partial sums of the 4 pooled quantities go to
// ppart[b][kind*F + f][tile][slot].
cudaError_t launch_k1_tc_train(const Geom& g, const float* x, const uint8_t* w16t, int FB, int n_groups,
                               const float* tprm, float* ppart, int* done, cudaStream_t stream) {
  cudaError_t err;
  const int n_sm = sm_count(&err);
  if (err != cudaSuccess) return err;
  const int grid = tc::pair_grid(n_sm, n_groups, (long long)g.B * g.n_tiles);
  const TcTrainArgs ta{tprm, n_groups * FB};
  const TcReady rdy{nullptr, 1, nullptr, nullptr};
  const TcMap tm{done, nullptr, nullptr};
  const int nslot = tc::slots_per_thread(g.K, g.H);
  if (FB == 16 && nslot <= 3) {
    const int smem = tc::smem_plan(96, g.Kp, g.SL, 1, 3).total;
    if (g.Kp == 26 * tc::KSTEP) return launch_inst<96, 3, 1, 26>(g, x, w16t, nullptr, ppart, n_groups, grid, smem, stream, ta, rdy, tm);
    return launch_inst<96, 3, 1, 0>(g, x, w16t, nullptr, ppart, n_groups, grid, smem, stream, ta, rdy, tm);
  }
  if (FB == 8 && nslot <= 5) {
    const int smem = tc::smem_plan(48, g.Kp, g.SL, 1, 5).total;
    return launch_inst<48, 5, 1, 0>(g, x, w16t, nullptr, ppart, n_groups, grid, smem, stream, ta, rdy, tm);
  }
  return cudaErrorNotSupported;
}

cudaError_t launch_k1_tc(const Geom& g, const float* x, const uint8_t* w16, const float* cprm, float* ppart,
                         int tc_cg, int tc_groups, const int* tc_perm, const int* tc_zones, int* done, cudaStream_t stream,
                         const int* ready, int clips_per_flag, long long* perf, int* err_word) {
  const TcReady rdy{ready, clips_per_flag < 1 ? 1 : clips_per_flag, perf, err_word};
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
    case 112: return launch_cg<112>(nslot, g, x, w16, cprm, ppart, tc_groups, grid, smem, stream, rdy, tm);
    case 128: return launch_cg<128>(nslot, g, x, w16, cprm, ppart, tc_groups, grid, smem, stream, rdy, tm);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace leafk
