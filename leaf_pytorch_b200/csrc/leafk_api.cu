// extern "C" surface of libleafk.so (declared in include/leafk.h).  Host-side planning only:
// geometry, workspace carving, kernel selection and launches on the caller's stream.
#include "../../include/leafk.h"
#include "leafk_common.cuh"
#include "k1_tc_layout.cuh"
#include "k2_pcen_args.cuh"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace leafk {
// k0_banks.cu
void launch_k0(const float* kernel, const float* pool_w, int F, int K, int Kp, int C2p, float* cprm,
               float* w32, float* g32, uint8_t* w16, int tc_cg, int tc_groups, int* tc_perm, int* tc_zones,
               float prune_c, float prune_c3, int* done, int n_done, cudaStream_t stream);
// k1_fp32.cu
cudaError_t launch_k1_fp32(const Geom& g, const float* x, const float* w32, const float* g32,
                           float* ppart, cudaStream_t stream);
constexpr int F32_TILE = 512;
// k1_tc.cu
bool k1_tc_supported(const Geom& g, const char** why);
cudaError_t launch_k1_tc(const Geom& g, const float* x, const uint8_t* w16, const float* cprm,
                         float* ppart, int tc_cg, int tc_groups, const int* tc_perm, const int* tc_zones, int* done,
                         cudaStream_t stream, const int* ready, int clips_per_flag, long long* perf, int* err_word);
constexpr int TC_TILE = 1024;
// k2_pcen.cu
cudaError_t launch_k2(const Geom& g, const float* ppart, const PcenArgs& a, cudaStream_t stream);
// prep.cu
void geom_apply_prep(const leafk_config* cfg, Geom* g);
// bwd.cu
int train_supported(int F, int K, int H);
size_t train_workspace_bytes(const leafk_config* cfg, int B, int T);
int forward_train_run(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T, float* out,
                      float* saved, void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t backward_saved_workspace_bytes(const leafk_config* cfg, int B, int T, int want_grad_x);
int backward_saved_run(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
                       const float* grad_out, const float* saved, const leafk_grads* grads, float* grad_x,
                       void* workspace, size_t workspace_bytes, cudaStream_t stream);
size_t bwd_workspace_bytes(const leafk_config* cfg, int B, int T);
int bwd_run(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
            const float* grad_out, const float* saved_p, const leafk_grads* grads, float* grad_x,
            void* workspace, size_t workspace_bytes, cudaStream_t stream);

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};   // process-wide: the backward runs on autograd's own thread

// Optional per-kernel timing (bench.py's roofline leg): between leafk_profile_begin() and
// leafk_profile_end() every forward records 4 events on its stream (start, after K0, K1, K2).
struct ProfRec { cudaEvent_t ev[4]; };
thread_local bool g_prof_on = false;
thread_local ProfRec g_prof[1024];
thread_local int g_prof_n = 0;
void prof_mark(int which, cudaStream_t stream) {
  if (!g_prof_on || g_prof_n >= 1024) return;
  if (which == 0)
    for (int i = 0; i < 4; ++i) cudaEventCreate(&g_prof[g_prof_n].ev[i]);
  cudaEventRecord(g_prof[g_prof_n].ev[which], stream);
  if (which == 3) ++g_prof_n;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches += n; }

static int pick_algo(const leafk_config* cfg, const Geom& g) {
  const int want = cfg->algo & 15;                       // low bits: kernel choice; high bits: flags
  if (want == LEAFK_ALGO_FP32) return LEAFK_ALGO_FP32;
  const char* why = nullptr;
  const bool ok = k1_tc_supported(g, &why);
  if (want == LEAFK_ALGO_TC) return ok ? LEAFK_ALGO_TC : -1;
  return ok ? LEAFK_ALGO_TC : LEAFK_ALGO_FP32;
}

// Fill the geometry for producing frames [n_begin, n_begin+n_count) of clips of length T_total.
static int make_geom(const leafk_config* cfg, int B, long long ldx, long long T_total, long long t_off,
                     int T_win, int n_begin, int n_count, int tile_len, Geom* out) {
  if (!cfg) return fail(LEAFK_EINVAL, "null config");
  if (cfg->F < 1 || cfg->K < 2 || cfg->H < 1) return fail(LEAFK_EINVAL, "bad F/K/H (%d,%d,%d)", cfg->F, cfg->K, cfg->H);
  if (cfg->input_format != LEAFK_INPUT_F32 && cfg->input_format != LEAFK_INPUT_S16)
    return fail(LEAFK_EINVAL, "unknown input_format %d", cfg->input_format);
  if (cfg->output_format != LEAFK_OUTPUT_F32 && cfg->output_format != LEAFK_OUTPUT_BF16)
    return fail(LEAFK_EINVAL, "unknown output_format %d", cfg->output_format);
  if (B < 1 || T_total < 1 || T_win < 1) return fail(LEAFK_EINVAL, "bad B/T (%d,%lld,%d)", B, T_total, T_win);
  if (T_total > (1LL << 30)) return fail(LEAFK_EINVAL, "clip too long (%lld samples)", T_total);
  Geom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.F = cfg->F; g.K = cfg->K; g.H = cfg->H;
  g.padL = cfg->K / 2 + (cfg->K % 2) - 1;            // utils.py:9
  g.padR = cfg->K / 2;
  g.C2 = 2 * cfg->F;
  g.C2p = (g.C2 + 7) / 8 * 8;
  g.Kp = (cfg->K + 15) / 16 * 16;
  g.T_total = T_total; g.t_off = t_off; g.T_win = T_win; g.ldx = ldx;
  geom_apply_prep(cfg, &g);
  g.N_total = (int)((T_total + g.padL + g.padR - g.K) / g.H + 1);
  if (n_begin < 0 || n_count < 1 || n_begin + n_count > g.N_total)
    return fail(LEAFK_EINVAL, "frame range [%d,%d) outside [0,%d)", n_begin, n_begin + n_count, g.N_total);
  g.n_begin = n_begin; g.n_count = n_count;
  long long lo = (long long)n_begin * g.H - g.padL;
  long long hi = (long long)(n_begin + n_count - 1) * g.H - g.padL + g.K;   // exclusive
  g.te_lo = lo < 0 ? 0 : lo;
  g.te_hi = hi > T_total ? T_total : hi;
  g.x_fmt = cfg->input_format == LEAFK_INPUT_S16 ? 1 : 0;
  g.TL = tile_len;
  g.n_tiles = (int)((g.te_hi - g.te_lo + tile_len - 1) / tile_len);
  g.SL = (tile_len + g.K - 2) / g.H + 1;
  // samples of the clip the frames depend on must be inside the window
  long long xlo = g.te_lo - g.padL, xhi = g.te_hi - 1 - g.padL + g.K - 1;   // inclusive
  if (xlo < 0) xlo = 0;
  if (xhi > T_total - 1) xhi = T_total - 1;
  if (xlo < t_off || xhi >= t_off + T_win)
    return fail(LEAFK_EWINDOW, "window [%lld,%lld) does not cover needed samples [%lld,%lld]", t_off,
                t_off + T_win, xlo, xhi);
  *out = g;
  return LEAFK_OK;
}

static void carve(const Geom& g, int max_tiles_fp32, int max_tiles_tc, Workspace* w, int* tc_cg, int* tc_groups) {
  const int sl_tc = (TC_TILE + g.K - 2) / g.H + 1;
  tc::channel_groups(g.C2, g.Kp, sl_tc, tc::slots_per_thread(g.K, g.H), tc_groups, tc_cg, g.K, g.H);
  size_t off = 0;
  // first, so that its address does not depend on the shapes: 16 ints (int 0 = asynchronous error word), then the
  // per-clip completion counters
  w->off_done = off;  off += align256(sizeof(int) * ((size_t)g.B + 16));
  w->off_cprm = off; off += align256(sizeof(float) * 8 * g.F);
  w->off_w32 = off;  off += align256(sizeof(float) * (size_t)g.Kp * g.C2p);
  w->off_g32 = off;  off += align256(sizeof(float) * (size_t)g.K * g.F);
  w->off_w16 = off;  off += align256(tc::b_group_bytes(*tc_cg, g.Kp) * (size_t)*tc_groups);
  w->off_tcmap = off; off += align256(sizeof(int) * (size_t)*tc_groups * (*tc_cg / 2 + tc::ZONE_INTS + g.Kp / tc::KSTEP));
  w->off_ppart = off;
  const int sl32 = (F32_TILE + g.K - 2) / g.H + 1, sltc = (TC_TILE + g.K - 2) / g.H + 1;
  size_t a = (size_t)max_tiles_fp32 * sl32, b = (size_t)max_tiles_tc * sltc;
  off += align256(sizeof(float) * (size_t)g.B * g.F * (a > b ? a : b));
  w->off_flags = off; off += 256;                      // 32 slice-ready flags (leafk_forward_host) + 2 perf counters
  w->total = off;
}

static int tiles_for(const leafk_config* cfg, int n_frames, int tile_len) {
  // upper bound of the e-range length for n_frames frames
  long long len = (long long)(n_frames - 1) * cfg->H + cfg->K;
  return (int)((len + tile_len - 1) / tile_len);
}

}  // namespace leafk

using namespace leafk;

extern "C" {

int leafk_version(void) { return LEAFK_VERSION; }
const char* leafk_last_error(void) { return g_err; }

int leafk_num_frames(int T, int K, int H) {
  if (T < 1 || K < 2 || H < 1) return 0;
  const int padL = K / 2 + (K % 2) - 1, padR = K / 2;
  return (T + padL + padR - K) / H + 1;
}

void leafk_same_padding(int K, int* pad_left, int* pad_right) {
  if (pad_left) *pad_left = K / 2 + (K % 2) - 1;
  if (pad_right) *pad_right = K / 2;
}

size_t leafk_workspace_bytes(const leafk_config* cfg, int B, int n_frames) {
  if (!cfg || cfg->F < 1 || cfg->K < 2 || cfg->H < 1 || B < 1 || n_frames < 1) return 0;
  Geom g;
  memset(&g, 0, sizeof(g));
  g.B = B; g.F = cfg->F; g.K = cfg->K; g.H = cfg->H;
  g.C2 = 2 * cfg->F; g.C2p = (g.C2 + 7) / 8 * 8; g.Kp = (cfg->K + 15) / 16 * 16;
  Workspace w;
  int cg, ng;
  carve(g, tiles_for(cfg, n_frames, F32_TILE), tiles_for(cfg, n_frames, TC_TILE), &w, &cg, &ng);
  return w.total;
}

// Shared implementation.  flag_slices > 0: the tensor-core kernel waits per clip on slice-ready flags (in the
// workspace, `clips_per_flag` clips each) that the caller sets with stream-ordered writes.
static int forward_impl(const leafk_config* cfg, const leafk_params* prm, const float* x_win, int B,
                        long long ldx, long long T_total, long long t_off, int T_win, int n_begin,
                        int n_count, const float* ema_state_in, float* ema_state_out, float* out,
                        float* saved_p, long long ldo_b, long long ldo_f, void* workspace,
                        size_t workspace_bytes, cudaStream_t stream, int clips_per_flag, int** flags_out,
                        int* algo_out) {
  if (!cfg || !prm || !x_win || !out || !workspace) return fail(LEAFK_EINVAL, "null pointer argument");
  if (!prm->kernel || !prm->pool_w) return fail(LEAFK_EINVAL, "null Gabor / pooling parameter");
  if (cfg->compression && (!prm->alpha || !prm->delta || !prm->root || !prm->ema_w))
    return fail(LEAFK_EINVAL, "compression=1 needs alpha, delta, root, ema_w");
  Geom g;
  int rc = make_geom(cfg, B, ldx, T_total, t_off, T_win, n_begin, n_count, TC_TILE, &g);
  if (rc) return rc;
  const int algo = pick_algo(cfg, g);
  if (algo < 0) {
    const char* why = "";
    k1_tc_supported(g, &why);
    return fail(LEAFK_EINVAL, "LEAFK_ALGO_TC unsupported for this geometry: %s", why);
  }
  if (algo == LEAFK_ALGO_FP32) {
    rc = make_geom(cfg, B, ldx, T_total, t_off, T_win, n_begin, n_count, F32_TILE, &g);
    if (rc) return rc;
  }
  Workspace w;
  int tc_cg, tc_groups;
  carve(g, algo == LEAFK_ALGO_FP32 ? g.n_tiles : 0, algo == LEAFK_ALGO_TC ? g.n_tiles : 0, &w, &tc_cg, &tc_groups);
  if (w.total > workspace_bytes)
    return fail(LEAFK_EWORKSPACE, "workspace %zu bytes < %zu needed", workspace_bytes, w.total);
  uint8_t* base = (uint8_t*)workspace;
  float* cprm = (float*)(base + w.off_cprm);
  float* w32 = (float*)(base + w.off_w32);
  float* g32 = (float*)(base + w.off_g32);
  uint8_t* w16 = base + w.off_w16;
  float* ppart = (float*)(base + w.off_ppart);
  int* tc_perm = (int*)(base + w.off_tcmap);
  int* tc_zones = tc_perm + (size_t)tc_groups * (tc_cg / 2);
  const float prune_c = (cfg->algo & LEAFK_TC_NOPRUNE) ? 0.f : tc::PRUNE_C;
  const float prune_c3 = (cfg->algo & LEAFK_TC_NOPRUNE) ? 0.f : tc::PRUNE_C3;
  int* err_word = (int*)(base + w.off_done);
  int* done = err_word + 16;

  prof_mark(0, stream);
  cudaError_t err;
  if (cfg->algo & LEAFK_REUSE_BANKS) {
    // banks, sort and schedule are still in the workspace (same parameters): only the per-clip counters are reset
    err = cudaMemsetAsync(err_word, 0, sizeof(int) * ((size_t)g.B + 16), stream);
    if (err != cudaSuccess) return fail(LEAFK_ECUDA, "counter reset: %s", cudaGetErrorString(err));
  } else {
    launch_k0(prm->kernel, prm->pool_w, g.F, g.K, g.Kp, g.C2p, cprm, w32, g32,
              algo == LEAFK_ALGO_TC ? w16 : nullptr, tc_cg, tc_groups, tc_perm, tc_zones, prune_c, prune_c3,
              err_word, g.B + 16, stream);             // also zeroes the error word and the per-clip counters
    err = cudaGetLastError();
    if (err != cudaSuccess) return fail(LEAFK_ECUDA, "k0 launch: %s", cudaGetErrorString(err));
  }
  prof_mark(1, stream);
  int* flags = (int*)(base + w.off_flags);
  if (flags_out) *flags_out = flags;
  if (algo_out) *algo_out = algo;
  if (algo == LEAFK_ALGO_TC)
    err = launch_k1_tc(g, x_win, w16, cprm, ppart, tc_cg, tc_groups, tc_perm, tc_zones, done, stream,
                       clips_per_flag > 0 ? flags : nullptr, clips_per_flag, g_prof_on ? (long long*)(flags + 32) : nullptr,
                       err_word);
  else
    err = launch_k1_fp32(g, x_win, w32, g32, ppart, stream);
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "k1 launch: %s", cudaGetErrorString(err));
  prof_mark(2, stream);
  PcenArgs a;
  memset(&a, 0, sizeof(a));
  a.err = err_word;
  a.out_bf16 = cfg->output_format == LEAFK_OUTPUT_BF16 ? 1 : 0;
  a.pool_b = prm->pool_b; a.alpha = prm->alpha; a.delta = prm->delta; a.root = prm->root;
  a.ema_w = prm->ema_w; a.ema_in = ema_state_in; a.ema_out = ema_state_out; a.out = out;
  a.saved_p = saved_p; a.ldo_b = ldo_b; a.ldo_f = ldo_f; a.pcen_floor = cfg->pcen_floor;
  a.clamp_min = cfg->clamp_min; a.compression = cfg->compression;
  a.done = (algo == LEAFK_ALGO_TC) ? done : nullptr;
  a.done_target = g.n_tiles * tc_groups * 8;              // epilogue warps x tiles x channel groups of a clip
  err = launch_k2(g, ppart, a, stream);
  if (err != cudaSuccess) return fail(LEAFK_ECUDA, "k2 launch: %s", cudaGetErrorString(err));
  prof_mark(3, stream);
  count_launch(3);
  return LEAFK_OK;
}

// workspace offset of the flags without running anything (for leafk_forward_host)
static int flags_location(const leafk_config* cfg, int B, int T, void* workspace, size_t workspace_bytes, int** flags,
                          int* algo) {
  const int N = leafk_num_frames(T, cfg->K, cfg->H);
  Geom g;
  int rc = make_geom(cfg, B, T, T, 0, T, 0, N, TC_TILE, &g);
  if (rc) return rc;
  *algo = pick_algo(cfg, g);
  if (*algo == LEAFK_ALGO_FP32) {
    rc = make_geom(cfg, B, T, T, 0, T, 0, N, F32_TILE, &g);
    if (rc) return rc;
  }
  Workspace w;
  int cg, ng;
  carve(g, *algo == LEAFK_ALGO_FP32 ? g.n_tiles : 0, *algo == LEAFK_ALGO_TC ? g.n_tiles : 0, &w, &cg, &ng);
  if (w.total > workspace_bytes) return fail(LEAFK_EWORKSPACE, "workspace %zu bytes < %zu needed", workspace_bytes, w.total);
  *flags = (int*)((uint8_t*)workspace + w.off_flags);
  return LEAFK_OK;
}

typedef int (*StreamWriteValue32Fn)(cudaStream_t, unsigned long long, unsigned int, unsigned int);
static StreamWriteValue32Fn stream_write_value32() {
  static StreamWriteValue32Fn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      fn = (StreamWriteValue32Fn)p;
    else
      (void)cudaGetLastError();
  }
  return fn;
}

int leafk_forward_window(const leafk_config* cfg, const leafk_params* prm, const float* x_win, int B,
                         long long ldx, long long T_total, long long t_off, int T_win, int n_begin,
                         int n_count, const float* ema_state_in, float* ema_state_out, float* out,
                         float* saved_p, long long ldo_b, long long ldo_f, void* workspace,
                         size_t workspace_bytes, void* stream) {
  return forward_impl(cfg, prm, x_win, B, ldx, T_total, t_off, T_win, n_begin, n_count, ema_state_in,
                      ema_state_out, out, saved_p, ldo_b, ldo_f, workspace, workspace_bytes, (cudaStream_t)stream, 0,
                      nullptr, nullptr);
}

int leafk_forward(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
                  float* out, float* saved_p, void* workspace, size_t workspace_bytes, void* stream) {
  if (!cfg) return fail(LEAFK_EINVAL, "null config");
  const int N = leafk_num_frames(T, cfg->K, cfg->H);
  if (N < 1) return fail(LEAFK_EINVAL, "bad T/K/H (%d,%d,%d)", T, cfg->K, cfg->H);
  return leafk_forward_window(cfg, prm, x, B, T, T, 0, T, 0, N, nullptr, nullptr, out, saved_p,
                              (long long)cfg->F * N, N, workspace, workspace_bytes, stream);
}

static size_t out_elem_bytes(const leafk_config* cfg) { return cfg->output_format == LEAFK_OUTPUT_BF16 ? 2 : 4; }

// Sliced fallback of leafk_forward_host: slice i = H2D on copy_stream -> kernels on stream -> D2H.
static int forward_host_sliced(const leafk_config* cfg, const leafk_params* prm, const float* x_host, int B, int T,
                               int N, float* out_host, int n_slices, float* dev_x, float* dev_out, void* workspace,
                               size_t workspace_bytes, cudaStream_t stream, cudaStream_t cstream) {
  cudaEvent_t up[16], done[16];
  int rc = LEAFK_OK;
  const bool two = (cstream != stream);
  int made = 0;
  for (int i = 0; i < n_slices && rc == LEAFK_OK; ++i) {
    const int b0 = (int)((long long)B * i / n_slices), b1 = (int)((long long)B * (i + 1) / n_slices);
    const int nb = b1 - b0;
    if (nb == 0) continue;
    cudaEventCreateWithFlags(&up[made], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&done[made], cudaEventDisableTiming);
    const size_t esz = cfg->input_format == LEAFK_INPUT_S16 ? 2 : 4;
    const size_t osz = out_elem_bytes(cfg);
    const size_t xoff = (size_t)b0 * T * esz, ooff = (size_t)b0 * cfg->F * N * osz;
    cudaError_t e = cudaMemcpyAsync((uint8_t*)dev_x + xoff, (const uint8_t*)x_host + xoff, esz * (size_t)nb * T,
                                    cudaMemcpyHostToDevice, cstream);
    if (e != cudaSuccess) { rc = fail(LEAFK_ECUDA, "H2D: %s", cudaGetErrorString(e)); ++made; break; }
    if (two) { cudaEventRecord(up[made], cstream); cudaStreamWaitEvent(stream, up[made], 0); }
    rc = leafk_forward(cfg, prm, (const float*)((const uint8_t*)dev_x + xoff), nb, T, (float*)((uint8_t*)dev_out + ooff),
                       nullptr, workspace, workspace_bytes, stream);
    if (rc == LEAFK_OK) {
      if (two) { cudaEventRecord(done[made], stream); cudaStreamWaitEvent(cstream, done[made], 0); }
      e = cudaMemcpyAsync((uint8_t*)out_host + ooff, (const uint8_t*)dev_out + ooff,
                          osz * (size_t)nb * cfg->F * N, cudaMemcpyDeviceToHost, two ? cstream : stream);
      if (e != cudaSuccess) rc = fail(LEAFK_ECUDA, "D2H: %s", cudaGetErrorString(e));
    }
    ++made;
  }
  if (two && made > 0 && rc == LEAFK_OK) {             // make `stream` wait for the last D2H
    cudaEvent_t fin;
    cudaEventCreateWithFlags(&fin, cudaEventDisableTiming);
    cudaEventRecord(fin, cstream);
    cudaStreamWaitEvent(stream, fin, 0);
    cudaEventDestroy(fin);
  }
  for (int i = 0; i < made; ++i) { cudaEventDestroy(up[i]); cudaEventDestroy(done[i]); }
  return rc;
}

int leafk_forward_host(const leafk_config* cfg, const leafk_params* prm, const float* x_host, int B, int T,
                       float* out_host, int n_slices, float* dev_x, float* dev_out, void* workspace,
                       size_t workspace_bytes, void* stream_, void* copy_stream_, int* status_host) {
  if (!cfg || !prm || !x_host || !out_host || !dev_x || !dev_out || !workspace)
    return fail(LEAFK_EINVAL, "null pointer argument");
  if (cfg->prep) return fail(LEAFK_EINVAL, "host-buffer calls take prepared batches (cfg->prep must be NULL)");
  cudaStream_t stream = (cudaStream_t)stream_, cstream = (cudaStream_t)copy_stream_;
  const int N = leafk_num_frames(T, cfg->K, cfg->H);
  if (N < 1 || B < 1) return fail(LEAFK_EINVAL, "bad B/T");
  if (n_slices < 1) n_slices = 1;
  if (n_slices > B) n_slices = B;
  if (n_slices > 32) n_slices = 32;
  int* flags = nullptr;
  int algo = 0;
  int rc = flags_location(cfg, B, T, workspace, workspace_bytes, &flags, &algo);
  if (rc) return rc;
  StreamWriteValue32Fn write32 = stream_write_value32();
  if (algo != LEAFK_ALGO_TC || cstream == stream || write32 == nullptr || n_slices == 1)
  {
    if (status_host != nullptr) *status_host = 0;      // no flags on this path: nothing can time out
    return forward_host_sliced(cfg, prm, x_host, B, T, N, out_host, n_slices > 16 ? 16 : n_slices, dev_x, dev_out,
                               workspace, workspace_bytes, stream, cstream);
  }

  // Pipelined path: ONE persistent launch of the tensor-core kernel over the whole batch; its producers wait
  // per clip on slice-ready flags that follow each slice of the H2D copy in copy_stream order.
  const int clips_per_flag = (B + n_slices - 1) / n_slices;
  const int n_flags = (B + clips_per_flag - 1) / clips_per_flag;
  cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(int) * 32, stream);
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "flag reset: %s", cudaGetErrorString(e));
  cudaEvent_t reset_done;
  cudaEventCreateWithFlags(&reset_done, cudaEventDisableTiming);
  cudaEventRecord(reset_done, stream);
  cudaStreamWaitEvent(cstream, reset_done, 0);         // flag writes must not be overtaken by the reset
  cudaEventDestroy(reset_done);
  for (int s = 0; s < n_flags; ++s) {
    const int b0 = s * clips_per_flag, b1 = (b0 + clips_per_flag < B) ? b0 + clips_per_flag : B;
    const size_t esz = cfg->input_format == LEAFK_INPUT_S16 ? 2 : 4;
    e = cudaMemcpyAsync((uint8_t*)dev_x + (size_t)b0 * T * esz, (const uint8_t*)x_host + (size_t)b0 * T * esz,
                        esz * (size_t)(b1 - b0) * T, cudaMemcpyHostToDevice, cstream);
    if (e != cudaSuccess) return fail(LEAFK_ECUDA, "H2D: %s", cudaGetErrorString(e));
    if (write32(cstream, (unsigned long long)(uintptr_t)(flags + s), 1u, 0u) != 0)
      return fail(LEAFK_ECUDA, "cuStreamWriteValue32 failed");
  }
  rc = forward_impl(cfg, prm, dev_x, B, T, T, 0, T, 0, N, nullptr, nullptr, dev_out, nullptr, (long long)cfg->F * N, N,
                    workspace, workspace_bytes, stream, clips_per_flag, nullptr, nullptr);
  if (rc) return rc;
  e = cudaMemcpyAsync(out_host, dev_out, out_elem_bytes(cfg) * (size_t)B * cfg->F * N, cudaMemcpyDeviceToHost, stream);
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "D2H: %s", cudaGetErrorString(e));
  if (status_host != nullptr) {                         // the asynchronous error word travels with the result
    e = cudaMemcpyAsync(status_host, workspace, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e != cudaSuccess) return fail(LEAFK_ECUDA, "status D2H: %s", cudaGetErrorString(e));
  }
  return LEAFK_OK;
}

void* leafk_event_create(void) {
  cudaEvent_t ev = nullptr;
  if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  return (void*)ev;
}
void leafk_event_destroy(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); }
int leafk_event_synchronize(void* ev) {
  if (!ev) return fail(LEAFK_EINVAL, "null event");
  cudaError_t e = cudaEventSynchronize((cudaEvent_t)ev);
  return e == cudaSuccess ? LEAFK_OK : fail(LEAFK_ECUDA, "event sync: %s", cudaGetErrorString(e));
}

int leafk_forward_host_async(const leafk_config* cfg, const leafk_params* prm, const float* x_host, int B, int T,
                             float* out_host, int n_slices, float* dev_x, float* dev_out, void* workspace,
                             size_t workspace_bytes, void* stream_, void* copy_stream_, void* d2h_stream_,
                             void* ev_compute_done_, void* ev_out_ready_, int* status_host) {
  if (!cfg || !prm || !x_host || !out_host || !dev_x || !dev_out || !workspace || !ev_compute_done_ || !ev_out_ready_)
    return fail(LEAFK_EINVAL, "null pointer argument");
  if (cfg->prep) return fail(LEAFK_EINVAL, "host-buffer calls take prepared batches (cfg->prep must be NULL)");
  cudaStream_t stream = (cudaStream_t)stream_, cstream = (cudaStream_t)copy_stream_, dstream = (cudaStream_t)d2h_stream_;
  cudaEvent_t ev_compute_done = (cudaEvent_t)ev_compute_done_, ev_out_ready = (cudaEvent_t)ev_out_ready_;
  if (cstream == stream || dstream == stream || cstream == dstream)
    return fail(LEAFK_EINVAL, "leafk_forward_host_async needs three distinct streams");
  const int N = leafk_num_frames(T, cfg->K, cfg->H);
  if (N < 1 || B < 1) return fail(LEAFK_EINVAL, "bad B/T");
  if (n_slices < 1) n_slices = 1;
  if (n_slices > B) n_slices = B;
  if (n_slices > 32) n_slices = 32;
  int* flags = nullptr;
  int algo = 0;
  int rc = flags_location(cfg, B, T, workspace, workspace_bytes, &flags, &algo);
  if (rc) return rc;
  StreamWriteValue32Fn write32 = stream_write_value32();
  if (algo != LEAFK_ALGO_TC || write32 == nullptr)
    return fail(LEAFK_EINVAL, "leafk_forward_host_async needs the tensor-core kernel and cuStreamWriteValue32");
  const size_t esz = cfg->input_format == LEAFK_INPUT_S16 ? 2 : 4;
  // copy stream: the previous use of this buffer set (dev_x, flags) must have been consumed
  cudaStreamWaitEvent(cstream, ev_compute_done, 0);             // no-op for a never-recorded event
  cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(int) * 32, cstream);
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "flag reset: %s", cudaGetErrorString(e));
  cudaEvent_t reset_done;
  cudaEventCreateWithFlags(&reset_done, cudaEventDisableTiming);
  cudaEventRecord(reset_done, cstream);
  cudaStreamWaitEvent(stream, reset_done, 0);                  // K1 must not poll flags of the previous use
  cudaEventDestroy(reset_done);
  const int clips_per_flag = (B + n_slices - 1) / n_slices;
  const int n_flags = (B + clips_per_flag - 1) / clips_per_flag;
  for (int s = 0; s < n_flags; ++s) {
    const int b0 = s * clips_per_flag, b1 = (b0 + clips_per_flag < B) ? b0 + clips_per_flag : B;
    e = cudaMemcpyAsync((uint8_t*)dev_x + (size_t)b0 * T * esz, (const uint8_t*)x_host + (size_t)b0 * T * esz,
                        esz * (size_t)(b1 - b0) * T, cudaMemcpyHostToDevice, cstream);
    if (e != cudaSuccess) return fail(LEAFK_ECUDA, "H2D: %s", cudaGetErrorString(e));
    if (write32(cstream, (unsigned long long)(uintptr_t)(flags + s), 1u, 0u) != 0)
      return fail(LEAFK_ECUDA, "cuStreamWriteValue32 failed");
  }
  // compute stream: the previous D2H out of dev_out must be over before K2 overwrites it
  cudaStreamWaitEvent(stream, ev_out_ready, 0);
  rc = forward_impl(cfg, prm, dev_x, B, T, T, 0, T, 0, N, nullptr, nullptr, dev_out, nullptr, (long long)cfg->F * N, N,
                    workspace, workspace_bytes, stream, clips_per_flag, nullptr, nullptr);
  if (rc) return rc;
  cudaEventRecord(ev_compute_done, stream);
  cudaStreamWaitEvent(dstream, ev_compute_done, 0);
  e = cudaMemcpyAsync(out_host, dev_out, out_elem_bytes(cfg) * (size_t)B * cfg->F * N, cudaMemcpyDeviceToHost, dstream);
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "D2H: %s", cudaGetErrorString(e));
  if (status_host != nullptr) {                         // the asynchronous error word travels with the result
    e = cudaMemcpyAsync(status_host, workspace, sizeof(int), cudaMemcpyDeviceToHost, dstream);
    if (e != cudaSuccess) return fail(LEAFK_ECUDA, "status D2H: %s", cudaGetErrorString(e));
  }
  cudaEventRecord(ev_out_ready, dstream);
  return LEAFK_OK;
}

size_t leafk_backward_workspace_bytes(const leafk_config* cfg, int B, int T) { return bwd_workspace_bytes(cfg, B, T); }

int leafk_train_supported(int F, int K, int H) { return train_supported(F, K, H); }
size_t leafk_train_workspace_bytes(const leafk_config* cfg, int B, int T) { return train_workspace_bytes(cfg, B, T); }
int leafk_forward_train(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T, float* out,
                        float* saved, void* workspace, size_t workspace_bytes, void* stream) {
  if (cfg && cfg->output_format != LEAFK_OUTPUT_F32) return fail(LEAFK_EINVAL, "the training forward emits float32 features");
  return forward_train_run(cfg, prm, x, B, T, out, saved, workspace, workspace_bytes, (cudaStream_t)stream);
}
size_t leafk_backward_saved_workspace_bytes(const leafk_config* cfg, int B, int T, int want_grad_x) {
  return backward_saved_workspace_bytes(cfg, B, T, want_grad_x);
}
int leafk_backward_saved(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
                         const float* grad_out, const float* saved, const leafk_grads* grads, float* grad_x,
                         void* workspace, size_t workspace_bytes, void* stream) {
  return backward_saved_run(cfg, prm, x, B, T, grad_out, saved, grads, grad_x, workspace, workspace_bytes,
                            (cudaStream_t)stream);
}

int leafk_async_status(const void* workspace) {
  if (!workspace) return fail(LEAFK_EINVAL, "null pointer argument");
  int word = 0;
  cudaError_t e = cudaMemcpy(&word, workspace, sizeof(int), cudaMemcpyDeviceToHost);   // synchronous by design
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "status read: %s", cudaGetErrorString(e));
  return leafk_status_message(word);
}

int leafk_status_message(int word) {
  if (word == LEAFK_ASYNC_H2D_TIMEOUT)
    return fail(LEAFK_ETIMEOUT, "a slice of the host-to-device copy never signalled ready (stalled copy); features invalid");
  if (word == LEAFK_ASYNC_K1_TIMEOUT)
    return fail(LEAFK_ETIMEOUT, "the PCEN kernel gave up waiting for the Gabor kernel's per-clip counters; features invalid");
  if (word != 0) return fail(LEAFK_ECUDA, "unknown asynchronous error word %d", word);
  return LEAFK_OK;
}

int leafk_backward(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
                   const float* grad_out, const float* saved_p, const leafk_grads* grads, float* grad_x,
                   void* workspace, size_t workspace_bytes, void* stream) {
  return bwd_run(cfg, prm, x, B, T, grad_out, saved_p, grads, grad_x, workspace, workspace_bytes,
                 (cudaStream_t)stream);
}

int leafk_tc_supported(int F, int K, int H) {
  if (F < 1 || K < 2 || H < 1) return 0;
  Geom g;
  memset(&g, 0, sizeof(g));
  g.F = F; g.K = K; g.H = H; g.C2 = 2 * F; g.Kp = (K + 15) / 16 * 16; g.TL = TC_TILE;
  g.SL = (TC_TILE + K - 2) / H + 1;
  const char* why = nullptr;
  return k1_tc_supported(g, &why) ? 1 : 0;
}

int leafk_describe_plan(int F, int K, int H, int* forward_groups, int* forward_channels_per_group,
                        int* train_filters_per_group, int* frame_slots) {
  if (F < 1 || K < 2 || H < 1) return fail(LEAFK_EINVAL, "bad F/K/H (%d,%d,%d)", F, K, H);
  const int Kp = (K + 15) / 16 * 16, SL = (TC_TILE + K - 2) / H + 1, nslot = tc::slots_per_thread(K, H);
  int ng = 0, cg = 0;
  const bool ok = nslot <= 5 && Kp <= 2048 && tc::channel_groups(2 * F, Kp, SL, nslot, &ng, &cg, K, H);
  if (forward_groups) *forward_groups = ok ? ng : 0;
  if (forward_channels_per_group) *forward_channels_per_group = ok ? cg : 0;
  if (train_filters_per_group) *train_filters_per_group = train_supported(F, K, H) ? tc::train_filters_per_group(Kp, SL, nslot) : 0;
  if (frame_slots) *frame_slots = nslot;
  return LEAFK_OK;
}

void leafk_profile_begin(void) {
  g_prof_on = true;
  g_prof_n = 0;
}

int leafk_profile_end(float* ms_k0, float* ms_k1, float* ms_k2) {
  g_prof_on = false;
  double acc[3] = {0, 0, 0};
  const int n = g_prof_n;
  for (int r = 0; r < n; ++r) {
    cudaEventSynchronize(g_prof[r].ev[3]);
    for (int i = 0; i < 3; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, g_prof[r].ev[i], g_prof[r].ev[i + 1]);
      acc[i] += ms;
    }
    for (int i = 0; i < 4; ++i) cudaEventDestroy(g_prof[r].ev[i]);
  }
  g_prof_n = 0;
  if (n > 0) {
    if (ms_k0) *ms_k0 = (float)(acc[0] / n);
    if (ms_k1) *ms_k1 = (float)(acc[1] / n);
    if (ms_k2) *ms_k2 = (float)(acc[2] / n);
  }
  return n;
}

int leafk_profile_k1_clock(const leafk_config* cfg, int B, int T, const void* workspace, size_t workspace_bytes,
                           long long* cycles, long long* nanoseconds) {
  if (!cfg || !workspace || !cycles || !nanoseconds) return fail(LEAFK_EINVAL, "null pointer argument");
  int* flags = nullptr;
  int algo = 0;
  int rc = flags_location(cfg, B, T, const_cast<void*>(workspace), workspace_bytes, &flags, &algo);
  if (rc) return rc;
  long long host[2] = {0, 0};
  cudaError_t e = cudaMemcpy(host, flags + 32, sizeof(host), cudaMemcpyDeviceToHost);   // synchronous: profiling only
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "perf read: %s", cudaGetErrorString(e));
  *cycles = host[0];
  *nanoseconds = host[1];
  return LEAFK_OK;
}

int leafk_profile_tc_schedule(const leafk_config* cfg, int B, int T, const void* workspace, size_t workspace_bytes,
                              int* n_groups, int* channels_per_group, int* n_ksteps, int* codes, int codes_capacity) {
  if (!cfg || !workspace || !n_groups || !channels_per_group || !n_ksteps || !codes)
    return fail(LEAFK_EINVAL, "null pointer argument");
  const int N = leafk_num_frames(T, cfg->K, cfg->H);
  Geom g;
  int rc = make_geom(cfg, B, T, T, 0, T, 0, N, TC_TILE, &g);
  if (rc) return rc;
  if (pick_algo(cfg, g) != LEAFK_ALGO_TC) return fail(LEAFK_EINVAL, "the tensor-core kernel does not run for this config");
  Workspace w;
  int cg, ng;
  carve(g, 0, g.n_tiles, &w, &cg, &ng);
  if (w.total > workspace_bytes) return fail(LEAFK_EWORKSPACE, "workspace %zu bytes < %zu needed", workspace_bytes, w.total);
  const int ks = g.Kp / tc::KSTEP;
  if (codes_capacity < ng * ks) return fail(LEAFK_EINVAL, "codes_capacity %d < %d", codes_capacity, ng * ks);
  const int* dev = (const int*)((const uint8_t*)workspace + w.off_tcmap) + (size_t)ng * (cg / 2) + (size_t)ng * tc::ZONE_INTS;
  cudaError_t e = cudaMemcpy(codes, dev, sizeof(int) * (size_t)ng * ks, cudaMemcpyDeviceToHost);   // synchronous: profiling only
  if (e != cudaSuccess) return fail(LEAFK_ECUDA, "schedule read: %s", cudaGetErrorString(e));
  *n_groups = ng; *channels_per_group = cg; *n_ksteps = ks;
  return LEAFK_OK;
}

long long leafk_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

}  // extern "C"
