/* leafk.h -- C ABI of libleafk.so, the Blackwell (sm_100a) LEAF frontend kernels.
 *
 * This is the drop-in boundary for ONE hot path of SarthakYadav/leaf-pytorch:
 * leaf_pytorch.frontend.Leaf.forward (reference leaf_pytorch/frontend.py:78-89) and its
 * parameter-gradient backward (autograd of the same lines, driven from train.py:258).
 * Every entry point takes plain pointers and sizes; no torch types.  All `const float*`
 * / `float*` arguments are DEVICE pointers unless the name ends in `_host`.  Nothing here
 * allocates device memory: scratch comes in through `workspace` (size from
 * leafk_workspace_bytes) so the caller's allocator (PyTorch's) owns every byte.
 * Kernels are enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises
 * the device.  Return value: 0 on success, a negative LEAFK_E* code otherwise, with a
 * thread-local message available from leafk_last_error().
 *
 * Notation: B clips, T samples per clip, F filters, K taps, H hop, N = (T-1)/H + 1 frames.
 */
#ifndef LEAFK_H_
#define LEAFK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LEAFK_VERSION 200

#define LEAFK_OK 0
#define LEAFK_EINVAL (-1)     /* bad shape / null pointer / unsupported geometry            */
#define LEAFK_EWORKSPACE (-2) /* workspace too small                                         */
#define LEAFK_ECUDA (-3)      /* a CUDA runtime call or kernel launch failed                 */
#define LEAFK_EWINDOW (-4)    /* streaming window does not cover the samples the frames need */
#define LEAFK_ETIMEOUT (-5)   /* a kernel gave up waiting (stalled host copy); reported by leafk_async_status */

/* Waveform sample type.  With LEAFK_INPUT_S16 every `const float* x` argument points to int16_t samples
 * (same shapes and strides, in elements). */
#define LEAFK_INPUT_F32 0
#define LEAFK_INPUT_S16 1

/* Feature element type written by the forward calls (leafk_forward_train always writes float32). */
#define LEAFK_OUTPUT_F32 0
#define LEAFK_OUTPUT_BF16 1   /* `out` points to bfloat16 elements, same strides in elements: features for a bf16 backbone
                                 (the step after the path, reference models/classifier.py:15-17) */

/* Which conv kernel computes the Gabor filterbank stage. */
#define LEAFK_ALGO_AUTO 0
#define LEAFK_ALGO_FP32 1     /* direct FP32-FMA correlation (CUDA cores)                    */
#define LEAFK_ALGO_TC 2       /* tcgen05 Toeplitz GEMM, fp16 hi/lo split (3 products), fp32 accumulate */
/* Flag OR-ed into `algo`, forward tensor-core kernel only (testing / A-B measurement).  By default the kernel
 * skips, per k-step of 16 taps, the filters whose Gaussian envelope has decayed below exp(-5.5^2/2) = 2.7e-7 of its
 * peak over the whole k-step (|tau| > ceil(5.5 sigma)): the rounding level of the fp16 hi/lo split itself; and beyond
 * ceil(3.7 sigma) (envelope < 1.1e-3) it runs only the main product x_hi*W_hi, not the two 2^-11 correction products.
 * With this flag every filter runs all three products over all taps. */
#define LEAFK_TC_NOPRUNE 32

/* Flag OR-ed into `algo`, forward calls: the caller guarantees that `workspace` still holds what the bank prologue
 * wrote during an earlier forward with the SAME parameters, config and workspace (e.g. the previous chunk of a
 * chunked long clip, leafk_forward_window): the prologue kernel is skipped. */
#define LEAFK_REUSE_BANKS 64

/* Learnable parameters of the frontend, in the reference's state_dict layout.
 *   kernel   (F,2)  _complex_conv._kernel      reference convolution.py:58
 *   pool_w   (F)    _pooling.weights (1,1,F,1) reference pooling.py:18-20
 *   pool_b   (F)    _pooling._bias  or NULL    reference pooling.py:21-22
 *   alpha,delta,root (F)  _compression.*       reference postprocessing.py:52-54
 *   ema_w    (F)    _compression.ema._weights  reference postprocessing.py:11
 * For compression == 0 (Leaf(pcen_compression=False), frontend.py:74-75) the last four may be NULL. */
typedef struct leafk_params {
  const float* kernel;
  const float* pool_w;
  const float* pool_b;
  const float* alpha;
  const float* delta;
  const float* root;
  const float* ema_w;
} leafk_params;

typedef struct leafk_grads {
  float* kernel; /* (F,2) */
  float* pool_w; /* (F)   */
  float* pool_b; /* (F) or NULL */
  float* alpha;
  float* delta;
  float* root;
  float* ema_w;
} leafk_grads;

/* Optional per-clip preparation applied while the kernels stage the waveform (nothing is written back): what the
 * reference's data pipeline does on the CPU before Leaf.forward -- PadToSize('wrap') / CenterCrop / RandomCrop
 * (utilities/data/raw_transforms.py:121-160), the collate function's zero padding (utilities/data/utils.py:8-28) and
 * PeakNormalization(only_too_loud_sounds) (raw_transforms.py:334-344).  Sample i (0 <= i < T) of prepared clip b is
 * raw[b*ld + (i + start[b])] when 0 <= i + start[b] < length[b]; outside that range it is, by pad_mode,
 *   LEAFK_PAD_ZERO   0 (the collate function's zero padding)
 *   LEAFK_PAD_WRAP   the index taken modulo length[b]                        (np.pad 'wrap': PadToSize_NP, raw_transforms.py:143-160)
 *   LEAFK_PAD_EDGE   the first / last sample of the clip                      (F.pad 'replicate': what PadToSize(mode='wrap') does, :162-183)
 *   LEAFK_PAD_VALUE  pad_value[b], e.g. the clip's minimum                   (PadToSize(mode='constant') pads with signal.min(), :172)
 * and the value is divided by divisor[b].  Any pointer may be NULL (start 0, length T, divisor 1, pad value 0).
 * All arrays are DEVICE pointers of B entries. */
#define LEAFK_PAD_ZERO 0
#define LEAFK_PAD_WRAP 1
#define LEAFK_PAD_EDGE 2
#define LEAFK_PAD_VALUE 3
typedef struct leafk_clip_prep {
  const int* start;
  const int* length;
  const float* divisor;
  long long ld;   /* row stride of the raw waveform buffer in samples; 0 = T */
  int pad_mode;
  const float* pad_value;
} leafk_clip_prep;

/* Static configuration: what Leaf.__init__ derives (frontend.py:38-39, 65-73, 84). */
typedef struct leafk_config {
  int F;           /* n_filters                                                     */
  int K;           /* int(sample_rate*window_len//1000 + 1)     frontend.py:38      */
  int H;           /* int(sample_rate*window_stride//1000)      frontend.py:39      */
  float pcen_floor;/* 1e-12                                     frontend.py:70      */
  float clamp_min; /* 1e-5                                      frontend.py:84      */
  int compression; /* 1: PCEN (frontend.py:65-73)  0: none (frontend.py:74-75)      */
  int algo;        /* LEAFK_ALGO_*                                                  */
  int input_format;/* LEAFK_INPUT_F32 (reference layout) or LEAFK_INPUT_S16: 16-bit PCM, converted
                      in the kernel as s/32768 (what soundfile hands the reference's data pipeline,
                      utilities/data/utils.py:136-157); halves the HBM / PCIe bytes of the waveform */
  int output_format;/* LEAFK_OUTPUT_F32 (reference) or LEAFK_OUTPUT_BF16 */
  const leafk_clip_prep* prep; /* NULL (reference: the batch is used as it is) or the per-clip preparation above; honoured
                      by leafk_forward, leafk_forward_window, leafk_forward_train and leafk_backward (host-buffer calls
                      and the waveform gradient reject it) */
} leafk_config;

int leafk_version(void);
const char* leafk_last_error(void);

/* Frames the pooling stage emits for T samples: (T + padL + padR - K)/H + 1 (pooling.py:36-41). */
int leafk_num_frames(int T, int K, int H);

/* 'same' padding of a K-tap correlation: (K/2 + K%2 - 1, K/2)  (utils.py:5-10). */
void leafk_same_padding(int K, int* pad_left, int* pad_right);

/* Bytes of device scratch leafk_forward / leafk_forward_window / leafk_backward need for B clips
 * when producing `n_frames` frames per clip. */
size_t leafk_workspace_bytes(const leafk_config* cfg, int B, int n_frames);

/* Peak-normalisation divisors of the B prepared clips described by cfg->prep (its `divisor` member is ignored here):
 * divisor_out[b] = max_i |sample i| when that exceeds 1 (only_too_loud != 0; the reference's setting) or is non-zero
 * (only_too_loud == 0), else 1.  One pass over the waveform; use the result as leafk_clip_prep.divisor. */
int leafk_peak_divisors(const leafk_config* cfg, const float* x, int B, int T, int only_too_loud,
                        float* divisor_out, void* stream);
/* minimum_out[b] = smallest raw sample of clip b over [0, length[b]) (cfg->prep gives start / length / ld; the pad value
 * of PadToSize(mode='constant')). */
int leafk_clip_minimum(const leafk_config* cfg, const float* x, int B, int T, float* minimum_out, void* stream);

/* Whole-clip forward: replaces Leaf.forward (frontend.py:78-89).
 *   x    (B,1,T) contiguous fp32                      -> out (B,F,N) contiguous fp32
 *   saved_p  optional (B,F,N): floored pooled energies max(pool(.),clamp_min) kept for backward. */
int leafk_forward(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
                  float* out, float* saved_p, void* workspace, size_t workspace_bytes, void* stream);

/* Streaming / chunked forward with carried PCEN state (BASELINE.json configs[4]; no reference
 * analogue: the reference always restarts the EMA at frame 0, postprocessing.py:15).
 * Produces frames [n_begin, n_begin+n_count) of clips whose full length is T_total, reading a
 * window of each clip: x_win[b*ldx + i] is sample (t_off + i) of clip b, i in [0,T_win).  The
 * window must contain every in-clip sample those frames depend on
 * ([n_begin*H - 2*padL, (n_begin+n_count-1)*H - 2*padL + 2K - 2] clipped to [0,T_total)), else
 * LEAFK_EWINDOW.  ema_state_in (B,F): smoother state after frame n_begin-1, or NULL to start
 * from the first produced frame as the reference does; ema_state_out (B,F) or NULL.
 * out[b*ldo_b + f*ldo_f + (n - n_begin)]; saved_p uses the same strides. */
int leafk_forward_window(const leafk_config* cfg, const leafk_params* prm, const float* x_win,
                         int B, long long ldx, long long T_total, long long t_off, int T_win,
                         int n_begin, int n_count, const float* ema_state_in,
                         float* ema_state_out, float* out, float* saved_p, long long ldo_b,
                         long long ldo_f, void* workspace, size_t workspace_bytes, void* stream);

/* Parameter gradients of sum(out * grad_out): replaces autograd through frontend.py:78-89
 * (train.py:258).  x (B,1,T), grad_out (B,F,N), saved_p (B,F,N) from leafk_forward.  Gradients
 * are written (not accumulated).  grad_x: optional (B,1,T) gradient w.r.t. the float32 waveform, or NULL
 * (train.py never needs it).  Works for every geometry: where the tensor-core training kernel applies it re-runs
 * leafk_forward_train into the workspace (prefer leafk_forward_train + leafk_backward_saved there: one pass less),
 * elsewhere a generic FP32 kernel computes the correlations. */
int leafk_backward(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
                   const float* grad_out, const float* saved_p, const leafk_grads* grads,
                   float* grad_x, void* workspace, size_t workspace_bytes, void* stream);
size_t leafk_backward_workspace_bytes(const leafk_config* cfg, int B, int T);

/* ---- training (reference train.py:233-265: forward, loss.backward()) -------------------------------------------
 * On geometries the tensor-core training kernel covers (leafk_train_supported) the forward that precedes a backward
 * is leafk_forward_train: besides the features it saves, per (clip, filter, frame), the floored pooled energy p and
 * three pooled bilinear forms Q_mu, Q_sigma, Q_poolw of the correlations with the derivative banks.  With those the
 * backward needs no correlation at all:  dL/dmu = 2 sum dp Q_mu,  dL/dsigma = 2 sum dp Q_sigma,
 * dL/ds = sum dp Q_poolw / (s^3 c^2)  (dp = gradient w.r.t. p).  Forward + backward cost three Gabor
 * correlations instead of one + three.
 *   saved  (4,B,F,N) float32: [0] = p, [1..3] = Q_mu, Q_sigma, Q_poolw
 * leafk_backward_saved: gradients are written (not accumulated).  x and grad_x (B,1,T) may both be NULL; when grad_x
 * is given (float32 waveforms only) x is needed and a generic FP32 pass adds dL/dx.
 * Any other geometry: leafk_forward (saved_p) + leafk_backward, which runs the generic FP32 backward. */
int leafk_train_supported(int F, int K, int H);
size_t leafk_train_workspace_bytes(const leafk_config* cfg, int B, int T);
int leafk_forward_train(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T, float* out,
                        float* saved, void* workspace, size_t workspace_bytes, void* stream);
size_t leafk_backward_saved_workspace_bytes(const leafk_config* cfg, int B, int T, int want_grad_x);
int leafk_backward_saved(const leafk_config* cfg, const leafk_params* prm, const float* x, int B, int T,
                         const float* grad_out, const float* saved, const leafk_grads* grads, float* grad_x,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Asynchronous error word of the LAST forward that used `workspace` (its first 4 bytes).  The
 * kernels never trap on a wait that depends on the host: a slice-ready flag that does not arrive within ~10 s
 * (stalled or failed H2D copy) makes the forward finish on whatever data is there and record the condition; this
 * call (a synchronous 4-byte read -- call it after synchronising the stream) returns LEAFK_ETIMEOUT then. */
int leafk_async_status(const void* workspace);
/* LEAFK_OK for word 0, else sets the thread's error message and returns LEAFK_ETIMEOUT (status_host words). */
int leafk_status_message(int word);

/* ---- the two stages the reference's constructor declares but does not implement (frontend.py:40-41, 62-63) -------
 * Stand-alone kernels around the fused frontend, forward and backward each; semantics of the original LEAF.
 * Pre-emphasis: y[b,t] = w2[0] x[b,t] + w2[1] x[b,t+1] with x[b,T] = 0 (learnable 2-tap 'same' correlation, initial
 * value (-0.97, 1)).  Backward: grad_x (B,T) or NULL, grad_w2 (2).
 * Mean/variance normalisation: every row of N frames is centred and scaled by its own biased standard deviation,
 * out = (v - mean) / sqrt(var + eps); stats (rows,2) receives (mean, 1/sqrt(var+eps)) for the backward. */
int leafk_preemp_forward(const float* x, const float* w2, int B, int T, float* y, void* stream);
size_t leafk_preemp_backward_workspace_bytes(void);
int leafk_preemp_backward(const float* x, const float* w2, const float* grad_y, int B, int T, float* grad_x,
                          float* grad_w2, void* workspace, size_t workspace_bytes, void* stream);
int leafk_instnorm_forward(const float* v, long long rows, int N, float eps, float* out, float* stats, void* stream);
int leafk_instnorm_backward(const float* v, const float* stats, const float* grad_out, long long rows, int N,
                            float* grad_v, void* stream);

/* End-to-end call on HOST buffers: x_host (B,1,T) and out_host (B,F,N) are host pointers
 * (pinned for full speed).  The H2D copy is enqueued on copy_stream in `n_slices` (<= 32) pieces,
 * each followed by a stream-ordered 32-bit flag write; ONE persistent launch of the tensor-core
 * kernel on `stream` consumes clips as their slice lands, so the PCIe transfer and the compute
 * overlap without per-slice launch overheads; PCEN and the D2H copy follow on `stream`.  When
 * the FP32 kernel is selected, copy_stream == stream, or the driver lacks cuStreamWriteValue32,
 * it falls back to per-slice launches.  dev_x (B*T samples of cfg->input_format), dev_out (B*F*N elements of
 * cfg->output_format) and workspace are device scratch; x_host / out_host hold the same element types (int16 PCM in
 * halves the upload, LEAFK_OUTPUT_BF16 halves the read-back).  The call returns after enqueueing; the caller
 * synchronises `stream`.
 * status_host (pinned host int, or NULL): receives the asynchronous error word together with the result (same stream
 * order as out_host) -- 0, or a code for leafk_status_message(): no extra synchronisation is needed to learn that a
 * slice of the copy stalled. */
int leafk_forward_host(const leafk_config* cfg, const leafk_params* prm, const float* x_host, int B,
                       int T, float* out_host, int n_slices, float* dev_x, float* dev_out,
                       void* workspace, size_t workspace_bytes, void* stream, void* copy_stream, int* status_host);

/* Fully asynchronous variant for serving loops that keep several batches in flight.  One call = one buffer SET
 * (dev_x, dev_out, workspace, the two events) -- use as many sets as batches in flight.  Ordering enforced
 * inside: H2D slices + ready flags on copy_stream (after the set's previous compute finished), the persistent
 * forward on `stream`, the D2H on d2h_stream (three distinct streams).  ev_compute_done / ev_out_ready are
 * (re)recorded by the call; out_host is valid once ev_out_ready has completed (leafk_event_synchronize).
 * Successive calls overlap: the H2D of batch i+1 runs under the kernels of batch i, the D2H of batch i under
 * the kernels of batch i+1.  Tensor-core kernel only. */
int leafk_forward_host_async(const leafk_config* cfg, const leafk_params* prm, const float* x_host, int B,
                             int T, float* out_host, int n_slices, float* dev_x, float* dev_out,
                             void* workspace, size_t workspace_bytes, void* stream, void* copy_stream,
                             void* d2h_stream, void* ev_compute_done, void* ev_out_ready, int* status_host);
void* leafk_event_create(void);            /* cudaEvent_t without timing, or NULL */
void leafk_event_destroy(void* ev);
int leafk_event_synchronize(void* ev);

/* 1 when LEAFK_ALGO_TC covers (F,K,H); LEAFK_ALGO_AUTO falls back to LEAFK_ALGO_FP32 otherwise. */
int leafk_tc_supported(int F, int K, int H);
/* Host-only description of how the tensor-core kernels would run (F,K,H): channel groups of the inference forward
 * (0 = not covered, the FP32 kernel runs) and channels per group (a multiple of 16; above 96 = the single-group "lean"
 * plan of F = 49..64), filters per group of the training forward (16 or 8; 0 = generic FP32 backward), and the frame
 * slots an 8-sample row touches.  Any output pointer may be NULL. */
int leafk_describe_plan(int F, int K, int H, int* forward_groups, int* forward_channels_per_group,
                        int* train_filters_per_group, int* frame_slots);

/* Per-kernel device timing for the roofline report: between begin and end every forward issued
 * by this thread records CUDA events around K0 (bank prologue), K1 (Gabor GEMM + pooling) and K2
 * (PCEN) on its stream (at most 1024 forwards).  leafk_profile_end synchronises those events and
 * returns the number of forwards seen and the mean milliseconds of each kernel. */
void leafk_profile_begin(void);
int leafk_profile_end(float* ms_k0, float* ms_k1, float* ms_k2);

/* While profiling is on (leafk_profile_begin), CTA 0 of the tensor-core kernel stores its SM-cycle count and
 * its wall time (globaltimer) in the workspace; this reads them back (synchronous copy) for the LAST forward
 * that used `workspace` with shapes (B,T).  cycles/nanoseconds = effective SM clock in GHz during K1. */
int leafk_profile_k1_clock(const leafk_config* cfg, int B, int T, const void* workspace, size_t workspace_bytes,
                           long long* cycles, long long* nanoseconds);

/* Profiling / tests only (synchronous copy): the support-pruning schedule k0 wrote for the LAST forward that used
 * `workspace` with shapes (B,T).  codes[g*n_ksteps + s] = (na1/16) | (na3/16) << 4 for channel group g and 16-tap
 * k-step s: na1 of the group's channels_per_group channels run there, na3 <= na1 of them all three split products
 * (the others only x_hi*W_hi).  The tensor work of a k-step is proportional to na1 + 2*na3. */
int leafk_profile_tc_schedule(const leafk_config* cfg, int B, int T, const void* workspace, size_t workspace_bytes,
                              int* n_groups, int* channels_per_group, int* n_ksteps, int* codes, int codes_capacity);

/* Introspection used by tests / bench: kernels launched by this process since the last reset. */
long long leafk_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* LEAFK_H_ */
