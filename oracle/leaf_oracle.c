/* leaf_oracle.c -- plain-C restatement of leaf_pytorch.frontend.Leaf.forward.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's CPU
 * legs may use it.  It exists as a second, torch-free opinion next to oracle/leaf_oracle.py (which
 * replays the reference's own ATen ops): straight loops, no library convolution, optional double
 * accumulation.  Each function cites the reference lines it follows (paths under /root/reference).
 *
 * Pinning: tests/test_oracle_golden.py::test_c_oracle_* compares it with the golden vectors that
 * tests/golden/make_golden.py generated from the real reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

static int pad_left(int K) { return K / 2 + (K % 2) - 1; }          /* utils.py:9 */
static int pad_right(int K) { return K / 2; }

int leaf_oracle_num_frames(int T, int K, int H) {                    /* pooling.py:36-41 */
  return (T + pad_left(K) + pad_right(K) - K) / H + 1;
}

/* convolution.py:15-22 + impulse_responses.py:5-16,66-71 + convolution.py:77-90.
 * bank[(2f+ri)*K + k], float32 operation order of the reference. */
static void gabor_bank(const float* kernel, int F, int K, float* bank) {
  const float root_2ln2 = sqrtf(2.0f * logf(2.0f));
  const float sig_lo = 4.0f * root_2ln2 / (float)M_PI, sig_hi = (float)K * root_2ln2 / (float)M_PI;
  const float sqrt_2pi = sqrtf(2.0f * (float)M_PI);
  for (int f = 0; f < F; ++f) {
    float mu = kernel[2 * f], sg = kernel[2 * f + 1];
    mu = mu < 0.f ? 0.f : (mu > (float)M_PI ? (float)M_PI : mu);
    sg = sg < sig_lo ? sig_lo : (sg > sig_hi ? sig_hi : sg);
    const float norm = 1.0f / (sqrt_2pi * sg);
    const float inv2s2 = 1.0f / (2.0f * (sg * sg));
    for (int k = 0; k < K; ++k) {
      const float tau = (float)(k - K / 2);
      const float env = expf(inv2s2 * (-(tau * tau)));
      const float ph = mu * tau;
      bank[(size_t)(2 * f) * K + k] = (norm * cosf(ph)) * env;
      bank[(size_t)(2 * f + 1) * K + k] = (norm * sinf(ph)) * env;
    }
  }
}

/* impulse_responses.py:74-80 */
static void gaussian_windows(const float* pool_w, int F, int K, float* win) {
  for (int f = 0; f < F; ++f) {
    float s = pool_w[f];
    const float lo = (float)(2.0 / (double)K);
    s = s < lo ? lo : (s > 0.5f ? 0.5f : s);
    const float den = (s * 0.5f) * (float)(K - 1);
    for (int k = 0; k < K; ++k) {
      const float r = ((float)k - (float)(0.5 * (double)(K - 1))) / den;
      win[(size_t)f * K + k] = expf(-0.5f * (r * r));
    }
  }
}

/* Whole path, frontend.py:78-89.  x (B,1,T) -> out (B,F,N); p_out (B,F,N) optional (floored pooled
 * energies).  use_double != 0: sums and PCEN carried in double (tolerance calibration). */
int leaf_oracle_forward(const float* x, int B, int T, const float* kernel, const float* pool_w,
                        const float* pool_b, const float* alpha, const float* delta, const float* root,
                        const float* ema_w, int F, int K, int H, int compression, int use_double,
                        float* out, float* p_out) {
  const int pl = pad_left(K), N = leaf_oracle_num_frames(T, K, H);
  float* bank = (float*)malloc(sizeof(float) * (size_t)2 * F * K);
  float* win = (float*)malloc(sizeof(float) * (size_t)F * K);
  if (!bank || !win) { free(bank); free(win); return -1; }
  gabor_bank(kernel, F, K, bank);
  gaussian_windows(pool_w, F, K, win);
  int rc = 0;
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int b = 0; b < B; ++b) {
    for (int f = 0; f < F; ++f) {
      const float* xb = x + (size_t)b * T;
      const float* wr = bank + (size_t)(2 * f) * K;
      const float* wi = wr + K;
      float* e = (float*)malloc(sizeof(float) * (size_t)T);
      double* pd = (double*)malloc(sizeof(double) * (size_t)N);
      if (!e || !pd) { free(e); free(pd); rc = -1; continue; }
      /* convolution.py:91-98 (zero 'same' padding, cross-correlation) + frontend.py:15-19 */
      for (int t = 0; t < T; ++t) {
        int k0 = pl - t; if (k0 < 0) k0 = 0;
        int k1 = T - 1 - t + pl; if (k1 > K - 1) k1 = K - 1;
        if (use_double) {
          double re = 0.0, im = 0.0;
          for (int k = k0; k <= k1; ++k) { const double v = xb[t - pl + k]; re += v * wr[k]; im += v * wi[k]; }
          e[t] = (float)(re * re + im * im);
        } else {
          float re = 0.f, im = 0.f;
          for (int k = k0; k <= k1; ++k) { const float v = xb[t - pl + k]; re += v * wr[k]; im += v * wi[k]; }
          e[t] = re * re + im * im;
        }
      }
      /* pooling.py:31-42 (+bias), frontend.py:84 (floor) */
      for (int n = 0; n < N; ++n) {
        const int t0 = n * H - pl;
        double acc = 0.0;
        float accf = 0.f;
        for (int k = 0; k < K; ++k) {
          const int t = t0 + k;
          if (t < 0 || t >= T) continue;
          if (use_double) acc += (double)win[(size_t)f * K + k] * e[t]; else accf += win[(size_t)f * K + k] * e[t];
        }
        double p = use_double ? acc : (double)accf;
        if (pool_b) p = use_double ? p + pool_b[f] : (double)((float)p + pool_b[f]);
        if (p < 1e-5f) p = 1e-5f;
        pd[n] = p;
        if (p_out) p_out[((size_t)b * F + f) * N + n] = (float)p;
      }
      float* o = out + ((size_t)b * F + f) * N;
      if (!compression) {
        for (int n = 0; n < N; ++n) o[n] = (float)pd[n];
      } else {
        /* postprocessing.py:13-28 and 62-69 */
        float w = ema_w[f]; w = w < 0.f ? 0.f : (w > 1.f ? 1.f : w);
        const float a = alpha[f] < 1.0f ? alpha[f] : 1.0f;
        const float r = root[f] > 1.0f ? root[f] : 1.0f;
        const float q = 1.0f / r;
        if (use_double) {
          double m = pd[0];
          for (int n = 0; n < N; ++n) {
            m = (double)w * pd[n] + (1.0 - (double)w) * m;
            o[n] = (float)(pow(pd[n] / pow(1e-12 + m, (double)a) + (double)delta[f], (double)q) - pow((double)delta[f], (double)q));
          }
        } else {
          float m = (float)pd[0];
          for (int n = 0; n < N; ++n) {
            const float pn = (float)pd[n];
            m = (w * pn) + ((1.0f - w) * m);
            o[n] = powf(pn / powf(1e-12f + m, a) + delta[f], q) - powf(delta[f], q);
          }
        }
      }
      free(e); free(pd);
    }
  }
  free(bank); free(win);
  return rc;
}
