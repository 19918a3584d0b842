/* Minimal C client of libleafk.so (include/leafk.h): what a non-Python host would write.
 *   gcc -I include -I /usr/local/cuda/include examples/c_abi_example.c -L leaf_pytorch_b200/lib -lleafk -L/usr/local/cuda/lib64 -lcudart \
 *       -Wl,-rpath,leaf_pytorch_b200/lib -o examples/c_abi_example
 * Device memory comes from the CUDA runtime here (from PyTorch's allocator in the Python binding). */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "leafk.h"

int main(void) {
  const int B = 4, T = 16000, F = 40;
  leafk_config cfg = {F, 401, 160, 1e-12f, 1e-5f, 1, LEAFK_ALGO_AUTO, LEAFK_INPUT_F32, LEAFK_OUTPUT_F32, /*prep=*/NULL};
  const int N = leafk_num_frames(T, cfg.K, cfg.H);
  printf("libleafk version %d, %d frames per clip, tensor-core kernel %s\n", leafk_version(), N,
         leafk_tc_supported(F, cfg.K, cfg.H) ? "available" : "not used");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { printf("no GPU: stopping after the host-only calls\n"); return 0; }

  float *x, *out, *par;          /* par: kernel(2F) | pool_w | pool_b | alpha | delta | root | ema_w */
  float* hx = (float*)malloc(sizeof(float) * B * T);
  float* hp = (float*)malloc(sizeof(float) * 8 * F);
  for (int i = 0; i < B * T; ++i) hx[i] = (float)((i * 2654435761u) >> 8 & 0xffff) / 65536.0f - 0.5f;
  for (int f = 0; f < F; ++f) {
    hp[2 * f] = 0.05f + 2.8f * f / F; hp[2 * f + 1] = 60.0f - 1.2f * f;             /* centre, width */
    hp[2 * F + f] = 0.4f; hp[3 * F + f] = 1.0f;                                      /* pool width, bias */
    hp[4 * F + f] = 0.96f; hp[5 * F + f] = 2.0f; hp[6 * F + f] = 2.0f; hp[7 * F + f] = 0.04f;
  }
  cudaMalloc((void**)&x, sizeof(float) * B * T);
  cudaMalloc((void**)&out, sizeof(float) * B * F * N);
  cudaMalloc((void**)&par, sizeof(float) * 8 * F);
  cudaMemcpy(x, hx, sizeof(float) * B * T, cudaMemcpyHostToDevice);
  cudaMemcpy(par, hp, sizeof(float) * 8 * F, cudaMemcpyHostToDevice);
  leafk_params prm = {par, par + 2 * F, par + 3 * F, par + 4 * F, par + 5 * F, par + 6 * F, par + 7 * F};
  size_t wsz = leafk_workspace_bytes(&cfg, B, N);
  void* ws;
  cudaMalloc(&ws, wsz);
  int rc = leafk_forward(&cfg, &prm, x, B, T, out, NULL, ws, wsz, /*stream=*/NULL);
  if (rc != LEAFK_OK) { printf("leafk_forward failed: %s\n", leafk_last_error()); return 1; }
  float* ho = (float*)malloc(sizeof(float) * B * F * N);
  cudaMemcpy(ho, out, sizeof(float) * B * F * N, cudaMemcpyDeviceToHost);
  printf("out[0,0,0..3] = %.6f %.6f %.6f %.6f\n", ho[0], ho[1], ho[2], ho[3]);

  /* one training step: the training forward saves (p, Q_mu, Q_sigma, Q_poolw); the backward runs no correlation */
  if (leafk_train_supported(F, cfg.K, cfg.H)) {
    float *saved, *gout, *gpar;
    void* tws; void* bws;
    const size_t bfn = (size_t)B * F * N;
    cudaMalloc((void**)&saved, sizeof(float) * 4 * bfn);
    cudaMalloc((void**)&gout, sizeof(float) * bfn);
    cudaMalloc((void**)&gpar, sizeof(float) * 8 * F);
    for (size_t i = 0; i < bfn; ++i) ho[i] = 1.0f;                                   /* d(sum out)/d out */
    cudaMemcpy(gout, ho, sizeof(float) * bfn, cudaMemcpyHostToDevice);
    size_t tsz = leafk_train_workspace_bytes(&cfg, B, T), bsz = leafk_backward_saved_workspace_bytes(&cfg, B, T, 0);
    cudaMalloc(&tws, tsz); cudaMalloc(&bws, bsz);
    leafk_grads g = {gpar, gpar + 2 * F, gpar + 3 * F, gpar + 4 * F, gpar + 5 * F, gpar + 6 * F, gpar + 7 * F};
    rc = leafk_forward_train(&cfg, &prm, x, B, T, out, saved, tws, tsz, NULL);
    if (rc == LEAFK_OK) rc = leafk_backward_saved(&cfg, &prm, NULL, B, T, gout, saved, &g, NULL, bws, bsz, NULL);
    if (rc != LEAFK_OK) { printf("training step failed: %s\n", leafk_last_error()); return 1; }
    cudaMemcpy(hp, gpar, sizeof(float) * 8 * F, cudaMemcpyDeviceToHost);
    printf("d(sum out)/d(centre, width) of filter 0 = %.6e %.6e\n", hp[0], hp[1]);
  }
  return 0;
}
