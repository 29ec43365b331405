/* C99 caller of the C ABI: b200ppo_gae (include/b200ppo.h; reference ppo.py:351-394) through the header,
 * with nothing but the CUDA runtime C API around it.  Usage: gae_abi_test <in.bin> <out.bin>
 *   in.bin : int32 T, int32 B, float lambda, float gamma, then rewards[T*B] f32, values[T*B] f32,
 *            last_value[B] f32, done[T*B] u8, truncation[T*B] u8
 *   out.bin: advantages[T*B] f32
 * tests/test_gpu_cabi.py feeds it the reference's own known-answer recipe (ppo_test.py:229-264). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cuda_runtime_api.h>

#include "b200ppo.h"

#define CHECK_CUDA(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 3; } } while (0)

static void* upload(const void* host, size_t bytes) {
  void* d = NULL;
  if (cudaMalloc(&d, bytes) != cudaSuccess) return NULL;
  if (cudaMemcpy(d, host, bytes, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
  return d;
}

int main(int argc, char** argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s in.bin out.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int32_t T = 0, B = 0;
  float lambda_ = 0.0f, gamma = 0.0f;
  if (fread(&T, 4, 1, f) != 1 || fread(&B, 4, 1, f) != 1 || fread(&lambda_, 4, 1, f) != 1 || fread(&gamma, 4, 1, f) != 1) return 2;
  const size_t n = (size_t)T * (size_t)B;
  float* rewards = (float*)malloc(4 * n);
  float* values = (float*)malloc(4 * n);
  float* last = (float*)malloc(4 * (size_t)B);
  uint8_t* done = (uint8_t*)malloc(n);
  uint8_t* trunc = (uint8_t*)malloc(n);
  float* adv = (float*)malloc(4 * n);
  if (!rewards || !values || !last || !done || !trunc || !adv) return 2;
  if (fread(rewards, 4, n, f) != n || fread(values, 4, n, f) != n || fread(last, 4, (size_t)B, f) != (size_t)B ||
      fread(done, 1, n, f) != n || fread(trunc, 1, n, f) != n) { fprintf(stderr, "short input\n"); return 2; }
  fclose(f);
  if (b200ppo_version() != 100) { fprintf(stderr, "unexpected library version\n"); return 4; }
  /* argument validation happens before any launch */
  if (b200ppo_gae(NULL, NULL, NULL, NULL, NULL, NULL, T, B, lambda_, gamma, NULL) != B200PPO_EINVAL) return 4;
  void* d_r = upload(rewards, 4 * n);
  void* d_v = upload(values, 4 * n);
  void* d_l = upload(last, 4 * (size_t)B);
  void* d_d = upload(done, n);
  void* d_t = upload(trunc, n);
  void* d_a = NULL;
  CHECK_CUDA(cudaMalloc(&d_a, 4 * n));
  if (!d_r || !d_v || !d_l || !d_d || !d_t) { fprintf(stderr, "device allocation / upload failed\n"); return 3; }
  cudaStream_t stream;
  CHECK_CUDA(cudaStreamCreate(&stream));
  const int rc = b200ppo_gae((void*)stream, (const float*)d_r, (const float*)d_v, (const float*)d_l,
                             (const uint8_t*)d_d, (const uint8_t*)d_t, T, B, lambda_, gamma, (float*)d_a);
  if (rc != 0) { fprintf(stderr, "b200ppo_gae: %s\n", b200ppo_error_string(rc)); return 5; }
  CHECK_CUDA(cudaStreamSynchronize(stream));
  CHECK_CUDA(cudaMemcpy(adv, d_a, 4 * n, cudaMemcpyDeviceToHost));
  f = fopen(argv[2], "wb");
  if (!f || fwrite(adv, 4, n, f) != n) { perror(argv[2]); return 2; }
  fclose(f);
  printf("b200ppo_gae through the C header: T=%d B=%d, %zu advantages written\n", (int)T, (int)B, n);
  return 0;
}
