/* Host-side packing probe: how fast can T threads gather the heavy atoms (12 of every 36 bytes of a water)
 * out of [N][3] float32 frames into a compact buffer?   gcc -O3 -march=native -fopenmp tools/pack_probe.c -o /tmp/pack_probe */
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
int main(int argc, char** argv) {
  const long n_w = 100000, n_sol = 14274, N = 3 * n_w + n_sol, F = 96;
  const long H = n_w + 7782;
  float* src = (float*)malloc(sizeof(float) * 3 * N * F);
  float* dst = (float*)malloc(sizeof(float) * 3 * H * F);
  int* heavy = (int*)malloc(sizeof(int) * H);
  long h = 0;
  for (long i = 0; i < n_sol && h < 7782; i += 2) heavy[h++] = (int)i;  /* every other solute atom */
  for (long w = 0; w < n_w; ++w) heavy[h++] = (int)(n_sol + 3 * w);
  for (long i = 0; i < 3 * N * F; ++i) src[i] = (float)i;
  memset(dst, 0, sizeof(float) * 3 * H * F);
  for (int T = 1; T <= 16; T *= 2) {
    omp_set_num_threads(T);
    double best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      const double t0 = now();
#pragma omp parallel for schedule(static)
      for (long c = 0; c < F * 16; ++c) {  /* 16 chunks per frame */
        const long f = c / 16, k = c % 16;
        const long a = H * k / 16, b = H * (k + 1) / 16;
        const float* s = src + 3 * N * f;
        float* d = dst + 3 * H * f;
        for (long i = a; i < b; ++i) {
          const long j = 3L * heavy[i];
          d[3 * i] = s[j]; d[3 * i + 1] = s[j + 1]; d[3 * i + 2] = s[j + 2];
        }
      }
      const double t = now() - t0;
      if (t < best) best = t;
    }
    printf("threads %2d: %.2f ms per 96 frames  (%.1f GB/s of source swept, %.1f GB/s written)\n", T, best * 1e3,
           sizeof(float) * 3.0 * N * F / best / 1e9, sizeof(float) * 3.0 * H * F / best / 1e9);
  }
  return 0;
}
