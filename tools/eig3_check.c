// Pins csrc/we_eig3.cuh (the restatement of LAPACK dgeev for symmetric 3x3 tensors) against the
// OpenBLAS that NumPy bundles, i.e. against what `numpy.linalg.eig` executes for the reference
// (maths/transformations.py:121-124).  Test infrastructure, not product code.
//   g++ -O2 -ffp-contract=off -mfma -x c++ tools/eig3_check.c -ldl -o /tmp/eig3_check
//   /tmp/eig3_check <libscipy_openblas64_.so> [n_samples] [seed]
// Draws rigid TIP3P waters with random orientation / position (float32 coordinates, fixture masses),
// builds the inertia tensor like MDAnalysis `moment_of_inertia`, and reports how many eigenvalue /
// eigenvector results are bit-identical and how many sign patterns agree.
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../waterentropy_b200/csrc/we_eig3.cuh"

typedef int64_t I;
typedef void (*dgeev_t)(const char*, const char*, I*, double*, I*, double*, double*, double*, I*, double*, I*,
                        double*, I*, I*);
static uint64_t rs = 88172645463325252ULL;
static double urand(void) { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (rs >> 11) * (1.0 / 9007199254740992.0); }
static double nrand(void) { return sqrt(-2 * log(urand() + 1e-300)) * cos(6.283185307179586 * urand()); }

static void water_tensor(double t[9]) {
  // random unit quaternion -> rotation; TIP3P geometry r_OH 0.9572 A, angle 104.52 deg
  double q[4], n = 0;
  for (int i = 0; i < 4; ++i) { q[i] = nrand(); n += q[i] * q[i]; }
  n = sqrt(n);
  for (int i = 0; i < 4; ++i) q[i] /= n;
  const double R[3][3] = {
      {1 - 2 * (q[2] * q[2] + q[3] * q[3]), 2 * (q[1] * q[2] - q[0] * q[3]), 2 * (q[1] * q[3] + q[0] * q[2])},
      {2 * (q[1] * q[2] + q[0] * q[3]), 1 - 2 * (q[1] * q[1] + q[3] * q[3]), 2 * (q[2] * q[3] - q[0] * q[1])},
      {2 * (q[1] * q[3] - q[0] * q[2]), 2 * (q[2] * q[3] + q[0] * q[1]), 1 - 2 * (q[1] * q[1] + q[2] * q[2])}};
  const double half = 0.5 * 104.52 * M_PI / 180.0, r = 0.9572;
  const double loc[3][3] = {{0, 0, 0}, {r * sin(half), 0, r * cos(half)}, {-r * sin(half), 0, r * cos(half)}};
  const double c0[3] = {urand() * 100, urand() * 100, urand() * 100};
  const double mass[3] = {16.0, 1.008, 1.008};
  float p[3][3];
  for (int a = 0; a < 3; ++a)
    for (int i = 0; i < 3; ++i)
      p[a][i] = (float)(c0[i] + R[i][0] * loc[a][0] + R[i][1] * loc[a][1] + R[i][2] * loc[a][2] + 0.02 * nrand());
  double com[3], mt = (mass[0] + mass[1]) + mass[2];
  for (int i = 0; i < 3; ++i) com[i] = (((double)p[0][i] * mass[0] + (double)p[1][i] * mass[1]) + (double)p[2][i] * mass[2]) / mt;
  double x[3], y[3], z[3];
  for (int a = 0; a < 3; ++a) { x[a] = (double)p[a][0] - com[0]; y[a] = (double)p[a][1] - com[1]; z[a] = (double)p[a][2] - com[2]; }
#define S3(e) ((e(0) + e(1)) + e(2))
#define XX(a) (mass[a] * (y[a] * y[a] + z[a] * z[a]))
#define YY(a) (mass[a] * (x[a] * x[a] + z[a] * z[a]))
#define ZZ(a) (mass[a] * (x[a] * x[a] + y[a] * y[a]))
#define XY(a) (mass[a] * x[a] * y[a])
#define XZ(a) (mass[a] * x[a] * z[a])
#define YZ(a) (mass[a] * y[a] * z[a])
  t[0] = S3(XX); t[4] = S3(YY); t[8] = S3(ZZ);
  t[1] = t[3] = -S3(XY); t[2] = t[6] = -S3(XZ); t[5] = t[7] = -S3(YZ);
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s libopenblas.so [n] [seed]\n", argv[0]); return 2; }
  void* H = dlopen(argv[1], RTLD_NOW);
  if (!H) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  dgeev_t dgeev = (dgeev_t)dlsym(H, "scipy_dgeev_64_");
  if (!dgeev) dgeev = (dgeev_t)dlsym(H, "dgeev_64_");
  if (!dgeev) { fprintf(stderr, "no dgeev symbol\n"); return 2; }
  const char* (*corename)(void) = (const char* (*)(void))dlsym(H, "scipy_openblas_get_corename64_");
  const long n = argc > 2 ? atol(argv[2]) : 1000000;
  if (argc > 3) rs ^= (uint64_t)atoll(argv[3]) * 0x9E3779B97F4A7C15ULL;
  // kernel family of the loaded OpenBLAS (argv[4] overrides): AVX-512 cores fuse, AVX2 cores do not
  const char* core = corename ? corename() : "SkylakeX";
  int flavour = (strstr(core, "Skylake") || strstr(core, "SKYLAKE") || strstr(core, "Cooper") || strstr(core, "Sapphire")) ? WE_EIG3_SKYLAKEX : WE_EIG3_HASWELL;
  if (argc > 4) flavour = atoi(argv[4]);
  I three = 3, lwork = -1, info = 0;
  double wq[1], a[9], wr[3], wi[3], vl[9], vr[9];
  dgeev("N", "V", &three, a, &three, wr, wi, vl, &three, vr, &three, wq, &lwork, &info);
  lwork = (I)wq[0];
  double* work = (double*)malloc(sizeof(double) * (size_t)lwork);
  long bit_w = 0, bit_v = 0, sign_ok = 0, order_ok = 0, unsupported = 0, close_ok = 0, axes_ok = 0;
  double worst = 0;
  for (long s = 0; s < n; ++s) {
    double t[9];
    water_tensor(t);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i + 3 * j] = t[3 * i + j];
    dgeev("N", "V", &three, a, &three, wr, wi, vl, &three, vr, &three, work, &lwork, &info);
    double mw[3], mv[9];
    const int st = we_dgeev3(t, mw, mv, flavour);
    if (st) { ++unsupported; continue; }
    const bool bw = memcmp(mw, wr, 24) == 0, bv = memcmp(mv, vr, 72) == 0;
    bit_w += bw; bit_v += bv;
    bool ord = true, sg = true, cl = true;
    for (int k = 0; k < 3; ++k) {
      if (fabs(mw[k] - wr[k]) > 1e-9 * fabs(wr[k])) ord = false;
      double dot = 0;
      for (int i = 0; i < 3; ++i) dot += mv[3 * k + i] * vr[3 * k + i];
      if (dot < 0) sg = false;
      for (int i = 0; i < 3; ++i) { const double e = fabs(mv[3 * k + i] - vr[3 * k + i]); if (e > 1e-12) cl = false; if (ord && dot > 0 && e > worst) worst = e; }
    }
    order_ok += ord; sign_ok += (ord && sg); close_ok += (ord && sg && cl);
    // what the recipe consumes: MDAnalysis principal_axes from both
    {
      double ax[3][3], moi[3];
      we_principal_axes(t, ax, moi, flavour);
      int o[3] = {0, 1, 2};
      for (int i = 1; i < 3; ++i) { const int v = o[i]; int j = i - 1; while (j >= 0 && wr[o[j]] > wr[v]) { o[j + 1] = o[j]; --j; } o[j + 1] = v; }
      double rx[3][3];
      for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) rx[i][k] = vr[3 * o[2 - i] + k];
      const double cx = rx[0][1] * rx[1][2] - rx[0][2] * rx[1][1], cy = rx[0][2] * rx[1][0] - rx[0][0] * rx[1][2], cz = rx[0][0] * rx[1][1] - rx[0][1] * rx[1][0];
      if (cx * rx[2][0] + cy * rx[2][1] + cz * rx[2][2] < 0) for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) rx[i][k] = -rx[i][k];
      bool same = true;
      for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) if (fabs(rx[i][k] - ax[i][k]) > 1e-12) same = false;
      axes_ok += same;
    }
    if (!(ord && sg) && (n - s) <= 0) printf("mismatch at %ld\n", s);
  }
  printf("{\"core\": \"%s\", \"flavour\": %d, \"samples\": %ld, \"unsupported\": %ld, \"eigenvalues_bit_identical\": %ld, "
         "\"eigenvectors_bit_identical\": %ld, \"same_order\": %ld, \"same_order_and_signs\": %ld, "
         "\"within_1e-12\": %ld, \"principal_axes_agree\": %ld, \"worst_abs_diff_when_signs_agree\": %.3g}\n",
         core, flavour, n, unsupported, bit_w, bit_v, order_ok, sign_ok, close_ok, axes_ok, worst);
  return 0;
}
