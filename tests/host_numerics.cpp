// Host build of waterentropy_b200/csrc/we_numerics.cuh for the CPU test-suite:
// the same source the CUDA kernels compile, so the float32/float64 replay can be
// pinned against libm / the oracle without a GPU.
// g++ -O2 -std=c++17 -fno-builtin -ffp-contract=off -shared -fPIC
#include "../waterentropy_b200/csrc/we_numerics.cuh"

static WePowTables tabs() { return WePowTables{&we_h_log2tab[0][0], we_h_exp2tab}; }

extern "C" {
void h_powf2(const float* x, int n, float* out) {
  WePowTables T = tabs();
  for (int i = 0; i < n; ++i) out[i] = we_powf2(x[i], T);
}
void h_powf2_fast(const float* x, int n, float* out) {
  WePowTables T = tabs();
  for (int i = 0; i < n; ++i) out[i] = we_powf2_fast(x[i], T);
}
void h_angle_cos(const float* abc, const float* dim, int n, float* out) {
  WePowTables T = tabs();
  for (int i = 0; i < n; ++i) out[i] = we_angle_cos(abc + 9 * i, abc + 9 * i + 3, abc + 9 * i + 6, dim, T);
}
void h_dist_ortho(const float* pairs, const float* dims, int n, double* out) {
  WeBox b;
  for (int t = 0; t < 3; ++t) { b.box[t] = dims[t]; b.inv[t] = (float)(1.0 / (double)dims[t]); }
  b.triclinic = 0;
  for (int i = 0; i < n; ++i) {
    const float* p = pairs + 6 * i;
    out[i] = we_dist_ortho(p[0], p[1], p[2], p[3], p[4], p[5], b);
  }
}
int h_angle_gt_90(float c) { return we_angle_gt_90(c) ? 1 : 0; }
}
extern "C" void h_pow2(const double* x, int n, double* out) {
  for (int i = 0; i < n; ++i) out[i] = we_pow2(x[i]);
}
