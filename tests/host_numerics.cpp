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

// ---- we_eig3.cuh: LAPACK dgeev replay for symmetric 3x3 tensors -----------------------------
#include "../waterentropy_b200/csrc/we_eig3.cuh"
extern "C" {
// out [n][25] = wr(3), vr(9), principal axes(9), MOI_axis(3), status -- same layout as we_test_eig3
void h_eig3(const double* t, int n, int flavour, double* out) {
  for (int i = 0; i < n; ++i) {
    double wr[3], vr[9], ax[3][3], moi[3];
    const int st = we_dgeev3(t + 9 * i, wr, vr, flavour);
    we_principal_axes(t + 9 * i, ax, moi, flavour);
    double* o = out + 25 * i;
    for (int k = 0; k < 3; ++k) o[k] = wr[k];
    for (int k = 0; k < 9; ++k) o[3 + k] = vr[k];
    for (int k = 0; k < 9; ++k) o[12 + k] = ax[k / 3][k % 3];
    for (int k = 0; k < 3; ++k) o[21 + k] = moi[k];
    o[24] = (double)st;
  }
}
// x87 dnrm2 emulation (integer soft-float) and the same sum on the host's real x87 unit
void h_x87_nrm2(const double* x, int n, int len, double* out) {
  for (int i = 0; i < n; ++i) out[i] = we_x87_nrm2(x + (size_t)len * i, len);
}
void h_long_double_nrm2(const double* x, int n, int len, double* out) {
  for (int i = 0; i < n; ++i) {
    long double acc = 0.0L;
    for (int k = 0; k < len; ++k) acc = acc + (long double)x[(size_t)len * i + k] * (long double)x[(size_t)len * i + k];
    out[i] = (double)__builtin_sqrtl(acc);
  }
}
}

// ---- we_multi_ua.cuh: the `multiple_UAs` force / torque item function, host build -------------
// Same argument list as the C ABI's we_molecule_force_torque_multi (minus the device ordinal), so the
// CPU suite can drive the product's own host-side packing (recipes/forces_torques.py) through it.
struct WeDiag { unsigned long long eig_fallbacks; };
#include "../waterentropy_b200/csrc/we_multi_ua.cuh"
extern "C" int h_molecule_force_torque_multi(int device, const float* pos, const float* frc, const double* masses,
                                             int n_pool, const int* mol_ptr, const int* mol_atoms, int n_mol,
                                             const int* ua_ptr, const int* ua_atom, const int* hv_ptr, const int* hv,
                                             const int* lt_ptr, const int* lt, const int* hin_ptr, const int* hin,
                                             const float* box3, int eig_flavour, double* out_mol, double* out_ua) {
  (void)device; (void)n_pool;
  WeMultiUA A;
  A.pos = pos; A.frc = frc; A.mass = masses; A.mol_ptr = mol_ptr; A.mol_atoms = mol_atoms; A.ua_ptr = ua_ptr;
  A.ua_atom = ua_atom; A.hv_ptr = hv_ptr; A.hv = hv; A.lt_ptr = lt_ptr; A.lt = lt; A.hin_ptr = hin_ptr; A.hin = hin;
  A.box[0] = box3[0]; A.box[1] = box3[1]; A.box[2] = box3[2];
  A.n_mol = n_mol; A.n_ua = ua_ptr[n_mol]; A.eig_flavour = eig_flavour;
  WeDiag d = {0};
  for (int item = 0; item < A.n_mol + A.n_ua; ++item) we_multi_ua_item(A, item, &d, out_mol, out_ua);
  return (int)d.eig_fallbacks;
}
