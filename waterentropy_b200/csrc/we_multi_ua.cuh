// `multiple_UAs` molecules: force / torque covariance of a molecule with more than one united atom.
//
// Reference: recipes/forces_torques.py:14-93 with maths/transformations.py:82-353 ("not applicable to
// water" upstream).  A molecule gets
//   * whole-molecule matrices: principal axes of ALL atoms (MDAnalysis principal_axes = dgeev replay,
//     we_eig3.cuh), moments = eigenvalues of the inertia tensor of its UNITED ATOMS (heavy-atom
//     positions, masses incl. the hydrogens bonded inside the molecule; transformations.py:82-101,138-163),
//     no halving (forces_torques.py:31-36);
//   * per united atom X a frame built from its bonds (get_bonded_axes, :286-353; get_custom_axes,
//     :166-228), "normalised" and given moments by get_custom_PI_MOI (:231-283), then torque about X and
//     rotated mass-weighted force of the WHOLE molecule in that frame.
// The float32 / float64 placement of every operation follows NumPy's promotion of the reference's
// expressions (positions, forces and box lengths are float32): get_vector and the bond normalisation run
// in float32, X-relative coordinates are float32 differences, everything after is fp64.  Summation orders
// are NumPy's (rows of an axis-0 reduction one after the other; pairwise blocks for 1-D sums of 8 or more).
// Quirks kept because outputs must match: get_vector adds the box length to EVERY negative component before
// the half-box test (maths/trig.py:150-153); get_flipped_axes divides column j by the norm of ROW j
// (:239-240); a united atom with one heavy and one light neighbour takes the ORIGIN as its third point
// (:337-340).
//
// One thread per item: item 0 of a molecule = the whole molecule, items 1.. = its united atoms.
#pragma once
#include "we_numerics.cuh"
#include "we_eig3.cuh"
// The per-item function also compiles for the host (tests/host_numerics.cpp runs it against the goldens
// without a GPU); the product only ever launches the kernel.
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WE_MU __device__ __forceinline__
#define WE_MU_FN __device__ __forceinline__
#else
#define WE_MU inline
#define WE_MU_FN inline
#endif

struct WeMultiUA {
  // atom pool: the molecules' atoms and every atom bonded to one of their united atoms
  const float* pos;      // [na][3]
  const float* frc;      // [na][3]
  const double* mass;    // [na]
  // molecules: CSR over pool indices, ascending atom index (the order MDAnalysis keeps a group in)
  const int* mol_ptr;    // [n_mol + 1]
  const int* mol_atoms;
  // united atoms (heavy atoms, 2 <= mass <= 999) of the molecules, molecule after molecule
  const int* ua_ptr;     // [n_mol + 1] first united atom of each molecule
  const int* ua_atom;    // [n_ua] pool index
  const int* hv_ptr;     // [n_ua + 1] CSR: heavy atoms bonded to it (anywhere in the system), ascending index
  const int* hv;
  const int* lt_ptr;     // [n_ua + 1] CSR: light atoms (1 <= mass <= 1.1) bonded to it, ascending index
  const int* lt;
  const int* hin_ptr;    // [n_ua + 1] CSR: light atoms bonded to it INSIDE the molecule (united-atom mass)
  const int* hin;
  float box[3];
  int n_mol, n_ua, eig_flavour;
};

// numpy's pairwise summation of a contiguous 1-D float64 array (umath/loops_utils.h: < 8 straight, <= 128
// eight accumulators, larger halves).  get(i) = element i.  The halving is walked with an explicit stack
// (device recursion would need a stack-size limit the caller cannot know).
template <class G>
WE_MU double we_np_block_sum(G get, int lo, int n) {
  if (n < 8) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = we_add(r, get(lo + i));
    return r;
  }
  double r[8];
  for (int k = 0; k < 8; ++k) r[k] = get(lo + k);
  int i = 8;
  for (; i < n - (n % 8); i += 8)
    for (int k = 0; k < 8; ++k) r[k] = we_add(r[k], get(lo + i + k));
  double res = we_add(we_add(we_add(r[0], r[1]), we_add(r[2], r[3])), we_add(we_add(r[4], r[5]), we_add(r[6], r[7])));
  for (; i < n; ++i) res = we_add(res, get(lo + i));
  return res;
}
template <class G>
WE_MU_FN double we_np_sum(G get, int lo0, int n0) {
  if (n0 <= 128) return we_np_block_sum(get, lo0, n0);
  int s_lo[28], s_n[28], s_state[28], sp = 0, nv = 0;
  double vals[28];
  s_lo[0] = lo0; s_n[0] = n0; s_state[0] = 0; sp = 1;
  while (sp > 0) {
    const int t = sp - 1, lo = s_lo[t], n = s_n[t];
    if (n <= 128) { vals[nv++] = we_np_block_sum(get, lo, n); --sp; continue; }
    int n2 = n / 2;
    n2 -= n2 % 8;
    if (s_state[t] == 0) { s_state[t] = 1; s_lo[sp] = lo; s_n[sp] = n2; s_state[sp] = 0; ++sp; }
    else if (s_state[t] == 1) { s_state[t] = 2; s_lo[sp] = lo + n2; s_n[sp] = n - n2; s_state[sp] = 0; ++sp; }
    else { vals[nv - 2] = we_add(vals[nv - 2], vals[nv - 1]); --nv; --sp; }
  }
  return vals[0];
}

// maths/trig.py:139-157 in float32
WE_MU float we_get_vector_f(float a, float b, float box) {
  float d = we_fsub(b, a);
  if (d < 0.f && d < we_fmul(0.5f, box)) d = we_fadd(d, box);
  if (d > 0.f && d > we_fmul(0.5f, box)) d = we_fsub(d, box);
  return d;
}
// the same with a float64 second point (np.zeros(3) - float32 position -> float64)
WE_MU double we_get_vector_d(float a, double b, float box) {
  double d = we_add(b, -(double)a);
  if (d < 0.0 && d < (double)we_fmul(0.5f, box)) d = we_add(d, (double)box);
  if (d > 0.0 && d > (double)we_fmul(0.5f, box)) d = we_add(d, -(double)box);
  return d;
}

WE_MU void we_cross(const double a[3], const double b[3], double c[3]) {
  c[0] = we_add(we_mul(a[1], b[2]), -we_mul(a[2], b[1]));
  c[1] = we_add(we_mul(a[2], b[0]), -we_mul(a[0], b[2]));
  c[2] = we_add(we_mul(a[0], b[1]), -we_mul(a[1], b[0]));
}

// inertia tensor of a group about its centre of mass (MDAnalysis moment_of_inertia) and the centre
template <class P, class M>
WE_MU_FN void we_group_tensor(int n, P pos, M mass, double com[3], double t[9]) {
  double s[3] = {0.0, 0.0, 0.0};
  for (int k = 0; k < n; ++k)
    for (int c = 0; c < 3; ++c) s[c] = we_add(s[c], we_mul((double)pos(k, c), mass(k)));
  const double mt = we_np_sum([&](int k) { return mass(k); }, 0, n);
  for (int c = 0; c < 3; ++c) com[c] = we_ddiv(s[c], mt);
  auto rel = [&](int k, int c) { return we_add((double)pos(k, c), -com[c]); };
  auto sq = [&](int k, int a, int b) {
    return we_mul(mass(k), we_add(we_mul(rel(k, a), rel(k, a)), we_mul(rel(k, b), rel(k, b))));
  };
  auto pr = [&](int k, int a, int b) { return we_mul(we_mul(mass(k), rel(k, a)), rel(k, b)); };
  t[0] = we_np_sum([&](int k) { return sq(k, 1, 2); }, 0, n);
  t[4] = we_np_sum([&](int k) { return sq(k, 0, 2); }, 0, n);
  t[8] = we_np_sum([&](int k) { return sq(k, 0, 1); }, 0, n);
  t[1] = t[3] = -we_np_sum([&](int k) { return pr(k, 0, 1); }, 0, n);
  t[2] = t[6] = -we_np_sum([&](int k) { return pr(k, 0, 2); }, 0, n);
  t[5] = t[7] = -we_np_sum([&](int k) { return pr(k, 1, 2); }, 0, n);
}

WE_MU void we_axes_of(const double t[9], int flavour, WeDiag* diag, double ax[3][3], double am[3]);

// torque about `centre` and rotated mass-weighted force of the molecule's atoms in the frame `ax`
// (maths/transformations.py:12-64).  centre_f32: the centre is a float32 position (united-atom frames:
// the difference is taken in float32), else the fp64 centre of mass.
WE_MU_FN void we_frame_force_torque(const WeMultiUA& A, int a0, int a1, const double ax[3][3], const double mi[3],
                                      bool centre_f32, const float cf[3], const double cd[3], double F[3], double tq[3]) {
  const double sq[3] = {sqrt(mi[0]), sqrt(mi[1]), sqrt(mi[2])};
  tq[0] = tq[1] = tq[2] = 0.0;
  float fs[3] = {0.f, 0.f, 0.f};
  for (int k = a0; k < a1; ++k) {
    const int a = A.mol_atoms[k];
    double p[3], g[3], rp[3], rg[3], cr[3];
    for (int c = 0; c < 3; ++c) {
      p[c] = centre_f32 ? (double)we_fsub(A.pos[3 * a + c], cf[c]) : we_add((double)A.pos[3 * a + c], -cd[c]);
      const float gf = A.frc[3 * a + c];
      g[c] = (double)gf;
      fs[c] = (k == a0) ? gf : we_fadd(fs[c], gf);  // float32 sum, row after row (transformations.py:48)
    }
    for (int i = 0; i < 3; ++i) {
      rp[i] = we_add(we_add(we_mul(p[0], ax[i][0]), we_mul(p[1], ax[i][1])), we_mul(p[2], ax[i][2]));
      rg[i] = we_add(we_add(we_mul(g[0], ax[i][0]), we_mul(g[1], ax[i][1])), we_mul(g[2], ax[i][2]));
    }
    we_cross(rp, rg, cr);
    for (int i = 0; i < 3; ++i) tq[i] = we_add(tq[i], we_ddiv(cr[i], sq[i]));
  }
  const double mt = we_np_sum([&](int k) { return A.mass[A.mol_atoms[a0 + k]]; }, 0, a1 - a0);
  const double msq = sqrt(mt);  // ** 0.5
  for (int i = 0; i < 3; ++i)
    F[i] = we_ddiv(we_add(we_add(we_mul((double)fs[0], ax[i][0]), we_mul((double)fs[1], ax[i][1])), we_mul((double)fs[2], ax[i][2])), msq);
}

WE_MU void we_axes_of(const double t[9], int flavour, WeDiag* diag, double ax[3][3], double am[3]) {
  if (we_principal_axes(t, ax, am, flavour) != WE_EIG3_OK) {
#ifdef __CUDA_ARCH__
    atomicAdd(&diag->eig_fallbacks, 1ULL);
    double I[3][3] = {{t[0], t[1], t[2]}, {t[3], t[4], t[5]}, {t[6], t[7], t[8]}};
    we_axes_jacobi(I, ax, am);
#else
    diag->eig_fallbacks += 1;  // host build (tests): counted, the dgeev result is kept
#endif
  }
}

// maths/transformations.py:166-228.  a: the united atom; b[0..nb): heavy neighbours; c: third point
// (float32 position, or the origin in fp64 when c_origin).
WE_MU_FN void we_custom_axes(const WeMultiUA& A, int a, const int* b, int nb, int c, bool c_origin, double ax[3][3]) {
  double axis1[3] = {0.0, 0.0, 0.0};
  for (int j = 0; j < nb; ++j) {
    float v[3];
    for (int k = 0; k < 3; ++k) v[k] = we_get_vector_f(A.pos[3 * a + k], A.pos[3 * b[j] + k], A.box[k]);
    const float d = we_fsqrt(we_fadd(we_fadd(we_fmul(v[0], v[0]), we_fmul(v[1], v[1])), we_fmul(v[2], v[2])));
    for (int k = 0; k < 3; ++k) axis1[k] = we_add(axis1[k], (double)we_fdiv(v[k], d));
  }
  const int from = nb > 2 ? b[0] : a;
  double acn[3];
  if (c_origin) {
    double v[3];
    for (int k = 0; k < 3; ++k) v[k] = we_get_vector_d(A.pos[3 * from + k], 0.0, A.box[k]);
    const double d = sqrt(we_add(we_add(we_mul(v[0], v[0]), we_mul(v[1], v[1])), we_mul(v[2], v[2])));
    for (int k = 0; k < 3; ++k) acn[k] = we_ddiv(v[k], d);
  } else {
    float v[3];
    for (int k = 0; k < 3; ++k) v[k] = we_get_vector_f(A.pos[3 * from + k], A.pos[3 * c + k], A.box[k]);
    const float d = we_fsqrt(we_fadd(we_fadd(we_fmul(v[0], v[0]), we_fmul(v[1], v[1])), we_fmul(v[2], v[2])));
    for (int k = 0; k < 3; ++k) acn[k] = (double)we_fdiv(v[k], d);
  }
  double axis2[3], axis3[3];
  if (nb > 2) we_cross(acn, axis1, axis2); else we_cross(axis1, acn, axis2);
  we_cross(axis1, axis2, axis3);
  for (int k = 0; k < 3; ++k) { ax[0][k] = axis1[k]; ax[1][k] = axis2[k]; ax[2][k] = axis3[k]; }
}

// out_mol [n_mol][24] = F(3), tq(3), outer(F, F) (9), outer(tq, tq) (9)
// out_ua  [n_ua][25]  = the same in the united atom's bonded-axes frame, then 1.0 if it has a frame, 0.0 if not
WE_MU_FN void we_multi_ua_item(const WeMultiUA& A, int item, WeDiag* diag, double* out_mol, double* out_ua) {
  double F[3], tq[3];
  double* o;
  if (item < A.n_mol) {
    // ---- whole molecule (forces_torques.py:38-55, transformations.py:104-135) ----
    const int m = item, a0 = A.mol_ptr[m], a1 = A.mol_ptr[m + 1];
    double com[3], t[9], ax[3][3], am[3], dummy[3][3];
    we_group_tensor(a1 - a0, [&](int k, int c) { return A.pos[3 * A.mol_atoms[a0 + k] + c]; },
                    [&](int k) { return A.mass[A.mol_atoms[a0 + k]]; }, com, t);
    we_axes_of(t, A.eig_flavour, diag, ax, am);
    // moments: inertia tensor of the united atoms about the same centre (transformations.py:138-163)
    double u[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = A.ua_ptr[m]; q < A.ua_ptr[m + 1]; ++q) {
      const int a = A.ua_atom[q];
      double mass = A.mass[a];
      for (int h = A.hin_ptr[q]; h < A.hin_ptr[q + 1]; ++h) mass = we_add(mass, A.mass[A.hin[h]]);
      const double dx = we_add((double)A.pos[3 * a], -com[0]), dy = we_add((double)A.pos[3 * a + 1], -com[1]),
                   dz = we_add((double)A.pos[3 * a + 2], -com[2]);
      const double xx = we_pow2(fabs(dx)), yy = we_pow2(fabs(dy)), zz = we_pow2(fabs(dz));
      u[0] = we_add(u[0], we_mul(we_add(yy, zz), mass));
      u[4] = we_add(u[4], we_mul(we_add(xx, zz), mass));
      u[8] = we_add(u[8], we_mul(we_add(xx, yy), mass));
      u[1] = we_add(u[1], -we_mul(we_mul(dx, dy), mass));
      u[2] = we_add(u[2], -we_mul(we_mul(dx, dz), mass));
      u[5] = we_add(u[5], -we_mul(we_mul(dy, dz), mass));
    }
    u[3] = u[1]; u[6] = u[2]; u[7] = u[5];
    we_axes_of(u, A.eig_flavour, diag, dummy, am);
    we_frame_force_torque(A, a0, a1, ax, am, false, nullptr, com, F, tq);
    o = out_mol + (size_t)m * 24;
  } else {
    // ---- one united atom (forces_torques.py:60-86, transformations.py:286-353) ----
    const int q = item - A.n_mol;
    int m = 0;
    while (A.ua_ptr[m + 1] <= q) ++m;
    const int a = A.ua_atom[q];
    const int* hv = A.hv + A.hv_ptr[q];
    const int* lt = A.lt + A.lt_ptr[q];
    const int nh = A.hv_ptr[q + 1] - A.hv_ptr[q], nl = A.lt_ptr[q + 1] - A.lt_ptr[q];
    o = out_ua + (size_t)q * 25;
    double ax[3][3];
    bool have = true;
    if (nh == 1 && nl == 1) we_custom_axes(A, a, hv, 1, -1, true, ax);          // :337-340 overrides :326-332
    else if (nh == 1 && nl >= 1) we_custom_axes(A, a, hv, 1, lt[0], false, ax);  // :326-332
    else if (nh == 2) we_custom_axes(A, a, hv, 1, hv[1], false, ax);             // :319-325
    else if (nh > 2) we_custom_axes(A, a, hv, nh, hv[1], false, ax);             // :333-336
    else if (nh == 0) {                                                         // :341-343
      double com[3], t[9], am[3];
      auto at = [&](int k) { return k == 0 ? a : lt[k - 1]; };
      we_group_tensor(1 + nl, [&](int k, int c) { return A.pos[3 * at(k) + c]; }, [&](int k) { return A.mass[at(k)]; }, com, t);
      we_axes_of(t, A.eig_flavour, diag, ax, am);
    } else have = false;  // bonded to one other united atom only (forces_torques.py:66)
    if (!have) {
      for (int i = 0; i < 25; ++i) o[i] = 0.0;
      return;
    }
    // get_custom_PI_MOI: column j over the norm of row j; no flip (the vector from the atom to itself is 0)
    double n2[3], PI[3][3], mi[3] = {0.0, 0.0, 0.0};
    for (int i = 0; i < 3; ++i) n2[i] = sqrt(we_add(we_add(we_mul(ax[i][0], ax[i][0]), we_mul(ax[i][1], ax[i][1])), we_mul(ax[i][2], ax[i][2])));
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) PI[i][j] = we_ddiv(ax[i][j], n2[j]);
    const float cf[3] = {A.pos[3 * a], A.pos[3 * a + 1], A.pos[3 * a + 2]};
    const int n_all = 1 + nh + nl;
    for (int k = 0; k < n_all; ++k) {
      const int b = k == 0 ? a : (k <= nh ? hv[k - 1] : lt[k - 1 - nh]);
      const double co[3] = {(double)we_fsub(A.pos[3 * b], cf[0]), (double)we_fsub(A.pos[3 * b + 1], cf[1]), (double)we_fsub(A.pos[3 * b + 2], cf[2])};
      const double mass = A.mass[b];
      for (int i = 0; i < 3; ++i) {
        double cr[3];
        we_cross(PI[i], co, cr);
        const double comp = we_add(we_add(we_mul(we_mul(cr[0], cr[0]), mass), we_mul(we_mul(cr[1], cr[1]), mass)), we_mul(we_mul(cr[2], cr[2]), mass));
        mi[i] = we_add(mi[i], comp);
      }
    }
    we_frame_force_torque(A, A.mol_ptr[m], A.mol_ptr[m + 1], PI, mi, true, cf, nullptr, F, tq);
    o[24] = 1.0;
  }
  for (int c = 0; c < 3; ++c) { o[c] = F[c]; o[3 + c] = tq[c]; }
  for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) {
    o[6 + 3 * c + d] = we_mul(F[c], F[d]);   // get_covariance_matrix(ft, 1): outer * 1 ** 2
    o[15 + 3 * c + d] = we_mul(tq[c], tq[d]);
  }
}

#ifdef __CUDACC__
// (the item function takes the argument block by reference: with the lambdas of we_group_tensor capturing the
// kernel PARAMETER itself by reference, the sm_100a build faulted with an illegal address)
__global__ void k_force_torque_multi(WeMultiUA A, WeDiag* diag, double* __restrict__ out_mol, double* __restrict__ out_ua) {
  const int item = blockIdx.x * blockDim.x + threadIdx.x;
  if (item >= A.n_mol + A.n_ua) return;
  we_multi_ua_item(A, item, diag, out_mol, out_ua);
}
#endif
