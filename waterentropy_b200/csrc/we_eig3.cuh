// Eigen-decomposition of a symmetric 3x3 moment-of-inertia tensor that follows LAPACK `dgeev`
// operation by operation (host + device).
//
// Why: the reference takes its principal axes from MDAnalysis `principal_axes`
// (= `numpy.linalg.eig` + descending sort + one global right-hand flip) and the eigenvalues from
// `LA.eig` (/root/reference/waterEntropy/maths/transformations.py:113-133).  For a symmetric
// input every returned eigenvector is +-(a Schur vector), and WHICH sign comes out is decided by
// the Householder reduction and by how many Francis sweeps `dlahqr` runs before it deflates - not
// by any geometric rule.  The off-diagonal elements of the force / torque covariance matrices
// (recipes/forces_torques.py:50-53) depend on those signs, so matching them means replaying the
// algorithm:  dgebal (no-op for a symmetric matrix) -> dgehd2 -> dorg2r -> dlahqr (+ dlanv2) ->
// dtrevc3 (blocked back-transform) -> normalise.
//
// The restatement follows LAPACK 3.11 as bundled with OpenBLAS 0.3.30 (numpy 2.3.5, the version
// the reference pins in pyproject.toml:25).  The LAPACK routines are plain IEEE double arithmetic in
// source order (the library's Fortran objects contain no FMA); the BLAS calls inside them go through
// OpenBLAS kernels whose rounding was pinned empirically against the library (tools/eig3_check.c):
//   dnrm2   x87: squares and sums in 64-bit-mantissa extended precision, fsqrt, one final
//           rounding to double                                    -> we_x87_nrm2 (integer soft-float)
//   dgemv N w_i = fma(a_i2, v_2, a_i1 v_1)         dgemv T  w_j = a_1j v_1 + a_2j v_2 (v_1 = 1)
//   dger    A_ij = fma(x_i, alpha y_j, A_ij)       daxpy    y_i = fma(a, x_i, y_i)
//           (these three fuse in the AVX-512 "SkylakeX" kernels only; the AVX2 "Haswell" kernels, which
//           AMD Zen hosts also run, multiply and add separately: `flavour` selects the family)
//   drot    x' = fma(c, x, s y), y' = fma(c, y, -(s x))
//   dgemm   c = fma(a_3, b_3, fma(a_2, b_2, a_1 b_1))             dscal  x_i = a x_i
// Compile with -fmad=false (nvcc) / -ffp-contract=off (gcc): only the explicit fma() calls fuse.
//
// Not replayed (status != 0, the caller falls back to its own diagonalisation): exact zeros that
// make dgebal permute, complex or nearly equal eigenvalue pairs in dlanv2, over/underflow scaling.
// None of them can occur for a water molecule (three distinct moments of order 1 amu A^2).
#pragma once
#include "we_numerics.cuh"

// OpenBLAS kernel families whose rounding is restated (dgemv N / dger / daxpy differ between them)
#define WE_EIG3_SKYLAKEX 0  // AVX-512 hosts (the build container, where the goldens were made)
#define WE_EIG3_HASWELL 1   // AVX2 hosts (Haswell .. , AMD Zen)

#define WE_EIG3_OK 0
#define WE_EIG3_UNSUPPORTED 1   // input pattern whose LAPACK path is not restated
#define WE_EIG3_NOCONV 2        // dlahqr did not converge (never observed)

// ---- x87 extended precision, positive numbers only: value = m * 2^(e - 63), m in [2^63, 2^64) ----
struct WeX87 { uint64_t m; int e; };

WE_HD void we_mul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#ifdef __CUDA_ARCH__
  hi = __umul64hi(a, b); lo = a * b;
#else
  const unsigned __int128 p = (unsigned __int128)a * b;
  hi = (uint64_t)(p >> 64); lo = (uint64_t)p;
#endif
}

// round the 128-bit magnitude (hi:lo), hi's bit 63 or bit 62 set, to a 64-bit mantissa (RNE);
// `sticky` = bits already lost below lo.  e_top = exponent if bit 127 is the leading bit.
WE_HD WeX87 we_x87_round128(uint64_t hi, uint64_t lo, bool sticky, int e_top) {
  WeX87 r;
  uint64_t m, rest;
  bool half;
  if (hi >> 63) { m = hi; half = (lo >> 63) != 0; rest = lo << 1; r.e = e_top; }
  else { m = (hi << 1) | (lo >> 63); half = ((lo >> 62) & 1) != 0; rest = lo << 2; r.e = e_top - 1; }
  const bool more = (rest != 0) || sticky;
  if (half && (more || (m & 1))) { ++m; if (m == 0) { m = 0x8000000000000000ULL; ++r.e; } }
  r.m = m;
  return r;
}

WE_HD WeX87 we_x87_from_abs(double x) {  // |x|, normal or zero
  const uint64_t u = we_d2u(x) & 0x7fffffffffffffffULL;
  WeX87 r;
  if (u == 0) { r.m = 0; r.e = 0; return r; }
  int ex = (int)(u >> 52);
  uint64_t mant = u & 0xfffffffffffffULL;
  if (ex == 0) {  // subnormal: normalise
    int sh = 0;
    while (!(mant >> 52)) { mant <<= 1; ++sh; }
    ex = 1 - sh;
  } else mant |= 1ULL << 52;
  r.m = mant << 11;
  r.e = ex - 1023;
  return r;
}

WE_HD WeX87 we_x87_sqr(WeX87 a) {
  WeX87 r;
  if (a.m == 0) { r.m = 0; r.e = 0; return r; }
  uint64_t hi, lo;
  we_mul64(a.m, a.m, hi, lo);
  return we_x87_round128(hi, lo, false, 2 * a.e + 1);
}

WE_HD WeX87 we_x87_add(WeX87 a, WeX87 b) {
  if (a.m == 0) return b;
  if (b.m == 0) return a;
  if (a.e < b.e) { const WeX87 t = a; a = b; b = t; }
  const int d = a.e - b.e;
  // work at half scale so the sum cannot carry out of 128 bits: A = a.m << 63, B = (b.m << 63) >> d
  uint64_t ahi = a.m >> 1, alo = a.m << 63;
  uint64_t bhi, blo;
  bool sticky = false;
  if (d == 0) { bhi = b.m >> 1; blo = b.m << 63; }
  else if (d < 64) {
    const uint64_t h = b.m >> 1, l = b.m << 63;
    bhi = h >> d; blo = (l >> d) | (h << (64 - d));
    sticky = (l << (64 - d)) != 0;
  } else if (d < 127) {
    const uint64_t h = b.m >> 1, l = b.m << 63;
    const int s = d - 64;
    bhi = 0; blo = s ? (h >> s) : h;
    sticky = (l != 0) || (s ? ((h << (64 - s)) != 0) : false);
  } else { bhi = 0; blo = 0; sticky = true; }
  const uint64_t slo = alo + blo;
  const uint64_t shi = ahi + bhi + (slo < alo ? 1ULL : 0ULL);
  // A has bit 126 set; the sum has bit 126 or bit 127 set.  Leading bit 127 <=> exponent a.e + 1.
  return we_x87_round128(shi, slo, sticky, a.e + 1);
}

WE_HD WeX87 we_x87_sqrt(WeX87 a) {
  WeX87 r;
  if (a.m == 0) { r.m = 0; r.e = 0; return r; }
  // N = a.m << 64 (a.e odd) or a.m << 63 (a.e even): the exponent left over is even
  uint64_t nhi, nlo;
  if (a.e & 1) { nhi = a.m; nlo = 0; r.e = (a.e - 63) / 2 + 31; }
  else { nhi = a.m >> 1; nlo = a.m << 63; r.e = a.e / 2; }
  // root = floor(sqrt(N)): double estimate, one Newton correction, exact fix-up
  uint64_t root;
  {
    const double est = sqrt((double)nhi) * 4294967296.0;  // sqrt(nhi * 2^64)
    root = (est >= 18446744073709549568.0) ? 0xfffffffffffff800ULL : (uint64_t)est;
    uint64_t qhi, qlo;
    we_mul64(root, root, qhi, qlo);
    // diff = N - root^2 (signed, |diff| < 2^80)
    const uint64_t dlo = nlo - qlo;
    const uint64_t dhi = nhi - qhi - (nlo < qlo ? 1ULL : 0ULL);
    const double diff = (double)(int64_t)dhi * 18446744073709551616.0 + (double)dlo;
    const double corr = diff / (2.0 * (double)root);
    root = (uint64_t)((int64_t)root + (int64_t)corr);
  }
  for (int it = 0; it < 64; ++it) {  // exact: root^2 <= N < (root + 1)^2
    uint64_t qhi, qlo;
    we_mul64(root, root, qhi, qlo);
    if (qhi > nhi || (qhi == nhi && qlo > nlo)) { --root; continue; }
    // rem = N - root^2 ; next square = root^2 + 2 root + 1 <= N  <=>  rem >= 2 root + 1
    const uint64_t rlo = nlo - qlo;
    const uint64_t rhi = nhi - qhi - (nlo < qlo ? 1ULL : 0ULL);
    const uint64_t t_hi = root >> 63, t_lo = (root << 1) | 1ULL;  // 2 root + 1 (129-bit safe: root < 2^64)
    if (rhi > t_hi || (rhi == t_hi && rlo >= t_lo)) { ++root; continue; }
    // round to nearest: up iff N > (root + 1/2)^2  <=>  rem > root
    if (rhi != 0 || rlo > root) { ++root; if (root == 0) { root = 0x8000000000000000ULL; ++r.e; } }
    break;
  }
  r.m = root;
  return r;
}

WE_HD double we_x87_to_double(WeX87 a) {  // fstp qword: round the 64-bit mantissa to 53 bits (RNE)
  if (a.m == 0) return 0.0;
  uint64_t m53 = a.m >> 11;
  const uint64_t rem = a.m & 0x7ffULL;
  int e = a.e;
  if (rem > 0x400ULL || (rem == 0x400ULL && (m53 & 1))) { ++m53; if (m53 >> 53) { m53 >>= 1; ++e; } }
  const int be = e + 1023;
  if (be <= 0 || be >= 2047) return (be <= 0) ? 0.0 : we_u2d(0x7ff0000000000000ULL);
  return we_u2d(((uint64_t)be << 52) | (m53 & 0xfffffffffffffULL));
}

// OpenBLAS x86_64 dnrm2 (kernel/x86_64/nrm2.S) for n < 8: acc = 0; acc += x_i * x_i in extended
// precision, fsqrt, store as double.
WE_HD double we_x87_nrm2(const double* x, int n) {
  WeX87 acc; acc.m = 0; acc.e = 0;
  for (int i = 0; i < n; ++i) acc = we_x87_add(acc, we_x87_sqr(we_x87_from_abs(x[i])));
  return we_x87_to_double(we_x87_sqrt(acc));
}

WE_HD double we_sub(double a, double b) { return we_add(a, -b); }
WE_HD double we_copysign(double mag, double sgn) {
  return we_u2d((we_d2u(mag) & 0x7fffffffffffffffULL) | (we_d2u(sgn) & 0x8000000000000000ULL));
}
#define WE_LAPACK_SAFMIN 2.2250738585072014e-308  // dlamch('S') = 2^-1022
#define WE_LAPACK_ULP 2.220446049250313e-16       // dlamch('P') = 2^-52
#define WE_LAPACK_EPS 1.1102230246251565e-16      // dlamch('E') = 2^-53

// DLAPY2 (LAPACK 3.11): sqrt(x^2 + y^2) without unnecessary overflow
WE_HD double we_dlapy2(double x, double y) {
  const double xa = fabs(x), ya = fabs(y);
  const double w = xa > ya ? xa : ya, z = xa > ya ? ya : xa;
  if (z == 0.0) return w;
  const double q = we_ddiv(z, w);
  return we_mul(w, we_dsqrt(we_add(1.0, we_mul(q, q))));
}

// DLARFG: elementary reflector; x has n - 1 entries (unit stride)
WE_HD void we_dlarfg(int n, double& alpha, double* x, double& tau, int& status) {
  if (n <= 1) { tau = 0.0; return; }
  const double xnorm = we_x87_nrm2(x, n - 1);
  if (xnorm == 0.0) { tau = 0.0; return; }
  const double beta = -we_copysign(we_dlapy2(alpha, xnorm), alpha);
  if (fabs(beta) < WE_LAPACK_SAFMIN / WE_LAPACK_EPS) { status = WE_EIG3_UNSUPPORTED; }
  tau = we_ddiv(we_sub(beta, alpha), beta);
  const double sc = we_ddiv(1.0, we_sub(alpha, beta));
  for (int i = 0; i < n - 1; ++i) x[i] = we_mul(sc, x[i]);
  alpha = beta;
}

// DLANV2: Schur factorisation of a real 2x2 block; only the real-eigenvalue branch is restated
WE_HD void we_dlanv2(double& a, double& b, double& c, double& d, double& rt1r, double& rt2r,
                     double& cs, double& sn, int& status) {
  if (c == 0.0) { cs = 1.0; sn = 0.0; }
  else if (b == 0.0) {
    cs = 0.0; sn = 1.0;
    const double t = d; d = a; a = t; b = -c; c = 0.0;
  } else if (we_sub(a, d) == 0.0 && (we_copysign(1.0, b) != we_copysign(1.0, c))) { cs = 1.0; sn = 0.0; }
  else {
    const double temp = we_sub(a, d);
    double p = we_mul(0.5, temp);
    const double ab = fabs(b), ac = fabs(c);
    const double bcmax = ab > ac ? ab : ac;
    const double bcmis = we_mul(we_mul(ab > ac ? ac : ab, we_copysign(1.0, b)), we_copysign(1.0, c));
    const double scale = fabs(p) > bcmax ? fabs(p) : bcmax;
    double z = we_add(we_mul(we_ddiv(p, scale), p), we_mul(we_ddiv(bcmax, scale), bcmis));
    if (z >= 4.0 * WE_LAPACK_ULP) {
      z = we_add(p, we_copysign(we_mul(we_dsqrt(scale), we_dsqrt(z)), p));
      a = we_add(d, z);
      d = we_sub(d, we_mul(we_ddiv(bcmax, z), bcmis));
      const double tau = we_dlapy2(c, z);
      cs = we_ddiv(z, tau);
      sn = we_ddiv(c, tau);
      b = we_sub(b, c);
      c = 0.0;
    } else {
      status = WE_EIG3_UNSUPPORTED;  // complex or (almost) equal eigenvalues
      cs = 1.0; sn = 0.0;
    }
  }
  rt1r = a; rt2r = d;
  if (c != 0.0) status = WE_EIG3_UNSUPPORTED;
}

// y + a x as the host's OpenBLAS level-1/2 kernels round it: one fused multiply-add in the AVX-512
// (SkylakeX) kernels, multiply then add in the AVX2 (Haswell / Zen) kernels
WE_HD double we_axpy1(int flavour, double a, double x, double y) {
  return flavour == WE_EIG3_SKYLAKEX ? we_fma(a, x, y) : we_add(y, we_mul(a, x));
}

WE_HD void we_drot1(double& x, double& y, double c, double s) {
  const double xn = we_fma(c, x, we_mul(s, y));
  const double yn = we_fma(c, y, -we_mul(s, x));
  x = xn; y = yn;
}

// numpy.linalg.eig of the symmetric 3x3 matrix `t` (row-major; only exact symmetry is assumed).
// wr[k], vr[3 * k + i] = component i of the k-th returned eigenvector (LAPACK's column k).
WE_HD int we_dgeev3(const double* t, double* wr, double* vr, int flavour) {
  int status = WE_EIG3_OK;
  double a[9], z[9];
#define A_(i, j) a[((i) - 1) + 3 * ((j) - 1)]
#define Z_(i, j) z[((i) - 1) + 3 * ((j) - 1)]
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) a[i + 3 * j] = t[3 * i + j];
  // ---- dgebal('B'): a symmetric matrix is never scaled (row norm == column norm).  It is permuted
  // when a row / column has only zeros off the diagonal: not restated.
  if ((A_(1, 2) == 0.0 && A_(1, 3) == 0.0) || (A_(1, 2) == 0.0 && A_(2, 3) == 0.0) ||
      (A_(1, 3) == 0.0 && A_(2, 3) == 0.0))
    status = WE_EIG3_UNSUPPORTED;
  // ---- dgehrd -> dgehd2, i = 1 (i = 2 has a reflector of order 1: tau = 0) -----------------
  double tau1, v2;
  {
    double alpha = A_(2, 1);
    v2 = A_(3, 1);
    we_dlarfg(2, alpha, &v2, tau1, status);
    A_(2, 1) = alpha;
    if (tau1 != 0.0) {
      const double mt = -tau1;
      // dlarf('Right', 3, 2, v, tau, A(1,2)): w = C v (dgemv N), C -= tau w v^T (dger)
      int lastc = 3;
      while (lastc > 0 && A_(lastc, 2) == 0.0 && A_(lastc, 3) == 0.0) --lastc;
      double w[3];
      for (int i = 1; i <= lastc; ++i) w[i - 1] = we_axpy1(flavour, A_(i, 3), v2, A_(i, 2));
      {
        const double t1 = mt, t2 = we_mul(mt, v2);  // alpha * y_j with y = (1, v2)
        for (int i = 1; i <= lastc; ++i) A_(i, 2) = we_axpy1(flavour, w[i - 1], t1, A_(i, 2));
        for (int i = 1; i <= lastc; ++i) A_(i, 3) = we_axpy1(flavour, w[i - 1], t2, A_(i, 3));
      }
      // dlarf('Left', 2, 2, v, tau, A(2,2)): w = C^T v (dgemv T), C -= tau v w^T (dger)
      lastc = 2;
      while (lastc > 0 && A_(2, 1 + lastc) == 0.0 && A_(3, 1 + lastc) == 0.0) --lastc;
      for (int j = 1; j <= lastc; ++j) w[j - 1] = we_add(A_(2, 1 + j), we_mul(A_(3, 1 + j), v2));
      for (int j = 1; j <= lastc; ++j) {
        const double tj = we_mul(mt, w[j - 1]);
        A_(2, 1 + j) = we_axpy1(flavour, 1.0, tj, A_(2, 1 + j));
        A_(3, 1 + j) = we_axpy1(flavour, v2, tj, A_(3, 1 + j));
      }
    }
  }
  // ---- dorghr -> dorg2r(2, 2, 2): Q = H(1) ---------------------------------------------------
  for (int i = 0; i < 9; ++i) z[i] = 0.0;
  Z_(1, 1) = 1.0; Z_(2, 2) = 1.0; Z_(3, 3) = 1.0;
  if (tau1 != 0.0) {
    const double tq = we_mul(-tau1, v2);          // w = v2 (dgemv T of the column (0, 1))
    Z_(2, 3) = tq;                                // fma(1, tq, 0)
    Z_(3, 3) = we_axpy1(flavour, v2, tq, 1.0);
    Z_(3, 2) = we_mul(-tau1, v2);                 // dscal
    Z_(2, 2) = we_sub(1.0, tau1);
  }
  // ---- dhseqr -> dlahqr(wantt, wantz, n = 3, ilo = 1, ihi = 3) -------------------------------
  A_(3, 1) = 0.0;
  {
    const double safmin = WE_LAPACK_SAFMIN, ulp = WE_LAPACK_ULP;
    const double smlnum = we_mul(safmin, we_ddiv(3.0, ulp));
    const int i1 = 1, i2 = 3;
    int kdefl = 0;
    int i = 3;
    bool failed = false;
    while (i >= 1 && !failed) {
      int l = 1;
      bool split = false;
      for (int its = 0; its <= 300; ++its) {
        int k;
        for (k = i; k > l; --k) {
          const double hk = fabs(A_(k, k - 1));
          if (hk <= smlnum) break;
          double tst = we_add(fabs(A_(k - 1, k - 1)), fabs(A_(k, k)));
          if (tst == 0.0) {
            if (k - 2 >= 1) tst = we_add(tst, fabs(A_(k - 1, k - 2)));
            if (k + 1 <= 3) tst = we_add(tst, fabs(A_(k + 1, k)));
          }
          if (hk <= we_mul(ulp, tst)) {
            const double hkk1 = fabs(A_(k - 1, k));
            const double ab = hk > hkk1 ? hk : hkk1, ba = hk > hkk1 ? hkk1 : hk;
            const double dkk = fabs(A_(k, k)), ddf = fabs(we_sub(A_(k - 1, k - 1), A_(k, k)));
            const double aa = dkk > ddf ? dkk : ddf, bb = dkk > ddf ? ddf : dkk;
            const double s = we_add(aa, ab);
            const double lhs = we_mul(ba, we_ddiv(ab, s));
            const double r0 = we_mul(ulp, we_mul(bb, we_ddiv(aa, s)));
            if (lhs <= (smlnum > r0 ? smlnum : r0)) break;
          }
        }
        l = k;
        if (l > 1) A_(l, l - 1) = 0.0;
        if (l >= i - 1) { split = true; break; }
        ++kdefl;
        double h11, h21, h12, h22;
        if (kdefl % 20 == 0) {
          const double s = we_add(fabs(A_(i, i - 1)), fabs(A_(i - 1, i - 2)));
          h11 = we_add(we_mul(0.75, s), A_(i, i)); h12 = we_mul(-0.4375, s); h21 = s; h22 = h11;
        } else if (kdefl % 10 == 0) {
          const double s = we_add(fabs(A_(l + 1, l)), fabs(A_(l + 2, l + 1)));
          h11 = we_add(we_mul(0.75, s), A_(l, l)); h12 = we_mul(-0.4375, s); h21 = s; h22 = h11;
        } else { h11 = A_(i - 1, i - 1); h21 = A_(i, i - 1); h12 = A_(i - 1, i); h22 = A_(i, i); }
        double rt1r, rt1i, rt2r, rt2i;
        {
          const double s = we_add(we_add(we_add(fabs(h11), fabs(h12)), fabs(h21)), fabs(h22));
          if (s == 0.0) { rt1r = rt1i = rt2r = rt2i = 0.0; }
          else {
            h11 = we_ddiv(h11, s); h21 = we_ddiv(h21, s); h12 = we_ddiv(h12, s); h22 = we_ddiv(h22, s);
            const double tr = we_ddiv(we_add(h11, h22), 2.0);
            const double det = we_sub(we_mul(we_sub(h11, tr), we_sub(h22, tr)), we_mul(h12, h21));
            const double rtdisc = we_dsqrt(fabs(det));
            if (det >= 0.0) {
              rt1r = we_mul(tr, s); rt2r = rt1r; rt1i = we_mul(rtdisc, s); rt2i = -rt1i;
            } else {
              rt1r = we_add(tr, rtdisc); rt2r = we_sub(tr, rtdisc);
              if (fabs(we_sub(rt1r, h22)) <= fabs(we_sub(rt2r, h22))) { rt1r = we_mul(rt1r, s); rt2r = rt1r; }
              else { rt2r = we_mul(rt2r, s); rt1r = rt2r; }
              rt1i = 0.0; rt2i = 0.0;
            }
          }
        }
        // look for two consecutive small subdiagonal elements; with l = 1, i = 3: m = 1 only
        double v[3];
        int m;
        for (m = i - 2; m >= l; --m) {
          double h21s = fabs(A_(m + 1, m));
          double s = we_add(we_add(fabs(we_sub(A_(m, m), rt2r)), fabs(rt2i)), h21s);
          h21s = we_ddiv(A_(m + 1, m), s);
          v[0] = we_sub(we_add(we_mul(h21s, A_(m, m + 1)),
                               we_mul(we_sub(A_(m, m), rt1r), we_ddiv(we_sub(A_(m, m), rt2r), s))),
                        we_mul(rt1i, we_ddiv(rt2i, s)));
          v[1] = we_mul(h21s, we_sub(we_sub(we_add(A_(m, m), A_(m + 1, m + 1)), rt1r), rt2r));
          v[2] = we_mul(h21s, A_(m + 2, m + 1));
          s = we_add(we_add(fabs(v[0]), fabs(v[1])), fabs(v[2]));
          v[0] = we_ddiv(v[0], s); v[1] = we_ddiv(v[1], s); v[2] = we_ddiv(v[2], s);
          if (m == l) break;
          const double h00 = fabs(A_(m - 1, m - 1)), h10 = fabs(A_(m, m - 1)), h11a = fabs(A_(m, m));
          if (we_mul(h10, we_add(fabs(v[1]), fabs(v[2]))) <=
              we_mul(we_mul(ulp, fabs(v[0])), we_add(we_add(h00, h11a), fabs(h22))))
            break;
        }
        if (m < l) m = l;
        // double-shift QR step
        for (int k2 = m; k2 <= i - 1; ++k2) {
          const int nr = (3 < i - k2 + 1) ? 3 : (i - k2 + 1);
          if (k2 > m) for (int q = 0; q < nr; ++q) v[q] = A_(k2 + q, k2 - 1);
          double t1;
          we_dlarfg(nr, v[0], v + 1, t1, status);
          if (k2 > m) {
            A_(k2, k2 - 1) = v[0];
            A_(k2 + 1, k2 - 1) = 0.0;
            if (k2 < i - 1) A_(k2 + 2, k2 - 1) = 0.0;
          } else if (m > l) {
            A_(k2, k2 - 1) = we_mul(A_(k2, k2 - 1), we_sub(1.0, t1));
          }
          const double vv2 = v[1], t2 = we_mul(t1, vv2);
          if (nr == 3) {
            const double vv3 = v[2], t3 = we_mul(t1, vv3);
            for (int j = k2; j <= i2; ++j) {
              const double sum = we_add(we_add(A_(k2, j), we_mul(vv2, A_(k2 + 1, j))), we_mul(vv3, A_(k2 + 2, j)));
              A_(k2, j) = we_sub(A_(k2, j), we_mul(sum, t1));
              A_(k2 + 1, j) = we_sub(A_(k2 + 1, j), we_mul(sum, t2));
              A_(k2 + 2, j) = we_sub(A_(k2 + 2, j), we_mul(sum, t3));
            }
            const int jmax = (k2 + 3 < i) ? k2 + 3 : i;
            for (int j = i1; j <= jmax; ++j) {
              const double sum = we_add(we_add(A_(j, k2), we_mul(vv2, A_(j, k2 + 1))), we_mul(vv3, A_(j, k2 + 2)));
              A_(j, k2) = we_sub(A_(j, k2), we_mul(sum, t1));
              A_(j, k2 + 1) = we_sub(A_(j, k2 + 1), we_mul(sum, t2));
              A_(j, k2 + 2) = we_sub(A_(j, k2 + 2), we_mul(sum, t3));
            }
            for (int j = 1; j <= 3; ++j) {
              const double sum = we_add(we_add(Z_(j, k2), we_mul(vv2, Z_(j, k2 + 1))), we_mul(vv3, Z_(j, k2 + 2)));
              Z_(j, k2) = we_sub(Z_(j, k2), we_mul(sum, t1));
              Z_(j, k2 + 1) = we_sub(Z_(j, k2 + 1), we_mul(sum, t2));
              Z_(j, k2 + 2) = we_sub(Z_(j, k2 + 2), we_mul(sum, t3));
            }
          } else if (nr == 2) {
            for (int j = k2; j <= i2; ++j) {
              const double sum = we_add(A_(k2, j), we_mul(vv2, A_(k2 + 1, j)));
              A_(k2, j) = we_sub(A_(k2, j), we_mul(sum, t1));
              A_(k2 + 1, j) = we_sub(A_(k2 + 1, j), we_mul(sum, t2));
            }
            for (int j = i1; j <= i; ++j) {
              const double sum = we_add(A_(j, k2), we_mul(vv2, A_(j, k2 + 1)));
              A_(j, k2) = we_sub(A_(j, k2), we_mul(sum, t1));
              A_(j, k2 + 1) = we_sub(A_(j, k2 + 1), we_mul(sum, t2));
            }
            for (int j = 1; j <= 3; ++j) {
              const double sum = we_add(Z_(j, k2), we_mul(vv2, Z_(j, k2 + 1)));
              Z_(j, k2) = we_sub(Z_(j, k2), we_mul(sum, t1));
              Z_(j, k2 + 1) = we_sub(Z_(j, k2 + 1), we_mul(sum, t2));
            }
          }
        }
      }
      if (!split) { failed = true; status = WE_EIG3_NOCONV; break; }
      if (l == i) wr[i - 1] = A_(i, i);
      else {  // l == i - 1: a pair has converged; standardise the 2x2 block
        double cs, sn;
        we_dlanv2(A_(i - 1, i - 1), A_(i - 1, i), A_(i, i - 1), A_(i, i), wr[i - 2], wr[i - 1], cs, sn, status);
        if (i2 > i) for (int j = i + 1; j <= i2; ++j) we_drot1(A_(i - 1, j), A_(i, j), cs, sn);
        for (int j = i1; j <= i - 2; ++j) we_drot1(A_(j, i - 1), A_(j, i), cs, sn);
        for (int j = 1; j <= 3; ++j) we_drot1(Z_(j, i - 1), Z_(j, i), cs, sn);
      }
      kdefl = 0;
      i = l - 1;
    }
  }
  A_(3, 1) = 0.0;  // dhseqr: dlaset('L', n - 2, n - 2, 0, 0, h(3,1))
  if (status == WE_EIG3_NOCONV) { for (int q = 0; q < 9; ++q) vr[q] = z[q]; return status; }
  // ---- dtrevc3('R', 'B'): eigenvectors of T, back-transformed with one dgemm -------------------
  {
    const double ulp = WE_LAPACK_ULP;
    const double smlnum = we_mul(WE_LAPACK_SAFMIN, we_ddiv(3.0, ulp));
    const double bignum = we_ddiv(we_sub(1.0, ulp), smlnum);
    const double cn[3] = {0.0, fabs(A_(1, 2)), we_add(fabs(A_(1, 3)), fabs(A_(2, 3)))};
    double bm[9];  // column ki - 1 = eigenvector ki of T
    for (int ki = 3; ki >= 1; --ki) {
      double* x = bm + 3 * (ki - 1);
      const double w = A_(ki, ki);
      const double smin = (we_mul(ulp, fabs(w)) > smlnum) ? we_mul(ulp, fabs(w)) : smlnum;
      x[0] = x[1] = x[2] = 0.0;
      x[ki - 1] = 1.0;
      for (int k = 1; k < ki; ++k) x[k - 1] = -A_(k, ki);
      for (int j = ki - 1; j >= 1; --j) {
        // dlaln2, 1x1 real: (T(j,j) - w) X = scale * x(j)
        const double smlnum2 = 2.0 * WE_LAPACK_SAFMIN, bignum2 = we_ddiv(1.0, smlnum2);
        const double smini = smin > smlnum2 ? smin : smlnum2;
        double csr = we_sub(A_(j, j), w);
        double cnorm = fabs(csr);
        if (cnorm < smini) { csr = smini; cnorm = smini; }
        const double bnorm = fabs(x[j - 1]);
        double scale = 1.0;
        if (cnorm < 1.0 && bnorm > 1.0) { if (bnorm > we_mul(bignum2, cnorm)) scale = we_ddiv(1.0, bnorm); }
        double X = we_ddiv(we_mul(x[j - 1], scale), csr);
        const double xnorm = fabs(X);
        if (xnorm > 1.0) {
          if (cn[j - 1] > we_ddiv(bignum, xnorm)) { X = we_ddiv(X, xnorm); scale = we_ddiv(scale, xnorm); }
        }
        if (scale != 1.0) for (int k = 0; k < ki; ++k) x[k] = we_mul(scale, x[k]);
        x[j - 1] = X;
        for (int k = 1; k < j; ++k) x[k - 1] = we_axpy1(flavour, -X, A_(k, j), x[k - 1]);  // daxpy
      }
    }
    for (int c = 0; c < 3; ++c) {
      double col[3];
      for (int i = 0; i < 3; ++i)
        col[i] = we_fma(z[i + 6], bm[3 * c + 2], we_fma(z[i + 3], bm[3 * c + 1], we_mul(z[i], bm[3 * c])));
      int ii = 0;
      if (fabs(col[1]) > fabs(col[ii])) ii = 1;
      if (fabs(col[2]) > fabs(col[ii])) ii = 2;
      const double remax = we_ddiv(1.0, fabs(col[ii]));
      for (int i = 0; i < 3; ++i) col[i] = we_mul(remax, col[i]);
      // ---- dgeev: normalise to unit Euclidean norm --------------------------------------------
      const double scl = we_ddiv(1.0, we_x87_nrm2(col, 3));
      for (int i = 0; i < 3; ++i) vr[3 * c + i] = we_mul(scl, col[i]);
    }
  }
#undef A_
#undef Z_
  return status;
}

// MDAnalysis `principal_axes` on top of we_dgeev3: rows = eigenvectors by descending
// eigenvalue (numpy argsort of three values = insertion sort, stable), all three negated when the
// triple is left-handed.  moi = eigenvalues by descending magnitude (transformations.py:131-133).
WE_HD int we_principal_axes(const double* tensor, double ax[3][3], double moi[3], int flavour) {
  double wr[3], vr[9];
  const int st = we_dgeev3(tensor, wr, vr, flavour);
  int ord[3] = {0, 1, 2};  // ascending, stable insertion sort
  for (int i = 1; i < 3; ++i) {
    const int o = ord[i];
    int j = i - 1;
    while (j >= 0 && wr[ord[j]] > wr[o]) { ord[j + 1] = ord[j]; --j; }
    ord[j + 1] = o;
  }
  for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) ax[i][k] = vr[3 * ord[2 - i] + k];
  const double cx = we_sub(we_mul(ax[0][1], ax[1][2]), we_mul(ax[0][2], ax[1][1]));
  const double cy = we_sub(we_mul(ax[0][2], ax[1][0]), we_mul(ax[0][0], ax[1][2]));
  const double cz = we_sub(we_mul(ax[0][0], ax[1][1]), we_mul(ax[0][1], ax[1][0]));
  const double hand = we_add(we_add(we_mul(cx, ax[2][0]), we_mul(cy, ax[2][1])), we_mul(cz, ax[2][2]));
  if (hand < 0.0) for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) ax[i][k] = -ax[i][k];
  int oa[3] = {0, 1, 2};  // ascending |wr|, stable
  for (int i = 1; i < 3; ++i) {
    const int o = oa[i];
    int j = i - 1;
    while (j >= 0 && fabs(wr[oa[j]]) > fabs(wr[o])) { oa[j + 1] = oa[j]; --j; }
    oa[j + 1] = o;
  }
  for (int i = 0; i < 3; ++i) moi[i] = wr[oa[2 - i]];
  return st;
}
