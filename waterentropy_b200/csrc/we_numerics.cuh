// Bit-exact scalar arithmetic of the waterEntropy hot path (host + device).
//
// The reference computes RAD / HB angles in *float32* through NumPy
// (/root/reference/waterEntropy/maths/trig.py:103-122) and NumPy's scalar
// `x**2` is glibc `powf(x, 2.0f)`, which is NOT always the correctly rounded
// x*x (it differs in ~0.08 % of inputs).  Shell membership must match the
// reference bit-for-bit, so the device replays that sequence exactly:
//   * we_powf2()      = glibc 2.39 powf(x, 2.0f)  (log2/exp2 tables of
//                       sysdeps/ieee754/flt-32/e_powf.c; checked against libm
//                       for all 1.4e8 floats in [0.01, 1000], FMA or not)
//   * we_angle_cos*() = maths/trig.py:112-121 in float32, no FMA contraction
//   * we_dist_ortho() = MDAnalysis calc_distances.h `_calc_distance_array_ortho`
//                       (float32 difference widened to double, round()-based
//                       minimum image, fp64 sqrt)
//   * we_dist_tric()  = wrap-into-cell + 27-image search (parity unpinned)
// Everything here must be compiled with -fmad=false (nvcc) /
// -ffp-contract=off (gcc): explicit intrinsics are used on the device where
// it matters, the flag protects the rest.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#ifndef __CUDACC__
#define WE_HD inline
#define WE_HOST_ONLY 1
#else
#define WE_HD __host__ __device__ __forceinline__
#endif

// ---- glibc powf tables -----------------------------------------------------
#define WE_POWF_LOG2_TAB \
  {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2}, \
  {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2}, {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2}, \
  {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3}, \
  {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4}, \
  {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1.0000000000000p+0, 0x0.0p+0}, \
  {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4}, {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3}, \
  {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3}, {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2}, \
  {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2}, {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}

#define WE_EXP2F_TAB \
  0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, \
  0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, \
  0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL, \
  0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL, \
  0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL, \
  0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, \
  0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, \
  0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL

#ifdef __CUDACC__
__device__ __constant__ double we_d_log2tab[16][2] = {WE_POWF_LOG2_TAB};
__device__ __constant__ unsigned long long we_d_exp2tab[32] = {WE_EXP2F_TAB};
#endif
static const double we_h_log2tab[16][2] = {WE_POWF_LOG2_TAB};
static const unsigned long long we_h_exp2tab[32] = {WE_EXP2F_TAB};

struct WePowTables {  // pointers to whichever copy is addressable (shared memory on device)
  const double* log2tab;              // [32] (invc, logc) pairs
  const unsigned long long* exp2tab;  // [32]
};

WE_HD uint32_t we_f2u(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
WE_HD float we_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
WE_HD uint64_t we_d2u(double f) {
#ifdef __CUDA_ARCH__
  return (uint64_t)__double_as_longlong(f);
#else
  uint64_t u; memcpy(&u, &f, 8); return u;
#endif
}
WE_HD double we_u2d(uint64_t u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double f; memcpy(&f, &u, 8); return f;
#endif
}

// mul / add / mul-add without contraction
WE_HD double we_mul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
WE_HD double we_add(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
WE_HD double we_mad(double a, double b, double c) { return we_add(we_mul(a, b), c); }
WE_HD float we_fmul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
WE_HD float we_fadd(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
WE_HD float we_fsub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
WE_HD float we_fdiv(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
WE_HD float we_fsqrt(float a) {
#ifdef __CUDA_ARCH__
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
WE_HD double we_dsqrt(double a) {
#ifdef __CUDA_ARCH__
  return __dsqrt_rn(a);
#else
  return sqrt(a);
#endif
}
WE_HD double we_ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}

// glibc powf(x, 2.0f) for x >= 0 (or NaN/inf), following e_powf.c:
//   log2_inline(x) with the 16-entry table, ylogx = 2*log2(x), exp2_inline().
WE_HD float we_powf2(float x, const WePowTables& T) {
  uint32_t ix = we_f2u(x);
  if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
    // zero, subnormal, inf, nan or negative
    if ((ix << 1) == 0) return 0.0f;                      // +-0 ** 2 = +0
    if (ix == 0x7f800000u || ix == 0xff800000u) return we_u2f(0x7f800000u);  // inf
    if ((ix & 0x7fffffffu) > 0x7f800000u) return x + x;   // nan
    if (ix & 0x80000000u) { ix &= 0x7fffffffu; }          // even power: sign drops
    if (ix < 0x00800000u) {                               // subnormal: normalise
      ix = we_f2u(we_u2f(ix) * 0x1p23f);
      ix &= 0x7fffffffu;
      ix -= 23u << 23;
    }
  }
  uint32_t tmp = ix - 0x3f330000u;
  int i = (int)((tmp >> 19) & 15u);
  uint32_t top = tmp & 0xff800000u;
  uint32_t iz = ix - top;
  int k = ((int32_t)top) >> 23;
  double invc = T.log2tab[2 * i], logc = T.log2tab[2 * i + 1];
  double z = (double)we_u2f(iz);
  double r = we_mad(z, invc, -1.0);
  double y0 = we_add(logc, (double)k);
  const double A0 = 0x1.27616c9496e0bp-2, A1 = -0x1.71969a075c67ap-2, A2 = 0x1.ec70a6ca7baddp-2,
               A3 = -0x1.7154748bef6c8p-1, A4 = 0x1.71547652ab82bp+0;
  double r2 = we_mul(r, r);
  double y = we_mad(A0, r, A1);
  double p = we_mad(A2, r, A3);
  double r4 = we_mul(r2, r2);
  double q = we_mad(A4, r, y0);
  q = we_mad(p, r2, q);
  y = we_mad(y, r4, q);
  double xd = we_mul(2.0, y);  // ylogx
  // overflow / underflow of the float result (e_powf.c: |ylogx| >= 126)
  if (xd > 0x1.fffffffd1d571p+6) return we_u2f(0x7f800000u);
  if (xd <= -150.0) return 0.0f;
  const double SHIFT = 0x1.8p+47;  // 0x1.8p52 / 32
  double kd = we_add(xd, SHIFT);
  uint64_t ki = we_d2u(kd);
  kd = we_add(kd, -SHIFT);
  double rr = we_add(xd, -kd);
  uint64_t t = T.exp2tab[ki & 31u];
  t += ki << (52 - 5);
  double s = we_u2d(t);
  const double C0 = 0x1.c6af84b912394p-5, C1 = 0x1.ebfce50fac4f3p-3, C2 = 0x1.62e42ff0c52d6p-1;
  double zz = we_mad(C0, rr, C1);
  double rr2 = we_mul(rr, rr);
  double yy = we_mad(C2, rr, 1.0);
  yy = we_mad(zz, rr2, yy);
  yy = we_mul(yy, s);
  return (float)yy;
}

// Fast path of we_powf2.  Enumerating EVERY float in [2^-60, 2^60] (1,006,632,961
// values, tests/exhaustive_powf2.cpp) shows glibc powf(x, 2) differs from the
// correctly rounded x*x only when the 29 mantissa bits of the exact double product
// that float rounding discards lie within 888,704 (< 2^20) of the rounding
// midpoint 2^28.  Outside that band RN(x*x) IS powf(x, 2); inside it (0.8 % of
// inputs) the full table emulation runs.
WE_HD float we_powf2_fast(float x, const WePowTables& T) {
  const uint32_t ix = we_f2u(x);
  if (ix - 0x21800000u < 0x5d800000u - 0x21800000u) {  // 2^-60 <= x < 2^60
    const double p = (double)x * (double)x;            // exact in double
    const uint32_t low = (uint32_t)we_d2u(p) & 0x1fffffffu;
    const uint32_t d = low > 0x10000000u ? low - 0x10000000u : 0x10000000u - low;
    if (d >= (1u << 20)) return (float)p;
  }
  return we_powf2(x, T);
}

// |a-b| with the reference's "box lengths only" wrap, float32
WE_HD float we_wrap_abs(float a, float b, float dim) {
  float v = fabsf(we_fsub(a, b));
  if (v > we_fmul(0.5f, dim)) v = we_fsub(v, dim);
  return v;
}

// float32 separation used by get_angle: sqrtf((x^2 + y^2) + z^2)
WE_HD float we_angle_sep(float ax, float ay, float az, float bx, float by, float bz,
                         const float* dim) {
  float x = we_wrap_abs(ax, bx, dim[0]);
  float y = we_wrap_abs(ay, by, dim[1]);
  float z = we_wrap_abs(az, bz, dim[2]);
  return we_fsqrt(we_fadd(we_fadd(we_fmul(x, x), we_fmul(y, y)), we_fmul(z, z)));
}

// same with pre-computed 0.5f*dim (exact: a power-of-two scaling)
WE_HD float we_angle_sep_h(float ax, float ay, float az, float bx, float by, float bz,
                           const float* dim, const float* half) {
  float x = fabsf(we_fsub(ax, bx)); if (x > half[0]) x = we_fsub(x, dim[0]);
  float y = fabsf(we_fsub(ay, by)); if (y > half[1]) y = we_fsub(y, dim[1]);
  float z = fabsf(we_fsub(az, bz)); if (z > half[2]) z = we_fsub(z, dim[2]);
  return we_fsqrt(we_fadd(we_fadd(we_fmul(x, x), we_fmul(y, y)), we_fmul(z, z)));
}

// squared separation: the argument of the square root above, same operations (exact variant), and
// with the sum contracted (fast-path variant, relative error <= 2 ulp)
WE_HD float we_angle_sep2_exact_h(float ax, float ay, float az, float bx, float by, float bz,
                                  const float* dim, const float* half) {
  float x = fabsf(we_fsub(ax, bx)); if (x > half[0]) x = we_fsub(x, dim[0]);
  float y = fabsf(we_fsub(ay, by)); if (y > half[1]) y = we_fsub(y, dim[1]);
  float z = fabsf(we_fsub(az, bz)); if (z > half[2]) z = we_fsub(z, dim[2]);
  return we_fadd(we_fadd(we_fmul(x, x), we_fmul(y, y)), we_fmul(z, z));
}
WE_HD float we_angle_sep2_h(float ax, float ay, float az, float bx, float by, float bz,
                            const float* dim, const float* half) {
  float x = fabsf(we_fsub(ax, bx)); if (x > half[0]) x = we_fsub(x, dim[0]);
  float y = fabsf(we_fsub(ay, by)); if (y > half[1]) y = we_fsub(y, dim[1]);
  float z = fabsf(we_fsub(az, bz)); if (z > half[2]) z = we_fsub(z, dim[2]);
  return fmaf(x, x, fmaf(y, y, we_fmul(z, z)));
}

// cosine from the three float32 separations and two pre-computed powf squares:
//   (powf(dac,2) - powf(dbc,2) - powf(dba,2)) / ((-2*dbc)*dba)
WE_HD float we_angle_cos_from(float dac, float dbc, float dbc2, float dba, float dba2,
                              const WePowTables& T) {
  float num = we_fsub(we_fsub(we_powf2_fast(dac, T), dbc2), dba2);
  float den = we_fmul(we_fmul(-2.0f, dbc), dba);
  return we_fdiv(num, den);
}

// full get_angle(a, b, c, dims[:3])   (b is the vertex)
WE_HD float we_angle_cos(const float* a, const float* b, const float* c, const float* dim,
                         const WePowTables& T) {
  float dba = we_angle_sep(a[0], a[1], a[2], b[0], b[1], b[2], dim);
  float dbc = we_angle_sep(c[0], c[1], c[2], b[0], b[1], b[2], dim);
  float dac = we_angle_sep(c[0], c[1], c[2], a[0], a[1], a[2], dim);
  return we_angle_cos_from(dac, dbc, we_powf2(dbc, T), dba, we_powf2(dba, T), T);
}

// float(degrees(arccos(c))) > 90 evaluated in float32  <=>  -1 <= c <= -2^-24
WE_HD bool we_angle_gt_90(float c) { return (c >= -1.0f) && (c <= -5.9604644775390625e-08f); }

// ---- glibc pow(x, 2.0) in double precision ------------------------------------
// The reference squares float64 scalars with `**2` (neighbours/RAD.py:53-57 `(1/r)**2`, analysis/HB.py:96
// `dist**2`), which NumPy hands to libm pow().  pow(x, 2.0) is correctly rounded in ~99.92 % of the cases
// and one ulp off x*x otherwise, so the device follows glibc (>= 2.28, e_pow.c as built for FMA CPUs;
// the contraction pattern below is the one of libm 2.39's __pow_fma, read from its disassembly)
// step by step: log_inline() -> (hi, lo), y = 2 -> exp_inline().  Valid for
// finite normal x > 0 with |2 ln x| < 512, which covers every distance and inverse distance (Angstrom)
// the path produces.  tests/host_numerics.cpp checks the host build against libm on 10^7 arguments; the
// GPU parity suite checks the device build against NumPy.
#include "we_pow64_tables.h"
#ifdef __CUDACC__
__device__ const double we_d_powlog[128][3] = {WE_POW_LOG_TAB};
__device__ const unsigned long long we_d_exptab[256] = {WE_EXP_TAB};
#endif
static const double we_h_powlog[128][3] = {WE_POW_LOG_TAB};
static const unsigned long long we_h_exptab[256] = {WE_EXP_TAB};

WE_HD double we_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return fma(a, b, c);
#endif
}

WE_HD double we_pow2(double x) {
#ifdef __CUDA_ARCH__
  const double(*LT)[3] = we_d_powlog;
  const unsigned long long* ET = we_d_exptab;
#else
  const double(*LT)[3] = we_h_powlog;
  const unsigned long long* ET = we_h_exptab;
#endif
  const double A[7] = WE_POW_A;
  const double C[4] = WE_EXP_C;
  // log_inline: x = 2^k z, z in [0x1.69555p-1, 0x1.69555p0), log(x) = k ln2 + log(c) + log1p(z/c - 1)
  const uint64_t ix = we_d2u(x);
  const uint64_t tmp = ix - 0x3fe6955500000000ULL;
  const int i = (int)((tmp >> 45) & 127);
  const int k = (int)((int64_t)tmp >> 52);
  const double z = we_u2d(ix - (tmp & (0xfffULL << 52)));
  const double kd = (double)k;
  const double invc = LT[i][0], logc = LT[i][1], logctail = LT[i][2];
  const double r = we_fma(z, invc, -1.0);
  const double t1 = we_fma(kd, WE_POW_LN2HI, logc);
  const double t2 = we_add(t1, r);
  const double lo1 = we_fma(kd, WE_POW_LN2LO, logctail);
  const double lo2 = we_add(we_add(t1, -t2), r);
  const double ar = we_mul(A[0], r);
  const double ar2 = we_mul(r, ar);
  const double ar3 = we_mul(r, ar2);
  const double hi = we_add(t2, ar2);
  const double lo3 = we_fma(ar, r, -ar2);
  const double lo4 = we_add(we_add(t2, -hi), ar2);
  double q = we_fma(ar2, we_fma(r, A[6], A[5]), we_fma(r, A[4], A[3]));
  q = we_fma(ar2, q, we_fma(r, A[2], A[1]));
  const double lo = we_fma(ar3, q, we_add(we_add(we_add(lo1, lo2), lo3), lo4));
  const double y = we_add(hi, lo);
  const double tail = we_add(we_add(hi, -y), lo);
  // y = 2: ehi = 2*log(x) and its rounding error are exact
  const double ehi = we_mul(2.0, y);
  const double elo = we_fma(2.0, tail, we_fma(2.0, y, -ehi));
  // exp_inline: exp(ehi + elo) = 2^(kk/128) exp(rr)
  double kd2 = we_fma(WE_EXP_INVLN2N, ehi, WE_EXP_SHIFT);
  const uint64_t ki = we_d2u(kd2);
  kd2 = we_add(kd2, -(WE_EXP_SHIFT));
  double rr = we_fma(kd2, WE_EXP_NEGLN2LON, we_fma(kd2, WE_EXP_NEGLN2HIN, ehi));
  rr = we_add(rr, elo);
  const uint64_t idx = 2 * (ki & 127);
  const double tl = we_u2d(ET[idx]);
  const double scale = we_u2d(ET[idx + 1] + (ki << 45));
  const double r2 = we_mul(rr, rr);
  double s = we_fma(r2, we_fma(rr, C[1], C[0]), we_add(tl, rr));
  s = we_fma(we_mul(r2, r2), we_fma(rr, C[3], C[2]), s);
  return we_fma(scale, s, scale);
}

// per-frame box description (built on the host by we_submit, see we_api.cu)
struct WeBox {
  float box[3];   // dimensions[:3]
  float inv[3];   // (float)(1.0 / (double)box[i])
  float m[9];     // float32 lower-triangular cell matrix (rows = a, b, c)
  int triclinic;
  // cell grid in fractional coordinates
  int nc[3];      // cells per fractional axis
  int wide[3];    // half-width (in cells) of the wide (10 A) search per axis
  int sw[3];      // half-width (in cells) of the first-pass search per axis
  double hinv[9]; // inverse cell matrix: fractional_j = sum_i r_i * hinv[3*i + j]
  float hinvf[9]; // float32 copy (candidate prefilter only)
  double wfac[3]; // slanted wrap bounds: {m3/m4, (m4*m6 - m3*m7)/(m4*m8), m7/m8}
  float rc_eff;   // radius guaranteed covered by the +-1 cell search
  int pre_ok;     // triclinic float prefilter valid (10 A < half the smallest cell width)
  // cell geometry for the scan's row culling (k_rad): y extent of a cell row m4/n1, y shift per z cell m7/n2,
  // z extent m8/n2, x shift per y cell m3/n1, x shift per z cell m6/n2, x cells per Angstrom n0/m0
  float cg[6];
};

WE_HD double we_round_half_away(double s) {
#ifdef __CUDA_ARCH__
  return round(s);
#else
  return round(s);
#endif
}

// MDAnalysis orthorhombic minimum-image distance, fp64 from float32 inputs
WE_HD double we_dist_ortho(float rx, float ry, float rz, float cx, float cy, float cz,
                           const WeBox& b) {
  double dx = (double)we_fsub(cx, rx);
  double dy = (double)we_fsub(cy, ry);
  double dz = (double)we_fsub(cz, rz);
  double s;
  s = we_mul((double)b.inv[0], dx); dx = we_mul((double)b.box[0], we_add(s, -we_round_half_away(s)));
  s = we_mul((double)b.inv[1], dy); dy = we_mul((double)b.box[1], we_add(s, -we_round_half_away(s)));
  s = we_mul((double)b.inv[2], dz); dz = we_mul((double)b.box[2], we_add(s, -we_round_half_away(s)));
  return we_dsqrt(we_add(we_add(we_mul(dx, dx), we_mul(dy, dy)), we_mul(dz, dz)));
}

// wrap a coordinate into the primary triclinic cell: c axis, then b axis, then a axis, each
// against the slanted bounds of the parallelepiped (MDAnalysis `_triclinic_pbc`); fp64, then float32.
// wf = {m3/m4, (m4*m6 - m3*m7)/(m4*m8), m7/m8} in double (WeBox.wfac, built on the host).
WE_HD void we_wrap_tric(float x, float y, float z, const float* m, const double* wf, float* out) {
  double c0 = x, c1 = y, c2 = z, s;
  s = floor(we_ddiv(c2, (double)m[8]));
  c0 = we_add(c0, -we_mul(s, (double)m[6])); c1 = we_add(c1, -we_mul(s, (double)m[7])); c2 = we_add(c2, -we_mul(s, (double)m[8]));
  s = floor(we_ddiv(we_add(c1, -we_mul(c2, wf[2])), (double)m[4]));
  c0 = we_add(c0, -we_mul(s, (double)m[3])); c1 = we_add(c1, -we_mul(s, (double)m[4]));
  s = floor(we_ddiv(we_add(c0, -we_add(we_mul(c1, wf[0]), we_mul(c2, wf[1]))), (double)m[0]));
  c0 = we_add(c0, -we_mul(s, (double)m[0]));
  out[0] = (float)c0; out[1] = (float)c1; out[2] = (float)c2;
}

// triclinic minimum image over the 27 neighbouring images (inputs already wrapped)
WE_HD double we_dist_tric(float rx, float ry, float rz, float cx, float cy, float cz,
                          const WeBox& b) {
  double dx = (double)we_fsub(cx, rx);
  double dy = (double)we_fsub(cy, ry);
  double dz = (double)we_fsub(cz, rz);
  const float* m = b.m;
  double best = 3.4028234663852886e+38;  // FLT_MAX
  for (int ix = -1; ix < 2; ++ix) {
    double rx0 = we_add(dx, we_mul((double)m[0], (double)ix));
    for (int iy = -1; iy < 2; ++iy) {
      double ry0 = we_add(rx0, we_mul((double)m[3], (double)iy));
      double ry1 = we_add(dy, we_mul((double)m[4], (double)iy));
      for (int iz = -1; iz < 2; ++iz) {
        double rz0 = we_add(ry0, we_mul((double)m[6], (double)iz));
        double rz1 = we_add(ry1, we_mul((double)m[7], (double)iz));
        double rz2 = we_add(dz, we_mul((double)m[8], (double)iz));
        double dsq = we_add(we_add(we_mul(rz0, rz0), we_mul(rz1, rz1)), we_mul(rz2, rz2));
        if (dsq < best) best = dsq;
      }
    }
  }
  return we_dsqrt(best);
}

// One-image evaluation of the triclinic distance.  The 27-image minimum of
// we_dist_tric() is attained at the image chosen by rounding the fractional
// separation whenever that minimum is below half the smallest perpendicular cell
// width (any other lattice translate is then longer), and its dsq is computed by
// the same operation sequence, so the result is bit-identical for every candidate
// that can pass a d <= 10 A test (b.pre_ok certifies 10 A is below that bound).
// Candidates beyond the bound return a value > 10 A, which is all callers need.
WE_HD double we_dist_tric_one(float rx, float ry, float rz, float cx, float cy, float cz,
                              const WeBox& b) {
  const double dx = (double)we_fsub(cx, rx);
  const double dy = (double)we_fsub(cy, ry);
  const double dz = (double)we_fsub(cz, rz);
  const double s0 = dx * b.hinv[0] + dy * b.hinv[3] + dz * b.hinv[6];
  const double s1 = dy * b.hinv[4] + dz * b.hinv[7];
  const double s2 = dz * b.hinv[8];
  double ix = -rint(s0), iy = -rint(s1), iz = -rint(s2);
  ix = fmin(fmax(ix, -1.0), 1.0); iy = fmin(fmax(iy, -1.0), 1.0); iz = fmin(fmax(iz, -1.0), 1.0);
  const float* m = b.m;
  const double rx0 = we_add(dx, we_mul((double)m[0], ix));
  const double ry0 = we_add(rx0, we_mul((double)m[3], iy));
  const double ry1 = we_add(dy, we_mul((double)m[4], iy));
  const double rz0 = we_add(ry0, we_mul((double)m[6], iz));
  const double rz1 = we_add(ry1, we_mul((double)m[7], iz));
  const double rz2 = we_add(dz, we_mul((double)m[8], iz));
  return we_dsqrt(we_add(we_add(we_mul(rz0, rz0), we_mul(rz1, rz1)), we_mul(rz2, rz2)));
}

// distance used by the kernels: exact for everything that can satisfy d <= 10 A
WE_HD double we_dist_fast(float rx, float ry, float rz, float cx, float cy, float cz, const WeBox& b) {
  if (!b.triclinic) return we_dist_ortho(rx, ry, rz, cx, cy, cz, b);
  if (b.pre_ok) return we_dist_tric_one(rx, ry, rz, cx, cy, cz, b);
  return we_dist_tric(rx, ry, rz, cx, cy, cz, b);
}

WE_HD double we_dist(float rx, float ry, float rz, float cx, float cy, float cz, const WeBox& b) {
  return b.triclinic ? we_dist_tric(rx, ry, rz, cx, cy, cz, b) : we_dist_ortho(rx, ry, rz, cx, cy, cz, b);
}
