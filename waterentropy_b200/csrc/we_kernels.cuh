// Hand-written sm_100a kernels for the waterEntropy per-frame hot path.
//
// One launch of each kernel processes a *batch* of frames.
//   K1  k_bin / k_scan / k_scatter   cell list over heavy atoms: cells of edge >= r/2 in
//                    fractional coordinates, float4 (x, y, z, atom) records in cell order
//   K2  k_rad        persistent, one warp per centre of a work list: flattened warp-dense
//                    cell scan -> float32 prefilter -> exact fp64 distances -> exact top-25
//                    -> RAD blocking test with the reference's float32 cosine
//       marks in k_rad + k_seed / k_list1 / k_list2   the work lists: solute atoms -> interfacial
//                    waters -> their shell members (the closure the reference memoises lazily)
//   K3  k_hb         persistent over the centres whose HBs are needed: 8 lanes per donor
//                    hydrogen, acceptor = argmin over the donor's RAD shell
//   K4  k_label      persistent, warp per water: labels, HB labels, sorted key, hash-table
//                    insert, record of the water (also into mapped pinned host memory)
//   K5  k_cov        principal-axis force / torque, outer products
//   K6  (fused in k_label / k_cov)  atomics into the label hash table and warp-aggregated
//                    fp64 atomics into the dense per-bin covariance sums
// Reference lines each kernel replaces are cited at the kernel.
#pragma once
#include <cuda_runtime.h>
#include "we_numerics.cuh"
#include "we_eig3.cuh"

#define WE_MAX_NBR 25
#define WE_GATE_NBR 20
#define WE_MAX_EXCL 8
#define WE_CAND_CAP 128
#define WE_CUTOFF 10.0
#ifndef WE_RAD_WARPS
#define WE_RAD_WARPS 8
#endif
#define WE_CELLS_PER_RADIUS 2  // cell edge = search radius / 2, searched +-2 cells
// resident blocks per SM the register allocator must allow for (tuned on B200)
#ifndef WE_RAD_MINBLOCKS
#define WE_RAD_MINBLOCKS 3
#endif
#ifndef WE_RAD_MINBLOCKS_GATED
#define WE_RAD_MINBLOCKS_GATED 4
#endif
#ifndef WE_HB_MINBLOCKS
#define WE_HB_MINBLOCKS 3
#endif
#ifndef WE_LABEL_MINBLOCKS
#define WE_LABEL_MINBLOCKS 4
#endif

// status codes (per centre / per donor), also the C-ABI error codes
#define WE_ST_OK 0
#define WE_ST_DUPLICATE 1
#define WE_ST_NO_NEIGHBOURS 2
#define WE_ST_OVERFLOW 3

// label prefix codes (analysis/shell_labels.py:41-83)
#define WE_P_NEAREST 0  // "0_RES"
#define WE_P_PLAIN 1    // "RES"
#define WE_P_FIRST 2    // "1_RES"
#define WE_P_SECOND 3   // "2_RES"
#define WE_P_OTHER 4    // "X_RES"

struct WeTopoDev {
  int n_atoms, n_heavy, n_don, n_centres, n_solute_ua;
  int n_classes, n_bins, recipe;
  const int* heavy_idx;           // [H] atom index of heavy slot
  const int* heavy_slot;          // [N] heavy slot of atom or -1
  const int* excl;                // [H][WE_MAX_EXCL] bonded heavy atoms (atom idx), -1 padded
  const int* resid;               // [N]
  const int* resname_id;          // [N]
  const int* type_id;             // [N]
  const unsigned char* is_water;  // [N]
  const double* charges;          // [N]
  const double* masses;           // [N]
  const int* don_ptr;             // [H+1] CSR over donors of each heavy slot
  const int* don_h;               // [ND] donor hydrogen (atom idx)
  const int* don_x;               // [ND] heavy slot it is bonded to
  const int* solute_ua;           // [S] heavy slots of solute united atoms
  const int* centres;             // [C] heavy slots of water centres
  const int* frag_ptr;            // [C+1] CSR: atoms of the centre's molecule
  const int* frag_atoms;
  const int* bin_of_atom;         // [N] covariance pair-bin of the atom's (resname,resid)
  const int* centre_class;        // [N] class of the centre's resname or -1
  const unsigned char* is_solute; // [H] heavy slot is a solute united atom (pass-0 centre of the interfacial recipe)
  const unsigned char* slot_flag; // [H] bit 0 water, bit 1 solute united atom, bits 2.. bonded heavy atoms
};

struct WeDiag {
  unsigned long long wide_searches;
  unsigned long long shrink_retries;
  unsigned long long table_entries;
  unsigned long long fast_path_mismatches;  // debug contexts: float32 fast-path decisions the exact replay contradicts
  unsigned long long eig_fallbacks;         // molecules whose inertia tensor took a LAPACK path that is not restated
  unsigned long long light_fetched;         // packed uploads: light atoms read from the caller's pinned buffer
  int err_code, err_frame, err_atom, pad;
};

struct WeLabelEntry {  // 512 bytes
  unsigned long long hash;        // 0 = empty
  unsigned long long hash2;       // independent second hash (collision check)
  unsigned long long first_seen_inv;  // max(~(frame_id * n_atoms + atom)): first insert wins
  unsigned long long shell_count;
  int nearest_resid;
  int nearest_resname_id;
  int n_labels;
  int n_w;
  unsigned short labels[32];      // sorted codes (prefix << 13 | resname_id)
  unsigned int donates[WE_MAX_NBR];   // indexed by position of first occurrence
  unsigned int accepts[WE_MAX_NBR];
  unsigned int pad2[50];
};
static_assert(sizeof(WeLabelEntry) == 512, "label entry must be 512 bytes");

// Bounds checks of our own (compute-sanitizer is closed on this pool): a -DWE_CHECKED build records the
// source line of the first violated index bound and we_sync reports it; the parity suite is run once per
// round against that build (profiles/).  The default build compiles the checks away.
#ifdef WE_CHECKED
__device__ int we_check_line = 0;
#define WE_ASSERT(cond) do { if (!(cond)) atomicCAS(&we_check_line, 0, __LINE__); } while (0)
#else
#define WE_ASSERT(cond) ((void)0)
#endif

__device__ __forceinline__ void we_set_error(WeDiag* diag, int code, int frame, int atom) {
  if (atomicCAS(&diag->err_code, 0, code) == 0) {
    diag->err_frame = frame;
    diag->err_atom = atom;
  }
}

// ============================================================================
// K1: cell list.  Replaces `select_atoms("mass 2 to 999 ...")` + brute-force
// `capped_distance` (maths/trig.py:59-67, 26-34) as the candidate generator.
// ============================================================================
// cells per edge of a cell group of the gating pre-stage (k_gate)
#define WE_GRP_X 2
#define WE_GRP_Y 2
#define WE_GRP_Z 1
// `packed` (optional): the heavy atoms of the batch as the host packed them, [frame][heavy slot][3]; the
// kernel then also writes them to their places in the full [frame][atom][3] layout `pos_out`, which is what
// the later kernels read (the hydrogens they need follow through k_fetch_light).
__global__ void k_bin(WeTopoDev T, const float* __restrict__ pos, const WeBox* __restrict__ boxes,
                      int* __restrict__ cell_count, int* __restrict__ cell_of,
                      int* __restrict__ rank_of, float4* __restrict__ tmp4, int ncap,
                      unsigned char* __restrict__ grp_flag, const float* __restrict__ packed,
                      float* __restrict__ pos_out) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  int f = blockIdx.y;
  if (h >= T.n_heavy) return;
  const WeBox& b = boxes[f];
  int atom = T.heavy_idx[h];
  float p[3];
  if (packed) {
    const float* s = packed + ((size_t)f * T.n_heavy + h) * 3;
    float* d = pos_out + ((size_t)f * T.n_atoms + atom) * 3;
    p[0] = s[0]; p[1] = s[1]; p[2] = s[2];
    d[0] = p[0]; d[1] = p[1]; d[2] = p[2];
  } else {
    const float* s = pos + ((size_t)f * T.n_atoms + atom) * 3;
    p[0] = s[0]; p[1] = s[1]; p[2] = s[2];
  }
  float q[3] = {p[0], p[1], p[2]};
  if (b.triclinic) we_wrap_tric(p[0], p[1], p[2], b.m, b.wfac, q);
  double s[3];
  for (int j = 0; j < 3; ++j) {
    double v = (double)q[0] * b.hinv[j] + (double)q[1] * b.hinv[3 + j] + (double)q[2] * b.hinv[6 + j];
    // triclinic: q is already wrapped, so v is in [0,1) up to float rounding; it is CLAMPED (below)
    // rather than re-wrapped so that the stored coordinate and the cell stay on the same side of
    // the boundary (k_rad's per-cell image shift relies on it).  Orthorhombic coordinates are kept
    // raw (the reference's distance formula needs them) and are wrapped here for binning only.
    if (!b.triclinic) v -= floor(v);
    s[j] = v;
  }
  int c0 = min((int)(s[0] * b.nc[0]), b.nc[0] - 1);
  int c1 = min((int)(s[1] * b.nc[1]), b.nc[1] - 1);
  int c2 = min((int)(s[2] * b.nc[2]), b.nc[2] - 1);
  c0 = max(c0, 0); c1 = max(c1, 0); c2 = max(c2, 0);
  int cell = (c2 * b.nc[1] + c1) * b.nc[0] + c0;
  size_t o = (size_t)f * T.n_heavy + h;
  WE_ASSERT(cell >= 0 && cell < ncap && b.nc[0] <= 1023 && b.nc[1] <= 1023 && b.nc[2] <= 1023);
  cell_of[o] = (c2 << 20) | (c1 << 10) | c0;  // packed coordinates (<= 1023 cells per axis)
  rank_of[o] = atomicAdd(&cell_count[(size_t)f * (ncap + 1) + cell], 1);
  tmp4[o] = make_float4(q[0], q[1], q[2], __int_as_float(atom));
  if (grp_flag && T.is_solute[h]) {  // the WE_GRP^3 cell group of a solute united atom has gating work (k_gate)
    const int g0 = (b.nc[0] + WE_GRP_X - 1) / WE_GRP_X, g1 = (b.nc[1] + WE_GRP_Y - 1) / WE_GRP_Y;
    grp_flag[(size_t)f * (ncap + 1) + ((size_t)(c2 / WE_GRP_Z) * g1 + c1 / WE_GRP_Y) * g0 + c0 / WE_GRP_X] = 1;
  }
}

// exclusive scan of the per-frame cell counts, in place; [ncell] receives the total.
// One 1024-thread block per frame; every thread owns 8 consecutive cells per sweep.
#define WE_SCAN_PER_THREAD 8
__global__ void k_scan(const WeBox* __restrict__ boxes, int* __restrict__ cell_count, int ncap) {
  __shared__ int s_part[32];
  __shared__ int s_carry;
  const int f = blockIdx.x;
  const WeBox& b = boxes[f];
  const int ncell = b.nc[0] * b.nc[1] * b.nc[2];
  int* cnt = cell_count + (size_t)f * (ncap + 1);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  const int span = blockDim.x * WE_SCAN_PER_THREAD;
  for (int base = 0; base < ncell + 1; base += span) {
    const int i0 = base + tid * WE_SCAN_PER_THREAD;
    int v[WE_SCAN_PER_THREAD];
    int sum = 0;
    if (i0 + WE_SCAN_PER_THREAD <= ncell) {  // 32-byte aligned: ncap + 1 is a multiple of 8
      const int4 a = *reinterpret_cast<const int4*>(cnt + i0), c4 = *reinterpret_cast<const int4*>(cnt + i0 + 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = c4.x; v[5] = c4.y; v[6] = c4.z; v[7] = c4.w;
#pragma unroll
      for (int k = 0; k < WE_SCAN_PER_THREAD; ++k) sum += v[k];
    } else {
#pragma unroll
      for (int k = 0; k < WE_SCAN_PER_THREAD; ++k) { v[k] = (i0 + k < ncell) ? cnt[i0 + k] : 0; sum += v[k]; }
    }
    int incl = sum;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_part[w] = incl;
    __syncthreads();
    if (w == 0) {
      const int pv = (lane < (blockDim.x >> 5)) ? s_part[lane] : 0;
      int pi = pv;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, pi, o);
        if (lane >= o) pi += t;
      }
      s_part[lane] = pi - pv;  // exclusive prefix of warp totals
    }
    __syncthreads();
    int run = s_carry + s_part[w] + incl - sum;  // exclusive prefix of this thread's first cell
    if (i0 + WE_SCAN_PER_THREAD <= ncell) {
      int4 a, c4;
      a.x = run; run += v[0]; a.y = run; run += v[1]; a.z = run; run += v[2]; a.w = run; run += v[3];
      c4.x = run; run += v[4]; c4.y = run; run += v[5]; c4.z = run; run += v[6]; c4.w = run; run += v[7];
      *reinterpret_cast<int4*>(cnt + i0) = a;
      *reinterpret_cast<int4*>(cnt + i0 + 4) = c4;
    } else {
#pragma unroll
      for (int k = 0; k < WE_SCAN_PER_THREAD; ++k) {
        if (i0 + k <= ncell) cnt[i0 + k] = run;
        run += v[k];
      }
    }
    __syncthreads();
    if (tid == blockDim.x - 1) s_carry = run;
    __syncthreads();
  }
}

// sflag (optional): per position of the cell-sorted array, bit 0 = water, bit 1 = solute united atom,
// bits 2.. = number of bonded heavy atoms (what k_gate needs of a candidate without chasing its atom index)
__global__ void k_scatter(int n_heavy, const WeBox* __restrict__ boxes, const int* __restrict__ cell_start,
                          const int* __restrict__ cell_of, const int* __restrict__ rank_of,
                          const float4* __restrict__ tmp4, float4* __restrict__ sorted4, int ncap,
                          const unsigned char* __restrict__ slot_flag, unsigned char* __restrict__ sflag) {
  int h = blockIdx.x * blockDim.x + threadIdx.x;
  int f = blockIdx.y;
  if (h >= n_heavy) return;
  size_t o = (size_t)f * n_heavy + h;
  const int pc = cell_of[o];
  const int cell = ((pc >> 20) * boxes[f].nc[1] + ((pc >> 10) & 1023)) * boxes[f].nc[0] + (pc & 1023);
  int dst = cell_start[(size_t)f * (ncap + 1) + cell] + rank_of[o];
  WE_ASSERT(cell >= 0 && cell < ncap && dst >= 0 && dst < n_heavy);
  sorted4[(size_t)f * n_heavy + dst] = tmp4[o];
  if (sflag) sflag[(size_t)f * n_heavy + dst] = slot_flag[h];
}

// ============================================================================
// K2: RAD shell, one warp per centre.
//   candidates : maths/trig.py:45-68 (heavy, not i, not bonded to i, d <= 10)
//   ordering   : maths/trig.py:36-38 (stable sort by distance -> ties by index)
//   RAD test   : analysis/RAD.py:39-71 with the float32 cosine of trig.py:103-122
//   gate       : recipes/interfacial_solvent.py:50-56 (water among the 20 nearest)
//   nearest non-like : analysis/shell_labels.py:90-105
// ============================================================================
struct WeRowInfo {  // per-warp description of up to 32 non-empty cell rows, compacted
  int4 ri[32];      // beg0, len0, beg1, first dense candidate index of the row
  float4 rp[32];    // centre minus image shift: x (first run), y, z, x (second run)
};

// Asynchronous global -> shared copies (LDGSTS): the candidate windows of the cell scan are staged one
// window ahead of the arithmetic, so the scan no longer waits on every window's gather.
__device__ __forceinline__ void we_cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void we_cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void we_cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

#ifndef WE_RAD_DYNAMIC
#define WE_RAD_DYNAMIC 1  // k_rad: 1 = warps claim centres from a counter, 0 = static common stride
#endif
#ifndef WE_CULL
#define WE_CULL 1  // cell scan: skip the rows / x cells of the searched block that the search sphere cannot reach
#endif
#define WE_CULL_MARGIN 0.05f  // Angstrom; the float32 prefilter itself keeps r + 0.02
#ifndef WE_PREFETCH
#define WE_PREFETCH 0  // candidate-window prefetch of the cell scan: 0 none, 1 cp.async ring, 2 registers
#endif

struct WeRadSmem {
#if WE_PREFETCH == 1
  float4 pre[WE_RAD_WARPS][64];  // two candidate windows per warp, filled by cp.async one window ahead
#else
  float4 pre[WE_RAD_WARPS][1];
#endif
  double d[WE_RAD_WARPS][WE_CAND_CAP];
  int idx[WE_RAD_WARPS][WE_CAND_CAP];
  int src[WE_RAD_WARPS][WE_CAND_CAP];
  float f[WE_RAD_WARPS][WE_CAND_CAP + 4];
  int top[WE_RAD_WARPS][32];
  int stage[WE_RAD_WARPS][64];
  WeRowInfo rows[WE_RAD_WARPS];
  float4 p4[WE_RAD_WARPS][32];  // ranked candidate k: position, separation e_k from the centre (float32);
                                // before the ranking: float32 d^2 of the prefilter survivors (gated pass)
  float4 q4[WE_RAD_WARPS][32];  // float32 e_k^2, 1 / r_k^2, 2 e_k
  double q[WE_RAD_WARPS][32];   // r_k, the exact fp64 distance of ranked candidate k
  double log2tab[32];
  unsigned long long exp2tab[32];
  int pref[130];  // dynamic work distribution: first item of every frame of the batch (<= 128 frames)
};

// ---------------------------------------------------------------------------
// work lists: which heavy-atom centres need a RAD shell in this frame.  The
// reference memoises shells lazily (analysis/RAD.py:86); the device computes
// exactly the same closure in three dependent passes:
//   pass 0  solute united atoms (interfacial) / every water centre (bulk)
//   pass 1  interfacial waters marked by pass 0 (k_list1)
//   pass 2  shell members of the interfacial waters not yet computed (k_list2)
// ---------------------------------------------------------------------------
struct WeWork {
  const int* list;   // heavy slots; frame f starts at list + f * stride
  const int* count;  // per-frame count (nullptr -> static_count for every frame)
  int stride;
  int static_count;
};

// Exact phase of the candidate search: lane `lane` < take owns staged candidate
// stage[(head + lane) & 63] (an index into the cell-sorted array).  Applies the
// reference's candidate rule (not the centre, not bonded to it), the duplicate-
// coordinate test and the bit-exact fp64 distance; appends survivors (d <= r).
__device__ __forceinline__ int we_flush(const WeBox& b, const float4* __restrict__ sorted,
                                        const int* stage, int head, int take, double r, float mx,
                                        float my, float mz, int my_atom, const int* ex, double* sd,
                                        float* sf, int* si, int* ss, int n, int lane, bool& dup) {
  bool keep = false;
  double dd = 0.0;
  int ai = -1, qi = -1;
  if (lane < take) {
    qi = stage[(head + lane) & 63];
    WE_ASSERT(qi >= 0);
    const float4 c = sorted[qi];
    ai = __float_as_int(c.w);
    bool skip = (ai == my_atom);
#pragma unroll
    for (int e = 0; e < WE_MAX_EXCL; ++e) skip |= (ex[e] == ai);
    if (!skip) {
      if (c.x == mx && c.y == my && c.z == mz) dup = true;
      dd = we_dist_fast(mx, my, mz, c.x, c.y, c.z, b);
      keep = (dd <= r);
    }
  }
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (keep) {
    const int p = n + __popc(m & ((1u << lane) - 1u));
    if (p < WE_CAND_CAP) { sd[p] = dd; sf[p] = (float)dd; si[p] = ai; ss[p] = qi; }
  }
  return n + __popc(m);
}

// visit the cell rows within +-(wx,wy,wz) cells of (c0,c1,c2); keep candidates with d <= r.
// Rows (runs of cells contiguous in x) are described one per lane, a warp scan turns their
// lengths into one dense candidate index space, and all 32 lanes walk that space together:
// Phase A (coalesced float4 loads): float32 prefilter.  For triclinic boxes the stored
//   coordinates are wrapped, every visited cell has one definite periodic image and the test is
//   |c + shift(row) - me|^2 <= (r + 0.02)^2; orthorhombic coordinates are kept raw (the
//   reference's distance formula needs them) and use the rounding nearest image instead.
//   Survivors are warp-compacted into a 64-entry ring in shared memory.
// Phase B (we_flush): every 32 staged survivors get the exact fp64 treatment, all lanes busy.
// Returns the number kept (may exceed WE_CAND_CAP -> caller retries).
template <bool defer>
__device__ __forceinline__ int we_collect(const WeBox& b, const int* __restrict__ cstart,
                                          const float4* __restrict__ sorted, int c0, int c1, int c2,
                                          int wx, int wy, int wz, double r, float mx, float my,
                                          float mz, int my_atom, const int* ex, double* sd, float* sf,
                                          int* si, int* ss, int* stage, WeRowInfo* rows, int lane,
                                          bool& dup, float* sg, float4* ring) {
  int n = 0, head = 0, cnt = 0;
  const int n0 = b.nc[0], n1 = b.nc[1], n2 = b.nc[2];
  const bool all_z = (2 * wz + 1 > n2), all_y = (2 * wy + 1 > n1), all_x = (2 * wx + 1 > n0);
  const int ny = all_y ? n1 : 2 * wy + 1, nz = all_z ? n2 : 2 * wz + 1;
  const int nrows = ny * nz;
  const float rf = (float)r + 0.02f;  // float prefilter radius (exact test follows in fp64)
  const float rf2 = rf * rf;
  const bool simple = b.triclinic && !(all_x || all_y || all_z);
#if WE_CULL
  // row culling (below) needs one definite periodic image per visited cell on every axis
  const bool cull = (2 * wx + 1 < n0) && (2 * wy + 1 < n1) && (2 * wz + 1 < n2);
#else
  const bool cull = false;
#endif
  const float cull_r = (float)r + WE_CULL_MARGIN;
  const float cull_r2 = cull_r * cull_r;
  // the centre inside the box (orthorhombic coordinates are stored raw; triclinic ones are wrapped already)
  float cx_w = mx, cy_w = my, cz_w = mz;
  if (cull && !b.triclinic) {
    cx_w = mx - floorf(mx * b.inv[0]) * b.box[0];
    cy_w = my - floorf(my * b.inv[1]) * b.box[1];
    cz_w = mz - floorf(mz * b.inv[2]) * b.box[2];
  }
  const int mode = !b.triclinic ? 0 : (b.pre_ok ? 1 : 2);
  for (int row0 = 0; row0 < nrows; row0 += 32) {
    // ---- one lane per row: cell ranges and image shift --------------------------------------
    const int row = row0 + lane;
    int beg0 = 0, len0 = 0, beg1 = 0, len1 = 0;
    float px0 = mx, px1 = mx, py = my, pz = mz;
    if (row < nrows) {
      const int iy = row % ny, iz = row / ny;
      int zc = all_z ? iz : c2 + iz - wz;
      const int kz = all_z ? 0 : (zc < 0 ? -1 : (zc >= n2 ? 1 : 0));
      zc -= kz * n2;
      int yc = all_y ? iy : c1 + iy - wy;
      const int ky = all_y ? 0 : (yc < 0 ? -1 : (yc >= n1 ? 1 : 0));
      yc -= ky * n1;
      const int rowbase = (zc * n1 + yc) * n0;
      // centre minus the image shift of this row (lower-triangular cell: rows a, b, c)
      py = my - (ky * b.m[4] + kz * b.m[7]);
      pz = mz - kz * b.m[8];
      px0 = mx - (ky * b.m[3] + kz * b.m[6]);
      px1 = px0;
      if (all_x) { beg0 = cstart[rowbase]; len0 = cstart[rowbase + n0] - beg0; }
      else {
        int lo = c0 - wx, hi = c0 + wx;
        if (cull) {
          // The searched block of cells is a parallelepiped around a sphere (17 % of its volume in a
          // 60/60/90 degree box): drop the rows the sphere does not reach and trim the others in x.
          // Lower bound of the distance between the centre and the row's tube (infinite in x), from the
          // bounding rectangle of the tube's cross-section in the y-z plane (lower-triangular cell: z depends
          // on the c index only, y on b and c); what remains of the radius bounds |x - x_centre|, and the
          // x cells follow after subtracting the range of the row's shear s1 m3 + s2 m6.  WE_CULL_MARGIN
          // covers float rounding and atoms clamped into edge cells (k_bin); it exceeds the prefilter's own
          // margin, so exactly the candidates the prefilter would reject are skipped.
          const int yu = c1 + iy - wy, zu = c2 + iz - wz;  // unwrapped cell indices of the row
          const float zl = zu * b.cg[2], yz0 = zu * b.cg[1], yl = yu * b.cg[0];
          const float gz = fmaxf(0.f, fmaxf(zl - cz_w, cz_w - (zl + b.cg[2])));
          const float gy = fmaxf(0.f, fmaxf(yl + fminf(yz0, yz0 + b.cg[1]) - cy_w, cy_w - (yl + b.cg[0] + fmaxf(yz0, yz0 + b.cg[1]))));
          const float rem = cull_r2 - (gy * gy + gz * gz);
          if (rem < 0.f) hi = lo - 1;
          else {
            const float hx = sqrtf(rem);
            const float sa = yu * b.cg[3], sb = zu * b.cg[4];
            const float shlo = fminf(sa, sa + b.cg[3]) + fminf(sb, sb + b.cg[4]);
            const float shhi = fmaxf(sa, sa + b.cg[3]) + fmaxf(sb, sb + b.cg[4]);
            lo = max(lo, (int)floorf((cx_w - hx - shhi) * b.cg[5]));
            hi = min(hi, (int)floorf((cx_w + hx - shlo) * b.cg[5]));
          }
        }
        if (lo <= hi) {
          // the part of [lo, hi] inside the box, and the part in the neighbouring image (wrapped cells)
          const int a0 = max(lo, 0), a1 = min(hi, n0 - 1);
          if (a0 <= a1) { beg0 = cstart[rowbase + a0]; len0 = cstart[rowbase + a1 + 1] - beg0; }
          if (lo < 0) {
            beg1 = cstart[rowbase + lo + n0]; len1 = cstart[rowbase + min(hi, -1) + n0 + 1] - beg1; px1 = px0 + b.m[0];
          } else if (hi >= n0) {
            beg1 = cstart[rowbase + max(lo, n0) - n0]; len1 = cstart[rowbase + hi - n0 + 1] - beg1; px1 = px0 - b.m[0];
          }
        }
      }
    }
    const int len = len0 + len1;
    WE_ASSERT(len0 >= 0 && len1 >= 0 && beg0 >= 0 && beg1 >= 0);
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    // non-empty rows, compacted: their first dense indices are strictly increasing
    const unsigned nonempty = __ballot_sync(0xffffffffu, len > 0);
    const int nr = __popc(nonempty);
    __syncwarp();
    if (len > 0) {
      const int ci = __popc(nonempty & ((1u << lane) - 1u));
      rows->ri[ci] = make_int4(beg0, len0, beg1, incl - len);
      rows->rp[ci] = make_float4(px0, py, pz, px1);
    }
    __syncwarp();
    const int coff = (lane < nr) ? rows->ri[lane].w : 0x7fffffff;  // lane r holds the start of compact row r
    // ---- dense walk over the candidates of these rows ---------------------------------------
    // Row of dense index q = base + lane: rows that start at or before `base` are counted by one
    // ballot, rows that start inside the 32-wide window set bit (start - base) of a warp-wide OR.
    // locate: the candidate of dense index base + lane (position in the cell-sorted array, its row);
    // test: float32 prefilter of its coordinates.  Split so that the gather of the NEXT window can be in
    // flight (cp.async into the warp's shared-memory ring) while this window is tested and pushed.
    auto locate = [&](int base, int& qi, int& rinfo) -> bool {
      const int rel = coff - base;
      const unsigned starts = __reduce_or_sync(0xffffffffu, (rel > 0 && rel < 32) ? (1u << rel) : 0u);
      const int r0 = __popc(__ballot_sync(0xffffffffu, rel <= 0)) - 1;
      const int q = base + lane;
      if (q >= total) return false;
      const int ro = r0 + __popc(starts & (0xffffffffu >> (31 - lane)));
      const int4 ri = rows->ri[ro];
      const int t = q - ri.w;
      const bool second = t >= ri.y;
      qi = second ? ri.z + (t - ri.y) : ri.x + t;
      WE_ASSERT(ro >= 0 && ro < 32 && t >= 0 && qi >= 0);
      rinfo = ro | (second ? 0x40000000 : 0);
      return true;
    };
    auto test = [&](const float4 c, int rinfo, float& d2) -> bool {
      if (simple) {
        const float4 rp = rows->rp[rinfo & 0xffff];
        const float ux = c.x - ((rinfo & 0x40000000) ? rp.w : rp.x);
        const float uy = c.y - rp.y, uz = c.z - rp.z;
        d2 = __fmaf_rn(ux, ux, __fmaf_rn(uy, uy, uz * uz));
        return d2 <= rf2;
      }
      const float ux = c.x - mx, uy = c.y - my, uz = c.z - mz;
      if (mode == 0) {
        const float fx = __fmaf_rn(-b.box[0], rintf(ux * b.inv[0]), ux);
        const float fy = __fmaf_rn(-b.box[1], rintf(uy * b.inv[1]), uy);
        const float fz = __fmaf_rn(-b.box[2], rintf(uz * b.inv[2]), uz);
        d2 = __fmaf_rn(fx, fx, __fmaf_rn(fy, fy, fz * fz));
        return d2 <= rf2;
      }
      if (mode == 1) {
        float s0 = ux * b.hinvf[0] + uy * b.hinvf[3] + uz * b.hinvf[6];
        float s1 = uy * b.hinvf[4] + uz * b.hinvf[7];
        float s2 = uz * b.hinvf[8];
        s0 -= rintf(s0); s1 -= rintf(s1); s2 -= rintf(s2);
        const float fx = s0 * b.m[0] + s1 * b.m[3] + s2 * b.m[6];
        const float fy = s1 * b.m[4] + s2 * b.m[7];
        const float fz = s2 * b.m[8];
        d2 = fx * fx + fy * fy + fz * fz;
        return d2 <= rf2;
      }
      d2 = -1.0f;  // no float32 estimate for this box: the caller must not use the deferred mode
      return true;
    };
    auto push = [&](bool pass, int qi, float d2) {
      const unsigned m = __ballot_sync(0xffffffffu, pass);
      if (defer) {  // survivors are only listed (index into the cell-sorted array, float32 d^2)
        const int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (pass && pos < WE_CAND_CAP) { ss[pos] = qi; sg[pos] = d2; }
        cnt += __popc(m);
        return;
      }
      if (pass) stage[(head + cnt + __popc(m & ((1u << lane) - 1u))) & 63] = qi;
      cnt += __popc(m);
      if (cnt >= 32) {
        __syncwarp();
        n = we_flush(b, sorted, stage, head, 32, r, mx, my, mz, my_atom, ex, sd, sf, si, ss, n, lane, dup);
        head = (head + 32) & 63;
        cnt -= 32;
        __syncwarp();
      }
    };
#if WE_PREFETCH == 1
    {
      int qa = 0, ra = 0, buf = 0;
      bool va = locate(0, qa, ra);
      if (va) we_cp_async16(&ring[lane], &sorted[qa]);
      we_cp_async_commit();
      for (int base = 0; base < total; base += 32) {
        int qn = 0, rn = 0;
        bool vn = false;
        if (base + 32 < total) {
          vn = locate(base + 32, qn, rn);
          if (vn) we_cp_async16(&ring[(buf ^ 1) * 32 + lane], &sorted[qn]);
        }
        we_cp_async_commit();
        we_cp_async_wait<1>();  // this window has landed; a lane reads back only what it copied itself
        float da = 0.f;
        bool pa = false;
        if (va) pa = test(ring[buf * 32 + lane], ra, da);
        push(pa, qa, da);
        qa = qn; ra = rn; va = vn; buf ^= 1;
      }
      we_cp_async_wait<0>();
    }
#elif WE_PREFETCH == 2
    {
      int qa = 0, ra = 0;
      bool va = locate(0, qa, ra);
      float4 ca = make_float4(0.f, 0.f, 0.f, 0.f);
      if (va) ca = sorted[qa];
      for (int base = 0; base < total; base += 32) {
        int qn = 0, rn = 0;
        bool vn = false;
        float4 cn = make_float4(0.f, 0.f, 0.f, 0.f);
        if (base + 32 < total) {
          vn = locate(base + 32, qn, rn);
          if (vn) cn = sorted[qn];  // in flight while this window is tested and pushed
        }
        float da = 0.f;
        bool pa = false;
        if (va) pa = test(ca, ra, da);
        push(pa, qa, da);
        qa = qn; ra = rn; va = vn; ca = cn;
      }
    }
#else
    (void)ring;
    for (int base = 0; base < total; base += 32) {
      int qa = 0, ra = 0;
      float da = 0.f;
      bool pa = locate(base, qa, ra);
      if (pa) pa = test(sorted[qa], ra, da);
      push(pa, qa, da);
    }
#endif
  }
  if (defer) { __syncwarp(); return cnt; }
  if (cnt > 0) {
    __syncwarp();
    n = we_flush(b, sorted, stage, head, cnt, r, mx, my, mz, my_atom, ex, sd, sf, si, ss, n, lane, dup);
  }
  dup = __any_sync(0xffffffffu, dup);
  return n;
}

// Exact phase over the survivors listed by a deferred we_collect (ss[0..cnt) = indices into the
// cell-sorted array).  In place: chunk i reads ss[32 i + lane] before anything is written, and the
// compacted outputs land at indices <= the ones read.
__device__ __forceinline__ int we_exact_list(const WeBox& b, const float4* __restrict__ sorted, int cnt,
                                             double r, float mx, float my, float mz, int my_atom,
                                             const int* ex, double* sd, float* sf, int* si, int* ss,
                                             int lane, bool& dup) {
  int n = 0;
  for (int i0 = 0; i0 < cnt; i0 += 32) {
    const int take = min(32, cnt - i0);
    bool keep = false;
    double dd = 0.0;
    int ai = -1, qi = -1;
    if (lane < take) {
      qi = ss[i0 + lane];
      const float4 c = sorted[qi];
      ai = __float_as_int(c.w);
      bool skip = (ai == my_atom);
#pragma unroll
      for (int e = 0; e < WE_MAX_EXCL; ++e) skip |= (ex[e] == ai);
      if (!skip) {
        if (c.x == mx && c.y == my && c.z == mz) dup = true;
        dd = we_dist_fast(mx, my, mz, c.x, c.y, c.z, b);
        keep = (dd <= r);
      }
    }
    __syncwarp();
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int p = n + __popc(m & ((1u << lane) - 1u));
      if (p < WE_CAND_CAP) { sd[p] = dd; sf[p] = (float)dd; si[p] = ai; ss[p] = qi; }  // p <= its source index < cap
    }
    n += __popc(m);
    __syncwarp();
  }
  dup = __any_sync(0xffffffffu, dup);
  return n;
}

struct WeRadOut {
  int* shell_n; int* shell_idx; int* nn; int* flags; int* nbr_idx; double* nbr_dist;
};

// What a pass of k_rad marks for the next work lists (bytes, plain stores; k_list1 / k_list2 compact them):
//   interf : pass 0 of the interfacial recipe - water members of the shell of a solute united atom that
//            has a water among its 20 nearest candidates (recipes/interfacial_solvent.py:50-65)
//   hb     : centres whose hydrogen bonds the reference evaluates - the centre and every shell member
//            (analysis/HB.py:142-169); bulk recipe: only for pure shells (recipes/bulk_water.py:49-51)
//   shell  : shell members whose own RAD shell the labels need (analysis/shell_labels.py:55-76)
struct WeMarks {
  unsigned char* interf; unsigned char* hb; unsigned char* shell;
  const long long* frame_ids;  // non-null: report search errors of the centres of this pass
  int bulk;
};

struct WeRadSmem;
__device__ __forceinline__ void we_rad_finish(const WeTopoDev& T, const float* __restrict__ fpos, const WeBox& b,
                                              const float4* __restrict__ sorted, int f, int h, int my_atom,
                                              float rx, float ry, float rz, int n, int status, bool dup,
                                              const WeRadOut& O, int* gate_only_state, const WeMarks& M,
                                              WeDiag* diag, WeRadSmem& S, const WePowTables& PT, int lane, int w);

// One centre (heavy slot h of frame f) on one warp.  GATED: the gating pass of the interfacial recipe
// (gate_only_state != nullptr) - a separate instantiation, so the other passes do not carry its code.
template <bool GATED>
__device__ __forceinline__ void we_rad_one(const WeTopoDev& T, const float* __restrict__ fpos,
                                           const WeBox& b, const int* __restrict__ cstart,
                                           const int* __restrict__ cell_of_f,
                                           const float4* __restrict__ sorted,
                                           const float4* __restrict__ tmp4_f, float search_radius,
                                           int f, int h, const WeRadOut& O, int* gate_only_state,
                                           const WeMarks& M, WeDiag* diag, WeRadSmem& S, const WePowTables& PT, int lane,
                                           int w) {
  const int my_atom = T.heavy_idx[h];
  // raw coordinates (angles, duplicate test) and distance coordinates (k_bin wrapped them already)
  const float rx = fpos[3 * my_atom], ry = fpos[3 * my_atom + 1], rz = fpos[3 * my_atom + 2];
  const float4 mq4 = tmp4_f[h];
  const float mq[3] = {mq4.x, mq4.y, mq4.z};
  int ex[WE_MAX_EXCL];
#pragma unroll
  for (int e = 0; e < WE_MAX_EXCL; ++e) ex[e] = T.excl[(size_t)h * WE_MAX_EXCL + e];
  const int cell = cell_of_f[h];
  const int c0 = cell & 1023, c1 = (cell >> 10) & 1023, c2 = cell >> 20;
  static_assert(sizeof(float4) * 32 >= sizeof(float) * WE_CAND_CAP, "p4 doubles as the float32 d^2 list");
  float* sg = reinterpret_cast<float*>(S.p4[w]);
  double* sd = S.d[w];
  float* sf = S.f[w];
  int* si = S.idx[w];
  int* ss = S.src[w];

  // ---- candidate collection with adaptive radius -------------------------
  const double r_cover = fmin((double)b.rc_eff, WE_CUTOFF);
  double r = fmin((double)search_radius, r_cover);
  bool dup = false;
  int n = 0;
  int status = WE_ST_OK;
  bool wide = false;
  for (int iter = 0; iter < 64; ++iter) {
    dup = false;
    if (GATED && !wide && iter == 0 && (!b.triclinic || b.pre_ok)) {
      // Solute atom of the gating pass (recipes/interfacial_solvent.py:50-56): most are buried.  List the
      // float32 survivors first; if at least 20 non-excluded solute atoms are closer (by a safe margin)
      // than the nearest water in range and than the search radius, no water can be among the 20
      // nearest candidates and the exact distances, the ranking and the RAD test are never needed.
      const int nf = we_collect<true>(b, cstart, sorted, c0, c1, c2, b.sw[0], b.sw[1], b.sw[2], r, mq[0], mq[1], mq[2], my_atom, ex, sd, sf, si, ss, S.stage[w], &S.rows[w], lane, dup, sg, S.pre[w]);
      if (nf > WE_CAND_CAP) {
        n = nf;  // overflow: the shrink / retry logic below takes over with the interleaved mode
      } else {
        bool buried = false;
        if (r <= r_cover) {  // everything closer than r has been seen
          float wmin = (float)(r * r);
          bool zero = false;
          int solute_d2_bits[WE_CAND_CAP / 32];
#pragma unroll
          for (int k = 0; k < WE_CAND_CAP / 32; ++k) {
            const int i = k * 32 + lane;
            solute_d2_bits[k] = 0x7f7fffff;  // FLT_MAX: not a solute candidate
            if (i < nf) {
              const int ai = __float_as_int(sorted[ss[i]].w);
              bool skip = (ai == my_atom);
#pragma unroll
              for (int e = 0; e < WE_MAX_EXCL; ++e) skip |= (ex[e] == ai);
              if (!skip) {
                const float d2 = sg[i];
                zero |= !(d2 > 0.f);  // identical coordinates (or NaN): the exact path must see this centre
                if (T.is_water[ai]) wmin = fminf(wmin, d2);
                else solute_d2_bits[k] = __float_as_int(d2);
              }
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) wmin = fminf(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
          // strictly inside the radius and closer than the nearest water by 1e-3 relative in d^2
          // (the float32 image shift of a triclinic row is good to ~1e-4 relative in d^2)
          const float lim = wmin * (1.0f - 1.0e-3f);
          int closer = 0;
#pragma unroll
          for (int k = 0; k < WE_CAND_CAP / 32; ++k)
            closer += __popc(__ballot_sync(0xffffffffu, __int_as_float(solute_d2_bits[k]) < lim));
          buried = (closer >= WE_GATE_NBR) && !__any_sync(0xffffffffu, zero);
        }
        if (buried) {
          const size_t hob = (size_t)f * T.n_heavy + h;
          if (lane == 0) { O.shell_n[hob] = 0; O.nn[hob] = -1; O.flags[hob] = 0; gate_only_state[hob] = 0; }
          __syncwarp();
          return;
        }
        dup = false;
        n = we_exact_list(b, sorted, nf, r, mq[0], mq[1], mq[2], my_atom, ex, sd, sf, si, ss, lane, dup);
      }
    }
    else if (!wide) n = we_collect<false>(b, cstart, sorted, c0, c1, c2, b.sw[0], b.sw[1], b.sw[2], r, mq[0], mq[1], mq[2], my_atom, ex, sd, sf, si, ss, S.stage[w], &S.rows[w], lane, dup, sg, S.pre[w]);
    else n = we_collect<false>(b, cstart, sorted, c0, c1, c2, b.wide[0], b.wide[1], b.wide[2], r, mq[0], mq[1], mq[2], my_atom, ex, sd, sf, si, ss, S.stage[w], &S.rows[w], lane, dup, sg, S.pre[w]);
    if (n > WE_CAND_CAP) {
      if (n >= WE_MAX_NBR && r > 0.5) { r *= 0.85; if (lane == 0) atomicAdd(&diag->shrink_retries, 1ULL); continue; }
      status = WE_ST_OVERFLOW; break;
    }
    if (n >= WE_MAX_NBR) break;
    if (r >= WE_CUTOFF) break;                 // everything within the cut-off has been seen
    if (!wide && r < r_cover) { r = r_cover; continue; }
    if (!wide) { wide = true; if (lane == 0) atomicAdd(&diag->wide_searches, 1ULL); }
    r = fmin(r * 1.3, WE_CUTOFF);
  }
  we_rad_finish(T, fpos, b, sorted, f, h, my_atom, rx, ry, rz, n, status, dup, O, gate_only_state, M, diag, S, PT, lane, w);
}

// Second half of a centre: the candidates (d <= r, exact fp64 distances sd, atom indices si, positions in
// the cell-sorted array ss) are listed in the warp's shared-memory scratch; rank them, run the RAD test,
// write the outputs and the marks.
__device__ __forceinline__ void we_rad_finish(const WeTopoDev& T, const float* __restrict__ fpos, const WeBox& b,
                                              const float4* __restrict__ sorted, int f, int h, int my_atom,
                                              float rx, float ry, float rz, int n, int status, bool dup,
                                              const WeRadOut& O, int* gate_only_state, const WeMarks& M,
                                              WeDiag* diag, WeRadSmem& S, const WePowTables& PT, int lane, int w) {
  double* sd = S.d[w];
  float* sf = S.f[w];
  int* si = S.idx[w];
  int* ss = S.src[w];
  __syncwarp();
  if (status == WE_ST_OK && n > WE_CAND_CAP) status = WE_ST_OVERFLOW;
  if (status == WE_ST_OK && dup) status = WE_ST_DUPLICATE;
  if (status == WE_ST_OK && n == 0) status = WE_ST_NO_NEIGHBOURS;
  const size_t ho = (size_t)f * T.n_heavy + h;
  if (status != WE_ST_OK) {
    if (lane == 0) {
      O.shell_n[ho] = 0; O.nn[ho] = -1; O.flags[ho] = status << 8;
      if (M.frame_ids) we_set_error(diag, status, (int)M.frame_ids[f], my_atom);
      if (M.hb && !M.bulk) M.hb[ho] = 1;
    }
    if (O.nbr_idx && lane < WE_MAX_NBR) O.nbr_idx[ho * WE_MAX_NBR + lane] = -1;
    return;
  }
  // ---- ranking by (distance, atom index) ------------------------------------------
  // float32 keys first: float(d) is monotone in d, so ranks from float keys are exact
  // unless two of the 32 smallest keys tie; then the fp64 (d, index) order decides.
  const int n25 = min(n, WE_MAX_NBR);
  const int npad = (n + 3) & ~3;
  if (lane < npad - n) sf[n + lane] = 3.0e38f;
  __syncwarp();
  unsigned seen = 0u;
  for (int c = lane; c < n; c += 32) {
    const float fc = sf[c];
    int rank = 0;
    for (int m = 0; m < npad; m += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(sf + m);
      rank += (k4.x < fc) + (k4.y < fc) + (k4.z < fc) + (k4.w < fc);
    }
    if (rank < 32) { S.top[w][rank] = c; seen |= 1u << rank; }
  }
  seen = __reduce_or_sync(0xffffffffu, seen);
  const int want = min(n, 32);
  if (__popc(seen) != want) {  // float tie among the leaders: exact path
    for (int c = lane; c < n; c += 32) {
      const double dc = sd[c];
      const int ic = si[c];
      int rank = 0;
      for (int m = 0; m < n; ++m) {
        const double dm = sd[m];
        rank += (dm < dc) || (dm == dc && si[m] < ic);
      }
      if (rank < 32) S.top[w][rank] = c;
    }
  }
  __syncwarp();
  // ---- solute atoms without a water among their 20 nearest candidates need no RAD shell
  // (recipes/interfacial_solvent.py:50-56).  Their state goes back to "not computed", so
  // that k_list2 queues them again should an interfacial water hold them in its shell.
  if (gate_only_state) {
    bool wat0 = false;
    if (lane < min(n25, WE_GATE_NBR)) wat0 = T.is_water[si[S.top[w][lane]]] != 0;
    if (!__any_sync(0xffffffffu, wat0)) {
      if (lane == 0) { O.shell_n[ho] = 0; O.nn[ho] = -1; O.flags[ho] = 0; gate_only_state[ho] = 0; }
      __syncwarp();
      return;
    }
  }
  // ---- per-candidate precomputation ------------------------------------------
  const float half[3] = {0.5f * b.box[0], 0.5f * b.box[1], 0.5f * b.box[2]};
  const bool act = lane < n25;
  int cidx = -1;
  double dj = 0.0;
  float pjx = 0.f, pjy = 0.f, pjz = 0.f, ej = 0.f, sj = 0.f, qjf = 0.f;
  if (act) {
    const int c = S.top[w][lane];
    cidx = si[c];
    dj = sd[c];
    if (b.triclinic) {
      pjx = fpos[3 * cidx]; pjy = fpos[3 * cidx + 1]; pjz = fpos[3 * cidx + 2];
    } else {
      const float4 c4 = sorted[ss[c]];
      pjx = c4.x; pjy = c4.y; pjz = c4.z;
    }
    ej = we_angle_sep_h(pjx, pjy, pjz, rx, ry, rz, b.box, half);
    // float32 stand-ins for the fast path of the blocking test (relative error <= 3e-7); the exact
    // powf(e, 2) and (1/r)**2 = libm pow(1/r, 2.0) are evaluated only by the pairs that need them
    sj = we_fmul(ej, ej);
    const float rinv = __fdividef(1.0f, (float)dj);
    qjf = we_fmul(rinv, rinv);
    S.p4[w][lane] = make_float4(pjx, pjy, pjz, ej);
    S.q4[w][lane] = make_float4(sj, qjf, we_fmul(2.0f, ej), 0.f);
    S.q[w][lane] = dj;
  }
  if (O.nbr_idx && lane < WE_MAX_NBR) {
    O.nbr_idx[ho * WE_MAX_NBR + lane] = act ? cidx : -1;
    O.nbr_dist[ho * WE_MAX_NBR + lane] = act ? dj : 0.0;
  }
  __syncwarp();
  // ---- RAD blocking test: lane j against closer candidates k = 0..j-1 ---------
  // j is blocked by k when q_j < q_k cos(jik), i.e. cos > rho = q_j / q_k  (RAD.py:63-67).
  // Fast path, float32: with s2 = |p_k - p_j|^2 (same min-image components as get_angle),
  // A = e_k^2 + e_j^2 - s2 and D = 2 e_k e_j the reference's cosine is A / D to within 5e-6
  // (it rounds sqrt, powf and the division; A and D see the same inputs), so the sign of
  // A q_k - q_j D decides unless it is within 1e-3 q_k D of zero (>100x that error); those
  // pairs, and anything non-finite, replay the reference's float32/fp64 arithmetic exactly.
  // Debug contexts replay every pair and count contradictions (we_diag.fast_path_mismatches).
  // (evaluating two k per iteration or two scan candidates per lane was measured: no gain)
  bool blocked = false;
  bool alive = act && lane > 0;
  const bool check = (O.nbr_idx != nullptr);
  for (int k = 0; k < n25 - 1; ++k) {
    if (!__any_sync(0xffffffffu, alive)) break;
    if (alive && k < lane) {
      const float4 pk = S.p4[w][k], qk = S.q4[w][k];
      float x = fabsf(we_fsub(pk.x, pjx)); if (x > half[0]) x = we_fsub(x, b.box[0]);
      float y = fabsf(we_fsub(pk.y, pjy)); if (y > half[1]) y = we_fsub(y, b.box[1]);
      float z = fabsf(we_fsub(pk.z, pjz)); if (z > half[2]) z = we_fsub(z, b.box[2]);
      const float s2 = __fmaf_rn(x, x, __fmaf_rn(y, y, we_fmul(z, z)));
      const float A = we_fsub(we_fadd(qk.x, sj), s2);
      const float D = we_fmul(qk.z, ej);
      const float df = we_fsub(we_fmul(A, qk.y), we_fmul(qjf, D));
      const bool decided = (D > 0.f) && (fabsf(df) > we_fmul(1.0e-3f, we_fmul(qk.y, D)));
      bool blk = df > 0.f;
      if (!decided || check) {
        const float dac = we_fsqrt(we_fadd(we_fadd(we_fmul(x, x), we_fmul(y, y)), we_fmul(z, z)));
        const float cs = we_angle_cos_from(dac, pk.w, we_powf2_fast(pk.w, PT), ej, we_powf2_fast(ej, PT), PT);
        bool eb = false;
        if (cs != cs) alive = false;  // NaN: stop testing, j stays in the shell (RAD.py:60-61)
        else eb = we_pow2(we_ddiv(1.0, dj)) < we_mul(we_pow2(we_ddiv(1.0, S.q[w][k])), (double)cs);
        if (decided && (eb != blk || cs != cs)) atomicAdd(&diag->fast_path_mismatches, 1ULL);
        blk = eb;
      }
      if (blk) { blocked = true; alive = false; }
    } else if (alive && k >= lane) {
      alive = false;
    }
  }
  // ---- outputs ------------------------------------------------------------------
  const bool member = act && !blocked;
  const unsigned mm = __ballot_sync(0xffffffffu, member);
  const int ns = __popc(mm);
  WE_ASSERT(ns <= WE_MAX_NBR && h >= 0 && h < T.n_heavy);
  if (member) O.shell_idx[ho * WE_MAX_NBR + __popc(mm & ((1u << lane) - 1u))] = cidx;
  bool nonlike = false, wat = false;
  if (act) {
    nonlike = member && (T.resname_id[cidx] != T.resname_id[my_atom]) && (T.type_id[cidx] != T.type_id[my_atom]);
    wat = (lane < WE_GATE_NBR) && T.is_water[cidx];
  }
  const unsigned nm = __ballot_sync(0xffffffffu, nonlike);
  const unsigned wm = __ballot_sync(0xffffffffu, wat);
  int nn = -1;
  if (nm) nn = __shfl_sync(0xffffffffu, cidx, __ffs(nm) - 1);
  if (lane == 0) {
    O.shell_n[ho] = ns;
    O.nn[ho] = nn;
    O.flags[ho] = (wm ? 1 : 0);
  }
  // ---- marks for the next work lists ----------------------------------------------
  const size_t fo = (size_t)f * T.n_heavy;
  if (M.interf && wm) {
    if (member && T.is_water[cidx]) M.interf[fo + T.heavy_slot[cidx]] = 1;
  }
  if (M.hb) {
    bool go = true;
    if (M.bulk) {
      const bool other = member && (T.resname_id[cidx] != T.resname_id[my_atom]);
      go = ns > 0 && !__any_sync(0xffffffffu, other);
    }
    if (go) {
      if (lane == 0) M.hb[ho] = 1;
      if (member) {
        const int sl = T.heavy_slot[cidx];
        M.hb[fo + sl] = 1;
        if (M.shell) M.shell[fo + sl] = 1;
      }
    }
  }
  __syncwarp();
}

// Persistent kernel: gridDim.x blocks of WE_RAD_WARPS warps walk the work list of
// every frame of the batch with a grid-stride loop (no empty blocks when the list
// is short, natural load balance when per-centre cost varies).
template <bool GATED>
__global__ void __launch_bounds__(WE_RAD_WARPS * 32, GATED ? WE_RAD_MINBLOCKS_GATED : WE_RAD_MINBLOCKS)
k_rad(WeTopoDev T, const float* __restrict__ pos, const WeBox* __restrict__ boxes,
      const int* __restrict__ cell_start, const int* __restrict__ cell_of,
      const float4* __restrict__ sorted4, const float4* __restrict__ tmp4, int ncap,
      float search_radius, WeWork work, int n_frames, WeRadOut O, int* gate_only_state, WeMarks M,
      WeDiag* diag, int* work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WeRadSmem& S = *reinterpret_cast<WeRadSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < 32) { S.log2tab[tid] = ((const double*)we_d_log2tab)[tid]; S.exp2tab[tid] = we_d_exp2tab[tid]; }
  __syncthreads();
  const WePowTables PT{S.log2tab, S.exp2tab};
  // The work lists of the frames of a batch form ONE item space (frame after frame).
#if WE_RAD_DYNAMIC
  // Warps claim the next centre from a counter (one atomic per centre, ~3,000 warp-instructions of work each):
  // the cost of a centre varies (retries, exact replays, early gate exits) and a static stride leaves the tail
  // of every pass to the unluckiest warps (measured: -3.5 % on the three passes of a C4 batch).
  if (tid == 0) {
    int acc = 0;
    for (int f = 0; f < n_frames; ++f) { S.pref[f] = acc; acc += work.count ? work.count[f] : work.static_count; }
    S.pref[n_frames] = acc;
  }
  __syncthreads();
  const int total = S.pref[n_frames];
  int f = 0;
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(work_counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    while (S.pref[f + 1] <= item) ++f;  // a warp's claims increase: f never goes back
    const int i = item - S.pref[f];
    WE_ASSERT(f < n_frames && i >= 0 && i < (work.count ? work.count[f] : work.static_count));
    const int* list = work.list + (size_t)f * work.stride;
    we_rad_one<GATED>(T, pos + (size_t)f * T.n_atoms * 3, boxes[f], cell_start + (size_t)f * (ncap + 1),
                      cell_of + (size_t)f * T.n_heavy, sorted4 + (size_t)f * T.n_heavy, tmp4 + (size_t)f * T.n_heavy,
                      search_radius, f, list[i], O, gate_only_state, M, diag, S, PT, lane, w);
  }
#else
  // The warps of the grid walk it with a common stride: `off` = items of the earlier frames, so the first item
  // of a frame goes to the warp after the one that took the last item of the frame before.  (Per-frame striding
  // from warp 0 left most of the grid idle whenever a frame's list is shorter than the grid - the
  // interfacial-water pass has ~1,500 centres per frame for 3,552 warps.)
  (void)work_counter;
  const int stride = gridDim.x * WE_RAD_WARPS;
  const int gw = blockIdx.x * WE_RAD_WARPS + w;
  int off = 0;
  for (int f = 0; f < n_frames; ++f) {
    const int cnt = work.count ? work.count[f] : work.static_count;
    int i = gw - off;
    if (i < 0) i += stride;
    off = (off + cnt) % stride;
    if (i >= cnt) continue;
    const int* list = work.list + (size_t)f * work.stride;
    const WeBox& b = boxes[f];
    const int* cstart = cell_start + (size_t)f * (ncap + 1);
    const float4* sorted = sorted4 + (size_t)f * T.n_heavy;
    const float* fpos = pos + (size_t)f * T.n_atoms * 3;
    const int* cof = cell_of + (size_t)f * T.n_heavy;
    const float4* t4 = tmp4 + (size_t)f * T.n_heavy;
    for (; i < cnt; i += stride)
      we_rad_one<GATED>(T, fpos, b, cstart, cof, sorted, t4, search_radius, f, list[i], O, gate_only_state, M, diag, S, PT, lane, w);
  }
#endif
}

// ============================================================================
// K4a, first stage of the gating pass of the interfacial recipe (recipes/interfacial_solvent.py:47-65):
// which solute united atoms are BURIED, decided per cell group.
//
// Most solute united atoms are buried: at least 20 other solute atoms are closer than the nearest water, so
// no water can be among their 20 nearest candidates and neither exact distances nor a ranking nor a RAD shell
// are ever needed.  Deciding that with the generic warp-per-centre search costs a full candidate scan with
// dependent global loads per atom (~1,500 warp-instructions).  Here one WARP takes a group of 2 x 2 x 1 cells
// that holds solute atoms, stages the group's whole candidate window (the group's cells +- 2 cells; every
// candidate shifted into the group's periodic image, water flag in the sign bit of its index) into shared
// memory ONCE, and classifies every solute atom of the group against the staged window with two sweeps of
// float32 distances from shared memory: the nearest water, then the solute atoms closer than it.  Buried
// atoms get their (empty) result here; the others are appended to the work list of the generic search
// (k_rad), which evaluates them exactly as before.  The test is conservative (1e-3 relative margin in d^2,
// bonded atoms subtracted whether or not they were counted, coincident coordinates never buried), and a
// group whose window or atom list exceeds the staging buffers hands all its atoms to the generic search, so
// the results are those of k_rad<true> in every case.
// ============================================================================
#define WE_GATE_STAGE_CAP 576
#define WE_GATE_CENTRE_CAP 64
#define WE_GATE_ROWS ((WE_GRP_Y + 4) * (WE_GRP_Z + 4))
#define WE_GATE_WARPS 4
static_assert(WE_GATE_ROWS <= 32, "one lane per window row");

__device__ __forceinline__ int we_append_block(bool want, int* counter, int* s_cnt);

__global__ void __launch_bounds__(256)
k_grplist(const WeBox* __restrict__ boxes, const unsigned char* __restrict__ grp_flag, int ncap,
          int* __restrict__ grp_list, int* __restrict__ grp_count) {
  __shared__ int s_cnt[8];
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  const WeBox& b = boxes[f];
  const int ng = ((b.nc[0] + WE_GRP_X - 1) / WE_GRP_X) * ((b.nc[1] + WE_GRP_Y - 1) / WE_GRP_Y) * ((b.nc[2] + WE_GRP_Z - 1) / WE_GRP_Z);
  const bool want = g < ng && grp_flag[(size_t)f * (ncap + 1) + g];
  const int slot = we_append_block(want, &grp_count[f], s_cnt);
  if (want) grp_list[(size_t)f * (ncap + 1) + slot] = g;
}

struct WeGateWarp {
  float4 cand[WE_GATE_STAGE_CAP];  // x, y, z in the group's image; w = position in the cell-sorted array | water << 31
  int4 ri[32];                     // per window row: beg0, len0, beg1, len1
  int4 rj[32];                     // first staged index, packed (ky, kz, yc, zc, x-wrap case), gbeg, gend
  int centres[WE_GATE_CENTRE_CAP]; // staged indices of the group's solute united atoms
  int open[WE_GATE_CENTRE_CAP];    // heavy slots of the ones that are not buried
  int n_centres;
};
struct WeGateSmem {
  WeGateWarp w[WE_GATE_WARPS];
  int pref[130];
};

__global__ void __launch_bounds__(WE_GATE_WARPS * 32)
k_gate(WeTopoDev T, const WeBox* __restrict__ boxes, const int* __restrict__ cell_start,
       const float4* __restrict__ sorted4, const unsigned char* __restrict__ sflag, int ncap, float search_radius,
       const int* __restrict__ grp_list, const int* __restrict__ grp_count, int n_frames, WeRadOut O,
       int* __restrict__ state, int* __restrict__ wl0, int* __restrict__ n0, int wl0_stride, int* work_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WeGateSmem& GS = *reinterpret_cast<WeGateSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  WeGateWarp& G = GS.w[w];
  if (tid == 0) {
    int acc = 0;
    for (int f = 0; f < n_frames; ++f) { GS.pref[f] = acc; acc += grp_count[f]; }
    GS.pref[n_frames] = acc;
  }
  __syncthreads();
  const int total = GS.pref[n_frames];
  constexpr int NWY = WE_GRP_Y + 4;
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(work_counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    int f = 0;
    while (GS.pref[f + 1] <= item) ++f;
    const int g = grp_list[(size_t)f * (ncap + 1) + item - GS.pref[f]];
    const WeBox& b = boxes[f];
    const int n0c = b.nc[0], n1c = b.nc[1], n2c = b.nc[2];
    const int ng0 = (n0c + WE_GRP_X - 1) / WE_GRP_X, ng1 = (n1c + WE_GRP_Y - 1) / WE_GRP_Y;
    const int gx = g % ng0, gy = (g / ng0) % ng1, gz = g / (ng0 * ng1);
    const int* cstart = cell_start + (size_t)f * (ncap + 1);
    const float4* sorted = sorted4 + (size_t)f * T.n_heavy;
    const unsigned char* sfl = sflag + (size_t)f * T.n_heavy;
    const size_t fo = (size_t)f * T.n_heavy;
    const int lo = gx * WE_GRP_X - 2, hi = gx * WE_GRP_X + WE_GRP_X + 1;
    const int xcase = lo < 0 ? 1 : (hi >= n0c ? 2 : 0);
    // ---- one lane per window row: cell ranges, image, the group's own members ---------------------
    int len = 0;
    if (lane < WE_GATE_ROWS) {
      const int iy = lane % NWY, iz = lane / NWY;
      int zc = gz * WE_GRP_Z - 2 + iz, yc = gy * WE_GRP_Y - 2 + iy;
      const int kz = zc < 0 ? -1 : (zc >= n2c ? 1 : 0);
      zc -= kz * n2c;
      const int ky = yc < 0 ? -1 : (yc >= n1c ? 1 : 0);
      yc -= ky * n1c;
      const int rowbase = (zc * n1c + yc) * n0c;
      int beg0, len0, beg1 = 0, len1 = 0;
      if (xcase == 1) {
        beg0 = cstart[rowbase]; len0 = cstart[rowbase + hi + 1] - beg0;
        beg1 = cstart[rowbase + lo + n0c]; len1 = cstart[rowbase + n0c] - beg1;
      } else if (xcase == 2) {
        beg0 = cstart[rowbase + lo]; len0 = cstart[rowbase + n0c] - beg0;
        beg1 = cstart[rowbase]; len1 = cstart[rowbase + hi - n0c + 1] - beg1;
      } else {
        beg0 = cstart[rowbase + lo]; len0 = cstart[rowbase + hi + 1] - beg0;
      }
      int gbeg = 0, gend = 0;
      if (iy >= 2 && iy < 2 + WE_GRP_Y && iz >= 2 && iz < 2 + WE_GRP_Z && ky == 0 && kz == 0) {
        const int x0 = gx * WE_GRP_X, x1 = min(x0 + WE_GRP_X, n0c);
        gbeg = cstart[rowbase + x0]; gend = cstart[rowbase + x1];
      }
      len = len0 + len1;
      G.ri[lane] = make_int4(beg0, len0, beg1, len1);
      G.rj[lane] = make_int4(0, (ky + 1) | ((kz + 1) << 2) | (yc << 4) | (zc << 14), gbeg, gend);
    }
    int incl = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int n_cand = __shfl_sync(0xffffffffu, incl, 31);
    if (lane < WE_GATE_ROWS) G.rj[lane].x = incl - len;
    if (lane == 0) G.n_centres = 0;
    __syncwarp();
    const bool fits = n_cand <= WE_GATE_STAGE_CAP;
    // ---- stage the window, two rows per sweep (half a warp each) -----------------------------------
    const float cw0 = b.box[0] / n0c, cw1 = b.box[1] / n1c, cw2 = b.box[2] / n2c;
    const int hw = lane >> 4, sub = lane & 15;
    for (int r0 = 0; r0 < WE_GATE_ROWS; r0 += 2) {
      const int r = r0 + hw;
      if (r >= WE_GATE_ROWS) continue;
      const int4 ri = G.ri[r];
      const int4 rj = G.rj[r];
      const int ky = (rj.y & 3) - 1, kz = ((rj.y >> 2) & 3) - 1, yc = (rj.y >> 4) & 1023, zc = (rj.y >> 14) & 1023;
      const float sy = ky * b.m[4] + kz * b.m[7], sz = kz * b.m[8], sxr = ky * b.m[3] + kz * b.m[6];
      for (int e = sub; e < ri.y + ri.w; e += 16) {
        const bool second = e >= ri.y;
        const int qi = second ? ri.z + (e - ri.y) : ri.x + e;
        const float4 c = sorted[qi];
        const unsigned fl = sfl[qi];
        float x = c.x, y = c.y, z = c.z;
        if (!b.triclinic) {  // raw coordinates: nearest image to the centre of the run / row
          float refx;
          if (xcase == 1) refx = second ? 0.5f * (lo + n0c + n0c) * cw0 : 0.5f * (hi + 1) * cw0;
          else if (xcase == 2) refx = second ? 0.5f * (hi - n0c + 1) * cw0 : 0.5f * (lo + n0c) * cw0;
          else refx = 0.5f * (lo + hi + 1) * cw0;
          x = __fmaf_rn(-b.box[0], rintf((x - refx) * b.inv[0]), x);
          y = __fmaf_rn(-b.box[1], rintf((y - (yc + 0.5f) * cw1) * b.inv[1]), y);
          z = __fmaf_rn(-b.box[2], rintf((z - (zc + 0.5f) * cw2) * b.inv[2]), z);
        }
        x += sxr + (second ? (xcase == 1 ? -b.m[0] : b.m[0]) : 0.f);
        y += sy; z += sz;
        WE_ASSERT(!fits || (rj.x + e >= 0 && rj.x + e < WE_GATE_STAGE_CAP && qi >= 0 && qi < T.n_heavy));
        if (fits) G.cand[rj.x + e] = make_float4(x, y, z, __int_as_float(qi | ((fl & 1u) ? 0x80000000 : 0)));
        if ((fl & 2u) && qi >= rj.z && qi < rj.w) {
          const int p = atomicAdd(&G.n_centres, 1);
          if (p < WE_GATE_CENTRE_CAP) G.centres[p] = fits ? (rj.x + e) : qi;
        }
      }
    }
    if (fits) {  // pad to a multiple of 32: far away, not water
      const int pad = ((n_cand + 31) & ~31) - n_cand;
      if (lane < pad) G.cand[n_cand + lane] = make_float4(1.0e18f, 1.0e18f, 1.0e18f, __int_as_float(0x7fffffff));
    }
    __syncwarp();
    const int n_centres = G.n_centres;
    int n_open = 0;
    if (!fits || n_centres > WE_GATE_CENTRE_CAP) {
      // an unusually dense window: every solute atom of the group goes to the generic search
      for (int r = 0; r < WE_GATE_ROWS; ++r) {
        const int4 rj = G.rj[r];
        for (int q0 = rj.z; q0 < rj.w; q0 += 32) {
          const int q = q0 + lane;
          const bool want = q < rj.w && (sfl[q] & 2u);
          const unsigned m = __ballot_sync(0xffffffffu, want);
          int base = 0;
          if (lane == 0 && m) base = atomicAdd(&n0[f], __popc(m));
          base = __shfl_sync(0xffffffffu, base, 0);
          if (want) wl0[(size_t)f * wl0_stride + base + __popc(m & ((1u << lane) - 1u))] = T.heavy_slot[__float_as_int(sorted[q].w)];
        }
      }
      continue;
    }
    // ---- every solute atom of the group against the staged window ----------------------------------
    const double r_cover = fmin((double)b.rc_eff, WE_CUTOFF);
    const double rr = fmin((double)search_radius, r_cover);
    for (int ci = 0; ci < n_centres; ++ci) {
      const int self = G.centres[ci];
      const float4 me = G.cand[self];
      const int qme = __float_as_int(me.w) & 0x7fffffff;
      const int n_bonded = sfl[qme] >> 2;
      // sweep A: float32 squared distances of the staged window, kept in registers (this lane's candidates):
      // the nearest water; solute candidates keep their d^2, waters become +inf.  The window is padded to a
      // multiple of 32 with far-away dummies, so there is no bounds test; the atom itself is a solute
      // candidate at distance zero: it is counted and taken off again, and any OTHER zero distance
      // (coincident coordinates) sends the atom to the exact path.
      float wmin = (float)(rr * rr);
      float dsol[WE_GATE_STAGE_CAP / 32];
      int nzero = 0;
#pragma unroll
      for (int k = 0; k < WE_GATE_STAGE_CAP / 32; ++k) {
        dsol[k] = __int_as_float(0x7f800000);
        if (k * 32 < n_cand) {  // warp-uniform
          const float4 c = G.cand[k * 32 + lane];
          const float ux = c.x - me.x, uy = c.y - me.y, uz = c.z - me.z;
          const float d2 = __fmaf_rn(ux, ux, __fmaf_rn(uy, uy, uz * uz));
          const bool wat = __float_as_int(c.w) < 0;
          wmin = fminf(wmin, wat ? d2 : wmin);
          dsol[k] = wat ? dsol[k] : d2;
          nzero += d2 > 0.f ? 0 : 1;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) wmin = fminf(wmin, __shfl_xor_sync(0xffffffffu, wmin, o));
      // sweep B: solute atoms closer than the nearest water by 1e-3 relative in d^2 (the float32 image shift
      // is good to ~1e-4): compares on the registers, one warp reduction
      const float lim = wmin * (1.0f - 1.0e-3f);
      int mine = 0;
#pragma unroll
      for (int k = 0; k < WE_GATE_STAGE_CAP / 32; ++k) mine += dsol[k] < lim ? 1 : 0;
      const int both = __reduce_add_sync(0xffffffffu, mine | (nzero << 16));
      const int closer = (both & 0xffff) - 1;  // without the atom itself
      const bool zero = (both >> 16) != 1;     // a zero distance besides its own
      const bool buried = (closer - n_bonded >= WE_GATE_NBR) && !zero;
      if (lane == 0) {
        const int h = T.heavy_slot[__float_as_int(sorted[qme].w)];
        if (buried) { O.shell_n[fo + h] = 0; O.nn[fo + h] = -1; O.flags[fo + h] = 0; state[fo + h] = 0; }
        else G.open[n_open] = h;
      }
      n_open += buried ? 0 : 1;
    }
    __syncwarp();
    if (n_open) {  // one append per group
      int base = 0;
      if (lane == 0) base = atomicAdd(&n0[f], n_open);
      base = __shfl_sync(0xffffffffu, base, 0);
      WE_ASSERT(n_open <= WE_GATE_CENTRE_CAP && base + n_open <= wl0_stride);
      for (int i = lane; i < n_open; i += 32) wl0[(size_t)f * wl0_stride + base + i] = G.open[i];
    }
    __syncwarp();
  }
}

// ============================================================================
// K3: hydrogen-bond acceptor of donor hydrogens; 8 lanes per donor, 4 donors per warp.
//   analysis/HB.py:96-139 (get_shell_HB_acceptors), :66-93 (get_HB_terms),
//   maths/trig.py:71-100 (candidates = heavy atoms of X's shell within 10 A of D,
//   sorted by distance).  Sequential "rc < current and angle > 90" scan ==
//   argmin over qualifying candidates of (rc, d, index); default = closest.
// Only donors of centres whose HBs the recipe touches (the HB work list) are evaluated.
// ============================================================================
__device__ __forceinline__ bool we_lex_less(double a0, double a1, int a2, double b0, double b1, int b2) {
  if (a0 < b0) return true;
  if (a0 > b0) return false;
  if (a1 < b1) return true;
  if (a1 > b1) return false;
  return a2 < b2;
}

// Persistent over the list of centres whose HBs the recipe evaluates.  Half a warp per centre,
// 8 lanes per donor hydrogen (two donors at a time), lanes = shell members in chunks of 8.
__global__ void __launch_bounds__(256, WE_HB_MINBLOCKS)
k_hb(WeTopoDev T, const float* __restrict__ pos, const WeBox* __restrict__ boxes,
     const float4* __restrict__ tmp4, const int* __restrict__ shell_n,
     const int* __restrict__ shell_idx, const int* __restrict__ flags, WeWork work, int n_frames,
     int* __restrict__ acceptor, WeDiag* diag, int check) {
  __shared__ double s_log2tab[32];
  __shared__ unsigned long long s_exp2tab[32];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < 32) { s_log2tab[tid] = ((const double*)we_d_log2tab)[tid]; s_exp2tab[tid] = we_d_exp2tab[tid]; }
  __syncthreads();
  const WePowTables PT{s_log2tab, s_exp2tab};
  const int sub = lane & 7, grp = (lane >> 3) & 1, hw = lane >> 4;
  // one item space over the frames of the batch (see k_rad); a worker is half a warp
  const int stride = gridDim.x * (blockDim.x >> 5) * 2;
  const int worker = (blockIdx.x * (blockDim.x >> 5) + w) * 2 + hw;
  int off = 0;
  for (int f = 0; f < n_frames; ++f) {
    const int cnt = work.count ? work.count[f] : work.static_count;
    int i = worker - off;
    if (i < 0) i += stride;
    off = (off + cnt) % stride;
    if (!__any_sync(0xffffffffu, i < cnt)) continue;
    const int* list = work.list + (size_t)f * work.stride;
    const WeBox& b = boxes[f];
    const float* fpos = pos + (size_t)f * T.n_atoms * 3;
    const size_t fo = (size_t)f * T.n_heavy;
    const float half[3] = {0.5f * b.box[0], 0.5f * b.box[1], 0.5f * b.box[2]};
    for (; __any_sync(0xffffffffu, i < cnt); i += stride) {
      const bool valid = i < cnt;
      int xs = 0, X = 0, nd = 0, d0 = 0, ns = 0, st = WE_ST_OK;
      if (valid) {
        xs = list[i];
        d0 = T.don_ptr[xs];
        nd = T.don_ptr[xs + 1] - d0;
        st = flags[fo + xs] >> 8;
        ns = (st == WE_ST_OK) ? shell_n[fo + xs] : 0;
        X = T.heavy_idx[xs];
      }
      const size_t ho = fo + xs;
      for (int it = 0; __any_sync(0xffffffffu, 2 * it < nd); ++it) {
        const int dn = 2 * it + grp;
        const bool need = valid && dn < nd;
        const int d = d0 + dn;
        float xx = 0.f, xy = 0.f, xz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, dba = 0.f, dba2 = 0.f, dba2f = 0.f;
        float dq[3] = {0.f, 0.f, 0.f};
        double qD = 0.0;
        const bool live = need && st == WE_ST_OK;
        if (live) {
          const int D = T.don_h[d];
          xx = fpos[3 * X]; xy = fpos[3 * X + 1]; xz = fpos[3 * X + 2];
          dx = fpos[3 * D]; dy = fpos[3 * D + 1]; dz = fpos[3 * D + 2];
          dq[0] = dx; dq[1] = dy; dq[2] = dz;
          if (b.triclinic) we_wrap_tric(dx, dy, dz, b.m, b.wfac, dq);
          dba = we_angle_sep_h(xx, xy, xz, dx, dy, dz, b.box, half);
          dba2 = we_powf2_fast(dba, PT);
          dba2f = we_fmul(dba, dba);
          qD = T.charges[D];
        }
        // group-local best over chunks of 8 shell members
        double bq0 = 1.0e300, bq1 = 1.0e300, bi0 = 1.0e300;  // qualifying (rc, dist); in-range dist
        int bq2 = 0x7fffffff, bi2 = 0x7fffffff;
        bool dup = false;
        const int ns_live = live ? ns : 0;
        const int ns_max = __reduce_max_sync(0xffffffffu, ns_live);
        for (int m0 = 0; m0 < ns_max; m0 += 8) {
          const int m = m0 + sub;
          if (m < ns_live) {
            const int A = shell_idx[ho * WE_MAX_NBR + m];
            const float ax = fpos[3 * A], ay = fpos[3 * A + 1], az = fpos[3 * A + 2];
            dup |= (ax == dx && ay == dy && az == dz);
            float aq0 = ax, aq1 = ay, aq2 = az;
            if (b.triclinic) { const float4 wq = tmp4[fo + T.heavy_slot[A]]; aq0 = wq.x; aq1 = wq.y; aq2 = wq.z; }
            const double dist = we_dist_fast(dq[0], dq[1], dq[2], aq0, aq1, aq2, b);
            if (dist <= WE_CUTOFF) {
              if (we_lex_less(dist, 0.0, A, bi0, 0.0, bi2)) { bi0 = dist; bi2 = A; }
              // angle X-D-A > 90 degrees  <=>  -1 <= cos <= -2^-24 (we_angle_gt_90).  Fast path as in the
              // RAD test: cos = A2 / D2 with A2 = dbc^2 + dba^2 - dac^2, D2 = 2 dbc dba from the squared
              // float32 separations (error <= 5e-6); compared through squares, so no sqrt, powf or
              // division unless cos is within 1e-3 of 0 or of -1.
              const float sbc = we_angle_sep2_h(ax, ay, az, dx, dy, dz, b.box, half);
              const float sac = we_angle_sep2_h(ax, ay, az, xx, xy, xz, b.box, half);
              const float A2 = we_fsub(we_fadd(sbc, dba2f), sac);
              const float D2sq = we_fmul(4.0f, we_fmul(sbc, dba2f));     // (2 dbc dba)^2
              const float A2sq = we_fmul(A2, A2);
              const bool off_zero = A2sq > we_fmul(1.0e-6f, D2sq);        // |cos| > 1e-3
              const bool decided = off_zero && (A2 > 0.f || A2sq < we_fmul(0.998f, D2sq));  // cos > -0.999
              bool gt90 = A2 < 0.f;
              if (!decided || check) {
                const float dbc = we_fsqrt(we_angle_sep2_exact_h(ax, ay, az, dx, dy, dz, b.box, half));
                const float dac = we_fsqrt(we_angle_sep2_exact_h(ax, ay, az, xx, xy, xz, b.box, half));
                const float cs = we_angle_cos_from(dac, dbc, we_powf2_fast(dbc, PT), dba, dba2, PT);
                const bool eg = we_angle_gt_90(cs);
                if (decided && eg != gt90) atomicAdd(&diag->fast_path_mismatches, 1ULL);
                gt90 = eg;
              }
              if (gt90) {  // dist**2 is libm pow(dist, 2.0): only evaluated past the angle test
                const double rc = we_ddiv(we_mul(qD, T.charges[A]), we_pow2(dist));
                if (rc < 99.0) {
                  if (we_lex_less(rc, dist, A, bq0, bq1, bq2)) { bq0 = rc; bq1 = dist; bq2 = A; }
                }
              }
            }
          }
        }
        // reduce within the 8-lane group
        for (int o = 4; o > 0; o >>= 1) {
          const double q0 = __shfl_xor_sync(0xffffffffu, bq0, o), q1 = __shfl_xor_sync(0xffffffffu, bq1, o);
          const int q2 = __shfl_xor_sync(0xffffffffu, bq2, o);
          const double i0d = __shfl_xor_sync(0xffffffffu, bi0, o);
          const int i2 = __shfl_xor_sync(0xffffffffu, bi2, o);
          const bool du = __shfl_xor_sync(0xffffffffu, (int)dup, o) != 0;
          if (we_lex_less(q0, q1, q2, bq0, bq1, bq2)) { bq0 = q0; bq1 = q1; bq2 = q2; }
          if (we_lex_less(i0d, 0.0, i2, bi0, 0.0, bi2)) { bi0 = i0d; bi2 = i2; }
          dup |= du;
        }
        if (need && sub == 0) {
          int out;
          if (st != WE_ST_OK) out = -1 - st;
          else if (dup) out = -1 - WE_ST_DUPLICATE;
          else if (bi2 == 0x7fffffff) out = -1 - WE_ST_NO_NEIGHBOURS;
          else out = (bq2 != 0x7fffffff) ? bq2 : bi2;
          acceptor[(size_t)f * T.n_don + d] = out;
        }
      }
    }
  }
}

// ============================================================================
// Work-list kernels.
// k_seed : pass-0 state (debug mode: every heavy atom is a pass-0 centre)
// k_rad itself marks what the next pass needs (WeMarks: plain byte stores from the warp that
// holds the shell in registers); the list kernels turn the marks into dense work lists:
// k_list1: K4a interfacial gating (recipes/interfacial_solvent.py:25-66): the water members of
//          the shells of gated solute united atoms form the pass-1 list.
// k_list2: shell members of the interfacial waters form the pass-2 list; the centres whose HBs
//          the reference evaluates (analysis/HB.py:142-169) form the HB list.
// ============================================================================
__global__ void k_seed(int n, const int* __restrict__ list, int n_heavy, int* __restrict__ state) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  if (i < n) state[(size_t)f * n_heavy + list[i]] = 1;
}

// Packed uploads (we_options.pos_upload): only the heavy atoms of a frame cross PCIe by DMA.  The light atoms
// the path reads - the donor hydrogens of the centres whose hydrogen bonds are evaluated (k_hb) and every atom
// of a recorded water (k_cov); both sets are inside the HB work list - are copied here, straight from the
// caller's pinned frame buffer (zero-copy reads, `*src_slot` = device alias of frame 0 of the batch) into
// their places in the device's full-layout positions.  Eight lanes per listed centre, one float per lane:
// the two hydrogens of a water are six consecutive floats, i.e. ONE load instruction and one or two 32-byte
// PCIe reads per water (a thread per atom and component-wise loads made three requests per atom).
// <= 32 registers: a block fits beside the persistent k_rad blocks of the pass it overlaps.
__global__ void __launch_bounds__(128, 16)
k_fetch_light(WeTopoDev T, const float* const* __restrict__ src_slot, float* __restrict__ pos_out, WeWork work,
              const int* __restrict__ centre_of_slot, int n_frames, WeDiag* diag) {
  const int f = blockIdx.y;
  if (f >= n_frames) return;
  const int cnt = work.count ? work.count[f] : work.static_count;
  const float* src = *src_slot + (size_t)f * T.n_atoms * 3;
  float* dst = pos_out + (size_t)f * T.n_atoms * 3;
  const int* list = work.list + (size_t)f * work.stride;
  const int sub = threadIdx.x & 7;
  int fetched = 0;
  for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < cnt; i += (gridDim.x * blockDim.x) >> 3) {
    const int xs = list[i];
    const int k = centre_of_slot[xs];
    const int a0 = k >= 0 ? T.frag_ptr[k] : 0, a1 = k >= 0 ? T.frag_ptr[k + 1] : 0;
    const int d0 = T.don_ptr[xs], d1 = T.don_ptr[xs + 1];
    // j-th light atom of the entry: the light atoms of a water centre's molecule (k_cov takes the whole
    // molecule), then the donor hydrogens (k_hb) that are not among them
    auto nth = [&](int j) -> int {
      for (int a = a0; a < a1; ++a) {
        const int at = T.frag_atoms[a];
        if (T.heavy_slot[at] >= 0) continue;
        if (j == 0) return at;
        --j;
      }
      for (int d = d0; d < d1; ++d) {
        const int at = T.don_h[d];
        bool in_frag = false;
        for (int a = a0; a < a1; ++a) in_frag |= (T.frag_atoms[a] == at);
        if (in_frag) continue;
        if (j == 0) return at;
        --j;
      }
      return -1;
    };
    for (int t = sub;; t += 8) {
      const int at = nth(t / 3);
      if (at < 0) break;
      const int c = t % 3;
      WE_ASSERT(at < T.n_atoms && T.heavy_slot[at] < 0);
      dst[3 * (size_t)at + c] = src[3 * (size_t)at + c];
      fetched += (c == 0);
    }
  }
  fetched = __reduce_add_sync(0xffffffffu, fetched);
  if ((threadIdx.x & 31) == 0 && fetched) atomicAdd(&diag->light_fetched, (unsigned long long)fetched);
}

// Block-aggregated append for 256-thread blocks: the threads with `want` get consecutive slots of a
// per-frame list whose length counter sees one atomicAdd per block (same-address atomics serialise).
__device__ __forceinline__ int we_append_block(bool want, int* counter, int* s_cnt /*[9]*/) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, want);
  if (lane == 0) s_cnt[w] = __popc(m);
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const int c = s_cnt[i]; s_cnt[i] = tot; tot += c; }
    const int base = tot ? atomicAdd(counter, tot) : 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s_cnt[i] += base;
  }
  __syncthreads();
  const int slot = s_cnt[w] + __popc(m & ((1u << lane) - 1u));
  __syncthreads();  // s_cnt is reused by the next append
  return want ? slot : -1;
}

// pass-1 list: the waters marked interfacial by pass 0 (state 2 = queued as interfacial water)
// (building the lists in cell order - neighbouring centres on neighbouring warps - was measured: the
// search passes gain 1-3 %, the list kernels and k_hb lose as much)
__global__ void __launch_bounds__(256)
k_list1(WeTopoDev T, const unsigned char* __restrict__ interfacial, int* __restrict__ state,
        int* __restrict__ wl1, int* __restrict__ n1) {
  __shared__ int s_cnt[8];
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  const size_t fo = (size_t)f * T.n_heavy;
  const int h = (c < T.n_centres) ? T.centres[c] : 0;
  const bool want = (c < T.n_centres) && interfacial[fo + h];
  const int slot = we_append_block(want, &n1[f], s_cnt);
  WE_ASSERT(!want || (slot >= 0 && slot < T.n_centres));
  if (want) { wl1[(size_t)f * T.n_centres + slot] = h; state[fo + h] = 2; }
}

// pass-2 list (marked shell members whose shell is not computed yet) and the HB list (marked
// centres that have donor hydrogens)
__global__ void __launch_bounds__(256)
k_list2(WeTopoDev T, const unsigned char* __restrict__ mark_shell,
        const unsigned char* __restrict__ mark_hb, int* __restrict__ state, int* __restrict__ wl2,
        int* __restrict__ n2, int* __restrict__ wlh, int* __restrict__ nh) {
  __shared__ int s_cnt[8];
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  const int f = blockIdx.y;
  const size_t fo = (size_t)f * T.n_heavy;
  const bool valid = h < T.n_heavy;
  const bool w2 = valid && mark_shell && mark_shell[fo + h] && state[fo + h] == 0;
  const bool wh = valid && mark_hb[fo + h] && T.don_ptr[h + 1] != T.don_ptr[h];
  const int p2 = we_append_block(w2, &n2[f], s_cnt);
  WE_ASSERT((!w2 || (p2 >= 0 && p2 < T.n_heavy)));
  if (w2) { wl2[fo + p2] = h; state[fo + h] = 1; }
  const int ph = we_append_block(wh, &nh[f], s_cnt);
  WE_ASSERT(!wh || (ph >= 0 && ph < T.n_heavy));
  if (wh) wlh[fo + ph] = h;
}

// ============================================================================
// K4b: shell labels, HB labels and the count-table update; one warp per water.
//   analysis/shell_labels.py:10-87, analysis/HB_labels.py:299-344 and :37-156,
//   recipes/interfacial_solvent.py:306-335, recipes/bulk_water.py:47-72.
// ============================================================================
__device__ __forceinline__ unsigned long long we_mix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

struct WeRecord { int atom; int nearest; int bin; int pad; };

// K6 for the integer label histograms: per-block staging table in shared memory.  Waters of one
// block that share a label key (the bulk recipe has a few dozen keys for all waters) add their
// counts with shared-memory atomics; the block folds each staged key into the global table once.
#define WE_STAGE_SLOTS 64
struct WeLabelStage {
  unsigned long long hash[WE_STAGE_SLOTS];       // 0 = free
  unsigned long long first_inv[WE_STAGE_SLOTS];
  int gslot[WE_STAGE_SLOTS];                     // slot in the global table (-1: insert failed)
  unsigned int shell_count[WE_STAGE_SLOTS];
  unsigned int don[WE_STAGE_SLOTS][WE_MAX_NBR];
  unsigned int acc[WE_STAGE_SLOTS][WE_MAX_NBR];
};

// One centre (heavy slot xs of frame f) on one warp.
__device__ __forceinline__ void we_label_one(const WeTopoDev& T, int f, int xs,
        const int* __restrict__ shell_n, const int* __restrict__ shell_idx,
        const int* __restrict__ nn_arr, const int* __restrict__ flags,
        const int* __restrict__ acceptor, const unsigned char* __restrict__ interfacial,
        const long long* __restrict__ frame_ids, WeLabelEntry* __restrict__ table,
        unsigned int table_mask, int* __restrict__ rec_count, WeRecord* __restrict__ rec,
        int2* __restrict__ rec_host, int* __restrict__ shell_host, WeLabelStage* stage,
        WeDiag* diag, int lane, int shells_only) {
  const size_t fo = (size_t)f * T.n_heavy;
  const size_t ho = fo + xs;
  const bool bulk = (T.recipe == 1);
  if (!bulk && !interfacial[ho]) return;
  const int X = T.heavy_idx[xs];
  const int frame_id = (int)frame_ids[f];
  const int fl = flags[ho];
  if ((fl >> 8) != WE_ST_OK) { if (lane == 0) we_set_error(diag, fl >> 8, frame_id, X); return; }
  const int ns = shell_n[ho];
  const bool act = lane < ns;
  int A = -1, as = -1;
  if (act) { A = shell_idx[ho * WE_MAX_NBR + lane]; as = T.heavy_slot[A]; }
  const int my_res = T.resname_id[X];
  if (shells_only) {
    // get_interfacial_shells (recipes/interfacial_solvent.py:131-146): the RAD shell of every interfacial
    // water that has a non-like member; no HBs, no labels, no neighbours' shells are evaluated there
    const int nn0 = nn_arr[ho];
    if (nn0 < 0) return;
    int r0 = 0;
    if (lane == 0) {
      r0 = atomicAdd(&rec_count[f], 1);
      WeRecord rr; rr.atom = X; rr.nearest = nn0; rr.bin = -1; rr.pad = 0;
      rec[(size_t)f * T.n_centres + r0] = rr;
      if (rec_host) rec_host[(size_t)f * T.n_centres + r0] = make_int2(X, nn0);
    }
    if (shell_host) {
      r0 = __shfl_sync(0xffffffffu, r0, 0);
      int* o = shell_host + ((size_t)f * T.n_centres + r0) * (WE_MAX_NBR + 1);
      if (lane == 0) o[0] = ns;
      if (act) o[1 + lane] = A;
    }
    return;
  }
  if (bulk) {
    // recipes/bulk_water.py:49-51: only shells made purely of the centre's resname
    const bool other = act && (T.resname_id[A] != my_res);
    if (ns == 0 || __any_sync(0xffffffffu, other)) return;
  }
  // HB step of the reference touches the centre and every shell member
  // (analysis/HB.py:142-169): their neighbour / HB failures propagate.
  int bad = 0;
  if (act) { const int mf = flags[fo + as] >> 8; if (mf != WE_ST_OK) bad = mf; }
  int don_cnt = 0, acc_cnt = 0;
  for (int d = T.don_ptr[xs]; d < T.don_ptr[xs + 1]; ++d) {
    const int a = acceptor[(size_t)f * T.n_don + d];
    if (a < 0) bad = -1 - a;
    if (act && a == A) ++don_cnt;
  }
  if (act && !bad) {
    for (int d = T.don_ptr[as]; d < T.don_ptr[as + 1]; ++d) {
      const int a = acceptor[(size_t)f * T.n_don + d];
      if (a < 0) bad = -1 - a;
      if (a == X) ++acc_cnt;
    }
  }
  const unsigned badm = __ballot_sync(0xffffffffu, bad != 0);
  if (badm) {
    const int code = __shfl_sync(0xffffffffu, bad, __ffs(badm) - 1);
    if (lane == 0) we_set_error(diag, code, frame_id, X);
    return;
  }
  int nn = -1, n_w = 0;
  unsigned int code = 0xffffu;
  if (bulk) {
    if (act) code = (WE_P_PLAIN << 13) | (unsigned)T.resname_id[A];
    n_w = ns;  // recipes/bulk_water.py:62
  } else {
    nn = nn_arr[ho];
    if (nn < 0) return;  // no non-like neighbour: not recorded (interfacial_solvent.py:315)
    const int nn_resid = T.resid[nn];
    bool same = false;
    if (act) {
      const int ra = T.resname_id[A];
      if (A == nn) code = (WE_P_NEAREST << 13) | (unsigned)ra;
      else if (ra != my_res) code = (WE_P_PLAIN << 13) | (unsigned)ra;
      else {
        same = true;
        const int nnn = nn_arr[fo + as];
        int pfx = WE_P_SECOND;
        if (nnn >= 0) pfx = (T.resid[nnn] == nn_resid) ? WE_P_FIRST : WE_P_OTHER;
        code = ((unsigned)pfx << 13) | (unsigned)ra;
      }
    }
    n_w = __popc(__ballot_sync(0xffffffffu, same));
  }
  // ---- canonical key: label codes sorted across the warp (payload: counts) -----
  unsigned int key = act ? ((code << 16) | ((unsigned)min(don_cnt, 255) << 8) | (unsigned)min(acc_cnt, 255)) : 0xffffffffu;
  for (int k = 2; k <= 32; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned int other = __shfl_xor_sync(0xffffffffu, key, j);
      const bool up = ((lane & k) == 0);
      const bool lower = ((lane & j) == 0);
      const unsigned int lo = min(key, other), hi = max(key, other);
      key = (up == lower) ? lo : hi;
    }
  }
  const unsigned int scode = key >> 16;  // sorted: lane p holds p-th smallest label
  const int sdon = (key >> 8) & 0xff, sacc = key & 0xff;
  const bool sact = lane < ns;
  // position of the first lane holding the same label
  const unsigned same_mask = __match_any_sync(0xffffffffu, scode);
  const int first_pos = __ffs(same_mask) - 1;
  // ---- hashes over (nearest resid, nearest resname, N_w, n, sorted codes) -------
  int key_resid, key_resname;
  if (bulk) { key_resid = -1; key_resname = my_res; }
  else { key_resid = T.resid[nn]; key_resname = T.resname_id[nn]; }
  unsigned long long h1 = sact ? we_mix64(((unsigned long long)scode << 8) ^ (unsigned long long)(lane + 1) * 0x100000001B3ULL) : 0ULL;
  unsigned long long h2 = sact ? we_mix64(((unsigned long long)scode << 24) ^ (unsigned long long)(lane + 7) * 0xC2B2AE3D27D4EB4FULL) : 0ULL;
  for (int o = 16; o > 0; o >>= 1) {
    h1 += __shfl_xor_sync(0xffffffffu, h1, o);
    h2 += __shfl_xor_sync(0xffffffffu, h2, o);
  }
  const unsigned long long head = ((unsigned long long)(unsigned)key_resid << 32) ^ ((unsigned long long)(unsigned)key_resname << 12) ^ ((unsigned long long)n_w << 6) ^ (unsigned long long)ns;
  h1 = we_mix64(h1 ^ we_mix64(head));
  h2 = we_mix64(h2 + we_mix64(head ^ 0xA5A5A5A5A5A5A5A5ULL));
  if (h1 == 0ULL) h1 = 1ULL;
  if (h2 == 0ULL) h2 = 1ULL;
  // ---- block-level staging: find / claim the key in shared memory (lane 0) --------------------
  int sidx = -1, claimed = 0;
  if (lane == 0) {
    unsigned int s = (unsigned int)h1 & (WE_STAGE_SLOTS - 1);
    for (int probe = 0; probe < 4; ++probe) {
      const unsigned long long old = atomicCAS(&stage->hash[s], 0ULL, h1);
      if (old == 0ULL) { sidx = (int)s; claimed = 1; break; }
      if (old == h1) { sidx = (int)s; break; }
      s = (s + 1) & (WE_STAGE_SLOTS - 1);
    }
  }
  sidx = __shfl_sync(0xffffffffu, sidx, 0);
  claimed = __shfl_sync(0xffffffffu, claimed, 0);
  const unsigned long long stamp_inv = ~((unsigned long long)frame_ids[f] * (unsigned long long)T.n_atoms + (unsigned long long)X);
  // ---- global insert / find (lane 0), open addressing: for newly staged keys and for waters
  // that found no staging slot --------------------------------------------------------------
  int slot = -1;
  if (sidx < 0 || claimed) {
    int won = 0;
    if (lane == 0) {
      unsigned int s = (unsigned int)h1 & table_mask;
      for (unsigned int probe = 0; probe <= table_mask; ++probe) {
        const unsigned long long old = atomicCAS(&table[s].hash, 0ULL, h1);
        if (old == 0ULL) { slot = (int)s; won = 1; break; }
        if (old == h1) { slot = (int)s; break; }
        s = (s + 1) & table_mask;
      }
      if (slot >= 0) {
        const unsigned long long old2 = atomicCAS(&table[slot].hash2, 0ULL, h2);
        if (old2 != 0ULL && old2 != h2) { we_set_error(diag, 5, frame_id, X); slot = -1; }
      } else {
        we_set_error(diag, 4, frame_id, X);  // table full
      }
      if (claimed) stage->gslot[sidx] = slot;
    }
    slot = __shfl_sync(0xffffffffu, slot, 0);
    won = __shfl_sync(0xffffffffu, won, 0);
    if (slot >= 0 && won) {
      WeLabelEntry* e = table + slot;
      e->labels[lane] = sact ? (unsigned short)scode : (unsigned short)0xffff;
      if (lane == 0) {
        e->nearest_resid = key_resid;
        e->nearest_resname_id = key_resname;
        e->n_labels = ns;
        e->n_w = n_w;
        atomicAdd(&diag->table_entries, 1ULL);
      }
    }
    if (slot < 0 && sidx < 0) return;  // not recorded: the error flag is set
  }
  // ---- counts: shared-memory stage, or straight to the global entry ----------------------------
  if (sidx >= 0) {
    if (lane == 0) { atomicAdd(&stage->shell_count[sidx], 1u); atomicMax(&stage->first_inv[sidx], stamp_inv); }
    if (sact && sdon) atomicAdd(&stage->don[sidx][first_pos], (unsigned int)sdon);
    if (sact && sacc) atomicAdd(&stage->acc[sidx][first_pos], (unsigned int)sacc);
  } else {
    WeLabelEntry* e = table + slot;
    if (lane == 0) { atomicAdd(&e->shell_count, 1ULL); atomicMax(&e->first_seen_inv, stamp_inv); }
    if (sact && sdon) atomicAdd(&e->donates[first_pos], (unsigned int)sdon);
    if (sact && sacc) atomicAdd(&e->accepts[first_pos], (unsigned int)sacc);
  }
  int r = 0;
  if (lane == 0) {
    r = atomicAdd(&rec_count[f], 1);
    WE_ASSERT(r >= 0 && r < T.n_centres);
    WeRecord rr;
    rr.atom = X;
    rr.nearest = bulk ? -1 : nn;
    const int cls = T.centre_class[X];
    rr.bin = bulk ? cls : (T.bin_of_atom[nn] * T.n_classes + cls);
    rr.pad = 0;
    rec[(size_t)f * T.n_centres + r] = rr;
    // compact copy straight into mapped pinned host memory (frame_solvent_indices)
    if (rec_host) rec_host[(size_t)f * T.n_centres + r] = make_int2(X, rr.nearest);
  }
  if (shell_host) {
    // RAD shell of the recorded water (distance order) for get_interfacial_shells:
    // [count, member 0 .. member 24] per record, in mapped pinned host memory
    r = __shfl_sync(0xffffffffu, r, 0);
    int* o = shell_host + ((size_t)f * T.n_centres + r) * (WE_MAX_NBR + 1);
    if (lane == 0) o[0] = ns;
    if (act) o[1 + lane] = A;
  }
}

// Persistent wrapper: grid-stride over the work list of every frame of the batch
// (interfacial recipe: the interfacial waters; bulk recipe: every water centre).
__global__ void __launch_bounds__(256, WE_LABEL_MINBLOCKS)
k_label(WeTopoDev T, const int* __restrict__ shell_n, const int* __restrict__ shell_idx,
        const int* __restrict__ nn_arr, const int* __restrict__ flags,
        const int* __restrict__ acceptor, const unsigned char* __restrict__ interfacial,
        const long long* __restrict__ frame_ids, WeLabelEntry* __restrict__ table,
        unsigned int table_mask, int* __restrict__ rec_count, WeRecord* __restrict__ rec,
        int2* __restrict__ rec_host, int* __restrict__ shell_host, WeWork work, int n_frames, WeDiag* diag,
        int shells_only) {
  __shared__ WeLabelStage stage;
  for (int t = threadIdx.x; t < (int)(sizeof(WeLabelStage) / 4); t += blockDim.x) reinterpret_cast<unsigned int*>(&stage)[t] = 0u;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int stride = gridDim.x * (blockDim.x >> 5);  // one item space over the frames of the batch (see k_rad)
  int off = 0;
  for (int f = 0; f < n_frames; ++f) {
    const int cnt = work.count ? work.count[f] : work.static_count;
    const int* list = work.list + (size_t)f * work.stride;
    int i = blockIdx.x * (blockDim.x >> 5) + w - off;
    if (i < 0) i += stride;
    off = (off + cnt) % stride;
    for (; i < cnt; i += stride)
      we_label_one(T, f, list[i], shell_n, shell_idx, nn_arr, flags, acceptor, interfacial, frame_ids, table,
                   table_mask, rec_count, rec, rec_host, shell_host, &stage, diag, lane, shells_only);
  }
  // ---- fold the staged keys into the global table ----------------------------------------------
  __syncthreads();
  for (int t = threadIdx.x; t < WE_STAGE_SLOTS * (2 * WE_MAX_NBR + 1); t += blockDim.x) {
    const int sl = t / (2 * WE_MAX_NBR + 1), i = t % (2 * WE_MAX_NBR + 1);
    if (stage.hash[sl] == 0ULL || stage.gslot[sl] < 0) continue;
    WeLabelEntry* e = table + stage.gslot[sl];
    if (i == 2 * WE_MAX_NBR) {
      atomicAdd(&e->shell_count, (unsigned long long)stage.shell_count[sl]);
      atomicMax(&e->first_seen_inv, stage.first_inv[sl]);
    } else if (i < WE_MAX_NBR) {
      if (stage.don[sl][i]) atomicAdd(&e->donates[i], stage.don[sl][i]);
    } else {
      if (stage.acc[sl][i - WE_MAX_NBR]) atomicAdd(&e->accepts[i - WE_MAX_NBR], stage.acc[sl][i - WE_MAX_NBR]);
    }
  }
}

// ============================================================================
// K5 + K6: force / torque covariance of each recorded water.
//   recipes/forces_torques.py:28-55, maths/transformations.py:12-79,104-135,
//   entropy/convariances.py:21-84 (sums here; means are taken on the host).
// Principal axes and moments replay numpy.linalg.eig (LAPACK dgeev as OpenBLAS executes it)
// operation by operation (we_eig3.cuh), so every eigenvector carries the sign the reference
// gets and the off-diagonal covariance elements match too.  Degenerate tensors whose LAPACK
// path is not restated (exact zeros, equal moments: never a water) fall back to a cyclic
// Jacobi diagonalisation with a fixed sign convention and are counted in we_diag.eig_fallbacks.
// ============================================================================
__device__ __forceinline__ void we_jacobi3(double a[3][3], double v[3][3], double ev[3]) {
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) v[i][j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
    if (off <= 1e-300 || off <= 1e-22 * diag) break;
    for (int p = 0; p < 2; ++p) {
      for (int q = p + 1; q < 3; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double cth = 1.0 / sqrt(t * t + 1.0);
        const double sth = t * cth;
        const double app = a[p][p], aqq = a[q][q];
        a[p][p] = app - t * apq;
        a[q][q] = aqq + t * apq;
        a[p][q] = a[q][p] = 0.0;
        const int r = 3 - p - q;
        const double arp = a[r][p], arq = a[r][q];
        a[r][p] = a[p][r] = cth * arp - sth * arq;
        a[r][q] = a[q][r] = sth * arp + cth * arq;
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = cth * vkp - sth * vkq;
          v[k][q] = sth * vkp + cth * vkq;
        }
      }
    }
  }
  ev[0] = a[0][0]; ev[1] = a[1][1]; ev[2] = a[2][2];
}

// Fallback for tensors whose dgeev path is not restated: axes = rows by descending eigenvalue,
// largest component of axes 0 and 1 positive, axis 2 = axis 0 x axis 1; am = moments by
// descending magnitude (transformations.py:131-133).
__device__ __noinline__ void we_axes_jacobi(double I[3][3], double ax[3][3], double am[3]) {
  double V[3][3], ev[3];
  we_jacobi3(I, V, ev);
  int o0 = 0, o1 = 1, o2 = 2;
  if (ev[o0] < ev[o1]) { int t = o0; o0 = o1; o1 = t; }
  if (ev[o1] < ev[o2]) { int t = o1; o1 = o2; o2 = t; }
  if (ev[o0] < ev[o1]) { int t = o0; o0 = o1; o1 = t; }
  const int ord[3] = {o0, o1, o2};
  for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) ax[i][k] = V[k][ord[i]];
  for (int i = 0; i < 2; ++i) {
    int big = 0;
    if (fabs(ax[i][1]) > fabs(ax[i][big])) big = 1;
    if (fabs(ax[i][2]) > fabs(ax[i][big])) big = 2;
    if (ax[i][big] < 0.0) { ax[i][0] = -ax[i][0]; ax[i][1] = -ax[i][1]; ax[i][2] = -ax[i][2]; }
  }
  ax[2][0] = ax[0][1] * ax[1][2] - ax[0][2] * ax[1][1];
  ax[2][1] = ax[0][2] * ax[1][0] - ax[0][0] * ax[1][2];
  ax[2][2] = ax[0][0] * ax[1][1] - ax[0][1] * ax[1][0];
  am[0] = ev[0]; am[1] = ev[1]; am[2] = ev[2];
  if (fabs(am[0]) < fabs(am[1])) { double t = am[0]; am[0] = am[1]; am[1] = t; }
  if (fabs(am[1]) < fabs(am[2])) { double t = am[1]; am[1] = am[2]; am[2] = t; }
  if (fabs(am[0]) < fabs(am[1])) { double t = am[0]; am[0] = am[1]; am[1] = t; }
}

// One molecule: rotated mass-weighted force F and inertia-weighted torque tq in its principal-axis
// frame (recipes/forces_torques.py:38-47, maths/transformations.py:12-64,104-135).  pos(k, i) /
// frc(k, i): float32 component i of atom k, mass(k): fp64.
template <class PosF, class FrcF, class MassF>
__device__ __forceinline__ void we_force_torque(int na, PosF pos, FrcF frc, MassF mass, int eig_flavour,
                                                WeDiag* diag, double F[3], double tq[3]) {
  // centre of mass (mass-weighted, fp64, no unwrapping: forces_torques.py:40)
  double mtot = 0.0, cx = 0.0, cy = 0.0, cz = 0.0;
  for (int k = 0; k < na; ++k) {
    const double m = mass(k);
    cx += (double)pos(k, 0) * m; cy += (double)pos(k, 1) * m; cz += (double)pos(k, 2) * m;
    mtot += m;
  }
  cx /= mtot; cy /= mtot; cz /= mtot;
  double I[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int k = 0; k < na; ++k) {
    const double m = mass(k);
    const double x = (double)pos(k, 0) - cx, y = (double)pos(k, 1) - cy, z = (double)pos(k, 2) - cz;
    I[0][0] += m * (y * y + z * z);
    I[1][1] += m * (x * x + z * z);
    I[2][2] += m * (x * x + y * y);
    I[0][1] -= m * x * y;
    I[0][2] -= m * x * z;
    I[1][2] -= m * y * z;
  }
  I[1][0] = I[0][1]; I[2][0] = I[0][2]; I[2][1] = I[1][2];
  double ax[3][3], am[3];
  {
    const double tens[9] = {I[0][0], I[0][1], I[0][2], I[1][0], I[1][1], I[1][2], I[2][0], I[2][1], I[2][2]};
    if (we_principal_axes(tens, ax, am, eig_flavour) != WE_EIG3_OK) {
      atomicAdd(&diag->eig_fallbacks, 1ULL);
      we_axes_jacobi(I, ax, am);
    }
  }
  const double sq[3] = {sqrt(am[0]), sqrt(am[1]), sqrt(am[2])};
  tq[0] = tq[1] = tq[2] = 0.0;
  float fsx = 0.f, fsy = 0.f, fsz = 0.f;
  for (int k = 0; k < na; ++k) {
    const double p[3] = {(double)pos(k, 0) - cx, (double)pos(k, 1) - cy, (double)pos(k, 2) - cz};
    const float g0 = frc(k, 0), g1 = frc(k, 1), g2 = frc(k, 2);
    const double g[3] = {(double)g0, (double)g1, (double)g2};
    double rp[3], rg[3];
    for (int i = 0; i < 3; ++i) {
      rp[i] = ax[i][0] * p[0] + ax[i][1] * p[1] + ax[i][2] * p[2];
      rg[i] = ax[i][0] * g[0] + ax[i][1] * g[1] + ax[i][2] * g[2];
    }
    tq[0] += (rp[1] * rg[2] - rp[2] * rg[1]) / sq[0];
    tq[1] += (rp[2] * rg[0] - rp[0] * rg[2]) / sq[1];
    tq[2] += (rp[0] * rg[1] - rp[1] * rg[0]) / sq[2];
    if (k == 0) { fsx = g0; fsy = g1; fsz = g2; }
    else { fsx = we_fadd(fsx, g0); fsy = we_fadd(fsy, g1); fsz = we_fadd(fsz, g2); }  // float32 sum (transformations.py:48)
  }
  const double msq = sqrt(mtot);
  for (int i = 0; i < 3; ++i) F[i] = (ax[i][0] * (double)fsx + ax[i][1] * (double)fsy + ax[i][2] * (double)fsz) / msq;
}

// Stand-alone form of the above for `recipes.forces_torques.get_forces_torques`: molecules given as
// compact arrays (CSR over atoms); out [n][24] = F(3), tq(3), F outer F / 4 (9), tq outer tq / 4 (9).
__global__ void k_force_torque(const float* __restrict__ pos, const float* __restrict__ frc,
                               const double* __restrict__ masses, const int* __restrict__ mol_ptr, int n_mol,
                               int eig_flavour, WeDiag* diag, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_mol) return;
  const int a0 = mol_ptr[i], a1 = mol_ptr[i + 1];
  double F[3], tq[3];
  we_force_torque(a1 - a0, [&](int k, int c) { return pos[3 * (a0 + k) + c]; },
                  [&](int k, int c) { return frc[3 * (a0 + k) + c]; }, [&](int k) { return masses[a0 + k]; },
                  eig_flavour, diag, F, tq);
  double* o = out + (size_t)i * 24;
  for (int c = 0; c < 3; ++c) { o[c] = F[c]; o[3 + c] = tq[c]; }
  for (int c = 0; c < 3; ++c) for (int d = 0; d < 3; ++d) {
    o[6 + 3 * c + d] = F[c] * F[d] * 0.25;       // get_covariance_matrix(ft, 0.5): outer * 0.5 ** 2
    o[15 + 3 * c + d] = tq[c] * tq[d] * 0.25;
  }
}

// grid = (blocks per frame, frames); the blocks of a frame walk its recorded waters with a
// grid-stride loop (about four warp-iterations each when every water is recorded).
// K6: per-block shared-memory staging of the covariance sums (16 bin slots), flushed with one
// global atomic per value per block; warps whose waters share one bin reduce with shuffles first.
#define WE_COV_SLOTS 16
__global__ void __launch_bounds__(128)
k_cov(WeTopoDev T, const float* __restrict__ pos, const float* __restrict__ frc,
      const int* __restrict__ rec_count, const WeRecord* __restrict__ rec,
      const int* __restrict__ centre_of_slot, double* __restrict__ cov_sum,
      unsigned long long* __restrict__ cov_cnt, double* __restrict__ ft_out,
      const float* __restrict__ fcomp, const int* __restrict__ fbase, int maxfrag, int n_frames,
      int eig_flavour, WeDiag* diag) {
  __shared__ int s_bin[WE_COV_SLOTS];
  __shared__ double s_sum[WE_COV_SLOTS][18];
  __shared__ unsigned long long s_cnt[WE_COV_SLOTS];
  const int lane = threadIdx.x & 31;
  for (int t = threadIdx.x; t < WE_COV_SLOTS * 18; t += blockDim.x) s_sum[t / 18][t % 18] = 0.0;
  if (threadIdx.x < WE_COV_SLOTS) { s_bin[threadIdx.x] = -1; s_cnt[threadIdx.x] = 0ULL; }
  __syncthreads();
  const int f = blockIdx.y;
  {
  const int nrec = rec_count[f];
  for (int r0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); r0 < nrec; r0 += gridDim.x * blockDim.x) {
  const int r = r0 + lane;
  const bool act = r < nrec;
  double val[18];
#pragma unroll
  for (int i = 0; i < 18; ++i) val[i] = 0.0;
  int bin = -1;
  if (act) {
    const WeRecord rr = rec[(size_t)f * T.n_centres + r];
    bin = rr.bin;
    WE_ASSERT(bin >= 0 && bin < T.n_bins);
    const int c = centre_of_slot[T.heavy_slot[rr.atom]];
    const int a0 = T.frag_ptr[c], a1 = T.frag_ptr[c + 1];
    const float* fp = pos + (size_t)f * T.n_atoms * 3;
    // forces: the whole frame (eager upload) or only the recorded molecules, gathered by the
    // host in record order (lazy upload: fcomp[(fbase[f] + r) * maxfrag + k][3])
    const float* ff = fcomp ? fcomp + ((size_t)(fbase[f] + r) * maxfrag) * 3 : frc + (size_t)f * T.n_atoms * 3;
    double F[3], tq[3];
    we_force_torque(a1 - a0,
                    [&](int k, int i) { return fp[3 * T.frag_atoms[a0 + k] + i]; },
                    [&](int k, int i) { return ff[3 * (fcomp ? k : T.frag_atoms[a0 + k]) + i]; },
                    [&](int k) { return T.masses[T.frag_atoms[a0 + k]]; }, eig_flavour, diag, F, tq);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
      val[i * 3 + j] = F[i] * F[j] * 0.25;
      val[9 + i * 3 + j] = tq[i] * tq[j] * 0.25;
    }
    if (ft_out) {
      double* o = ft_out + ((size_t)f * T.n_centres + r) * 6;
      o[0] = F[0]; o[1] = F[1]; o[2] = F[2]; o[3] = tq[0]; o[4] = tq[1]; o[5] = tq[2];
    }
  }
  // ---- K6: accumulate ---------------------------------------------------------------
  const unsigned actm = __ballot_sync(0xffffffffu, act);
  const int lead = __ffs(actm) - 1;
  const int b0 = __shfl_sync(0xffffffffu, bin, lead);
  const bool uniform = __all_sync(0xffffffffu, !act || bin == b0);
  if (uniform) {
    double mine = 0.0;  // after the butterflies lane i keeps the warp total of value i
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      double v = val[i];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == i) mine = v;
    }
    const int slot = b0 & (WE_COV_SLOTS - 1);
    int owner = 0;
    if (lane == 0) { const int old = atomicCAS(&s_bin[slot], -1, b0); owner = (old == -1 || old == b0); }
    owner = __shfl_sync(0xffffffffu, owner, 0);
    if (owner) {
      if (lane < 18) atomicAdd(&s_sum[slot][lane], mine);
      else if (lane == 18) atomicAdd(&s_cnt[slot], (unsigned long long)__popc(actm));
    } else {
      if (lane < 18) atomicAdd(&cov_sum[(size_t)b0 * 18 + lane], mine);
      else if (lane == 18) atomicAdd(&cov_cnt[b0], (unsigned long long)__popc(actm));
    }
  } else if (act) {
#pragma unroll
    for (int i = 0; i < 18; ++i) atomicAdd(&cov_sum[(size_t)bin * 18 + i], val[i]);
    atomicAdd(&cov_cnt[bin], 1ULL);
  }
  }  // records
  }
  __syncthreads();
  for (int t = threadIdx.x; t < WE_COV_SLOTS * 19; t += blockDim.x) {
    const int slot = t / 19, i = t % 19;
    const int bb = s_bin[slot];
    if (bb < 0) continue;
    if (i < 18) { if (s_sum[slot][i] != 0.0) atomicAdd(&cov_sum[(size_t)bb * 18 + i], s_sum[slot][i]); }
    else if (s_cnt[slot]) atomicAdd(&cov_cnt[bb], s_cnt[slot]);
  }
}
