// Orientational entropy per residue on the device, straight from the label count table.
//
// Reference: entropy/orientations.py:298-370 (`Orientations.get_orientational_entropy_from_dict`) on top of
// `WaterOrientCalculator` (:22-238).  For every group of label entries (by nearest resid + resname, or by
// resname alone with the entries of all resids MERGED: counts add, N_w of the first insert stays,
// analysis/HB_labels.py:58,71) the reference visits the shell types in Python's sorted order of the label
// tuples and folds eight running averages.  Seven of them are folded against the ALREADY UPDATED total count
// (:337-360), so the result depends on the visiting order; the device therefore reproduces that order:
//   k_orient_keys    one sort record per entry: group, the entry's labels as ranks in Python's string order
//                    (host-made table over the label alphabet, topology only), sorted ascending
//   k_orient_bitonic global bitonic sort of the records by (group, label-rank tuple) - tuples compare like
//                    Python's (element-wise, a prefix is smaller)
//   k_orient_rows    per run of equal (group, tuple): merged counts -> pD_i, pA_i, Nc_eff, pbias, S_orient
//   k_orient_fold    per group: the running averages in sorted order
// Integer inputs, fp64 arithmetic in the reference's operation order; `log` / `pow` are CUDA's (<= 2 ulp),
// so rows agree with the host classes to ~1e-15 relative, tested at rtol 1e-9.
#pragma once
#include <cuda_runtime.h>
#include "we_kernels.cuh"

struct WeOrientRec {  // 64 bytes
  int resid;
  int resname;
  unsigned short key[25];  // label ranks + 1, ascending, 0 = past the end
  unsigned short n;
  int src;                 // index into the compact entry array, -1 = padding
};
static_assert(sizeof(WeOrientRec) == 64, "sort record must be 64 bytes");

struct WeOrientRow {  // one merged shell type
  double s_hb, n_eff, bias, s_nc, s_nw;
  int n, n_w;
  long long w;  // shell_count; < 0: not the head of a run
};

__device__ __forceinline__ int we_orient_cmp_group(const WeOrientRec& a, const WeOrientRec& b) {
  if (a.resname != b.resname) return a.resname < b.resname ? -1 : 1;
  if (a.resid != b.resid) return a.resid < b.resid ? -1 : 1;
  return 0;
}
__device__ __forceinline__ int we_orient_cmp(const WeOrientRec& a, const WeOrientRec& b) {
  const int g = we_orient_cmp_group(a, b);
  if (g) return g;
#pragma unroll
  for (int k = 0; k < 25; ++k)
    if (a.key[k] != b.key[k]) return a.key[k] < b.key[k] ? -1 : 1;
  if (a.src != b.src) return a.src < b.src ? -1 : 1;  // total order: the sort is deterministic
  return 0;
}
__device__ __forceinline__ bool we_orient_same_type(const WeOrientRec& a, const WeOrientRec& b) {
  if (we_orient_cmp_group(a, b)) return false;
  for (int k = 0; k < 25; ++k) if (a.key[k] != b.key[k]) return false;
  return true;
}

// label_rank[code]: rank of the label string of `code` (prefix << 13 | resname id) in Python's string order.
// by_resname: the group is the nearest residue NAME alone (resid is dropped from the record).
__global__ void k_orient_keys(const WeLabelEntry* __restrict__ e, int n, int n_pad, const unsigned short* __restrict__ label_rank,
                              int by_resname, WeOrientRec* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  WeOrientRec r;
  if (i >= n) {
    r.resid = 0x7fffffff; r.resname = 0x7fffffff; r.n = 0; r.src = -1;
    for (int k = 0; k < 25; ++k) r.key[k] = 0;
    rec[i] = r;
    return;
  }
  const WeLabelEntry& x = e[i];
  r.resid = by_resname ? 0 : x.nearest_resid;
  r.resname = x.nearest_resname_id;
  const int m = min(x.n_labels, 25);
  r.n = (unsigned short)m;
  r.src = i;
  for (int k = 0; k < 25; ++k) r.key[k] = 0;
  for (int k = 0; k < m; ++k) {  // insertion sort of the ranks
    const unsigned short v = (unsigned short)(label_rank[x.labels[k]] + 1);
    int j = k;
    while (j > 0 && r.key[j - 1] > v) { r.key[j] = r.key[j - 1]; --j; }
    r.key[j] = v;
  }
  rec[i] = r;
}

// one compare-exchange step of the bitonic network: partner distance j inside blocks of size k
__global__ void k_orient_bitonic(WeOrientRec* __restrict__ rec, int n_pad, int j, int k) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pad) return;
  const int p = i ^ j;
  if (p <= i) return;
  const WeOrientRec a = rec[i], b = rec[p];
  const bool up = (i & k) == 0;
  const int c = we_orient_cmp(a, b);
  if ((c > 0) == up && c != 0) { rec[i] = b; rec[p] = a; }
}

__device__ __forceinline__ double we_orient_entropy(double n_eff, double p_corr) {
  // R * max(0, ln(n^(3/2) * sqrt(pi) * p / 2))   (entropy/orientations.py:46-66, 209-217)
  double s = 0.0;
  if (n_eff != 0.0) s = log(pow(n_eff, 1.5) * pow(3.141592653589793, 0.5) * p_corr / 2.0);
  return fmax(s, 0.0) * 8.314;
}

// thread per sorted record; the head of a run of equal (group, label tuple) merges the run and evaluates it
__global__ void k_orient_rows(const WeLabelEntry* __restrict__ e, const WeOrientRec* __restrict__ rec, int n,
                              const unsigned short* __restrict__ label_rank, WeOrientRow* __restrict__ rows) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const WeOrientRec me = rec[i];
  WE_ASSERT(me.src >= 0 && me.src < n && me.n <= 25);  // padding records sort behind the n real ones
  if (i > 0 && we_orient_same_type(rec[i - 1], me)) { rows[i].w = -1; return; }
  // ---- merge the run (analysis/HB_labels.py:158-296): counts add, N_w of the earliest insert ----
  const WeLabelEntry& first = e[me.src];
  const int m = me.n;
  unsigned long long don[25], acc[25];
  for (int k = 0; k < 25; ++k) { don[k] = first.donates[k]; acc[k] = first.accepts[k]; }
  unsigned long long w = first.shell_count, seen = first.first_seen_inv;
  int n_w = first.n_w;
  for (int j = i + 1; j < n && we_orient_same_type(me, rec[j]); ++j) {
    WE_ASSERT(rec[j].src >= 0 && rec[j].src < n);
    const WeLabelEntry& x = e[rec[j].src];
    for (int k = 0; k < 25; ++k) { don[k] += x.donates[k]; acc[k] += x.accepts[k]; }
    w += x.shell_count;
    if (x.first_seen_inv > seen) { seen = x.first_seen_inv; n_w = x.n_w; }
  }
  // ---- distinct labels: the codes of an entry are sorted, equal labels are one run whose first position
  // holds the label's donate / accept counts; visit them in rank order (Counter over the sorted tuple) ----
  int pos[25], mult[25], nd = 0;
  for (int k = 0; k < m; ++k) {
    if (k == 0 || first.labels[k] != first.labels[k - 1]) { pos[nd] = k; mult[nd] = 1; ++nd; }
    else ++mult[nd - 1];
  }
  for (int a = 1; a < nd; ++a) {  // insertion sort by rank
    const int p = pos[a], mu = mult[a];
    const unsigned short rk = label_rank[first.labels[p]];
    int j = a;
    while (j > 0 && label_rank[first.labels[pos[j - 1]]] > rk) { pos[j] = pos[j - 1]; mult[j] = mult[j - 1]; --j; }
    pos[j] = p; mult[j] = mu;
  }
  unsigned long long tot_don = 0, tot_acc = 0;
  for (int a = 0; a < nd; ++a) { tot_don += don[pos[a]]; tot_acc += acc[pos[a]]; }
  double n_eff = 0.0, f_sum = 0.0, comp = 0.0;
  for (int a = 0; a < nd; ++a) {
    const double pd = tot_don ? (double)don[pos[a]] / (double)tot_don : 0.0;
    const double pa = tot_acc ? (double)acc[pos[a]] / (double)tot_acc : 0.0;
    double prod = 0.0;
    if (pd != 0.0 && pa != 0.0) {
      const double both = pd + pa;
      const double ppd = pd / both, ppa = pa / both;
      prod = ppd * ppa;
      n_eff += prod / 0.25 * (double)mult[a];
    }
    for (int c = 0; c < mult[a]; ++c) {  // sum() over the per-member list: CPython adds floats compensated
      const double t = f_sum + prod;
      comp += (fabs(f_sum) >= fabs(prod)) ? (f_sum - t) + prod : (prod - t) + f_sum;
      f_sum = t;
    }
  }
  const double bias = m > 0 ? (f_sum + comp) / (double)m : 0.0;
  WeOrientRow r;
  r.s_hb = we_orient_entropy(n_eff, bias);
  r.n_eff = n_eff;
  r.bias = bias;
  r.s_nc = we_orient_entropy((double)m, 0.25);
  r.s_nw = we_orient_entropy((double)n_w, 0.25);
  r.n = m;
  r.n_w = n_w;
  r.w = (long long)w;
  rows[i] = r;
}

// thread per sorted record; the head of a group folds its shell types in order (entropy/orientations.py:318-370)
// out_rows [g][8] = Sor, count, N_c, N_w, Nc_eff, pbias, Sor_Nc, Sor_Nw; out_keys [g][2] = resid, resname id
__global__ void k_orient_fold(const WeLabelEntry* __restrict__ e, const WeOrientRec* __restrict__ rec,
                              const WeOrientRow* __restrict__ rows, int n, int max_groups, int* __restrict__ n_groups,
                              double* __restrict__ out_rows, int* __restrict__ out_keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const WeOrientRec me = rec[i];
  if (i > 0 && we_orient_cmp_group(rec[i - 1], me) == 0) return;
  double s_hb = 0.0, s_nc = 0.0, s_nw = 0.0, n_eff = 0.0, bias = 0.0, n_c = 0.0, n_w = 0.0;
  long long total = 0;
  for (int j = i; j < n && we_orient_cmp_group(me, rec[j]) == 0; ++j) {
    const WeOrientRow r = rows[j];
    if (r.w < 0) continue;
    const double w = (double)r.w, t0 = (double)total;
    s_hb = (r.s_hb * w + s_hb * t0) / (double)(total + r.w);
    total += r.w;
    const double t1 = (double)total, den = (double)(total + r.w);  // upstream folds the rest against the updated total
    s_nc = (r.s_nc * w + s_nc * t1) / den;
    s_nw = (r.s_nw * w + s_nw * t1) / den;
    n_eff = (r.n_eff * w + n_eff * t1) / den;
    bias = (r.bias * w + bias * t1) / den;
    n_c = ((double)r.n * w + n_c * t1) / den;
    n_w = ((double)r.n_w * w + n_w * t1) / den;
  }
  const int g = atomicAdd(n_groups, 1);
  if (g >= max_groups) return;
  WE_ASSERT(me.src >= 0 && me.src < n);
  double* o = out_rows + (size_t)g * 8;
  o[0] = s_hb; o[1] = (double)total; o[2] = n_c; o[3] = n_w; o[4] = n_eff; o[5] = bias; o[6] = s_nc; o[7] = s_nw;
  const WeLabelEntry& x = e[me.src];
  out_keys[2 * g] = x.nearest_resid;
  out_keys[2 * g + 1] = x.nearest_resname_id;
}
