// C ABI + host runtime of waterentropy_b200 (see include/waterentropy_b200.h).
//
// Host side of the hot path: topology upload, per-frame box / cell-grid set-up,
// H2D staging through a ring of four slots on a copy stream, kernel batches on a compute
// stream, asynchronous D2H of the per-frame recorded-water lists.  There is no
// CPU implementation of any kernel in this library: without a CUDA device every
// entry point fails with WE_ERR_CUDA.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <sched.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <deque>
#include <string>
#include <map>
#include <vector>

#include "../../include/waterentropy_b200.h"
#include "we_kernels.cuh"
#include "we_multi_ua.cuh"
#include "we_orient.cuh"

static_assert(sizeof(we_label_entry) == sizeof(WeLabelEntry), "ABI entry layout");

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CK(call)                                                                         \
  do {                                                                                   \
    cudaError_t e_ = (call);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(WE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),   \
                  __FILE__, __LINE__);                                                   \
  } while (0)

extern "C" const char* we_last_error(void) { return g_err.c_str(); }
extern "C" int we_version(void) { return WE_VERSION; }

// ---------------------------------------------------------------------------
struct FrameDebug {
  std::vector<int> shell_n, shell_idx, nn, flags, acceptor, nbr_idx, rec;
  std::vector<double> nbr_dist, ft;
};

// Batches in flight: input staging buffers, record buffers and their events form a ring of WE_NBUF
// slots, so the device always has a few milliseconds of queued work while the host reads results.
#ifndef WE_NBUF_SLOTS
#define WE_NBUF_SLOTS 4
#endif
static const int WE_NBUF = WE_NBUF_SLOTS;

// ---------------------------------------------------------------------------
// Host threads of the packed upload (we_options.pos_upload): gather the heavy atoms of a batch of frames
// out of the caller's [n][n_atoms][3] buffer into pinned staging, [n][n_heavy][3].  A persistent pool; the
// calling thread works too.  (GPU-box host, 16 cores: 96 C4 frames in 4.3 ms = 85 GB/s of source swept.)
// ---------------------------------------------------------------------------
struct WePackPool {
  std::vector<std::thread> th;
  std::mutex m;
  std::condition_variable cv, cv_done;
  bool stop = false;
  unsigned long long gen = 0;
  int busy = 0;
  const float* src = nullptr; float* dst = nullptr; const int* heavy = nullptr;
  int H = 0, n_tasks = 0, slices = 1;
  size_t n3 = 0;
  std::atomic<int> next{0};
  explicit WePackPool(int n_threads) {
    for (int i = 1; i < n_threads; ++i) th.emplace_back([this] { loop(); });
  }
  ~WePackPool() {
    { std::lock_guard<std::mutex> l(m); stop = true; }
    cv.notify_all();
    for (auto& t : th) t.join();
  }
  void work() {
    for (;;) {
      const int t = next.fetch_add(1);
      if (t >= n_tasks) break;
      const int f = t / slices, k = t % slices;
      const int a = (int)((long long)H * k / slices), b = (int)((long long)H * (k + 1) / slices);
      const float* s = src + (size_t)f * n3;
      float* d = dst + (size_t)f * H * 3;
      for (int i = a; i < b; ++i) {
        const size_t j = 3 * (size_t)heavy[i];
        d[3 * (size_t)i] = s[j]; d[3 * (size_t)i + 1] = s[j + 1]; d[3 * (size_t)i + 2] = s[j + 2];
      }
    }
  }
  // A worker spins for `spin_us` after its last task before it sleeps on the condition variable: batches
  // follow each other within ~2 ms while a trajectory streams, and a sleeping thread of a virtual machine
  // takes 100s of microseconds to come back.
  long spin_us = 0;
  std::atomic<unsigned long long> gen_a{0};
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      bool got = false;
      if (spin_us > 0) {
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
          if (gen_a.load(std::memory_order_acquire) != seen) { got = true; break; }
          for (int i = 0; i < 64; ++i) __builtin_ia32_pause();
          if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(spin_us)) break;
        }
      }
      {
        std::unique_lock<std::mutex> l(m);
        if (!got) cv.wait(l, [&] { return stop || gen != seen; });
        if (stop) return;
        seen = gen;
      }
      work();
      { std::lock_guard<std::mutex> l(m); if (--busy == 0) cv_done.notify_one(); }
    }
  }
  void run(const float* src_, float* dst_, const int* heavy_, int H_, size_t n3_, int n_frames) {
    src = src_; dst = dst_; heavy = heavy_; H = H_; n3 = n3_;
    slices = 8;
    n_tasks = n_frames * slices;
    next.store(0);
    { std::lock_guard<std::mutex> l(m); busy = (int)th.size(); ++gen; gen_a.store(gen, std::memory_order_release); }
    cv.notify_all();
    work();
    std::unique_lock<std::mutex> l(m);
    cv_done.wait(l, [&] { return busy == 0; });
  }
};

struct we_ctx {
  int device = 0;
  we_options opt{};
  int N = 0, H = 0, ND = 0, C = 0, S = 0, B = 0, n_classes = 1, recipe = 0;
  std::vector<int> heavy_idx, don_x_atom, don_h_atom;
  std::vector<void*> allocs;      // device allocations
  std::vector<void*> pinned;      // pinned host allocations
  WeTopoDev T{};
  int* d_centre_of_slot = nullptr;
  // accumulators
  double* d_cov_sum = nullptr;
  unsigned long long* d_cov_cnt = nullptr;
  WeLabelEntry* d_table = nullptr;
  WeLabelEntry* d_compact = nullptr;
  size_t compact_cap = 0;
  int* d_compact_n = nullptr;
  unsigned int table_mask = 0;
  WeDiag* h_diag = nullptr;            // pinned mirror refreshed after every batch
  unsigned char* h_snap[4] = {nullptr, nullptr, nullptr, nullptr};  // pinned covariance snapshots
  cudaEvent_t ev_snap[4]{};
  cudaEvent_t ev_cov = nullptr;
  long long snap_counter = 0;
  // label-table sizing: a key is inserted per recorded water at most, so the recorded waters per frame seen
  // so far bound the growth of a batch; until a first batch has been drained only n_centres per frame does
  long long max_rec_frame = -1;           // most recorded waters in one drained frame (-1: none drained yet)
  unsigned long long frames_drained = 0;  // frames whose batch has certainly finished (their keys are in h_diag)
  WeDiag* d_diag = nullptr;
  // workspace
  bool ws_ready = false;
  int chunk = 0, ncap = 0;
  float cell_target = 7.0f;
  // first candidate radius of a centre, per k_rad pass, as a fraction of cell_target (the cells always cover
  // cell_target: a centre that finds fewer than 25 candidates inside the first radius rescans the same cells
  // at the covered radius before it widens the search)
  float r_first_frac[3] = {1.0f, 1.0f, 1.0f};
  float* d_pos[WE_NBUF] = {};
  // packed uploads: pinned staging and device copy of the heavy atoms [chunk][H][3]; per ring slot the device
  // alias of the caller's frame buffer (read by k_fetch_light through a pointer in device memory, so that the
  // launch can live in the batch's CUDA graph)
  float* h_pack[WE_NBUF] = {};
  float* d_pack[WE_NBUF] = {};
  const float** d_src[WE_NBUF] = {};
  const float** h_src = nullptr;
  WePackPool* pool = nullptr;
  int pack_threads = 0;
  unsigned long long h2d_bytes = 0, pack_us = 0;
  struct OrientScratch {  // device scratch of we_orient_entropy (raw cudaMalloc, freed by we_destroy)
    void *rank = nullptr, *rec = nullptr, *rows = nullptr, *out = nullptr, *keys = nullptr, *ng = nullptr;
    size_t rank_cap = 0, rec_cap = 0, rows_cap = 0, out_cap = 0, keys_cap = 0, ng_cap = 0;
  } orient;
  float* d_frc[WE_NBUF] = {};
  WeBox* d_boxes[WE_NBUF] = {};
  long long* d_fids[WE_NBUF] = {};
  WeBox* h_boxes[WE_NBUF] = {};
  long long* h_fids[WE_NBUF] = {};
  int *d_cell_count = nullptr, *d_cell_of = nullptr, *d_rank_of = nullptr;
  float4 *d_tmp4 = nullptr, *d_sorted4 = nullptr;
  int *d_shell_n = nullptr, *d_shell_idx = nullptr, *d_nn = nullptr, *d_flags = nullptr, *d_acceptor = nullptr;
  unsigned char* d_interf = nullptr;
  int *d_state = nullptr, *d_wl1 = nullptr, *d_wl2 = nullptr, *d_n1 = nullptr, *d_n2 = nullptr;
  int *d_hbq = nullptr /* mark bytes: HB centres, then pass-2 shells */, *d_wlh = nullptr, *d_nh = nullptr;
  int* d_zero = nullptr;
  size_t zero_ints = 0;
  // gating pass by cell group (k_gate)
  unsigned char *d_grp_flag = nullptr, *d_sflag = nullptr;
  int *d_grp_list = nullptr, *d_grp_count = nullptr, *d_work_counter = nullptr, *d_wl0 = nullptr, *d_n0 = nullptr;
  int gate_blocks = 0, use_gate = 1;
  bool cov_on_own_stream = false;  // the last batch's k_cov ran on s_cov (zero-copy forces)
  int hb_blocks = 0, label_blocks = 0;
  int *d_all_slots = nullptr;
  int rad_blocks = 0;
  int* d_rec_count[WE_NBUF] = {};
  WeRecord* d_rec[WE_NBUF] = {};
  // lazy force upload
  bool lazy = false;
  int maxfrag = 0;
  std::vector<int> h_centre_of_atom, h_frag_ptr, h_frag_atoms;
  float* h_fcomp[WE_NBUF] = {};
  float* d_fcomp[WE_NBUF] = {};
  int* h_fbase[WE_NBUF] = {};
  int* d_fbase[WE_NBUF] = {};
  bool lazy_pending[WE_NBUF] = {};
  const float* pend_frc[WE_NBUF] = {};
  const float* pend_dp[WE_NBUF] = {};
  cudaEvent_t ev_label[WE_NBUF]{}, ev_f[WE_NBUF]{};
  int* d_nbr_idx = nullptr;
  double* d_nbr_dist = nullptr;
  double* d_ft = nullptr;
  int* h_rec_count[WE_NBUF] = {};
  int2* h_rec[WE_NBUF] = {};  // mapped pinned memory written by k_label
  int* h_shell[WE_NBUF] = {};  // mapped pinned: RAD shells of recorded waters (keep_shells)
  std::vector<std::vector<int>> frame_shells;
  int pending_n[WE_NBUF] = {};
  cudaStream_t s_copy = nullptr, s_comp = nullptr, s_cov = nullptr, s_aux = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_in_ready[WE_NBUF]{}, ev_in_free[WE_NBUF]{}, ev_out_ready[WE_NBUF]{}, ev_p0[WE_NBUF]{};
  int copy_gate = 0, cov_stream = 2, use_graphs = 1, n_sm = 148, rad_blocks_gated = 0;
  // CUDA graphs of the two kernel runs of a batch (before / after the pass-0 event), per ring slot and
  // frame count; host-buffer submissions only (the device pointers of a slot are fixed)
  struct BatchGraph { cudaGraphExec_t exec = nullptr; unsigned long long launches = 0; };
  std::map<std::pair<long long, const void*>, BatchGraph> graphs;  // (slot / size / mode key, device-input pointer or null)
  bool out_pending[WE_NBUF] = {};
  // results on host
  std::vector<std::vector<int>> frame_records;  // per submitted frame: (atom, nearest) pairs sorted by atom
  std::vector<unsigned long long> sort_keys;
  std::vector<FrameDebug> debug;
  unsigned long long launches = 0, frames_done = 0, recorded = 0;
  long long chunk_counter = 0;
  // batches whose input buffers (the caller's host frames / device tensors) may still be read
  struct InFlight { unsigned long long frames_end; int buf; };
  std::deque<InFlight> inflight;
  unsigned long long frames_retired = 0;
  // optional per-kernel CUDA-event timing (we_profile)
  bool prof = false;
  struct ProfRec { int id; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_pending;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[WE_N_KERNELS] = {0};
  unsigned long long prof_n[WE_N_KERNELS] = {0};
  cudaEvent_t span_a = nullptr, span_b = nullptr;
  bool span_open = false;
};

static cudaEvent_t prof_event(we_ctx* c) {
  if (!c->prof_pool.empty()) { cudaEvent_t e = c->prof_pool.back(); c->prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
// record start/stop events around one kernel launch on the compute stream
struct ProfScope {
  we_ctx* c; int id; cudaEvent_t a = nullptr, b = nullptr;
  ProfScope(we_ctx* c_, int id_, cudaStream_t st) : c(c_), id(id_) {
    if (c->prof) { a = prof_event(c); b = prof_event(c); cudaEventRecord(a, st); }
  }
  void stop(cudaStream_t st) { if (c->prof) { cudaEventRecord(b, st); c->prof_pending.push_back({id, a, b}); } }
};
static void prof_resolve(we_ctx* c) {
  for (auto& r : c->prof_pending) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { c->prof_ms[r.id] += ms; c->prof_n[r.id] += 1; }
    c->prof_pool.push_back(r.a);
    c->prof_pool.push_back(r.b);
  }
  c->prof_pending.clear();
}

template <typename Tp>
static int dev_alloc(we_ctx* c, Tp** p, size_t n) {
  void* q = nullptr;
  if (n == 0) n = 1;
  CK(cudaMalloc(&q, n * sizeof(Tp)));
  c->allocs.push_back(q);
  *p = (Tp*)q;
  return 0;
}
template <typename Tp>
static int dev_upload(we_ctx* c, const Tp** p, const std::vector<Tp>& v) {
  Tp* q = nullptr;
  int rc = dev_alloc(c, &q, v.size());
  if (rc) return rc;
  if (!v.empty()) CK(cudaMemcpy(q, v.data(), v.size() * sizeof(Tp), cudaMemcpyHostToDevice));
  *p = q;
  return 0;
}
template <typename Tp>
static int pin_alloc(we_ctx* c, Tp** p, size_t n) {
  void* q = nullptr;
  if (n == 0) n = 1;
  CK(cudaMallocHost(&q, n * sizeof(Tp)));
  c->pinned.push_back(q);
  *p = (Tp*)q;
  return 0;
}

// ---------------------------------------------------------------------------
// per-frame box: MDAnalysis triclinic_vectors (float32) + cell grid
static void triclinic_vectors(const float* dim, float* m) {
  const double lx = dim[0], ly = dim[1], lz = dim[2];
  const double alpha = dim[3], beta = dim[4], gamma = dim[5];
  for (int i = 0; i < 9; ++i) m[i] = 0.f;
  const double cos_a = (alpha == 90.0) ? 0.0 : std::cos(alpha * (M_PI / 180.0));
  const double cos_b = (beta == 90.0) ? 0.0 : std::cos(beta * (M_PI / 180.0));
  double cos_g, sin_g;
  if (gamma == 90.0) { cos_g = 0.0; sin_g = 1.0; }
  else { const double g = gamma * (M_PI / 180.0); cos_g = std::cos(g); sin_g = std::sin(g); }
  m[0] = (float)lx;
  m[3] = (float)(ly * cos_g);
  m[4] = (float)(ly * sin_g);
  m[6] = (float)(lz * cos_b);
  m[7] = (float)(lz * (cos_a - cos_b * cos_g) / sin_g);
  m[8] = (float)std::sqrt(lz * lz - (double)m[6] * (double)m[6] - (double)m[7] * (double)m[7]);
}

static int make_box(const float* dims, float cell_target, int ncap, WeBox* b) {
  memset(b, 0, sizeof(WeBox));
  for (int i = 0; i < 3; ++i) {
    if (!(dims[i] > 0.f) || !std::isfinite(dims[i])) return -1;
    b->box[i] = dims[i];
    b->inv[i] = (float)(1.0 / (double)dims[i]);
  }
  b->triclinic = !(dims[3] == 90.0f && dims[4] == 90.0f && dims[5] == 90.0f);
  if (b->triclinic) triclinic_vectors(dims, b->m);
  else { b->m[0] = dims[0]; b->m[4] = dims[1]; b->m[8] = dims[2]; }
  const double ax = b->m[0], bx = b->m[3], by = b->m[4], cx = b->m[6], cy = b->m[7], cz = b->m[8];
  if (!(ax > 0 && by > 0 && cz > 0)) return -1;
  const double vol = ax * by * cz;
  // perpendicular widths  w_a = V / |b x c| ...
  const double bxc[3] = {by * cz, -bx * cz, bx * cy - by * cx};
  const double cxa[3] = {0.0, cz * ax, -cy * ax};
  const double axb[3] = {0.0, 0.0, ax * by};
  const double w[3] = {vol / std::sqrt(bxc[0] * bxc[0] + bxc[1] * bxc[1] + bxc[2] * bxc[2]),
                       vol / std::sqrt(cxa[0] * cxa[0] + cxa[1] * cxa[1] + cxa[2] * cxa[2]),
                       vol / std::sqrt(axb[0] * axb[0] + axb[1] * axb[1] + axb[2] * axb[2])};
  // cells of edge >= cell_target / WE_CELLS_PER_RADIUS, searched +-WE_CELLS_PER_RADIUS cells
  const double edge = (double)cell_target / WE_CELLS_PER_RADIUS;
  for (int i = 0; i < 3; ++i) b->nc[i] = std::max(1, (int)std::floor(w[i] / edge));
  while ((long long)b->nc[0] * b->nc[1] * b->nc[2] > (long long)ncap) {
    int big = 0;
    if (b->nc[1] > b->nc[big]) big = 1;
    if (b->nc[2] > b->nc[big]) big = 2;
    if (b->nc[big] <= 1) break;
    b->nc[big] -= 1;
  }
  double rc = 1e30, wmin = std::min(w[0], std::min(w[1], w[2]));
  for (int i = 0; i < 3; ++i) {
    const double cw = w[i] / b->nc[i];
    b->sw[i] = std::max(1, (int)std::ceil(((double)cell_target * (1.0 + 1e-5)) / cw));
    // radius guaranteed covered: sw cells on this axis, unless the search covers the whole axis
    if (2 * b->sw[i] + 1 <= b->nc[i]) rc = std::min(rc, b->sw[i] * cw * (1.0 - 1e-5));
    b->wide[i] = std::max(1, (int)std::ceil((WE_CUTOFF * (1.0 + 1e-5)) / cw));
  }
  b->rc_eff = (float)std::min(rc, 1e30);
  b->cg[0] = (float)(by / b->nc[1]); b->cg[1] = (float)(cy / b->nc[2]); b->cg[2] = (float)(cz / b->nc[2]);
  b->cg[3] = (float)(bx / b->nc[1]); b->cg[4] = (float)(cx / b->nc[2]); b->cg[5] = (float)(b->nc[0] / ax);
  // inverse of the lower-triangular cell matrix (row-vector convention r = s . M)
  double hi[9] = {1.0 / ax, 0, 0, -bx / (ax * by), 1.0 / by, 0, (bx * cy - by * cx) / (ax * by * cz), -cy / (by * cz), 1.0 / cz};
  for (int i = 0; i < 9; ++i) { b->hinv[i] = hi[i]; b->hinvf[i] = (float)hi[i]; }
  b->wfac[0] = (double)b->m[3] / (double)b->m[4];
  b->wfac[1] = ((double)b->m[4] * (double)b->m[6] - (double)b->m[3] * (double)b->m[7]) / ((double)b->m[4] * (double)b->m[8]);
  b->wfac[2] = (double)b->m[7] / (double)b->m[8];
  b->pre_ok = b->triclinic ? ((WE_CUTOFF + 0.05) < 0.5 * wmin) : 1;
  return 0;
}

// ---------------------------------------------------------------------------
extern "C" int we_create(const we_topology* t, const we_options* opt, we_ctx** out) {
  if (!t || !out) return fail(WE_ERR_INVALID, "null argument");
  we_ctx* c = new we_ctx();
  if (opt) c->opt = *opt;
  c->device = c->opt.device;
  int ndev = 0;
  cudaError_t e0 = cudaGetDeviceCount(&ndev);
  if (e0 != cudaSuccess || ndev == 0) {
    delete c;
    return fail(WE_ERR_CUDA, "no CUDA device available (%s); waterentropy_b200 has no CPU path",
                e0 == cudaSuccess ? "device count 0" : cudaGetErrorString(e0));
  }
  CK(cudaSetDevice(c->device));
  const int N = t->n_atoms;
  if (N <= 0) { delete c; return fail(WE_ERR_INVALID, "n_atoms must be positive"); }
  c->N = N;
  c->recipe = t->recipe;
  c->n_classes = std::max(1, t->n_classes);
  c->B = (t->recipe == WE_RECIPE_BULK) ? c->n_classes : std::max(1, t->n_pair_bins) * c->n_classes;
  std::vector<int> heavy_slot(N, -1);
  for (int i = 0; i < N; ++i) {
    if (t->masses[i] >= 2.0 && t->masses[i] <= 999.0) { heavy_slot[i] = (int)c->heavy_idx.size(); c->heavy_idx.push_back(i); }
    if (t->resname_id[i] < 0 || t->resname_id[i] >= 8192) { delete c; return fail(WE_ERR_INVALID, "resname_id out of range [0,8192)"); }
  }
  const int H = c->H = (int)c->heavy_idx.size();
  std::vector<std::vector<int>> adj(N);
  for (int k = 0; k < t->n_bonds; ++k) {
    const int a = t->bonds[2 * k], b = t->bonds[2 * k + 1];
    if (a < 0 || b < 0 || a >= N || b >= N || a == b) { delete c; return fail(WE_ERR_INVALID, "bad bond %d", k); }
    adj[a].push_back(b);
    adj[b].push_back(a);
  }
  std::vector<int> excl((size_t)H * WE_MAX_EXCL, -1), don_ptr(H + 1, 0), don_h, don_x;
  for (int h = 0; h < H; ++h) {
    const int i = c->heavy_idx[h];
    std::vector<int> nb = adj[i];
    std::sort(nb.begin(), nb.end());
    nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
    int ne = 0;
    for (int j : nb) {
      if (heavy_slot[j] >= 0) {
        if (ne >= WE_MAX_EXCL) { delete c; return fail(WE_ERR_INVALID, "atom %d has more than %d bonded heavy atoms", i, WE_MAX_EXCL); }
        excl[(size_t)h * WE_MAX_EXCL + ne++] = j;
      }
      // analysis/HB.py:111-115: bond (X, D) with D the higher index, 1.0 < m_D < 1.1, q_D > 0
      if (j > i && t->masses[j] > 1.0 && t->masses[j] < 1.1 && t->charges[j] > 0.0) { don_h.push_back(j); don_x.push_back(h); }
    }
    don_ptr[h + 1] = (int)don_h.size();
  }
  c->ND = (int)don_h.size();
  c->don_h_atom = don_h;
  c->don_x_atom.resize(don_x.size());
  for (size_t k = 0; k < don_x.size(); ++k) c->don_x_atom[k] = c->heavy_idx[don_x[k]];
  std::vector<int> centres, solute, centre_of_slot(H, -1);
  for (int k = 0; k < t->n_centres; ++k) {
    const int a = t->centres[k];
    if (a < 0 || a >= N || heavy_slot[a] < 0) { delete c; return fail(WE_ERR_INVALID, "centre %d is not a heavy atom", a); }
    if (t->centre_class[a] < 0 || t->centre_class[a] >= c->n_classes) { delete c; return fail(WE_ERR_INVALID, "centre %d has no class", a); }
    centre_of_slot[heavy_slot[a]] = k;
    centres.push_back(heavy_slot[a]);
  }
  for (int k = 0; k < t->n_solute_ua; ++k) {
    const int a = t->solute_ua[k];
    if (a < 0 || a >= N || heavy_slot[a] < 0) { delete c; return fail(WE_ERR_INVALID, "solute atom %d is not a heavy atom", a); }
    solute.push_back(heavy_slot[a]);
  }
  c->C = (int)centres.size();
  c->S = (int)solute.size();
  std::vector<int> frag_ptr(t->frag_ptr, t->frag_ptr + c->C + 1);
  std::vector<int> frag_atoms(t->frag_atoms, t->frag_atoms + frag_ptr[c->C]);
  c->h_frag_ptr = frag_ptr;
  c->h_frag_atoms = frag_atoms;
  c->h_centre_of_atom.assign(N, -1);
  for (int k = 0; k < c->C; ++k) {
    c->h_centre_of_atom[t->centres[k]] = k;
    c->maxfrag = std::max(c->maxfrag, frag_ptr[k + 1] - frag_ptr[k]);
  }
  if (t->recipe != WE_RECIPE_BULK) {
    for (int i = 0; i < N; ++i)
      if (heavy_slot[i] >= 0 && t->bin_of_atom[i] >= std::max(1, t->n_pair_bins)) { delete c; return fail(WE_ERR_INVALID, "bin_of_atom out of range"); }
  }
  WeTopoDev& T = c->T;
  T.n_atoms = N; T.n_heavy = H; T.n_don = c->ND; T.n_centres = c->C; T.n_solute_ua = c->S;
  T.n_classes = c->n_classes; T.n_bins = c->B; T.recipe = c->recipe;
  int rc = 0;
#define UP(field, vec) if ((rc = dev_upload(c, &T.field, vec))) { we_destroy(c); return rc; }
  UP(heavy_idx, c->heavy_idx)
  UP(heavy_slot, heavy_slot)
  UP(excl, excl)
  UP(resid, std::vector<int>(t->resid, t->resid + N))
  UP(resname_id, std::vector<int>(t->resname_id, t->resname_id + N))
  UP(type_id, std::vector<int>(t->type_id, t->type_id + N))
  UP(is_water, std::vector<unsigned char>(t->is_water, t->is_water + N))
  UP(charges, std::vector<double>(t->charges, t->charges + N))
  UP(masses, std::vector<double>(t->masses, t->masses + N))
  UP(don_ptr, don_ptr)
  UP(don_h, don_h)
  UP(don_x, don_x)
  UP(solute_ua, solute)
  UP(centres, centres)
  UP(frag_ptr, frag_ptr)
  UP(frag_atoms, frag_atoms)
  UP(bin_of_atom, std::vector<int>(t->bin_of_atom, t->bin_of_atom + N))
  UP(centre_class, std::vector<int>(t->centre_class, t->centre_class + N))
  {
    std::vector<unsigned char> is_solute(std::max(1, H), 0), slot_flag(std::max(1, H), 0);
    for (int hs : solute) is_solute[hs] = 1;
    for (int hs = 0; hs < H; ++hs) {
      int nb = 0;
      for (int e = 0; e < WE_MAX_EXCL; ++e) nb += excl[(size_t)hs * WE_MAX_EXCL + e] >= 0;
      slot_flag[hs] = (unsigned char)((t->is_water[c->heavy_idx[hs]] ? 1 : 0) | (is_solute[hs] << 1) | (nb << 2));
    }
    UP(is_solute, is_solute)
    UP(slot_flag, slot_flag)
  }
#undef UP
  const int* cos_p = nullptr;
  if ((rc = dev_upload(c, &cos_p, centre_of_slot))) { we_destroy(c); return rc; }
  c->d_centre_of_slot = (int*)cos_p;
  // accumulators
  if (c->opt.eig_flavour != WE_EIG_OPENBLAS_AVX512 && c->opt.eig_flavour != WE_EIG_OPENBLAS_AVX2) { delete c; return fail(WE_ERR_INVALID, "eig_flavour must be WE_EIG_OPENBLAS_AVX512 (0) or WE_EIG_OPENBLAS_AVX2 (1)"); }
  if (c->opt.shells_only && (t->recipe != WE_RECIPE_INTERFACIAL || !c->opt.keep_shells || c->opt.keep_debug)) {
    delete c;
    return fail(WE_ERR_INVALID, "shells_only needs the interfacial recipe with keep_shells and without keep_debug");
  }
  const int tl = c->opt.table_log2 > 0 ? c->opt.table_log2 : 17;
  if (tl < 4 || tl > 24) { we_destroy(c); return fail(WE_ERR_INVALID, "table_log2 must be in [4,24]"); }
  c->table_mask = (1u << tl) - 1u;
  if ((rc = dev_alloc(c, &c->d_cov_sum, (size_t)c->B * 18))) { we_destroy(c); return rc; }
  if ((rc = dev_alloc(c, &c->d_cov_cnt, (size_t)c->B))) { we_destroy(c); return rc; }
  {
    void* q = nullptr;
    if (cudaMalloc(&q, ((size_t)c->table_mask + 1) * sizeof(WeLabelEntry)) != cudaSuccess) { we_destroy(c); return fail(WE_ERR_CUDA, "label table allocation failed"); }
    c->d_table = (WeLabelEntry*)q;
  }
  if ((rc = dev_alloc(c, &c->d_compact_n, 1))) { we_destroy(c); return rc; }
  if ((rc = dev_alloc(c, &c->d_diag, 1))) { we_destroy(c); return rc; }
  {
    // the covariance kernel of batch k runs beside the kernels of batch k+1: it must only fill the
    // SM slots they leave free, so its stream has the lowest priority and the compute stream the highest
    int prio_lo = 0, prio_hi = 0;
    CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CK(cudaStreamCreateWithFlags(&c->s_copy, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithPriority(&c->s_comp, cudaStreamNonBlocking, prio_hi));
    CK(cudaStreamCreateWithPriority(&c->s_cov, cudaStreamNonBlocking, prio_lo));
    CK(cudaStreamCreateWithPriority(&c->s_aux, cudaStreamNonBlocking, prio_lo));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  }
  for (int i = 0; i < WE_NBUF; ++i) {
    CK(cudaEventCreateWithFlags(&c->ev_in_ready[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_in_free[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_out_ready[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_p0[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_label[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_f[i], cudaEventDisableTiming));
  }
  if ((rc = pin_alloc(c, &c->h_diag, 1))) { we_destroy(c); return rc; }
  memset(c->h_diag, 0, sizeof(WeDiag));
  if (c->opt.pos_upload < 0 || c->opt.pos_upload > 2) { we_destroy(c); return fail(WE_ERR_INVALID, "pos_upload must be 0 (auto), 1 (whole frames) or 2 (packed)"); }
  if (c->opt.pos_upload == 2 && (t->recipe != WE_RECIPE_INTERFACIAL || c->opt.keep_debug || c->opt.force_upload == 2)) {
    we_destroy(c);
    return fail(WE_ERR_INVALID, "pos_upload = 2 (packed) needs the interfacial recipe, without keep_debug and without force_upload = 2");
  }
  {
    cpu_set_t set;
    CPU_ZERO(&set);
    int avail = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
    c->pack_threads = std::max(1, std::min(16, avail));
    if (const char* g = getenv("WE_PACK_THREADS")) c->pack_threads = std::max(1, std::min(64, atoi(g)));
  }
  if (const char* g = getenv("WE_COPY_GATE")) c->copy_gate = atoi(g);
  if (const char* g = getenv("WE_COV_STREAM")) c->cov_stream = atoi(g);
  if (const char* g = getenv("WE_GRAPHS")) c->use_graphs = atoi(g);
  if (const char* g = getenv("WE_GATE")) c->use_gate = atoi(g);
  if (const char* g = getenv("WE_R_FIRST")) {  // tuning: "f0,f1,f2" fractions of the cell radius per pass
    float a = 1.f, b = 1.f, d = 1.f;
    const int k = sscanf(g, "%f,%f,%f", &a, &b, &d);
    if (k >= 1) c->r_first_frac[0] = a;
    if (k >= 2) c->r_first_frac[1] = b;
    if (k >= 3) c->r_first_frac[2] = d;
    for (float& v : c->r_first_frac) v = std::min(1.0f, std::max(0.3f, v));
  }
  rc = we_reset(c);
  if (rc) { we_destroy(c); return rc; }
  *out = c;
  return WE_OK;
}

extern "C" int we_reset(we_ctx* c) {
  if (!c) return fail(WE_ERR_INVALID, "null context");
  CK(cudaSetDevice(c->device));
  CK(cudaDeviceSynchronize());
  CK(cudaMemset(c->d_cov_sum, 0, (size_t)c->B * 18 * sizeof(double)));
  CK(cudaMemset(c->d_cov_cnt, 0, (size_t)c->B * sizeof(unsigned long long)));
  CK(cudaMemset(c->d_table, 0, ((size_t)c->table_mask + 1) * sizeof(WeLabelEntry)));
  CK(cudaMemset(c->d_diag, 0, sizeof(WeDiag)));
  c->frame_records.clear();
  c->frame_shells.clear();
  c->debug.clear();
  c->frames_done = 0;
  c->recorded = 0;
  c->inflight.clear();
  c->frames_retired = 0;
  for (int i = 0; i < WE_NBUF; ++i) { c->out_pending[i] = false; c->lazy_pending[i] = false; }
  c->max_rec_frame = -1;
  c->frames_drained = 0;
  if (c->h_diag) memset(c->h_diag, 0, sizeof(WeDiag));
  return WE_OK;
}

static void drop_graphs(we_ctx* c);
extern "C" void we_destroy(we_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (void* p : c->allocs) cudaFree(p);
  if (c->d_table) cudaFree(c->d_table);
  if (c->d_compact) cudaFree(c->d_compact);
  for (void* p : c->pinned) cudaFreeHost(p);
  drop_graphs(c);
  if (c->s_copy) cudaStreamDestroy(c->s_copy);
  if (c->s_comp) cudaStreamDestroy(c->s_comp);
  delete c->pool;
  c->pool = nullptr;
  for (void* q : {c->orient.rank, c->orient.rec, c->orient.rows, c->orient.out, c->orient.keys, c->orient.ng})
    if (q) cudaFree(q);
  c->orient = we_ctx::OrientScratch();
  if (c->s_cov) cudaStreamDestroy(c->s_cov);
  if (c->s_aux) cudaStreamDestroy(c->s_aux);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_cov) cudaEventDestroy(c->ev_cov);
  for (int i = 0; i < WE_NBUF; ++i) {
    if (c->ev_in_ready[i]) cudaEventDestroy(c->ev_in_ready[i]);
    if (c->ev_in_free[i]) cudaEventDestroy(c->ev_in_free[i]);
    if (c->ev_out_ready[i]) cudaEventDestroy(c->ev_out_ready[i]);
    if (c->ev_p0[i]) cudaEventDestroy(c->ev_p0[i]);
    if (c->ev_label[i]) cudaEventDestroy(c->ev_label[i]);
    if (c->ev_f[i]) cudaEventDestroy(c->ev_f[i]);
    if (i < 4 && c->ev_snap[i]) cudaEventDestroy(c->ev_snap[i]);
  }
  delete c;
}

// frames per kernel batch
static int batch_frames(const we_ctx* c) {
  int chunk = c->opt.chunk_frames;
  if (chunk <= 0) {
    // ~3.5 M heavy-atom centres per batch (32 frames on C4): the per-batch fixed cost (a dozen launches, the
    // tails of the persistent kernels) is ~0.14 ms on C4, 16 % of an 8-frame batch, 5 % of a 24-frame one;
    // 32-frame batches measured +3.5 % device-resident over 24, end to end unchanged; 48 hurts the overlap
    // of uploads and kernels (two batches per 96-frame submission)
    long long t = 3500000LL / std::max(1, c->H);
    if (c->opt.keep_shells) t /= 4;  // shells of the recorded waters are staged in pinned memory
    chunk = (int)std::max(1LL, std::min(128LL, t));
    if (chunk >= 8) chunk &= ~7;
  }
  return std::min(chunk, 128);  // per-batch frame tables in shared memory (k_gate, k_rad) hold 128 frames
}
extern "C" int we_batch_frames(we_ctx* c) { return c ? batch_frames(c) : fail(WE_ERR_INVALID, "null context"); }

static int ensure_workspace(we_ctx* c, const float* dims0, bool need_input_buffers) {
  if (c->ws_ready) return 0;
  const int H = c->H, N = c->N;
  const int chunk = batch_frames(c);
  c->chunk = chunk;
  c->cell_target = c->opt.search_radius > 0.f ? c->opt.search_radius : 6.6f;
  // cell capacity from the largest box expected
  WeBox b0;
  float d[6];
  memcpy(d, dims0, sizeof(d));
  const float grow = 1.25f;
  for (int i = 0; i < 3; ++i) d[i] = c->opt.max_box > 0.f ? std::max(c->opt.max_box, d[i]) : d[i] * grow;
  if (make_box(d, c->cell_target, 1 << 30, &b0)) return fail(WE_ERR_INVALID, "invalid box dimensions");
  long long ncell = (long long)b0.nc[0] * b0.nc[1] * b0.nc[2];
  c->ncap = (int)std::min<long long>(std::max<long long>(ncell, 8), 1LL << 26);
  c->ncap = ((c->ncap + 1 + 7) & ~7) - 1;  // per-frame stride ncap + 1 is a multiple of 8 ints (int4 scans)
  int rc = 0;
  const size_t F = (size_t)chunk;
#define AL(p, n) if ((rc = dev_alloc(c, &c->p, (n)))) return rc;
  // optional (force_upload = 2): upload only the forces of the recorded waters, gathered by the
  // host one batch later.  Halves PCIe traffic, but puts host-side gathers on the critical path;
  // measured on this pool it is within 5 % of the whole-frame upload and much more sensitive to
  // a busy host, so whole frames (copy engine only, no CPU touch) are the default.
  c->lazy = need_input_buffers && !c->opt.keep_debug && c->maxfrag > 0 && c->maxfrag <= 8 &&
            c->opt.force_upload == 2 && !c->opt.shells_only;
  if (need_input_buffers) {
    for (int i = 0; i < WE_NBUF; ++i) { AL(d_pos[i], F * N * 3) }  // d_frc: allocated on first whole-frame upload
  }
  if (c->lazy) {
    const size_t cap = F * std::max(1, c->C) * (size_t)c->maxfrag * 3;
    for (int i = 0; i < WE_NBUF; ++i) {
      AL(d_fcomp[i], cap) AL(d_fbase[i], F)
      if ((rc = pin_alloc(c, &c->h_fcomp[i], cap))) return rc;
      if ((rc = pin_alloc(c, &c->h_fbase[i], F))) return rc;
    }
  }
  for (int i = 0; i < WE_NBUF; ++i) {
    AL(d_boxes[i], F) AL(d_fids[i], F)
    if ((rc = pin_alloc(c, &c->h_boxes[i], F))) return rc;
    if ((rc = pin_alloc(c, &c->h_fids[i], F))) return rc;
    if ((rc = pin_alloc(c, &c->h_rec_count[i], F))) return rc;
    if (c->opt.keep_records || c->lazy || c->opt.keep_shells) { if ((rc = pin_alloc(c, &c->h_rec[i], F * std::max(1, c->C)))) return rc; }
    if (c->opt.keep_shells) { if ((rc = pin_alloc(c, &c->h_shell[i], F * std::max(1, c->C) * (WE_MAX_NBR + 1)))) return rc; }
  }
  AL(d_cell_of, F * H) AL(d_rank_of, F * H) AL(d_tmp4, F * H) AL(d_sorted4, F * H)
  AL(d_shell_n, F * H) AL(d_shell_idx, F * H * WE_MAX_NBR) AL(d_nn, F * H) AL(d_flags, F * H)
  AL(d_acceptor, F * std::max(1, c->ND))
  for (int i = 0; i < WE_NBUF; ++i) { AL(d_rec_count[i], F) AL(d_rec[i], F * std::max(1, c->C)) }
  AL(d_wlh, F * H) AL(d_wl1, F * std::max(1, c->C)) AL(d_wl2, F * H)
  {
    // everything that must be zero at the start of a batch lives in one block: one memset per batch
    const size_t n_cells = F * ((size_t)c->ncap + 1), n_fh = F * (size_t)H;
    c->zero_ints = n_cells + 2 * n_fh + 5 * F + 8 + (n_fh + 3) / 4 + (n_cells + 3) / 4;
    AL(d_zero, c->zero_ints)
    int* p = c->d_zero;
    c->d_cell_count = p; p += n_cells;
    c->d_state = p; p += n_fh;
    c->d_hbq = p; p += n_fh;
    c->d_nh = p; p += F;
    c->d_n1 = p; p += F;
    c->d_n2 = p; p += F;
    c->d_grp_count = p; p += F;
    c->d_n0 = p; p += F;
    c->d_work_counter = p; p += 8;
    c->d_interf = reinterpret_cast<unsigned char*>(p); p += (n_fh + 3) / 4;
    c->d_grp_flag = reinterpret_cast<unsigned char*>(p);
    AL(d_grp_list, n_cells)
    AL(d_sflag, n_fh)
    AL(d_wl0, F * (size_t)std::max(1, c->S))
  }
  if (c->opt.keep_debug) { AL(d_nbr_idx, F * H * WE_MAX_NBR) AL(d_nbr_dist, F * H * WE_MAX_NBR) AL(d_ft, F * std::max(1, c->C) * 6) }
#undef AL
  CK(cudaFuncSetAttribute(k_rad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WeRadSmem)));
  CK(cudaFuncSetAttribute(k_rad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WeRadSmem)));
  CK(cudaFuncSetAttribute(k_gate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WeGateSmem)));
  {
    std::vector<int> all(H);
    for (int i = 0; i < H; ++i) all[i] = i;
    const int* p = nullptr;
    if ((rc = dev_upload(c, &p, all))) return rc;
    c->d_all_slots = (int*)p;
    int per_sm = 0, sms = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rad<false>, WE_RAD_WARPS * 32, sizeof(WeRadSmem)));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    c->n_sm = std::max(1, sms);
    c->rad_blocks = std::max(1, per_sm) * std::max(1, sms);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rad<true>, WE_RAD_WARPS * 32, sizeof(WeRadSmem)));
    c->rad_blocks_gated = std::max(1, per_sm) * std::max(1, sms);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gate, WE_GATE_WARPS * 32, sizeof(WeGateSmem)));
    c->gate_blocks = std::max(1, per_sm) * std::max(1, sms);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_hb, 256, 0));
    c->hb_blocks = std::max(1, per_sm) * std::max(1, sms);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_label, 256, 0));
    c->label_blocks = std::max(1, per_sm) * std::max(1, sms);
  }
  c->ws_ready = true;
  return 0;
}

// move one finished chunk's recorded-water lists to host vectors
static int drain(we_ctx* c, int buf) {
  if (!c->out_pending[buf]) return 0;
  CK(cudaEventSynchronize(c->ev_out_ready[buf]));
  const int n = c->pending_n[buf];
  c->frames_drained += (unsigned long long)n;
  for (int f = 0; f < n; ++f) {
    const int cnt = c->h_rec_count[buf][f];
    c->recorded += (unsigned long long)cnt;
    c->max_rec_frame = std::max(c->max_rec_frame, (long long)cnt);
    if ((c->opt.keep_records || c->opt.keep_shells)) {
      // ascending atom index (recipes/interfacial_solvent.py:66,303): sort (atom, position) packed in
      // 64-bit keys - this runs on the submitting thread, between two batch submissions
      const int2* r = c->h_rec[buf] + (size_t)f * std::max(1, c->C);
      std::vector<unsigned long long>& order = c->sort_keys;
      order.resize((size_t)cnt);
      for (int k = 0; k < cnt; ++k) order[k] = ((unsigned long long)(unsigned int)r[k].x << 32) | (unsigned int)k;
      std::sort(order.begin(), order.end());
      std::vector<int> flat((size_t)cnt * 2);
      for (int k = 0; k < cnt; ++k) { const int2 v = r[order[k] & 0xffffffffULL]; flat[2 * k] = v.x; flat[2 * k + 1] = v.y; }
      c->frame_records.push_back(std::move(flat));
      if (c->opt.keep_shells) {
        const int W = WE_MAX_NBR + 1;
        const int* sh = c->h_shell[buf] + (size_t)f * std::max(1, c->C) * W;
        std::vector<int> fs((size_t)cnt * W);
        for (int k = 0; k < cnt; ++k) memcpy(fs.data() + (size_t)k * W, sh + (size_t)(order[k] & 0xffffffffULL) * W, W * sizeof(int));
        c->frame_shells.push_back(std::move(fs));
      }
    }
  }
  c->out_pending[buf] = false;
  return 0;
}

// ---- label table growth ---------------------------------------------------------------------
__global__ void k_rehash(const WeLabelEntry* __restrict__ old_t, unsigned int old_n,
                         WeLabelEntry* __restrict__ new_t, unsigned int new_mask) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= old_n) return;
  const unsigned long long h = old_t[i].hash;
  if (h == 0ULL) return;
  unsigned int s = (unsigned int)h & new_mask;
  while (atomicCAS(&new_t[s].hash, 0ULL, h) != 0ULL) s = (s + 1) & new_mask;
  const uint4* src = reinterpret_cast<const uint4*>(old_t + i);
  uint4* dst = reinterpret_cast<uint4*>(new_t + s);
#pragma unroll
  for (int k = 0; k < 32; ++k) dst[k] = src[k];
}

// Keep the open-addressing table below ~50 % load.  The table cannot grow while a batch runs, so before a
// batch of `n` frames is queued it must hold every key that batch - and the batches still in flight, whose
// keys the pinned mirror h_diag does not show yet - can add.  A batch adds at most one key per recorded
// water: 2 x (most recorded waters seen in one frame) per frame once a batch has been drained, n_centres per
// frame before that (submit_impl keeps the very first batch short for this reason).
static int grow_table_to(we_ctx* c, unsigned long long need);
static unsigned long long keys_per_frame_bound(const we_ctx* c) {
  // bulk recipe: a key is (water resname, shell size) - a few dozen in total, whatever the system
  if (c->recipe == WE_RECIPE_BULK) return 64;
  if (c->max_rec_frame < 0) return (unsigned long long)std::max(1, c->C);
  return std::min<unsigned long long>((unsigned long long)std::max(1, c->C), 2ULL * (unsigned long long)c->max_rec_frame + 64);
}
static int maybe_grow_table(we_ctx* c, int n) {
  const unsigned long long entries = c->h_diag->table_entries;
  const unsigned long long open_frames = c->frames_done - c->frames_drained + (unsigned long long)n;
  return grow_table_to(c, entries + keys_per_frame_bound(c) * open_frames + 1024);
}
static int grow_table_to(we_ctx* c, unsigned long long need) {
  unsigned long long cap = (unsigned long long)c->table_mask + 1;
  if (need * 2 <= cap) return 0;
  unsigned long long ncap = cap;
  while (need * 2 > ncap) ncap <<= 1;
  if (ncap > (1ULL << 24)) return fail(WE_ERR_TABLE_FULL, "label count table would exceed 2^24 entries");
  CK(cudaStreamSynchronize(c->s_comp));
  CK(cudaStreamSynchronize(c->s_cov));
  WeLabelEntry* nt = nullptr;
  CK(cudaMalloc((void**)&nt, ncap * sizeof(WeLabelEntry)));
  CK(cudaMemsetAsync(nt, 0, ncap * sizeof(WeLabelEntry), c->s_comp));
  k_rehash<<<(unsigned int)((cap + 255) / 256), 256, 0, c->s_comp>>>(c->d_table, (unsigned int)cap, nt, (unsigned int)(ncap - 1));
  c->launches++;
  CK(cudaStreamSynchronize(c->s_comp));
  CK(cudaFree(c->d_table));
  drop_graphs(c);  // k_label's table pointer and mask are baked into the captured launches
  c->d_table = nt;
  c->table_mask = (unsigned int)(ncap - 1);
  return 0;
}

// Lazy force upload, second half of a batch: wait for its label kernel, gather the forces of the
// recorded waters' atoms from the caller's host frames (record order), upload them compactly and
// launch the covariance kernel.  The recorded-water lists are moved to host vectors here too.
static int finish_lazy(we_ctx* c, int buf) {
  if (!c->lazy_pending[buf]) return 0;
  CK(cudaEventSynchronize(c->ev_label[buf]));
  const int n = c->pending_n[buf];
  const int C = std::max(1, c->C), N = c->N, mf = c->maxfrag;
  size_t total = 0;
  for (int f = 0; f < n; ++f) { c->h_fbase[buf][f] = (int)total; total += (size_t)c->h_rec_count[buf][f]; }
  float* out = c->h_fcomp[buf];
  for (int f = 0; f < n; ++f) {
    const int cnt = c->h_rec_count[buf][f];
    const int2* r = c->h_rec[buf] + (size_t)f * C;
    const float* ff = c->pend_frc[buf] + (size_t)f * N * 3;
    float* o = out + (size_t)c->h_fbase[buf][f] * mf * 3;
    for (int k = 0; k < cnt; ++k) {
      const int ce = c->h_centre_of_atom[r[k].x];
      const int a0 = c->h_frag_ptr[ce], a1 = c->h_frag_ptr[ce + 1];
      for (int a = a0; a < a1; ++a) {
        const float* src = ff + (size_t)c->h_frag_atoms[a] * 3;
        float* dst = o + ((size_t)k * mf + (a - a0)) * 3;
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2];
      }
    }
  }
  if (total) CK(cudaMemcpyAsync(c->d_fcomp[buf], out, total * mf * 3 * sizeof(float), cudaMemcpyHostToDevice, c->s_copy));
  CK(cudaMemcpyAsync(c->d_fbase[buf], c->h_fbase[buf], (size_t)n * sizeof(int), cudaMemcpyHostToDevice, c->s_copy));
  CK(cudaEventRecord(c->ev_f[buf], c->s_copy));
  // own stream: the covariance kernel of batch k must not queue behind the kernels of batch k+1
  // (its completion frees this slot's input buffers for the H2D copy of batch k+2)
  cudaStream_t st = c->s_cov;
  CK(cudaStreamWaitEvent(st, c->ev_label[buf], 0));
  CK(cudaStreamWaitEvent(st, c->ev_f[buf], 0));
  {
    ProfScope ps(c, WE_K_COV, st);
    k_cov<<<dim3((c->C + 511) / 512, n), 128, 0, st>>>(c->T, c->pend_dp[buf], nullptr, c->d_rec_count[buf], c->d_rec[buf], c->d_centre_of_slot,
                                       c->d_cov_sum, c->d_cov_cnt, nullptr, c->d_fcomp[buf], c->d_fbase[buf], mf, n, c->opt.eig_flavour, c->d_diag);
    ps.stop(st);
    c->launches += 1;
  }
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ev_in_free[buf], st));
  CK(cudaEventRecord(c->ev_out_ready[buf], st));
  c->lazy_pending[buf] = false;
  c->out_pending[buf] = true;
  return 0;
}

static void drop_graphs(we_ctx* c) {
  for (auto& kv : c->graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  c->graphs.clear();
}

// One run of kernels of a batch as a CUDA graph: the first time a (slot, frames, part) key is seen the
// launches between begin() and end() are captured from the compute stream and instantiated; afterwards
// begin() replays the graph and tells the caller to skip the launches.  While host->device uploads are
// in flight every individual launch costs 20-40 us extra (see the upload gating above); a graph is one
// submission for the whole run.
struct GraphRegion {
  we_ctx* c; cudaStream_t st; std::pair<long long, const void*> key; bool on; bool capturing = false; unsigned long long l0 = 0;
  GraphRegion(we_ctx* c_, cudaStream_t st_, long long key_, const void* ptr_, bool on_) : c(c_), st(st_), key(key_, ptr_), on(on_) {}
  // returns 1: replayed (skip the launches), 0: run the launches (captured if graphs are on), <0: error
  int begin() {
    if (!on) return 0;
    auto it = c->graphs.find(key);
    if (it != c->graphs.end()) {
      if (cudaGraphLaunch(it->second.exec, st) != cudaSuccess) return fail(WE_ERR_CUDA, "cudaGraphLaunch failed");
      c->launches += it->second.launches;
      return 1;
    }
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) { cudaGetLastError(); on = false; return 0; }
    capturing = true;
    l0 = c->launches;
    return 0;
  }
  int end() {
    if (!capturing) return 0;
    capturing = false;
    cudaGraph_t g = nullptr;
    if (cudaStreamEndCapture(st, &g) != cudaSuccess || !g) return fail(WE_ERR_CUDA, "stream capture failed: %s", cudaGetErrorString(cudaGetLastError()));
    we_ctx::BatchGraph bg;
    bg.launches = c->launches - l0;
    const cudaError_t e = cudaGraphInstantiate(&bg.exec, g, 0);
    cudaGraphDestroy(g);
    if (e != cudaSuccess) return fail(WE_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
    c->graphs[key] = bg;
    if (cudaGraphLaunch(bg.exec, st) != cudaSuccess) return fail(WE_ERR_CUDA, "cudaGraphLaunch failed");
    return 0;
  }
};

// the slot chunk_counter % WE_NBUF holds the oldest batch in flight
static int finish_lazy_all(we_ctx* c) {
  for (int i = 0; i < WE_NBUF; ++i) {
    const int rc = finish_lazy(c, (int)((c->chunk_counter + i) % WE_NBUF));
    if (rc) return rc;
  }
  return 0;
}

static int submit_impl(we_ctx* c, const float* pos, const float* frc, bool on_device, const float* dims,
                       const int64_t* frame_ids, int n_frames) {
  if (!c || !pos || !dims || !frame_ids) return fail(WE_ERR_INVALID, "null argument");
  if (!frc && !c->opt.shells_only) return fail(WE_ERR_INVALID, "null force buffer (only a shells_only context runs without forces)");
  if (n_frames <= 0) return WE_OK;
  CK(cudaSetDevice(c->device));
  int rc = ensure_workspace(c, dims, !on_device);
  if (rc) return rc;
  if (!on_device && !c->d_pos[0]) return fail(WE_ERR_INVALID, "context was first used with device buffers; create a new context for host frames");
  const int H = c->H, N = c->N, ND = c->ND, C = c->C, S = c->S;
  const size_t frame_floats = (size_t)N * 3;
  // Forces are only read by k_cov, for the recorded waters (1-2 % of the atoms in the interfacial
  // recipe).  If the caller's force buffer is pinned (UVA-mapped) the kernel reads exactly those
  // atoms straight from host memory over PCIe and the whole-frame upload is skipped.
  const float* frc_zero_copy = nullptr;
  const bool no_forces = c->opt.shells_only != 0;  // nothing downstream of the shells reads forces
  if (!no_forces && !on_device && !c->lazy && !c->opt.keep_debug &&
      (c->opt.force_upload == 3 || (c->opt.force_upload == 0 && c->recipe == WE_RECIPE_INTERFACIAL))) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, frc) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
      frc_zero_copy = (const float*)at.devicePointer;
    else
      cudaGetLastError();
    if (!frc_zero_copy && c->opt.force_upload == 3)
      return fail(WE_ERR_INVALID, "force_upload = 3 needs the force buffer in pinned (page-locked) host memory");
  }
  if (!no_forces && !on_device && !c->lazy && !frc_zero_copy && !c->d_frc[0]) {
    for (int i = 0; i < WE_NBUF; ++i) {
      if ((rc = dev_alloc(c, &c->d_frc[i], (size_t)c->chunk * N * 3))) return rc;
    }
  }
  // Positions: the neighbour search reads heavy atoms only (a third of a water box), and of the light atoms
  // the path reads only the donor hydrogens of the centres whose hydrogen bonds are evaluated and the atoms
  // of the recorded waters (~5 % of the waters on C4).  PACKED uploads: host threads gather the heavy atoms
  // into pinned staging, the copy engine moves those, k_bin scatters them into the device's full-layout
  // positions and k_fetch_light reads the light atoms straight from the caller's buffer - which therefore
  // has to be pinned.  A third of the PCIe bytes of a whole-frame copy; the end-to-end arm was upload-bound.
  const float* pos_zero_copy = nullptr;
  // MEASURED (C4, one B200, 16 host cores): 9.2-10.8 ms per 96-frame step against 7.4 ms for whole frames.  The
  // copy engine moves a third of the bytes (124 MB instead of 362 MB per step), but the light atoms arrive as
  // ~350,000 scattered 24-byte PCIe reads per batch (k_fetch_light: 0.5-1.0 ms beside a 0.45 ms search pass,
  // k_hb waits for it), every kernel runs slower under that traffic, and the host gather (3.7-5.3 ms per step
  // on all 16 cores) is as long as the DMA it replaces.  So packed uploads are what pos_upload = 2 asks for,
  // never what auto (0) picks; the path stays as the tested starting point for hosts whose PCIe feed per GPU
  // is far below one x16 link (eight ranks on one box get 21 GB/s each).
  if (!on_device && c->opt.pos_upload == 2) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, pos) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
      pos_zero_copy = (const float*)at.devicePointer;
    else
      cudaGetLastError();
    if (!pos_zero_copy && c->opt.pos_upload == 2)
      return fail(WE_ERR_INVALID, "pos_upload = 2 needs the position buffer in pinned (page-locked) host memory");
  }
  const bool packed = pos_zero_copy != nullptr;
  if (packed && !c->h_pack[0]) {
    for (int i = 0; i < WE_NBUF; ++i) {
      if ((rc = pin_alloc(c, &c->h_pack[i], (size_t)c->chunk * H * 3))) return rc;
      if ((rc = dev_alloc(c, &c->d_pack[i], (size_t)c->chunk * H * 3))) return rc;
      if ((rc = dev_alloc(c, &c->d_src[i], 1))) return rc;
    }
    if ((rc = pin_alloc(c, &c->h_src, WE_NBUF))) return rc;
    if (!c->pool) {
      c->pool = new WePackPool(c->pack_threads);
      if (const char* g = getenv("WE_PACK_SPIN_US")) c->pool->spin_us = atol(g);
    }
  }
  // Host buffers: a call that fits in one or two batches would upload everything before computing
  // anything, so (auto batch size only, and only when the pipeline is empty: a caller that streams
  // batch after batch already overlaps) it is cut into at least three batches of >= ~1 M centres each
  // and the uploads of the later ones hide behind the kernels of the earlier ones.
  int step = c->chunk;
  if (!on_device && c->opt.chunk_frames <= 0 && n_frames < 3 * c->chunk && c->inflight.empty()) {
    const int floor_frames = std::max(1, (int)(1000000LL / std::max(1, H)));
    step = std::min(c->chunk, std::max((n_frames + 2) / 3, floor_frames));
    if (step >= 8) step &= ~7;
  }
  for (int f0 = 0, n = 0; f0 < n_frames; f0 += n) {
    n = std::min(step, n_frames - f0);
    if (c->max_rec_frame < 0 && c->C > 0 && c->recipe == WE_RECIPE_INTERFACIAL && !c->opt.shells_only) {
      // No batch has been drained yet: only n_centres bounds the keys a frame can add.  The first batch is
      // as many frames as the table takes on that bound (at least one); it is waited for, and from its
      // recorded-water counts on the real bound applies.  One short batch + one wait per analysis.
      if (c->frames_done > c->frames_drained) {
        for (int i = 0; i < WE_NBUF; ++i) {
          const int b = (int)((c->chunk_counter + i) % WE_NBUF);
          if ((rc = finish_lazy(c, b))) return rc;
          if ((rc = drain(c, b))) return rc;
        }
      }
      if (c->max_rec_frame < 0) {
        const unsigned long long cap = (unsigned long long)c->table_mask + 1;
        n = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((unsigned long long)n, (cap / 2) / (unsigned long long)c->C));
      }
    }
    const int buf = (int)(c->chunk_counter % WE_NBUF);
    c->chunk_counter++;
    // staging buffers of this slot were last used WE_NBUF chunks ago
    if ((rc = finish_lazy(c, buf))) return rc;
    if ((rc = drain(c, buf))) return rc;
    CK(cudaEventSynchronize(c->ev_in_free[buf]));
    while (!c->inflight.empty() && c->inflight.front().buf == buf) {  // the slot's previous batch is over
      c->frames_retired = c->inflight.front().frames_end;
      c->inflight.pop_front();
    }
    // batches that have finished meanwhile: move their lists to the host now (keeps the bound below tight)
    for (int i = 1; i < WE_NBUF; ++i) {
      const int b = (buf + i) % WE_NBUF;
      if (c->lazy_pending[b]) break;
      if (!c->out_pending[b]) continue;
      if (cudaEventQuery(c->ev_out_ready[b]) != cudaSuccess) { cudaGetLastError(); break; }
      if ((rc = drain(c, b))) return rc;
    }
    if ((rc = maybe_grow_table(c, n))) return rc;
    for (int f = 0; f < n; ++f) {
      if (make_box(dims + (size_t)(f0 + f) * 6, c->cell_target, c->ncap, &c->h_boxes[buf][f]))
        return fail(WE_ERR_INVALID, "frame %lld: box dimensions must be positive and finite", (long long)frame_ids[f0 + f]);
      c->h_fids[buf][f] = (long long)frame_ids[f0 + f];
    }
    const float* dp;
    const float* df;
    if (!on_device) {
      // While a host->device DMA is in flight every kernel launch of the compute stream takes 20-40 us
      // longer (measured: the short kernels of a batch appear 2-3x slower, the 0.4 ms first neighbour
      // pass +3 %).  WE_COPY_GATE=1 therefore holds the upload of this batch back until the first pass
      // of the previous batch starts.  That paid while an upload was shorter than a batch's kernels; with
      // 32-frame batches on C4 the upload (2.2 ms) is as long as the kernels (2.1 ms), the gate only
      // idles the copy engine, and back-to-back uploads are faster (7.77 -> 7.40 ms per 96-frame step):
      // the default is ungated.
      if (c->copy_gate) CK(cudaStreamWaitEvent(c->s_copy, c->ev_p0[(buf + WE_NBUF - 1) % WE_NBUF], 0));
      if (packed) {
        // the slot's staging was last read by the DMA of the batch WE_NBUF batches ago (ev_in_free above)
        const auto tp0 = std::chrono::steady_clock::now();
        c->pool->run(pos + (size_t)f0 * frame_floats, c->h_pack[buf], c->heavy_idx.data(), H, frame_floats, n);
        c->pack_us += (unsigned long long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - tp0).count();
        CK(cudaMemcpyAsync(c->d_pack[buf], c->h_pack[buf], (size_t)n * H * 3 * sizeof(float), cudaMemcpyHostToDevice, c->s_copy));
        c->h_src[buf] = pos_zero_copy + (size_t)f0 * frame_floats;
        CK(cudaMemcpyAsync(c->d_src[buf], &c->h_src[buf], sizeof(float*), cudaMemcpyHostToDevice, c->s_copy));
        c->h2d_bytes += (unsigned long long)n * H * 3 * sizeof(float) + sizeof(float*);
      } else {
        CK(cudaMemcpyAsync(c->d_pos[buf], pos + (size_t)f0 * frame_floats, (size_t)n * frame_floats * sizeof(float), cudaMemcpyHostToDevice, c->s_copy));
        c->h2d_bytes += (unsigned long long)n * frame_floats * sizeof(float);
      }
      dp = c->d_pos[buf];
      if (no_forces) {
        df = nullptr;
      } else if (frc_zero_copy) {
        df = frc_zero_copy + (size_t)f0 * frame_floats;
      } else {
        if (!c->lazy) {
          CK(cudaMemcpyAsync(c->d_frc[buf], frc + (size_t)f0 * frame_floats, (size_t)n * frame_floats * sizeof(float), cudaMemcpyHostToDevice, c->s_copy));
          c->h2d_bytes += (unsigned long long)n * frame_floats * sizeof(float);
        }
        df = c->d_frc[buf];
      }
    } else {
      dp = pos + (size_t)f0 * frame_floats;
      df = no_forces ? nullptr : frc + (size_t)f0 * frame_floats;
    }
    CK(cudaMemcpyAsync(c->d_boxes[buf], c->h_boxes[buf], (size_t)n * sizeof(WeBox), cudaMemcpyHostToDevice, c->s_copy));
    CK(cudaMemcpyAsync(c->d_fids[buf], c->h_fids[buf], (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, c->s_copy));
    if (!on_device) c->h2d_bytes += (unsigned long long)n * (sizeof(WeBox) + sizeof(long long));
    CK(cudaEventRecord(c->ev_in_ready[buf], c->s_copy));
    cudaStream_t st = c->s_comp, cov_st = c->s_comp;
    CK(cudaStreamWaitEvent(st, c->ev_in_ready[buf], 0));
    const WeBox* bx = c->d_boxes[buf];
    const long long* fids = c->d_fids[buf];
    const bool interfacial = (c->recipe == WE_RECIPE_INTERFACIAL);
    const bool shells_only = c->opt.shells_only != 0;
    dim3 gH((H + 255) / 256, n);
    // burial pre-stage of the gating pass (k_gate): production interfacial contexts, every frame of the batch
    // with enough cells per axis for the group window (6 x 6 x 5 cells) to have one definite periodic image,
    // a +-2 cell first search and, for triclinic boxes, a valid float32 prefilter
    bool gate = interfacial && c->use_gate && !c->opt.keep_debug && S > 0 && C > 0;
    for (int f = 0; gate && f < n; ++f) {
      const WeBox& hb = c->h_boxes[buf][f];
      for (int i = 0; i < 3; ++i) gate = gate && hb.nc[i] >= 8 && hb.sw[i] <= 2;
      gate = gate && (!hb.triclinic || hb.pre_ok);
    }
    auto gkey_of = [](int b, int nn, bool g, bool pk) { return ((((long long)b * 4096 + nn) * 2 + (g ? 1 : 0)) * 2 + (pk ? 1 : 0)) * 2; };
    // Device-resident inputs: the caller's position pointer is baked into the captured launches, so it is part
    // of the key (a caller that streams through ever-new device buffers stops capturing at 64 graphs).
    const void* gptr = on_device ? (const void*)dp : nullptr;
    bool graphs_on = c->use_graphs && !c->prof && !c->opt.keep_debug;
    if (on_device && graphs_on && c->graphs.size() >= 64 && !c->graphs.count({gkey_of(buf, n, gate, packed), gptr}))
      graphs_on = false;
    const long long gkey = gkey_of(buf, n, gate, packed);
    GraphRegion g1(c, st, gkey, gptr, graphs_on), g2(c, st, gkey + 1, gptr, graphs_on);
    int replay = g1.begin();
    if (replay < 0) return replay;
    if (!replay) {
    CK(cudaMemsetAsync(c->d_zero, 0, c->zero_ints * sizeof(int), st));  // cells, state, queues, counters, flags
    CK(cudaMemsetAsync(c->d_rec_count[buf], 0, (size_t)n * sizeof(int), st));
    { ProfScope ps(c, WE_K_BIN, st);
      k_bin<<<gH, 256, 0, st>>>(c->T, dp, bx, c->d_cell_count, c->d_cell_of, c->d_rank_of, c->d_tmp4, c->ncap,
                                gate ? c->d_grp_flag : nullptr, packed ? c->d_pack[buf] : nullptr,
                                packed ? c->d_pos[buf] : nullptr);
      ps.stop(st); }
    { ProfScope ps(c, WE_K_SCAN, st);
      k_scan<<<n, 1024, 0, st>>>(bx, c->d_cell_count, c->ncap);
      ps.stop(st); }
    { ProfScope ps(c, WE_K_SCATTER, st);
      k_scatter<<<gH, 256, 0, st>>>(H, bx, c->d_cell_count, c->d_cell_of, c->d_rank_of, c->d_tmp4, c->d_sorted4, c->ncap,
                                    c->T.slot_flag, gate ? c->d_sflag : nullptr);
      ps.stop(st); }
    c->launches += 3;
    if (gate) {
      ProfScope ps(c, WE_K_GATE, st);
      k_grplist<<<dim3((c->ncap + 1 + 255) / 256, n), 256, 0, st>>>(bx, c->d_grp_flag, c->ncap, c->d_grp_list, c->d_grp_count);
      ps.stop(st);
      c->launches++;
    }
    }
    // pass-0 centres: every heavy atom (debug), solute united atoms (interfacial), waters (bulk)
    WeWork w0;
    w0.count = nullptr; w0.stride = 0;
    if (c->opt.keep_debug) { w0.list = c->d_all_slots; w0.static_count = H; }
    else if (interfacial) { w0.list = c->T.solute_ua; w0.static_count = S; }
    else { w0.list = c->T.centres; w0.static_count = C; }
    WeRadOut ro{c->d_shell_n, c->d_shell_idx, c->d_nn, c->d_flags, c->d_nbr_idx, c->d_nbr_dist};
    auto rad_grid = [&](long long items) {
      long long blocks = (items + WE_RAD_WARPS - 1) / WE_RAD_WARPS;
      return (int)std::max<long long>(1, std::min<long long>(blocks, c->rad_blocks));
    };
    auto grid_for = [&](long long items, int per_block, int cap) {
      long long blocks = (items + per_block - 1) / per_block;
      return (int)std::max<long long>(1, std::min<long long>(blocks, cap));
    };
    // marks written by the passes (bytes inside the zero block): see WeMarks
    unsigned char* mark_hb = reinterpret_cast<unsigned char*>(c->d_hbq);
    unsigned char* mark_shell = mark_hb + (size_t)c->chunk * H;
    const WeMarks no_marks{nullptr, nullptr, nullptr, nullptr, 0};
    // pass 0 of the recipe: solute united atoms mark interfacial waters / bulk waters mark their HB centres
    const WeMarks m0 = interfacial ? WeMarks{c->d_interf, nullptr, nullptr, fids, 0}
                                   : WeMarks{nullptr, mark_hb, nullptr, nullptr, 1};
    int rad_launch = 0;  // every k_rad launch of a batch claims from its own counter (zeroed with the batch's zero block)
    auto rad = [&](int id, const WeWork& wk, long long items, int* gate_state, const WeMarks& mk) {
      if (gate_state && gate) {
        // buried solute atoms are settled by cell group; the rest form the work list of the generic search
        ProfScope pg(c, WE_K_GATE, st);
        k_gate<<<c->gate_blocks, WE_GATE_WARPS * 32, sizeof(WeGateSmem), st>>>(
            c->T, bx, c->d_cell_count, c->d_sorted4, c->d_sflag, c->ncap, c->cell_target, c->d_grp_list, c->d_grp_count,
            n, ro, gate_state, c->d_wl0, c->d_n0, std::max(1, S), c->d_work_counter);
        pg.stop(st);
      }
      ProfScope ps(c, id, st);
      const float r_first = c->cell_target * c->r_first_frac[id == WE_K_RAD ? 0 : (id == WE_K_RAD1 ? 1 : 2)];
      if (gate_state && gate) {
        const WeWork open{c->d_wl0, c->d_n0, std::max(1, S), 0};
        k_rad<false><<<rad_grid(items), WE_RAD_WARPS * 32, sizeof(WeRadSmem), st>>>(
            c->T, dp, bx, c->d_cell_count, c->d_cell_of, c->d_sorted4, c->d_tmp4, c->ncap, r_first, open, n, ro,
            gate_state, mk, c->d_diag, c->d_work_counter + 1 + (rad_launch++ % 7));
        c->launches++;
      } else if (gate_state)
        k_rad<true><<<(int)std::max<long long>(1, std::min<long long>((items + WE_RAD_WARPS - 1) / WE_RAD_WARPS, c->rad_blocks_gated)), WE_RAD_WARPS * 32, sizeof(WeRadSmem), st>>>(
            c->T, dp, bx, c->d_cell_count, c->d_cell_of, c->d_sorted4, c->d_tmp4, c->ncap, r_first, wk, n, ro,
            gate_state, mk, c->d_diag, c->d_work_counter + 1 + (rad_launch++ % 7));
      else
        k_rad<false><<<rad_grid(items), WE_RAD_WARPS * 32, sizeof(WeRadSmem), st>>>(
            c->T, dp, bx, c->d_cell_count, c->d_cell_of, c->d_sorted4, c->d_tmp4, c->ncap, r_first, wk, n, ro,
            nullptr, mk, c->d_diag, c->d_work_counter + 1 + (rad_launch++ % 7));
      ps.stop(st);
      c->launches++;
    };
    if (w0.static_count > 0 && !replay) {
      ProfScope ps(c, WE_K_LIST2, st);
      k_seed<<<dim3((w0.static_count + 255) / 256, n), 256, 0, st>>>(w0.static_count, w0.list, H, c->d_state);
      ps.stop(st);
      c->launches++;
    }
    if ((rc = g1.end())) return rc;
    CK(cudaEventRecord(c->ev_p0[buf], st));
    replay = g2.begin();
    if (replay < 0) return replay;
    if (!replay) {
    if (w0.static_count > 0) {
      if (!c->opt.keep_debug) {
        rad(WE_K_RAD, w0, (long long)w0.static_count * n, interfacial ? c->d_state : nullptr, m0);
      } else {
        // debug contexts evaluate every heavy atom first (no marks), then the recipe's own pass 0
        rad(WE_K_RAD, w0, (long long)w0.static_count * n, nullptr, no_marks);
        WeWork wr{interfacial ? c->T.solute_ua : c->T.centres, nullptr, 0, interfacial ? S : C};
        if (wr.static_count > 0) rad(WE_K_RAD, wr, (long long)wr.static_count * n, nullptr, m0);
      }
    }
    if (interfacial && S > 0 && C > 0) {
      { ProfScope ps(c, WE_K_LIST1, st);
        k_list1<<<dim3((C + 255) / 256, n), 256, 0, st>>>(c->T, c->d_interf, c->d_state, c->d_wl1, c->d_n1);
        ps.stop(st); }
      WeWork w1{c->d_wl1, c->d_n1, C, 0};
      if (shells_only) {
        // get_interfacial_shells stops here: shells of the interfacial waters, nothing downstream
        rad(WE_K_RAD1, w1, (long long)C * n, nullptr, no_marks);
        c->launches += 1;
      } else {
      rad(WE_K_RAD1, w1, (long long)C * n, nullptr, WeMarks{nullptr, mark_hb, mark_shell, nullptr, 0});
      { ProfScope ps(c, WE_K_LIST2, st);
        k_list2<<<gH, 256, 0, st>>>(c->T, mark_shell, mark_hb, c->d_state, c->d_wl2, c->d_n2, c->d_wlh, c->d_nh);
        ps.stop(st); }
      if (packed) {
        // light atoms of the centres on the HB list (k_hb: donor hydrogens; k_cov: the recorded waters' atoms),
        // read over PCIe beside the last search pass: a fork / join through the auxiliary stream (also inside
        // the captured graph)
        CK(cudaEventRecord(c->ev_fork, st));
        CK(cudaStreamWaitEvent(c->s_aux, c->ev_fork, 0));
        { ProfScope ps(c, WE_K_FETCH, c->s_aux);
          k_fetch_light<<<dim3(64, n), 128, 0, c->s_aux>>>(c->T, c->d_src[buf], c->d_pos[buf], WeWork{c->d_wlh, c->d_nh, H, 0},
                                                           c->d_centre_of_slot, n, c->d_diag);
          ps.stop(c->s_aux); }
        CK(cudaEventRecord(c->ev_join, c->s_aux));
        c->launches += 1;
      }
      WeWork w2{c->d_wl2, c->d_n2, H, 0};
      rad(WE_K_RAD2, w2, (long long)H * n, nullptr, no_marks);
      if (packed) CK(cudaStreamWaitEvent(st, c->ev_join, 0));
      c->launches += 2;
      }
    } else if (!interfacial && C > 0) {
      ProfScope ps(c, WE_K_LIST2, st);
      k_list2<<<gH, 256, 0, st>>>(c->T, nullptr, mark_hb, c->d_state, c->d_wl2, c->d_n2, c->d_wlh, c->d_nh);
      ps.stop(st);
      c->launches++;
    }
    if (ND > 0 && !shells_only) {
      // HB work list: every heavy atom (debug) or the centres marked by the passes (k_list2)
      WeWork wh;
      if (c->opt.keep_debug) wh = WeWork{c->d_all_slots, nullptr, 0, H};
      else wh = WeWork{c->d_wlh, c->d_nh, H, 0};
      ProfScope ps(c, WE_K_HB, st);
      k_hb<<<grid_for((long long)H * n, 16, c->hb_blocks), 256, 0, st>>>(
          c->T, dp, bx, c->d_tmp4, c->d_shell_n, c->d_shell_idx, c->d_flags, wh, n, c->d_acceptor, c->d_diag,
          c->opt.keep_debug ? 1 : 0);
      ps.stop(st);
      c->launches++;
    }
    if (C > 0) {
      WeWork wlab;
      if (interfacial) wlab = WeWork{c->d_wl1, c->d_n1, C, 0};
      else wlab = WeWork{c->T.centres, nullptr, 0, C};
      { ProfScope ps(c, WE_K_LABEL, st);
        k_label<<<grid_for((long long)C * n, 8, c->label_blocks), 256, 0, st>>>(
            c->T, c->d_shell_n, c->d_shell_idx, c->d_nn, c->d_flags, c->d_acceptor, c->d_interf, fids,
            c->d_table, c->table_mask, c->d_rec_count[buf], c->d_rec[buf],
            (c->opt.keep_records || c->lazy || c->opt.keep_shells) ? c->h_rec[buf] : nullptr,
            c->opt.keep_shells ? c->h_shell[buf] : nullptr, wlab, n, c->d_diag, shells_only ? 1 : 0);
        ps.stop(st); }
      c->launches += 1;
    }
    CK(cudaMemcpyAsync(c->h_diag, c->d_diag, sizeof(WeDiag), cudaMemcpyDeviceToHost, st));  // table load mirror
    // per-frame record counts -> pinned staging (the records themselves are already in mapped memory)
    CK(cudaMemcpyAsync(c->h_rec_count[buf], c->d_rec_count[buf], (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
    }
    if ((rc = g2.end())) return rc;
    if (C > 0 && !shells_only) {
      if (!c->lazy) {
        // zero-copy forces: the kernel waits on PCIe reads, so it runs beside the next batch's kernels
        // (WE_COV_STREAM=2: also with forces resident on the device - the kernel is one long dependent fp64
        // chain per water, ~80 us whatever the count, and fits beside the next batch's cell-list kernels)
        cov_st = ((frc_zero_copy && c->cov_stream) || c->cov_stream >= 2) ? c->s_cov : st;
        c->cov_on_own_stream = (cov_st != st);
        if (cov_st != st) {
          CK(cudaEventRecord(c->ev_label[buf], st));
          CK(cudaStreamWaitEvent(cov_st, c->ev_label[buf], 0));
        }
        ProfScope ps(c, WE_K_COV, cov_st);
        // zero-copy: the kernel is bound by PCIe read latency, not by SM throughput; about two blocks per
        // SM keep as many reads in flight as the link takes and leave the SMs to the next batch
        const int cov_bx = (cov_st != st) ? std::max(1, std::min((C + 511) / 512, (2 * c->n_sm + n - 1) / n)) : (C + 511) / 512;
        k_cov<<<dim3(cov_bx, n), 128, 0, cov_st>>>(
            c->T, dp, df, c->d_rec_count[buf], c->d_rec[buf], c->d_centre_of_slot, c->d_cov_sum,
            c->d_cov_cnt, c->d_ft, nullptr, nullptr, 0, n, c->opt.eig_flavour, c->d_diag);
        ps.stop(cov_st);
        c->launches += 1;
      }
    }
    CK(cudaGetLastError());
    c->pending_n[buf] = n;
    c->frames_done += (unsigned long long)n;
    c->inflight.push_back({c->frames_done, buf});
    if (c->lazy && C > 0 && !shells_only) {
      // the covariance kernel of this batch runs one batch later, after the host has gathered
      // the forces of the recorded waters (finish_lazy); until then its inputs stay alive
      CK(cudaEventRecord(c->ev_label[buf], st));
      c->lazy_pending[buf] = true;
      c->pend_frc[buf] = frc + (size_t)f0 * frame_floats;
      c->pend_dp[buf] = dp;
      if ((rc = finish_lazy(c, (buf + WE_NBUF - 1) % WE_NBUF))) return rc;
    } else {
      if (cov_st != st) {  // both events must also cover the record-count copy queued on the compute stream
        CK(cudaEventRecord(c->ev_f[buf], st));
        CK(cudaStreamWaitEvent(cov_st, c->ev_f[buf], 0));
      }
      CK(cudaEventRecord(c->ev_in_free[buf], cov_st));
      CK(cudaEventRecord(c->ev_out_ready[buf], cov_st));
      c->out_pending[buf] = true;
    }
    if (c->opt.keep_debug) {
      CK(cudaStreamSynchronize(st));
      for (int i = 1; i <= WE_NBUF; ++i)  // oldest first: keeps frame order in frame_records
        if ((rc = drain(c, (buf + i) % WE_NBUF))) return rc;
      std::vector<int> rc_h(n);
      CK(cudaMemcpy(rc_h.data(), c->d_rec_count[buf], (size_t)n * sizeof(int), cudaMemcpyDeviceToHost));
      for (int f = 0; f < n; ++f) {
        FrameDebug D;
        D.shell_n.resize(H); D.shell_idx.resize((size_t)H * WE_MAX_NBR); D.nn.resize(H); D.flags.resize(H);
        D.acceptor.resize(ND); D.nbr_idx.resize((size_t)H * WE_MAX_NBR); D.nbr_dist.resize((size_t)H * WE_MAX_NBR);
        D.ft.resize((size_t)rc_h[f] * 6); D.rec.resize((size_t)rc_h[f] * 4);
        CK(cudaMemcpy(D.shell_n.data(), c->d_shell_n + (size_t)f * H, (size_t)H * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(D.shell_idx.data(), c->d_shell_idx + (size_t)f * H * WE_MAX_NBR, (size_t)H * WE_MAX_NBR * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(D.nn.data(), c->d_nn + (size_t)f * H, (size_t)H * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(D.flags.data(), c->d_flags + (size_t)f * H, (size_t)H * 4, cudaMemcpyDeviceToHost));
        if (ND) CK(cudaMemcpy(D.acceptor.data(), c->d_acceptor + (size_t)f * ND, (size_t)ND * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(D.nbr_idx.data(), c->d_nbr_idx + (size_t)f * H * WE_MAX_NBR, (size_t)H * WE_MAX_NBR * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(D.nbr_dist.data(), c->d_nbr_dist + (size_t)f * H * WE_MAX_NBR, (size_t)H * WE_MAX_NBR * 8, cudaMemcpyDeviceToHost));
        if (rc_h[f]) {
          CK(cudaMemcpy(D.ft.data(), c->d_ft + (size_t)f * C * 6, (size_t)rc_h[f] * 6 * 8, cudaMemcpyDeviceToHost));
          CK(cudaMemcpy(D.rec.data(), c->d_rec[buf] + (size_t)f * C, (size_t)rc_h[f] * sizeof(WeRecord), cudaMemcpyDeviceToHost));
        }
        c->debug.push_back(std::move(D));
      }
    }
  }
  return WE_OK;
}

extern "C" int we_submit(we_ctx* c, const float* pos, const float* frc, const float* dims,
                         const int64_t* frame_ids, int32_t n_frames) {
  return submit_impl(c, pos, frc, false, dims, frame_ids, n_frames);
}
extern "C" int we_submit_device(we_ctx* c, const float* d_pos, const float* d_frc, const float* dims,
                                const int64_t* frame_ids, int32_t n_frames) {
  return submit_impl(c, d_pos, d_frc, true, dims, frame_ids, n_frames);
}

// Everything submitted so far has updated the accumulators: the engine's streams are non-blocking, so
// the blocking copies of the read-out calls (legacy default stream) do not order behind them by themselves.
static int quiesce(we_ctx* c) {
  CK(cudaSetDevice(c->device));
  const int rc = finish_lazy_all(c);
  if (rc) return rc;
  CK(cudaStreamSynchronize(c->s_comp));
  CK(cudaStreamSynchronize(c->s_cov));
  return 0;
}

static int read_diag(we_ctx* c, WeDiag* d) {
  CK(cudaMemcpy(d, c->d_diag, sizeof(WeDiag), cudaMemcpyDeviceToHost));
  return 0;
}

// ---- f3: vibrational entropy of every covariance bin on the device ---------------------------------
// entropy/vibrations.py:142-192 with diagonalise=True: lambda = diagonal of the mean matrix in SI units,
// nu = sqrt(lambda / kT), a = hbar nu / kT, S = R (a / (e^a - 1) - ln(1 - e^-a)); zero where a is zero.
__global__ void k_vib_entropy(const double* __restrict__ cov_sum, const unsigned long long* __restrict__ cov_cnt,
                              int n_bins, double conversion, double kT, double hbar, double gas, double* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_bins * 6) return;
  const int b = t / 6, m = (t % 6) / 3, i = t % 3;
  const unsigned long long n = cov_cnt[b];
  double s = 0.0;
  if (n) {
    double lam = (cov_sum[(size_t)b * 18 + m * 9 + i * 4] / (double)n) * conversion;
    if (!(lam >= 0.0)) lam = 0.0;
    const double a = (hbar * sqrt(lam / kT)) / kT;
    if (a != 0.0) {
      const double bb = exp(a) - 1.0, cc = log(1.0 - exp(-a));
      if (bb != 0.0 && cc != 0.0) s = a / bb - cc;
    }
  }
  out[t] = s * gas;
}

extern "C" int we_vib_entropy(we_ctx* c, double temperature, double conversion, double* out) {
  if (!c || !out) return fail(WE_ERR_INVALID, "null argument");
  { const int rc = quiesce(c); if (rc) return rc; }
  double* d_out = nullptr;
  CK(cudaMalloc((void**)&d_out, (size_t)c->B * 6 * sizeof(double)));
  const double kB = 1.38064852e-23, hbar = 1.0545718e-34, gas = 8.314;  // entropy/vibrations.py:26-29
  k_vib_entropy<<<(c->B * 6 + 127) / 128, 128>>>(c->d_cov_sum, c->d_cov_cnt, c->B, conversion, kB * temperature, hbar, gas, d_out);
  c->launches++;
  cudaError_t e = cudaMemcpy(out, d_out, (size_t)c->B * 6 * sizeof(double), cudaMemcpyDeviceToHost);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(WE_ERR_CUDA, "we_vib_entropy failed: %s", cudaGetErrorString(e));
  return WE_OK;
}

// entropy/orientations.py:298-370 on the device: one row per group of label entries (we_orient.cuh).
static int compact_entries(we_ctx* c, int n);
extern "C" int we_orient_entropy(we_ctx* c, const uint16_t* label_rank, int32_t by_resname, int32_t max_groups,
                                 int32_t* out_keys, double* out_rows, int32_t* n_groups) {
  if (!c || !label_rank || !out_keys || !out_rows || !n_groups || max_groups < 0) return fail(WE_ERR_INVALID, "null argument");
  { const int rc = quiesce(c); if (rc) return rc; }
  const int n = we_n_label_entries(c);
  if (n < 0) return n;
  *n_groups = 0;
  if (n == 0) return WE_OK;
  { const int rc = compact_entries(c, n); if (rc) return rc; }
  int n_pad = 1;
  while (n_pad < n) n_pad <<= 1;
  // scratch kept with the context and grown on demand: the recipes call this twice per analysis (per residue
  // id, per residue name), and six cudaMalloc / cudaFree pairs per call cost more than the kernels
  cudaError_t e = cudaSuccess;
  auto grow = [&](void** p, size_t* cap, size_t bytes) {
    if (e != cudaSuccess || (*p && *cap >= bytes)) return;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = bytes + bytes / 2 + 64;
    e = cudaMalloc(p, *cap);
    if (e != cudaSuccess) { *p = nullptr; *cap = 0; }
  };
  we_ctx::OrientScratch& os = c->orient;
  grow((void**)&os.rank, &os.rank_cap, 65536 * sizeof(unsigned short));
  grow((void**)&os.rec, &os.rec_cap, (size_t)n_pad * sizeof(WeOrientRec));
  grow((void**)&os.rows, &os.rows_cap, (size_t)n * sizeof(WeOrientRow));
  grow((void**)&os.out, &os.out_cap, (size_t)max_groups * 8 * sizeof(double));
  grow((void**)&os.keys, &os.keys_cap, (size_t)max_groups * 2 * sizeof(int));
  grow((void**)&os.ng, &os.ng_cap, sizeof(int));
  if (e != cudaSuccess) return fail(WE_ERR_CUDA, "we_orient_entropy: scratch allocation failed: %s", cudaGetErrorString(e));
  unsigned short* d_rank = (unsigned short*)os.rank;
  WeOrientRec* d_rec = (WeOrientRec*)os.rec;
  WeOrientRow* d_rows = (WeOrientRow*)os.rows;
  double* d_out = (double*)os.out;
  int *d_keys = (int*)os.keys, *d_ng = (int*)os.ng;
  cudaStream_t st = c->s_comp;
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_rank, label_rank, 65536 * sizeof(unsigned short), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_ng, 0, sizeof(int), st);
  if (e == cudaSuccess) {
    const WeLabelEntry* ent = c->d_compact;
    k_orient_keys<<<(n_pad + 127) / 128, 128, 0, st>>>(ent, n, n_pad, d_rank, by_resname ? 1 : 0, d_rec);
    c->launches++;
    for (int k = 2; k <= n_pad; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) { k_orient_bitonic<<<(n_pad + 255) / 256, 256, 0, st>>>(d_rec, n_pad, j, k); c->launches++; }
    k_orient_rows<<<(n + 127) / 128, 128, 0, st>>>(ent, d_rec, n, d_rank, d_rows);
    k_orient_fold<<<(n + 127) / 128, 128, 0, st>>>(ent, d_rec, d_rows, n, max_groups, d_ng, d_out, d_keys);
    c->launches += 2;
    e = cudaStreamSynchronize(st);
  }
  int ng = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&ng, d_ng, sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && ng <= max_groups && ng > 0) {
    e = cudaMemcpy(out_rows, d_out, (size_t)ng * 8 * sizeof(double), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpy(out_keys, d_keys, (size_t)ng * 2 * sizeof(int), cudaMemcpyDeviceToHost);
  }
  if (e != cudaSuccess) return fail(WE_ERR_CUDA, "we_orient_entropy failed: %s", cudaGetErrorString(e));
#ifdef WE_CHECKED
  {
    int line = 0;
    CK(cudaMemcpyFromSymbol(&line, we_check_line, sizeof(int)));
    if (line) return fail(WE_ERR_INVALID, "WE_CHECKED build: index bound violated at line %d of we_orient.cuh / we_kernels.cuh", line);
  }
#endif
  *n_groups = ng;
  if (ng > max_groups) return fail(WE_ERR_INVALID, "max_groups %d < %d groups", max_groups, ng);
  return WE_OK;
}

// ---- input-buffer retirement ------------------------------------------------------------------
extern "C" int64_t we_frames_retired(we_ctx* c) {
  if (!c) return fail(WE_ERR_INVALID, "null context");
  cudaSetDevice(c->device);
  while (!c->inflight.empty()) {
    const we_ctx::InFlight& b = c->inflight.front();
    if (c->lazy_pending[b.buf]) break;  // its forces are still to be gathered from the caller's buffer
    if (cudaEventQuery(c->ev_in_free[b.buf]) != cudaSuccess) { cudaGetLastError(); break; }
    c->frames_retired = b.frames_end;
    c->inflight.pop_front();
  }
  return (int64_t)c->frames_retired;
}

extern "C" int we_wait_frames_retired(we_ctx* c, int64_t n_frames) {
  if (!c) return fail(WE_ERR_INVALID, "null context");
  CK(cudaSetDevice(c->device));
  if ((unsigned long long)std::max<int64_t>(0, n_frames) > c->frames_done)
    return fail(WE_ERR_INVALID, "only %llu frames were submitted", c->frames_done);
  while (!c->inflight.empty() && c->frames_retired < (unsigned long long)n_frames) {
    const we_ctx::InFlight b = c->inflight.front();
    int rc = finish_lazy(c, b.buf);
    if (rc) return rc;
    CK(cudaEventSynchronize(c->ev_in_free[b.buf]));
    c->frames_retired = b.frames_end;
    c->inflight.pop_front();
  }
  return WE_OK;
}

extern "C" int we_sync(we_ctx* c) {
  if (!c) return fail(WE_ERR_INVALID, "null context");
  CK(cudaSetDevice(c->device));
  {
    int rc0;
    if ((rc0 = finish_lazy_all(c))) return rc0;
  }
  // drain in submission order; each slot waits for its own batch only, so the host sorts the lists of
  // batch k while the device still runs the batches after it
  int rc;
  for (int i = 0; i < WE_NBUF; ++i)
    if ((rc = drain(c, (int)((c->chunk_counter + i) % WE_NBUF)))) return rc;
  CK(cudaStreamSynchronize(c->s_copy));
  CK(cudaStreamSynchronize(c->s_comp));
  CK(cudaStreamSynchronize(c->s_cov));
  c->inflight.clear();
  c->frames_retired = c->frames_done;
  prof_resolve(c);
#ifdef WE_CHECKED
  {
    int line = 0;
    CK(cudaMemcpyFromSymbol(&line, we_check_line, sizeof(int)));
    if (line) return fail(WE_ERR_INVALID, "WE_CHECKED build: index bound violated at we_kernels.cuh:%d", line);
  }
#endif
  WeDiag d;
  if ((rc = read_diag(c, &d))) return rc;
  switch (d.err_code) {
    case 0: return WE_OK;
    case WE_ST_DUPLICATE: return fail(WE_ERR_DUPLICATE_COORD, "Atom coordinates of atom %d in neighbour list (frame %d)", d.err_atom, d.err_frame);
    case WE_ST_NO_NEIGHBOURS: return fail(WE_ERR_NO_NEIGHBOURS, "not enough values to unpack: no neighbours within the cut-off of atom %d (frame %d)", d.err_atom, d.err_frame);
    case WE_ST_OVERFLOW: return fail(WE_ERR_CAND_OVERFLOW, "candidate buffer overflow at atom %d (frame %d)", d.err_atom, d.err_frame);
    case 4: return fail(WE_ERR_TABLE_FULL, "label count table full (raise table_log2)");
    case 5: return fail(WE_ERR_HASH_COLLISION, "label key hash collision at atom %d (frame %d)", d.err_atom, d.err_frame);
    default: return fail(WE_ERR_INVALID, "unknown device error %d", d.err_code);
  }
}

extern "C" int we_error_location(we_ctx* c, int64_t* frame_id, int32_t* atom) {
  WeDiag d;
  int rc = read_diag(c, &d);
  if (rc) return rc;
  if (frame_id) *frame_id = d.err_frame;
  if (atom) *atom = d.err_atom;
  return WE_OK;
}

extern "C" int we_get_diag(we_ctx* c, we_diag* out) {
  if (!c || !out) return fail(WE_ERR_INVALID, "null argument");
  WeDiag d;
  int rc = read_diag(c, &d);
  if (rc) return rc;
  out->wide_searches = d.wide_searches;
  out->shrink_retries = d.shrink_retries;
  out->table_entries = d.table_entries;
  out->kernel_launches = c->launches;
  out->frames_done = c->frames_done;
  out->recorded = c->recorded;
  out->fast_path_mismatches = d.fast_path_mismatches;
  out->eig_fallbacks = d.eig_fallbacks;
  out->h2d_dma_bytes = c->h2d_bytes;
  out->light_atoms_fetched = d.light_fetched;
  out->host_pack_us = c->pack_us;
  return WE_OK;
}

extern "C" int we_profile(we_ctx* c, int32_t enable) {
  if (!c) return fail(WE_ERR_INVALID, "null context");
  c->prof = enable != 0;
  if (enable) { for (int i = 0; i < WE_N_KERNELS; ++i) { c->prof_ms[i] = 0; c->prof_n[i] = 0; } }
  return WE_OK;
}
extern "C" int we_kernel_times(we_ctx* c, double* ms, uint64_t* launches) {
  if (!c || !ms || !launches) return fail(WE_ERR_INVALID, "null argument");
  for (int i = 0; i < WE_N_KERNELS; ++i) { ms[i] = c->prof_ms[i]; launches[i] = c->prof_n[i]; }
  return WE_OK;
}
// device-time span of the compute stream: start marks "now" in stream order, stop
// returns the elapsed milliseconds after everything queued so far has finished
extern "C" int we_span_start(we_ctx* c) {
  if (!c) return fail(WE_ERR_INVALID, "null context");
  CK(cudaSetDevice(c->device));
  if (!c->span_a) { CK(cudaEventCreate(&c->span_a)); CK(cudaEventCreate(&c->span_b)); }
  CK(cudaEventRecord(c->span_a, c->s_comp));
  c->span_open = true;
  return WE_OK;
}
extern "C" int we_span_stop(we_ctx* c, double* ms) {
  if (!c || !ms || !c->span_open) return fail(WE_ERR_INVALID, "span not started");
  // the span ends when EVERY stream of the context is done: the covariance kernels run on their own stream
  if (!c->ev_cov) CK(cudaEventCreateWithFlags(&c->ev_cov, cudaEventDisableTiming));
  CK(cudaEventRecord(c->ev_cov, c->s_cov));
  CK(cudaStreamWaitEvent(c->s_comp, c->ev_cov, 0));
  CK(cudaEventRecord(c->span_b, c->s_comp));
  CK(cudaEventSynchronize(c->span_b));
  float f = 0.f;
  CK(cudaEventElapsedTime(&f, c->span_a, c->span_b));
  *ms = f;
  c->span_open = false;
  return WE_OK;
}

// Trajectory ingestion: big-endian float32 blocks -> native float32 (+ the readers' unit conversion), threaded.
// One loop body, compiled twice: for the baseline ISA and for AVX2 (the byte swap then is one vpshufb per eight
// values); picked once per call with __builtin_cpu_supports.
#define WE_CONVERT_BODY                                                    \
  for (uint64_t i = 0; i < n; ++i) {                                       \
    uint32_t v;                                                            \
    memcpy(&v, s + 4 * i, 4);                                              \
    v = __builtin_bswap32(v);                                              \
    float f;                                                               \
    memcpy(&f, &v, 4);                                                     \
    if (op == 1) f = f * factor;                                           \
    else if (op == 2) f = f / factor;                                      \
    d[i] = f;                                                              \
  }
static void we_convert_base(const unsigned char* s, float* d, uint64_t n, int op, float factor) { WE_CONVERT_BODY }
__attribute__((target("avx2"))) static void we_convert_avx2(const unsigned char* s, float* d, uint64_t n, int op, float factor) { WE_CONVERT_BODY }

extern "C" int we_convert_be_f32(const void* const* src, float* const* dst, int32_t n_blocks, uint64_t n_each,
                                 int32_t op, float factor, int32_t n_threads) {
  if (!src || !dst || n_blocks < 0 || op < 0 || op > 2) return fail(WE_ERR_INVALID, "we_convert_be_f32: bad argument");
  if (n_blocks == 0 || n_each == 0) return WE_OK;
  if (n_threads <= 0) {
    cpu_set_t set;
    CPU_ZERO(&set);
    const int avail = (sched_getaffinity(0, sizeof(set), &set) == 0) ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
    n_threads = std::max(1, std::min(16, avail));
  }
  const uint64_t slice = 1u << 16;  // values per task
  const uint64_t per_block = (n_each + slice - 1) / slice;
  const uint64_t n_tasks = per_block * (uint64_t)n_blocks;
  n_threads = (int)std::min<uint64_t>((uint64_t)n_threads, n_tasks);
  std::atomic<uint64_t> next{0};
  const bool avx2 = __builtin_cpu_supports("avx2");
  auto work = [&]() {
    for (;;) {
      const uint64_t t = next.fetch_add(1);
      if (t >= n_tasks) break;
      const uint64_t b = t / per_block, i0 = (t % per_block) * slice, i1 = std::min(n_each, i0 + slice);
      const unsigned char* s = (const unsigned char*)src[b] + 4 * i0;
      float* d = dst[b] + i0;
      if (avx2) we_convert_avx2(s, d, i1 - i0, op, factor);
      else we_convert_base(s, d, i1 - i0, op, factor);
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < n_threads; ++i) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return WE_OK;
}

extern "C" int we_n_bins(we_ctx* c) { return c ? c->B : 0; }
extern "C" int we_n_heavy(we_ctx* c) { return c ? c->H : 0; }
extern "C" int we_n_donors(we_ctx* c) { return c ? c->ND : 0; }
extern "C" int we_copy_heavy_idx(we_ctx* c, int32_t* out) {
  if (!c || !out) return fail(WE_ERR_INVALID, "null argument");
  memcpy(out, c->heavy_idx.data(), c->heavy_idx.size() * sizeof(int));
  return WE_OK;
}
extern "C" int we_copy_donors(we_ctx* c, int32_t* dx, int32_t* dh) {
  if (!c || !dx || !dh) return fail(WE_ERR_INVALID, "null argument");
  memcpy(dx, c->don_x_atom.data(), c->don_x_atom.size() * sizeof(int));
  memcpy(dh, c->don_h_atom.data(), c->don_h_atom.size() * sizeof(int));
  return WE_OK;
}

extern "C" int we_copy_cov(we_ctx* c, double* cov_sum, uint64_t* cov_cnt) {
  if (!c || !cov_sum || !cov_cnt) return fail(WE_ERR_INVALID, "null argument");
  { const int rc = quiesce(c); if (rc) return rc; }
  CK(cudaMemcpy(cov_sum, c->d_cov_sum, (size_t)c->B * 18 * sizeof(double), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cov_cnt, c->d_cov_cnt, (size_t)c->B * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  return WE_OK;
}

extern "C" int we_snapshot_cov(we_ctx* c) {
  if (!c) return fail(WE_ERR_INVALID, "null context");
  CK(cudaSetDevice(c->device));
  int rc;
  // lazy-upload batches still waiting for their covariance kernel must run it first
  if ((rc = finish_lazy_all(c))) return rc;
  const int slot = (int)(c->snap_counter & 3);
  const size_t bs = (size_t)c->B * 18 * sizeof(double), bc = (size_t)c->B * sizeof(unsigned long long);
  if (!c->h_snap[slot]) {
    if ((rc = pin_alloc(c, &c->h_snap[slot], bs + bc))) return rc;
    CK(cudaEventCreateWithFlags(&c->ev_snap[slot], cudaEventDisableTiming));
  }
  // The tables are written by k_cov only.  When it runs on its own stream (zero-copy forces) the snapshot is
  // queued there, behind the last k_cov, and the compute stream - already busy with the next batches - is
  // not held up; otherwise everything is on the compute stream.
  cudaStream_t ss = c->cov_on_own_stream ? c->s_cov : c->s_comp;
  if (!c->ev_cov) CK(cudaEventCreateWithFlags(&c->ev_cov, cudaEventDisableTiming));
  if (c->cov_on_own_stream) {  // behind everything queued on the compute stream so far, without blocking it
    CK(cudaEventRecord(c->ev_cov, c->s_comp));
    CK(cudaStreamWaitEvent(c->s_cov, c->ev_cov, 0));
  } else {
    CK(cudaEventRecord(c->ev_cov, c->s_cov));
    CK(cudaStreamWaitEvent(c->s_comp, c->ev_cov, 0));
  }
  CK(cudaMemcpyAsync(c->h_snap[slot], c->d_cov_sum, bs, cudaMemcpyDeviceToHost, ss));
  CK(cudaMemcpyAsync(c->h_snap[slot] + bs, c->d_cov_cnt, bc, cudaMemcpyDeviceToHost, ss));
  CK(cudaEventRecord(c->ev_snap[slot], ss));
  return (int)(c->snap_counter++ & 0x7fffffff);
}

extern "C" int we_wait_cov(we_ctx* c, int32_t ticket, double* cov_sum, uint64_t* cov_cnt) {
  if (!c || !cov_sum || !cov_cnt) return fail(WE_ERR_INVALID, "null argument");
  if (ticket < 0 || (long long)ticket >= c->snap_counter || (long long)ticket + 4 < c->snap_counter)
    return fail(WE_ERR_INVALID, "snapshot ticket %d is not outstanding", ticket);
  const int slot = ticket & 3;
  CK(cudaEventSynchronize(c->ev_snap[slot]));
  const size_t bs = (size_t)c->B * 18 * sizeof(double), bc = (size_t)c->B * sizeof(unsigned long long);
  memcpy(cov_sum, c->h_snap[slot], bs);
  memcpy(cov_cnt, c->h_snap[slot] + bs, bc);
  return WE_OK;
}

extern "C" int we_device_tables(we_ctx* c, void** s, void** n) {
  if (!c) return fail(WE_ERR_INVALID, "null context");
  { const int rc = quiesce(c); if (rc) return rc; }  // the caller hands these to NCCL on its own stream
  if (s) *s = c->d_cov_sum;
  if (n) *n = c->d_cov_cnt;
  return WE_OK;
}

__global__ void k_compact(const WeLabelEntry* __restrict__ table, unsigned int n, WeLabelEntry* __restrict__ out, int* count) {
  const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (table[i].hash != 0ULL) {
    const int p = atomicAdd(count, 1);
    const uint4* src = reinterpret_cast<const uint4*>(table + i);
    uint4* dst = reinterpret_cast<uint4*>(out + p);
#pragma unroll
    for (int k = 0; k < 32; ++k) dst[k] = src[k];
  }
}

extern "C" int we_n_label_entries(we_ctx* c) {
  if (!c) return fail(WE_ERR_INVALID, "null context");
  { const int rc = quiesce(c); if (rc) return rc; }
  WeDiag d;
  int rc = read_diag(c, &d);
  if (rc) return rc;
  return (int)d.table_entries;
}

static int compact_entries(we_ctx* c, int n);
extern "C" int we_copy_label_entries(we_ctx* c, we_label_entry* out, int32_t capacity) {
  if (!c || !out) return fail(WE_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  const int n = we_n_label_entries(c);
  if (n < 0) return n;
  if (n > capacity) return fail(WE_ERR_INVALID, "capacity %d < %d entries", capacity, n);
  if (n == 0) return 0;
  const int rc = compact_entries(c, n);
  if (rc) return rc;
  CK(cudaMemcpy(out, c->d_compact, (size_t)n * sizeof(WeLabelEntry), cudaMemcpyDeviceToHost));
  return n;
}

// the table's occupied entries, packed into the context's compact buffer (kept between calls)
static int compact_entries(we_ctx* c, int n) {
  if ((size_t)n > c->compact_cap) {
    if (c->d_compact) cudaFree(c->d_compact);
    c->d_compact = nullptr;
    c->compact_cap = (size_t)n + (size_t)n / 4 + 1024;
    CK(cudaMalloc((void**)&c->d_compact, c->compact_cap * sizeof(WeLabelEntry)));
  }
  CK(cudaMemsetAsync(c->d_compact_n, 0, sizeof(int), c->s_comp));
  const unsigned int tn = c->table_mask + 1;
  if (n) { k_compact<<<(tn + 255) / 256, 256, 0, c->s_comp>>>(c->d_table, tn, c->d_compact, c->d_compact_n); c->launches++; }
  CK(cudaStreamSynchronize(c->s_comp));
  return 0;
}

extern "C" int we_label_entries_device(we_ctx* c, void** d_entries) {
  if (!c || !d_entries) return fail(WE_ERR_INVALID, "null argument");
  CK(cudaSetDevice(c->device));
  const int n = we_n_label_entries(c);
  if (n < 0) return n;
  const int rc = compact_entries(c, n);
  if (rc) return rc;
  *d_entries = c->d_compact;
  return n;
}

__global__ void k_merge(const WeLabelEntry* __restrict__ src, int n, WeLabelEntry* __restrict__ table,
                        unsigned int mask, WeDiag* diag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const WeLabelEntry& e = src[i];
  unsigned int s = (unsigned int)e.hash & mask;
  bool won = false;
  for (unsigned int probe = 0; probe <= mask; ++probe) {
    const unsigned long long old = atomicCAS(&table[s].hash, 0ULL, e.hash);
    if (old == 0ULL) { won = true; break; }
    if (old == e.hash) break;
    s = (s + 1) & mask;
  }
  WeLabelEntry* t = table + s;
  const unsigned long long old2 = atomicCAS(&t->hash2, 0ULL, e.hash2);
  if (old2 != 0ULL && old2 != e.hash2) { we_set_error(diag, 5, -1, -1); return; }
  if (won) {
    t->nearest_resid = e.nearest_resid; t->nearest_resname_id = e.nearest_resname_id;
    t->n_labels = e.n_labels; t->n_w = e.n_w;
    for (int k = 0; k < 32; ++k) t->labels[k] = e.labels[k];
    atomicAdd(&diag->table_entries, 1ULL);
  }
  atomicAdd(&t->shell_count, e.shell_count);
  atomicMax(&t->first_seen_inv, e.first_seen_inv);
  for (int k = 0; k < WE_MAX_NBR; ++k) {
    if (e.donates[k]) atomicAdd(&t->donates[k], e.donates[k]);
    if (e.accepts[k]) atomicAdd(&t->accepts[k], e.accepts[k]);
  }
}

extern "C" int we_merge_label_entries(we_ctx* c, const void* d_entries, int32_t n) {
  if (!c || (!d_entries && n > 0)) return fail(WE_ERR_INVALID, "null argument");
  if (n <= 0) return WE_OK;
  CK(cudaSetDevice(c->device));
  const int have = we_n_label_entries(c);
  if (have < 0) return have;
  int rc = grow_table_to(c, (unsigned long long)have + (unsigned long long)n + 1024);
  if (rc) return rc;
  k_merge<<<(n + 127) / 128, 128, 0, c->s_comp>>>((const WeLabelEntry*)d_entries, n, c->d_table, c->table_mask, c->d_diag);
  c->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->s_comp));
  return WE_OK;
}

// Frames of batches still in flight: wait for (only) the batches up to the one holding frame k.
static int drain_until(we_ctx* c, size_t have, int64_t k) {
  if (!c || k < 0 || (size_t)k < have || (unsigned long long)k >= c->frames_done) return 0;
  cudaSetDevice(c->device);
  for (int i = 0; i < WE_NBUF; ++i) {
    const int buf = (int)((c->chunk_counter + i) % WE_NBUF);
    int rc;
    if ((rc = finish_lazy(c, buf))) return rc;
    if ((rc = drain(c, buf))) return rc;
    if ((size_t)k < c->frame_records.size()) break;
  }
  return 0;
}

extern "C" int we_frame_record_count(we_ctx* c, int64_t k) {
  if (c && (c->opt.keep_records || c->opt.keep_shells)) { const int rc = drain_until(c, c->frame_records.size(), k); if (rc) return rc; }
  if (!c || k < 0 || k >= (int64_t)c->frame_records.size()) return fail(WE_ERR_INVALID, "frame ordinal out of range (keep_records set?)");
  return (int)(c->frame_records[(size_t)k].size() / 2);
}
extern "C" int we_copy_frame_records(we_ctx* c, int64_t k, int32_t* out, int32_t capacity) {
  const int n = we_frame_record_count(c, k);
  if (n < 0) return n;
  if (n > capacity) return fail(WE_ERR_INVALID, "capacity too small");
  memcpy(out, c->frame_records[(size_t)k].data(), (size_t)n * 2 * sizeof(int));
  return n;
}

// the recorded waters of `n_frames` consecutive submitted frames in one call: counts[f] pairs per frame,
// concatenated in `out` (capacity in pairs); returns the total number of pairs
extern "C" int64_t we_copy_frame_records_range(we_ctx* c, int64_t k0, int32_t n_frames, int32_t* counts, int32_t* out,
                                               int64_t capacity) {
  if (!c || !counts || (!out && capacity > 0) || n_frames < 0 || k0 < 0) return fail(WE_ERR_INVALID, "bad argument");
  if (n_frames == 0) return 0;
  if (c->opt.keep_records || c->opt.keep_shells) { const int rc = drain_until(c, c->frame_records.size(), k0 + n_frames - 1); if (rc) return rc; }
  if (k0 + n_frames > (int64_t)c->frame_records.size()) return fail(WE_ERR_INVALID, "frame ordinal out of range (keep_records set?)");
  int64_t total = 0;
  for (int f = 0; f < n_frames; ++f) {
    const std::vector<int>& r = c->frame_records[(size_t)(k0 + f)];
    counts[f] = (int32_t)(r.size() / 2);
    if (total + counts[f] > capacity) return fail(WE_ERR_INVALID, "capacity too small");
    if (!r.empty()) memcpy(out + 2 * total, r.data(), r.size() * sizeof(int));
    total += counts[f];
  }
  return total;
}

extern "C" int we_copy_frame_shells(we_ctx* c, int64_t k, int32_t* out, int32_t capacity) {
  if (c && c->opt.keep_shells) { const int rc = drain_until(c, c->frame_shells.size(), k); if (rc) return rc; }
  if (!c || k < 0 || k >= (int64_t)c->frame_shells.size()) return fail(WE_ERR_INVALID, "frame ordinal out of range (keep_shells set?)");
  const int n = (int)(c->frame_shells[(size_t)k].size() / (WE_MAX_NBR + 1));
  if (n > capacity) return fail(WE_ERR_INVALID, "capacity too small");
  memcpy(out, c->frame_shells[(size_t)k].data(), c->frame_shells[(size_t)k].size() * sizeof(int));
  return n;
}

static const void* dbg_field(we_ctx* c, int64_t k, int field, int64_t* bytes) {
  if (!c || k < 0 || k >= (int64_t)c->debug.size()) return nullptr;
  FrameDebug& D = c->debug[(size_t)k];
  switch (field) {
    case WE_DBG_SHELL_N: *bytes = D.shell_n.size() * 4; return D.shell_n.data();
    case WE_DBG_SHELL_IDX: *bytes = D.shell_idx.size() * 4; return D.shell_idx.data();
    case WE_DBG_NEAREST: *bytes = D.nn.size() * 4; return D.nn.data();
    case WE_DBG_FLAGS: *bytes = D.flags.size() * 4; return D.flags.data();
    case WE_DBG_ACCEPTOR: *bytes = D.acceptor.size() * 4; return D.acceptor.data();
    case WE_DBG_NBR_IDX: *bytes = D.nbr_idx.size() * 4; return D.nbr_idx.data();
    case WE_DBG_NBR_DIST: *bytes = D.nbr_dist.size() * 8; return D.nbr_dist.data();
    case WE_DBG_FT: *bytes = D.ft.size() * 8; return D.ft.data();
    case WE_DBG_REC: *bytes = D.rec.size() * 4; return D.rec.data();
    default: return nullptr;
  }
}
extern "C" int64_t we_debug_size(we_ctx* c, int64_t k, int32_t field) {
  int64_t bytes = 0;
  if (!dbg_field(c, k, field, &bytes)) { fail(WE_ERR_INVALID, "no debug data for frame %lld field %d (keep_debug set?)", (long long)k, field); return -1; }
  return bytes;
}
extern "C" int we_debug_copy(we_ctx* c, int64_t k, int32_t field, void* out, int64_t bytes) {
  int64_t have = 0;
  const void* p = dbg_field(c, k, field, &have);
  if (!p && have == 0 && !(c && k >= 0 && k < (int64_t)c->debug.size())) return fail(WE_ERR_INVALID, "no debug data");
  if (bytes < have) return fail(WE_ERR_INVALID, "buffer too small");
  if (have) memcpy(out, p, (size_t)have);
  return WE_OK;
}

// ---------------------------------------------------------------------------
// device-executed scalar hooks for the parity tests
__global__ void k_test_angle(const float* abc, const float* dim, int n, float* out) {
  __shared__ double s_l[32];
  __shared__ unsigned long long s_e[32];
  if (threadIdx.x < 32) { s_l[threadIdx.x] = ((const double*)we_d_log2tab)[threadIdx.x]; s_e[threadIdx.x] = we_d_exp2tab[threadIdx.x]; }
  __syncthreads();
  WePowTables PT{s_l, s_e};
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = abc + (size_t)i * 9;
  float d3[3] = {dim[0], dim[1], dim[2]};
  out[i] = we_angle_cos(p, p + 3, p + 6, d3, PT);
}
__global__ void k_test_powf2(const float* x, int n, float* out) {
  __shared__ double s_l[32];
  __shared__ unsigned long long s_e[32];
  if (threadIdx.x < 32) { s_l[threadIdx.x] = ((const double*)we_d_log2tab)[threadIdx.x]; s_e[threadIdx.x] = we_d_exp2tab[threadIdx.x]; }
  __syncthreads();
  WePowTables PT{s_l, s_e};
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = we_powf2(x[i], PT);
}
__global__ void k_test_pow2(const double* x, int n, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = we_pow2(x[i]);
}
__global__ void k_test_dist(const float* pairs, WeBox b, int n, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = pairs + (size_t)i * 6;
  if (b.triclinic) {
    float r[3], c[3];
    we_wrap_tric(p[0], p[1], p[2], b.m, b.wfac, r);
    we_wrap_tric(p[3], p[4], p[5], b.m, b.wfac, c);
    out[i] = we_dist_tric(r[0], r[1], r[2], c[0], c[1], c[2], b);
  } else {
    out[i] = we_dist_ortho(p[0], p[1], p[2], p[3], p[4], p[5], b);
  }
}

__global__ void k_test_eig3(const double* t, int n, int flavour, double* out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double wr[3], vr[9], ax[3][3], moi[3];
  const int st = we_dgeev3(t + (size_t)i * 9, wr, vr, flavour);
  we_principal_axes(t + (size_t)i * 9, ax, moi, flavour);
  double* o = out + (size_t)i * 25;
  for (int k = 0; k < 3; ++k) o[k] = wr[k];
  for (int k = 0; k < 9; ++k) o[3 + k] = vr[k];
  for (int k = 0; k < 9; ++k) o[12 + k] = ax[k / 3][k % 3];
  for (int k = 0; k < 3; ++k) o[21 + k] = moi[k];
  o[24] = (double)st;
}

template <typename Fn>
static int run_test(int device, const void* in, size_t in_bytes, void* out, size_t out_bytes, Fn launch) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(WE_ERR_CUDA, "no CUDA device available");
  CK(cudaSetDevice(device));
  void *d_in = nullptr, *d_out = nullptr;
  CK(cudaMalloc(&d_in, in_bytes ? in_bytes : 1));
  CK(cudaMalloc(&d_out, out_bytes ? out_bytes : 1));
  CK(cudaMemcpy(d_in, in, in_bytes, cudaMemcpyHostToDevice));
  launch(d_in, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, out_bytes, cudaMemcpyDeviceToHost);
  cudaFree(d_in);
  cudaFree(d_out);
  if (e != cudaSuccess) return fail(WE_ERR_CUDA, "test kernel failed: %s", cudaGetErrorString(e));
  return WE_OK;
}

extern "C" int we_test_angle_cos(int32_t device, const float* abc, const float* dim3v, int32_t n, float* out) {
  std::vector<float> in((size_t)n * 9 + 3);
  memcpy(in.data(), dim3v, 12);
  memcpy(in.data() + 3, abc, (size_t)n * 36);
  return run_test(device, in.data(), in.size() * 4, out, (size_t)n * 4, [&](void* di, void* dout) {
    k_test_angle<<<(n + 127) / 128, 128>>>((const float*)di + 3, (const float*)di, n, (float*)dout);
  });
}
extern "C" int we_test_powf2(int32_t device, const float* x, int32_t n, float* out) {
  return run_test(device, x, (size_t)n * 4, out, (size_t)n * 4, [&](void* di, void* dout) {
    k_test_powf2<<<(n + 127) / 128, 128>>>((const float*)di, n, (float*)dout);
  });
}
extern "C" int we_test_pow2(int32_t device, const double* x, int32_t n, double* out) {
  return run_test(device, x, (size_t)n * 8, out, (size_t)n * 8, [&](void* di, void* dout) {
    k_test_pow2<<<(n + 127) / 128, 128>>>((const double*)di, n, (double*)dout);
  });
}
extern "C" int we_test_distance(int32_t device, const float* pairs, const float* dims6, int32_t n, double* out) {
  WeBox b;
  if (make_box(dims6, 7.0f, 1 << 20, &b)) return fail(WE_ERR_INVALID, "invalid box");
  return run_test(device, pairs, (size_t)n * 24, out, (size_t)n * 8, [&](void* di, void* dout) {
    k_test_dist<<<(n + 127) / 128, 128>>>((const float*)di, b, n, (double*)dout);
  });
}
extern "C" int we_test_eig3(int32_t device, const double* tensors, int32_t n, int32_t flavour, double* out) {
  return run_test(device, tensors, (size_t)n * 72, out, (size_t)n * 25 * 8, [&](void* di, void* dout) {
    k_test_eig3<<<(n + 127) / 128, 128>>>((const double*)di, n, flavour, (double*)dout);
  });
}

// recipes/forces_torques.py:14-93 for `multiple_UAs` molecules handed over as compact arrays: one launch,
// one thread per molecule and per united atom (we_multi_ua.cuh).
extern "C" int we_molecule_force_torque_multi(int32_t device, const float* pos, const float* frc, const double* masses,
                                              int32_t n_pool, const int32_t* mol_ptr, const int32_t* mol_atoms, int32_t n_mol,
                                              const int32_t* ua_ptr, const int32_t* ua_atom, const int32_t* hv_ptr,
                                              const int32_t* hv, const int32_t* lt_ptr, const int32_t* lt,
                                              const int32_t* hin_ptr, const int32_t* hin, const float* box3,
                                              int32_t eig_flavour, double* out_mol, double* out_ua) {
  if (!pos || !frc || !masses || !mol_ptr || !mol_atoms || !ua_ptr || !ua_atom || !hv_ptr || !hv || !lt_ptr || !lt ||
      !hin_ptr || !hin || !box3 || !out_mol || !out_ua || n_mol < 0 || n_pool <= 0)
    return fail(WE_ERR_INVALID, "null argument");
  if (eig_flavour != WE_EIG_OPENBLAS_AVX512 && eig_flavour != WE_EIG_OPENBLAS_AVX2) return fail(WE_ERR_INVALID, "bad eig_flavour");
  if (n_mol == 0) return 0;
  const int n_ua = ua_ptr[n_mol];
  auto in_pool = [&](const int32_t* a, int n) { for (int i = 0; i < n; ++i) if (a[i] < 0 || a[i] >= n_pool) return false; return true; };
  auto csr_ok = [&](const int32_t* p, int n) { if (p[0] != 0) return false; for (int i = 0; i < n; ++i) if (p[i + 1] < p[i]) return false; return true; };
  if (!csr_ok(mol_ptr, n_mol) || !csr_ok(ua_ptr, n_mol) || !csr_ok(hv_ptr, n_ua) || !csr_ok(lt_ptr, n_ua) || !csr_ok(hin_ptr, n_ua))
    return fail(WE_ERR_INVALID, "malformed CSR offsets");
  for (int i = 0; i < n_mol; ++i) if (mol_ptr[i + 1] == mol_ptr[i]) return fail(WE_ERR_INVALID, "molecule %d has no atoms", i);
  if (!in_pool(mol_atoms, mol_ptr[n_mol]) || !in_pool(ua_atom, n_ua) || !in_pool(hv, hv_ptr[n_ua]) || !in_pool(lt, lt_ptr[n_ua]) ||
      !in_pool(hin, hin_ptr[n_ua]))
    return fail(WE_ERR_INVALID, "atom index outside the pool");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(WE_ERR_CUDA, "no CUDA device available; waterentropy_b200 has no CPU path");
  CK(cudaSetDevice(device));
  cudaError_t e = cudaSuccess;
  std::vector<void*> owned;
  auto up = [&](const void* src, size_t bytes) -> void* {
    void* d = nullptr;
    if (e == cudaSuccess) e = cudaMalloc(&d, bytes ? bytes : 4);
    if (e == cudaSuccess) owned.push_back(d);
    if (e == cudaSuccess && src && bytes) e = cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    return d;
  };
  WeMultiUA A;
  A.pos = (const float*)up(pos, (size_t)n_pool * 12);
  A.frc = (const float*)up(frc, (size_t)n_pool * 12);
  A.mass = (const double*)up(masses, (size_t)n_pool * 8);
  A.mol_ptr = (const int*)up(mol_ptr, ((size_t)n_mol + 1) * 4);
  A.mol_atoms = (const int*)up(mol_atoms, (size_t)mol_ptr[n_mol] * 4);
  A.ua_ptr = (const int*)up(ua_ptr, ((size_t)n_mol + 1) * 4);
  A.ua_atom = (const int*)up(ua_atom, (size_t)n_ua * 4);
  A.hv_ptr = (const int*)up(hv_ptr, ((size_t)n_ua + 1) * 4);
  A.hv = (const int*)up(hv, (size_t)hv_ptr[n_ua] * 4);
  A.lt_ptr = (const int*)up(lt_ptr, ((size_t)n_ua + 1) * 4);
  A.lt = (const int*)up(lt, (size_t)lt_ptr[n_ua] * 4);
  A.hin_ptr = (const int*)up(hin_ptr, ((size_t)n_ua + 1) * 4);
  A.hin = (const int*)up(hin, (size_t)hin_ptr[n_ua] * 4);
  A.box[0] = box3[0]; A.box[1] = box3[1]; A.box[2] = box3[2];
  A.n_mol = n_mol; A.n_ua = n_ua; A.eig_flavour = eig_flavour;
  double* d_mol = (double*)up(nullptr, (size_t)n_mol * 24 * 8);
  double* d_ua = (double*)up(nullptr, (size_t)(n_ua ? n_ua : 1) * 25 * 8);
  WeDiag* d_diag = (WeDiag*)up(nullptr, sizeof(WeDiag));
  WeDiag hd;
  memset(&hd, 0, sizeof(hd));
  if (e == cudaSuccess) e = cudaMemset(d_diag, 0, sizeof(WeDiag));
  if (e == cudaSuccess) {
    const int items = n_mol + n_ua;
    k_force_torque_multi<<<(items + 63) / 64, 64>>>(A, d_diag, d_mol, d_ua);
    e = cudaDeviceSynchronize();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out_mol, d_mol, (size_t)n_mol * 24 * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && n_ua) e = cudaMemcpy(out_ua, d_ua, (size_t)n_ua * 25 * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(&hd, d_diag, sizeof(WeDiag), cudaMemcpyDeviceToHost);
  for (void* d : owned) cudaFree(d);
  if (e != cudaSuccess) return fail(WE_ERR_CUDA, "we_molecule_force_torque_multi failed: %s", cudaGetErrorString(e));
  return (int)hd.eig_fallbacks;
}

// recipes/forces_torques.py:14-55 for molecules handed over as compact arrays (one launch, thread per
// molecule): the same device function k_cov uses.
extern "C" int we_molecule_force_torque(int32_t device, const float* pos, const float* frc, const double* masses,
                                        const int32_t* mol_ptr, int32_t n_mol, int32_t eig_flavour, double* out) {
  if (!pos || !frc || !masses || !mol_ptr || !out || n_mol < 0) return fail(WE_ERR_INVALID, "null argument");
  if (eig_flavour != WE_EIG_OPENBLAS_AVX512 && eig_flavour != WE_EIG_OPENBLAS_AVX2) return fail(WE_ERR_INVALID, "bad eig_flavour");
  if (n_mol == 0) return 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(WE_ERR_CUDA, "no CUDA device available; waterentropy_b200 has no CPU path");
  CK(cudaSetDevice(device));
  const size_t na = (size_t)mol_ptr[n_mol];
  for (int i = 0; i < n_mol; ++i) if (mol_ptr[i + 1] <= mol_ptr[i]) return fail(WE_ERR_INVALID, "molecule %d has no atoms", i);
  float *d_pos = nullptr, *d_frc = nullptr;
  double *d_m = nullptr, *d_out = nullptr;
  int* d_ptr = nullptr;
  WeDiag* d_diag = nullptr;
  cudaError_t e = cudaSuccess;
  auto al = [&](void** p, size_t bytes) { if (e == cudaSuccess) e = cudaMalloc(p, bytes ? bytes : 1); };
  al((void**)&d_pos, na * 12); al((void**)&d_frc, na * 12); al((void**)&d_m, na * 8);
  al((void**)&d_ptr, ((size_t)n_mol + 1) * 4); al((void**)&d_out, (size_t)n_mol * 24 * 8); al((void**)&d_diag, sizeof(WeDiag));
  WeDiag hd;
  memset(&hd, 0, sizeof(hd));
  if (e == cudaSuccess) e = cudaMemcpy(d_pos, pos, na * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_frc, frc, na * 12, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_m, masses, na * 8, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_ptr, mol_ptr, ((size_t)n_mol + 1) * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(d_diag, 0, sizeof(WeDiag));
  if (e == cudaSuccess) {
    k_force_torque<<<(n_mol + 127) / 128, 128>>>(d_pos, d_frc, d_m, d_ptr, n_mol, eig_flavour, d_diag, d_out);
    e = cudaDeviceSynchronize();
  }
  if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)n_mol * 24 * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(&hd, d_diag, sizeof(WeDiag), cudaMemcpyDeviceToHost);
  cudaFree(d_pos); cudaFree(d_frc); cudaFree(d_m); cudaFree(d_ptr); cudaFree(d_out); cudaFree(d_diag);
  if (e != cudaSuccess) return fail(WE_ERR_CUDA, "we_molecule_force_torque failed: %s", cudaGetErrorString(e));
  return (int)hd.eig_fallbacks;
}
