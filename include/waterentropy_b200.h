/*
 * waterentropy_b200 -- C ABI of the B200-native per-frame interfacial-solvent
 * analysis (drop-in for the hot path of jkalayan/waterEntropy 2.2.0).
 *
 * The reference is pure Python and has no FFI; its natural seam is the
 * per-frame worker that dask ships between processes plus the two recipe
 * entry points.  Each entry point below names the reference interface it
 * replaces (paths relative to /root/reference/waterEntropy/).  The Python
 * package `waterentropy_b200` binds these with ctypes (INTEGRATION.md shows the
 * stub a reference maintainer would add).
 *
 * Conventions: plain pointers and sizes, no torch / numpy types.  Every call
 * returns 0 on success or a negative WE_ERR_* code; `we_last_error()` gives the
 * text.  One context per GPU; a context is not thread-safe.  All host arrays
 * are owned by the caller; the library copies what it needs during the call
 * (frame buffers passed to `we_submit` must stay valid until the next
 * `we_sync`, or until `we_frames_retired` reports their batch as finished).
 */
#ifndef WATERENTROPY_B200_H
#define WATERENTROPY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WE_VERSION 202

/* recipes */
#define WE_RECIPE_INTERFACIAL 0 /* recipes/interfacial_solvent.py:273-345 (_entropy_per_step) */
#define WE_RECIPE_BULK 1        /* recipes/bulk_water.py:37-72 */

/* error codes (negative return values) */
#define WE_OK 0
#define WE_ERR_DUPLICATE_COORD -1 /* ValueError of maths/trig.py:40-42 */
#define WE_ERR_NO_NEIGHBOURS -2   /* implicit ValueError of maths/trig.py:36-38 */
#define WE_ERR_CAND_OVERFLOW -3   /* > 128 candidates inside the adaptive search radius */
#define WE_ERR_TABLE_FULL -4      /* label count table full: raise we_options.table_log2 */
#define WE_ERR_HASH_COLLISION -5  /* two label keys share a 128-bit hash (never observed) */
#define WE_ERR_INVALID -10        /* bad argument / topology */
#define WE_ERR_CUDA -11           /* CUDA runtime failure (no CPU fallback exists) */

#define WE_MAX_SHELL 25 /* analysis/RAD.py:39 */

/*
 * Frame-invariant system description.  Replaces what the reference
 * re-derives every frame through MDAnalysis selection strings
 * (utils/selections.py:8-105, recipes/interfacial_solvent.py:298-299); the
 * Python host builds it once (`waterentropy_b200.system.SystemTopology`).
 */
typedef struct we_topology {
    int32_t n_atoms;
    const double *masses;       /* [n_atoms] */
    const double *charges;      /* [n_atoms] */
    const int32_t *resid;       /* [n_atoms] */
    const int32_t *resname_id;  /* [n_atoms] index into the host's resname table (< 8192) */
    const int32_t *type_id;     /* [n_atoms] index into the host's atom-type table */
    const uint8_t *is_water;    /* [n_atoms] MDAnalysis "water" selection */
    int32_t n_bonds;
    const int32_t *bonds;       /* [n_bonds][2] */
    int32_t n_solute_ua;        /* utils/selections.py:34-56 + interfacial_solvent.py:39-43 */
    const int32_t *solute_ua;   /* [n_solute_ua] atom indices of solute united atoms */
    int32_t n_centres;          /* water heavy atoms that can be analysed */
    const int32_t *centres;     /* [n_centres] atom indices, ascending */
    const int32_t *frag_ptr;    /* [n_centres+1] CSR: atoms of each centre's molecule */
    const int32_t *frag_atoms;
    const int32_t *bin_of_atom; /* [n_atoms] covariance pair-bin of (resname, resid), -1 = none */
    const int32_t *centre_class;/* [n_atoms] class of a centre's resname, -1 = not a centre */
    int32_t n_pair_bins;
    int32_t n_classes;
    int32_t recipe;             /* WE_RECIPE_* */
} we_topology;

typedef struct we_options {
    int32_t device;          /* CUDA device ordinal */
    int32_t chunk_frames;    /* frames per kernel batch (0 = auto) */
    int32_t table_log2;      /* log2 of the initial capacity of the label count table (0 = 17); it grows by
                                device rehash when half full */
    int32_t keep_records;    /* 1: keep per-frame recorded-water lists (frame_solvent_indices) */
    int32_t keep_debug;      /* 1: keep all-centre shells / neighbours / acceptors / F,T of every frame */
    float search_radius;     /* first-pass candidate radius in A (0 = 6.6) */
    float max_box;           /* largest box edge expected in A (0 = take from first frame x 1.25) */
    int32_t keep_shells;     /* 1: also keep the RAD shell of every recorded water (get_interfacial_shells) */
    int32_t force_upload;    /* we_submit, how forces reach k_cov: 0 auto (zero-copy for the interfacial recipe
                                when the buffer is pinned, else whole frames), 1 whole frames, 2 recorded
                                waters only, gathered by the host one batch later, 3 zero-copy: k_cov reads
                                the recorded waters' forces straight from pinned host memory */
    int32_t shells_only;     /* 1: stop after the RAD shells of the interfacial waters (needs keep_shells, interfacial
                                recipe): what get_interfacial_shells evaluates (recipes/interfacial_solvent.py:104-146)
                                - no HBs, no labels, no neighbours' shells, no forces (frc may be NULL) */
    int32_t eig_flavour;     /* which OpenBLAS kernel family the principal-axis step replays (WE_EIG_*): the
                                reference's eigenvector signs are whatever numpy.linalg.eig returns on the
                                host it runs on (maths/transformations.py:121-124) */
    int32_t pos_upload;      /* we_submit, how positions reach the device: 0 / 1 whole frames (one DMA copy of
                                [n][n_atoms][3]); 2 packed (interfacial recipe, PINNED buffers): host threads gather
                                the heavy atoms (all the neighbour search reads: a third of a water box) into
                                pinned staging, only those are copied, and the few light atoms the path reads
                                (donor hydrogens of the centres whose hydrogen bonds are evaluated, atoms of the
                                recorded waters) are read straight from the caller's buffer by the device.
                                Same results; a third of the DMA bytes; measured SLOWER than whole frames on a
                                full x16 link, so never chosen automatically */
} we_options;

/* numpy.linalg.eig = LAPACK dgeev on OpenBLAS; its level-1/2 kernels round differently per CPU family,
 * which flips the eigenvector signs of ~1 molecule in 10^4 (csrc/we_eig3.cuh). */
#define WE_EIG_OPENBLAS_AVX512 0 /* "SkylakeX" kernels: Intel AVX-512 hosts (where the golden vectors were made) */
#define WE_EIG_OPENBLAS_AVX2 1   /* "Haswell" kernels: AVX2 hosts, AMD Zen */

/* Tuning knobs read from the environment by we_create (defaults are the measured optimum on B200):
 *   WE_COPY_GATE=1   hold the upload of a batch back until the previous batch's first neighbour pass starts
 *                    (default 0: uploads run back to back as soon as a slot is free)
 *   WE_COV_STREAM=0  k_cov on the compute stream; 1: own low-priority stream when it reads the forces zero-copy;
 *                    default 2: own stream always (it runs beside the next batch's cell-list kernels)
 *   WE_R_FIRST=a,b,c first candidate radius of the three k_rad passes as fractions of the cell radius (default 1)
 *   WE_PACK_THREADS=n  host threads that gather the heavy atoms of packed uploads (default: the cores this
 *                    process may run on, at most 16); WE_PACK_SPIN_US=t lets them spin t microseconds for the
 *                    next batch before they sleep
 *   WE_GRAPHS=0      launch every kernel individually (default 1: two CUDA graphs per batch for
 *                    host-buffer submissions)                                                          */
typedef struct we_ctx we_ctx;

typedef struct we_diag {
    uint64_t wide_searches;   /* centres that needed the 10 A wide search */
    uint64_t shrink_retries;
    uint64_t table_entries;
    uint64_t kernel_launches; /* kernels launched so far by this context */
    uint64_t frames_done;
    uint64_t recorded;        /* recorded (water, frame) pairs so far */
    uint64_t fast_path_mismatches; /* keep_debug contexts replay every float32 fast-path decision of the RAD
                                      test exactly and count contradictions: must stay 0 */
    uint64_t eig_fallbacks;   /* molecules whose inertia tensor took a LAPACK path that is not restated (exact
                                 zeros / equal moments; never a water): Jacobi axes with a fixed sign rule */
    uint64_t h2d_dma_bytes;   /* bytes the copy engine moved host -> device for frames (positions whole or packed,
                                 forces when uploaded whole, boxes, frame ids) */
    uint64_t light_atoms_fetched; /* packed uploads: light atoms the device read from the caller's pinned buffer */
    uint64_t host_pack_us;    /* packed uploads: wall time the submitting thread spent in the host gather */
} we_diag;

const char *we_last_error(void);
int we_version(void);

/* Build a context: uploads the topology, allocates per-GPU accumulators
 * (CovarianceCollection / HBLabelCollection state, entropy/convariances.py:10-19,
 * analysis/HB_labels.py:33-35) and the frame workspace. */
int we_create(const we_topology *topo, const we_options *opt, we_ctx **out);
void we_destroy(we_ctx *ctx);

/* Frames per kernel batch this context uses (we_options.chunk_frames, or the automatic size): a reader
 * that hands over whole batches keeps the device busiest. */
int we_batch_frames(we_ctx *ctx);

/* Analyse `n_frames` frames held in HOST memory (pinned or pageable):
 *   pos, frc  float32 [n_frames][n_atoms][3]   (ts.positions / ts.forces)
 *   dims      float32 [n_frames][6]            (ts.dimensions)
 *   frame_ids int64   [n_frames]               (ts.frame)
 * Replaces a `client.map(_entropy_per_step, ...)` over those frames
 * (recipes/interfacial_solvent.py:229-243) or the frame loop of
 * recipes/bulk_water.py:37.  Copies are asynchronous and double buffered;
 * accumulation happens on the device. */
int we_submit(we_ctx *ctx, const float *pos, const float *frc, const float *dims,
              const int64_t *frame_ids, int32_t n_frames);

/* Same, with pos / frc already resident in DEVICE memory (dims and frame_ids
 * stay on the host; they are 32 bytes per frame). */
int we_submit_device(we_ctx *ctx, const float *d_pos, const float *d_frc, const float *dims,
                     const int64_t *frame_ids, int32_t n_frames);

/* Input-buffer retirement: how many of the frames submitted so far no longer need their caller-owned
 * buffers (uploads done, and for zero-copy forces the covariance kernel has read them).  Lets a reader
 * recycle a small ring of pinned staging buffers instead of holding the trajectory until `we_sync`
 * (the reference holds one frame at a time, recipes/interfacial_solvent.py:285).  `we_frames_retired`
 * polls; `we_wait_frames_retired` blocks until at least `n_frames` are retired. */
int64_t we_frames_retired(we_ctx *ctx);
int we_wait_frames_retired(we_ctx *ctx, int64_t n_frames);

/* Wait for all submitted work; surfaces the reference's per-frame exceptions
 * as WE_ERR_DUPLICATE_COORD / WE_ERR_NO_NEIGHBOURS. */
int we_sync(we_ctx *ctx);
int we_get_diag(we_ctx *ctx, we_diag *out);
int we_error_location(we_ctx *ctx, int64_t *frame_id, int32_t *atom);

/* ---- trajectory ingestion (host only, no context) -----------------------------------
 * Big-endian float32 blocks of a trajectory file - the per-frame records of an AMBER NetCDF-3 file, the
 * XDR position / force blocks of a GROMACS TRR frame - converted to native float32 straight into the (pinned)
 * frame buffers we_submit takes, with the unit conversion MDAnalysis' readers apply, in float32 like theirs
 * (NCDFReader: forces *= 4.184; TRRReader: positions *= 10, forces /= 10).  This is the part of
 * `system.trajectory[index]` (recipes/interfacial_solvent.py:285) that touches every byte; host threads share it.
 *   src[b], dst[b]: block b (n_each values each);  op: 0 copy, 1 multiply by factor, 2 divide by factor;
 *   n_threads: 0 = the cores this process may run on (at most 16). */
int we_convert_be_f32(const void *const *src, float *const *dst, int32_t n_blocks, uint64_t n_each,
                      int32_t op, float factor, int32_t n_threads);

/* ---- results ------------------------------------------------------------------
 * The read-out calls below wait for every batch submitted so far (they synchronise the context's own
 * streams), so they never return partially accumulated tables; per-frame errors are reported by we_sync. */

/* Dense covariance table: sums over recorded waters of the two 3x3 outer
 * products (force, torque; recipes/forces_torques.py:50-53) and the counts
 * (entropy/convariances.py:21-84 keeps running means; mean = sum / count).
 *   cov_sum [n_bins][2][3][3] float64, cov_cnt [n_bins] uint64,
 *   n_bins = n_pair_bins * n_classes (interfacial) or n_classes (bulk). */
int we_n_bins(we_ctx *ctx);
int we_copy_cov(we_ctx *ctx, double *cov_sum, uint64_t *cov_cnt);
/* Stream-ordered snapshot of the covariance tables: `we_snapshot_cov` enqueues a copy into pinned
 * memory behind everything submitted so far and returns a ticket (>= 0) without waiting;
 * `we_wait_cov` waits for that copy only, so later batches keep running while the host reads the
 * result of an earlier one (4 snapshots may be outstanding). */
int we_snapshot_cov(we_ctx *ctx);
int we_wait_cov(we_ctx *ctx, int32_t ticket, double *cov_sum, uint64_t *cov_cnt);
/* device pointers of the same tables, for an NCCL all-reduce across ranks
 * (replaces CovarianceCollection.merge, entropy/convariances.py:86-108) */
int we_device_tables(we_ctx *ctx, void **d_cov_sum, void **d_cov_cnt);

/* Vibrational entropy of every covariance bin, computed on the device from the (all-reduced) tables:
 * entropy/vibrations.py:142-192 with diagonalise=True.  `conversion` = Vibrations.mda_conversion()
 * (entropy/vibrations.py:117-128).  out float64 [n_bins][2][3] = translational (force) and rotational (torque)
 * terms in J / (K mol); bins without samples give zeros. */
int we_vib_entropy(we_ctx *ctx, double temperature, double conversion, double *out);

/* Label count table (HBLabelCollection, analysis/HB_labels.py:15-156).
 * Each entry is 512 bytes, layout `we_label_entry`. */
typedef struct we_label_entry {
    uint64_t hash, hash2;
    uint64_t first_seen_inv;  /* ~min(frame_id * n_atoms + atom): N_w is frozen by the first insert */
    uint64_t shell_count;
    int32_t nearest_resid;      /* -1 for the bulk recipe */
    int32_t nearest_resname_id; /* bulk: the water resname */
    int32_t n_labels;
    int32_t n_w;
    uint16_t labels[32];        /* sorted codes: (prefix << 13) | resname_id; prefix 0:"0_" 1:"" 2:"1_" 3:"2_" 4:"X_" */
    uint32_t donates[25];       /* per label position (first occurrence) */
    uint32_t accepts[25];
    uint32_t pad[50];
} we_label_entry;
int we_n_label_entries(we_ctx *ctx);
int we_copy_label_entries(we_ctx *ctx, we_label_entry *out, int32_t capacity);
/* Cross-rank merge on the device (replaces HBLabelCollection.merge, analysis/HB_labels.py:158-296):
 * `we_label_entries_device` compacts this rank's entries into a device buffer owned by the context
 * (valid until the next call) and returns their number; after an all-gather of those buffers
 * `we_merge_label_entries` adds another rank's entries (device pointer) into this context's table:
 * counts add, N_w / first-seen follow the "first insert wins" rule. */
int we_label_entries_device(we_ctx *ctx, void **d_entries);
int we_merge_label_entries(we_ctx *ctx, const void *d_entries, int32_t n);

/* Per-frame recorded waters (frame_solvent_indices, recipes/interfacial_solvent.py:69-88);
 * needs we_options.keep_records.  `k` counts submitted frames from 0.
 * out = int32 [count][2] = (water atom index, nearest non-like atom index or -1), ascending. */
int we_frame_record_count(we_ctx *ctx, int64_t k);
int we_copy_frame_records(we_ctx *ctx, int64_t k, int32_t *out, int32_t capacity);
/* The same for `n_frames` consecutive frames in one call: counts int32 [n_frames] (pairs per frame), out
 * int32 [capacity][2] receives the frames' pairs back to back; returns the total number of pairs. */
int64_t we_copy_frame_records_range(we_ctx *ctx, int64_t k0, int32_t n_frames, int32_t *counts, int32_t *out,
                                    int64_t capacity);
/* RAD shells of the recorded waters of frame k, same order as we_copy_frame_records
 * (needs we_options.keep_shells): out = int32 [count][26] = (n, member 0 .. member 24). */
int we_copy_frame_shells(we_ctx *ctx, int64_t k, int32_t *out, int32_t capacity);

/* All-centre per-frame state (needs we_options.keep_debug); used by
 * get_interfacial_shells (recipes/interfacial_solvent.py:104-146) and by the
 * parity tests.  field: */
#define WE_DBG_SHELL_N 0   /* int32 [n_heavy] */
#define WE_DBG_SHELL_IDX 1 /* int32 [n_heavy][25] atom indices, distance order */
#define WE_DBG_NEAREST 2   /* int32 [n_heavy] nearest non-like atom or -1 */
#define WE_DBG_FLAGS 3     /* int32 [n_heavy] bit0 gate, bits 8.. status */
#define WE_DBG_ACCEPTOR 4  /* int32 [n_donors] acceptor atom or -(1+status) */
#define WE_DBG_NBR_IDX 5   /* int32 [n_heavy][25] sorted candidates */
#define WE_DBG_NBR_DIST 6  /* float64 [n_heavy][25] */
#define WE_DBG_FT 7        /* float64 [n_records][6] F(3), torque(3) in record order */
#define WE_DBG_REC 8       /* int32 [n_records][4] atom, nearest, bin, 0 (unsorted) */
int we_n_heavy(we_ctx *ctx);
int we_n_donors(we_ctx *ctx);
int we_copy_heavy_idx(we_ctx *ctx, int32_t *out);
int we_copy_donors(we_ctx *ctx, int32_t *don_x_atom, int32_t *don_h_atom);
int64_t we_debug_size(we_ctx *ctx, int64_t k, int32_t field);
int we_debug_copy(we_ctx *ctx, int64_t k, int32_t field, void *out, int64_t bytes);

/* ---- measurement ------------------------------------------------------------ */
/* kernel ids for we_kernel_times */
#define WE_K_BIN 0
#define WE_K_SCAN 1
#define WE_K_SCATTER 2
#define WE_K_RAD 3
#define WE_K_HB 4
#define WE_K_LIST1 5  /* k_list1: pass-1 list from the interfacial marks */
#define WE_K_LABEL 6
#define WE_K_COV 7
#define WE_K_RAD1 8   /* k_rad, pass 1: interfacial waters */
#define WE_K_RAD2 9   /* k_rad, pass 2: their shell members */
#define WE_K_LIST2 10 /* k_seed + k_list2 (pass-2 and HB lists) */
#define WE_K_GATE 11  /* k_gate (+ k_grplist): burial pre-stage of the gating pass, by cell group */
#define WE_K_FETCH 12 /* k_fetch_light (packed uploads): light atoms read from the caller's pinned buffer */
#define WE_N_KERNELS 13
/* Record CUDA events around every kernel launch on the compute stream (the
 * reference only has wall-clock prints, cli/runWaterEntropy.py:28-29,54). */
int we_profile(we_ctx *ctx, int32_t enable);
int we_kernel_times(we_ctx *ctx, double *ms /*[WE_N_KERNELS]*/, uint64_t *launches /*[WE_N_KERNELS]*/);
/* Device-time span on the compute stream (CUDA events, not wall clock). */
int we_span_start(we_ctx *ctx);
int we_span_stop(we_ctx *ctx, double *ms);

/* entropy/orientations.py:298-370 (`Orientations.get_orientational_entropy_from_dict` over
 * `WaterOrientCalculator`, :22-238) on the device, from this context's label table (after the cross-rank merge
 * on rank 0): one row per group of label entries, the shell types of a group folded in Python's sorted order of
 * their label tuples as upstream does.  label_rank uint16 [65536]: rank of every label code (prefix << 13 |
 * resname id) in Python's string order of the label strings (topology only; unused codes: anything).
 * by_resname = 0: groups are (nearest resid, nearest resname) = `resid_labelled_Sorient`; 1: groups are the
 * nearest resname alone with the entries of all resids merged = `resname_labelled_Sorient` (and both tables of
 * the bulk recipe).  out_keys int32 [max_groups][2] = resid, resname id of each group (unordered);
 * out_rows float64 [max_groups][8] = Sor, count, N_c, N_w, Nc_eff, pbias, Sor_Nc, Sor_Nw (:361-370). */
int we_orient_entropy(we_ctx *ctx, const uint16_t *label_rank, int32_t by_resname, int32_t max_groups,
                      int32_t *out_keys, double *out_rows, int32_t *n_groups);

/* recipes/forces_torques.py:14-55 (`get_forces_torques`, single-united-atom branch) for `n_mol` molecules
 * given as compact arrays: pos / frc float32 [n_atoms_total][3], masses float64 [n_atoms_total], mol_ptr
 * int32 [n_mol + 1] (CSR).  out float64 [n_mol][24] = rotated mass-weighted force (3), inertia-weighted torque
 * (3), force covariance matrix (9), torque covariance matrix (9) (maths/transformations.py:67-79, halve = 0.5).
 * Stateless; returns the number of molecules that needed the Jacobi fallback (>= 0) or an error. */
int we_molecule_force_torque(int32_t device, const float *pos, const float *frc, const double *masses,
                             const int32_t *mol_ptr, int32_t n_mol, int32_t eig_flavour, double *out);

/* recipes/forces_torques.py:14-93 for `multiple_UAs` molecules (more than one heavy atom in one residue;
 * maths/transformations.py:82-353: united-atom inertia tensor, get_bonded_axes, get_custom_axes,
 * get_custom_PI_MOI), `n_mol` molecules given as compact arrays over an atom POOL (the molecules' atoms plus
 * every atom bonded to one of their united atoms): pos / frc float32 [n_pool][3], masses float64 [n_pool];
 * mol_ptr / mol_atoms: CSR of pool indices per molecule, ascending atom index; ua_ptr [n_mol + 1]: first united
 * atom of each molecule in ua_atom (pool indices of its heavy atoms, ascending); per united atom three CSR
 * lists of pool indices, each ascending by atom index: hv = bonded heavy atoms (2 <= mass <= 999), lt = bonded
 * light atoms (1 <= mass <= 1.1), hin = the light ones that belong to the molecule; box3 = dimensions[:3].
 * out_mol float64 [n_mol][24]: whole-molecule force (3), torque (3), covariance matrices (9 + 9, no halving);
 * out_ua float64 [n_ua][25]: the same in the united atom's bonded-axes frame, then 1.0 / 0.0 = has a frame
 * (an atom bonded to exactly one other heavy atom and no hydrogen has none, forces_torques.py:66).
 * Stateless; returns the number of Jacobi fallbacks (>= 0) or an error. */
int we_molecule_force_torque_multi(int32_t device, const float *pos, const float *frc, const double *masses,
                                   int32_t n_pool, const int32_t *mol_ptr, const int32_t *mol_atoms, int32_t n_mol,
                                   const int32_t *ua_ptr, const int32_t *ua_atom, const int32_t *hv_ptr,
                                   const int32_t *hv, const int32_t *lt_ptr, const int32_t *lt,
                                   const int32_t *hin_ptr, const int32_t *hin, const float *box3,
                                   int32_t eig_flavour, double *out_mol, double *out_ua);

/* Clear the accumulators (start a new analysis with the same topology). */
int we_reset(we_ctx *ctx);

/* Scalar arithmetic hooks (device-executed) so the float32 angle / powf
 * replay and the distance formula can be pinned directly in tests:
 *   maths/trig.py:103-122 (get_angle), MDAnalysis capped_distance formula,
 *   and the float64 `x ** 2` of neighbours/RAD.py:53-57 / analysis/HB.py:96 (libm pow(x, 2.0)). */
int we_test_angle_cos(int32_t device, const float *abc /*[n][9]*/, const float *dim3, int32_t n, float *out);
int we_test_powf2(int32_t device, const float *x, int32_t n, float *out);
int we_test_pow2(int32_t device, const double *x, int32_t n, double *out);
int we_test_distance(int32_t device, const float *pairs /*[n][6]*/, const float *dims6, int32_t n, double *out);
/* numpy.linalg.eig / MDAnalysis principal_axes of symmetric 3x3 tensors (maths/transformations.py:113-133):
 * out [n][25] = wr(3), vr(9: eigenvector k at [3k..3k+2]), principal axes (9, rows), MOI_axis(3), status. */
int we_test_eig3(int32_t device, const double *tensors /*[n][9]*/, int32_t n, int32_t flavour, double *out);

#ifdef __cplusplus
}
#endif
#endif
