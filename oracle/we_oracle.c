/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the waterEntropy per-frame hot path.
 *
 * Plain-C restatement of the reference's *algorithm* (brute-force neighbour
 * search per centre, stable sort, O(k^2) RAD test, sequential HB acceptor
 * scan).  It is the checker for the CUDA path and the `cpu_baseline` / `--impl
 * reference` arm of bench.py; nothing in the product package may link, import
 * or call it.  Parity is PINNED: `tests/test_oracle_golden.py` checks it
 * against vectors produced by the unmodified reference (run through
 * oracle/mda_shim) and against the literal goldens in the reference's tests.
 *
 * Must be compiled with  -O2 -fno-builtin -ffp-contract=off  so that
 *   - powf/pow go to glibc libm (NumPy scalar `x**2` is libm powf/pow, NOT x*x),
 *   - no FMA contraction changes the float32/float64 rounding sequence.
 *
 * Reference lines restated (all under /root/reference/waterEntropy/):
 *   maths/trig.py:11-42    get_neighbourlist      -> sorted_neighbours()
 *   maths/trig.py:45-68    get_sorted_neighbours  -> candidate rule (heavy, not i, not bonded to i)
 *   maths/trig.py:103-122  get_angle              -> angle_cos()   (float32!)
 *   analysis/RAD.py:12-71  get_RAD_neighbours     -> rad_shell()
 *   analysis/HB.py:96-139  get_shell_HB_acceptors -> we_oracle_hb()
 *   analysis/HB.py:66-93   get_HB_terms           -> inside we_oracle_hb()
 *   recipes/interfacial_solvent.py:47-56  "water among the 20 nearest" gate
 * MDAnalysis arithmetic (not under /root/reference; mdanalysis>=2.10):
 *   lib/src/calc_distances.h _calc_distance_array_ortho + minimum_image
 *   (float32 difference widened to double, round()-based minimum image);
 *   triclinic: wrap into cell + 27-image search (parity UNPINNED).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>

#define WE_MAX_NBR 25
#define WE_GATE_NBR 20
#define WE_CUTOFF 10.0

enum { ST_OK = 0, ST_DUPLICATE = 1, ST_NO_NEIGHBOURS = 2 };

typedef struct {
    int triclinic;
    float box[3];      /* dimensions[:3] as float32 */
    float inv[3];      /* (float)(1.0/(double)box[i]) */
    float m[9];        /* float32 lower-triangular box matrix (triclinic only) */
} box_t;

/* MDAnalysis lib.mdamath.triclinic_vectors, float32 result */
static void triclinic_vectors(const float *dim, float *m)
{
    double lx = dim[0], ly = dim[1], lz = dim[2];
    double alpha = dim[3], beta = dim[4], gamma = dim[5];
    memset(m, 0, 9 * sizeof(float));
    double cos_a = (alpha == 90.0) ? 0.0 : cos(alpha * (M_PI / 180.0));
    double cos_b = (beta == 90.0) ? 0.0 : cos(beta * (M_PI / 180.0));
    double cos_g, sin_g;
    if (gamma == 90.0) { cos_g = 0.0; sin_g = 1.0; }
    else { double g = gamma * (M_PI / 180.0); cos_g = cos(g); sin_g = sin(g); }
    m[0] = (float)lx;
    m[3] = (float)(ly * cos_g);
    m[4] = (float)(ly * sin_g);
    m[6] = (float)(lz * cos_b);
    m[7] = (float)(lz * (cos_a - cos_b * cos_g) / sin_g);
    m[8] = (float)sqrt(lz * lz - (double)m[6] * (double)m[6] - (double)m[7] * (double)m[7]);
}

static void make_box(const float *dims, box_t *b)
{
    for (int t = 0; t < 3; ++t) {
        b->box[t] = dims[t];
        b->inv[t] = (float)(1.0 / (double)dims[t]);
    }
    b->triclinic = !(dims[3] == 90.0f && dims[4] == 90.0f && dims[5] == 90.0f);
    if (b->triclinic) triclinic_vectors(dims, b->m);
}

/* wrap one coordinate into the primary triclinic cell: c axis, then b axis, then a axis, each
 * against the slanted bounds of the parallelepiped (MDAnalysis calc_distances.h::_triclinic_pbc) */
static void wrap_triclinic(const float *in, const float *m, float *out)
{
    double c0 = in[0], c1 = in[1], c2 = in[2], s;
    const double b_z = (double)m[7] / (double)m[8];
    const double a_y = (double)m[3] / (double)m[4];
    const double a_z = ((double)m[4] * (double)m[6] - (double)m[3] * (double)m[7]) / ((double)m[4] * (double)m[8]);
    s = floor(c2 / (double)m[8]);
    c0 -= s * (double)m[6]; c1 -= s * (double)m[7]; c2 -= s * (double)m[8];
    s = floor((c1 - c2 * b_z) / (double)m[4]);
    c0 -= s * (double)m[3]; c1 -= s * (double)m[4];
    s = floor((c0 - (c1 * a_y + c2 * a_z)) / (double)m[0]);
    c0 -= s * (double)m[0];
    out[0] = (float)c0; out[1] = (float)c1; out[2] = (float)c2;
}

/* distance between reference coordinate r and configuration coordinate c.
 * For triclinic boxes both must already be wrapped. */
static double pbc_distance(const float *r, const float *c, const box_t *b)
{
    double dx[3];
    for (int t = 0; t < 3; ++t) dx[t] = (double)(c[t] - r[t]); /* float32 subtraction */
    if (!b->triclinic) {
        for (int t = 0; t < 3; ++t) {
            double s = (double)b->inv[t] * dx[t];
            dx[t] = (double)b->box[t] * (s - round(s));
        }
        return sqrt((dx[0] * dx[0] + dx[1] * dx[1]) + dx[2] * dx[2]);
    }
    const float *m = b->m;
    double best = (double)FLT_MAX;
    for (int ix = -1; ix < 2; ++ix) {
        double rx = dx[0] + (double)m[0] * ix;
        for (int iy = -1; iy < 2; ++iy) {
            double ry0 = rx + (double)m[3] * iy;
            double ry1 = dx[1] + (double)m[4] * iy;
            for (int iz = -1; iz < 2; ++iz) {
                double rz0 = ry0 + (double)m[6] * iz;
                double rz1 = ry1 + (double)m[7] * iz;
                double rz2 = dx[2] + (double)m[8] * iz;
                double dsq = (rz0 * rz0 + rz1 * rz1) + rz2 * rz2;
                if (dsq < best) best = dsq;
            }
        }
    }
    return sqrt(best);
}

/* maths/trig.py:103-122 -- everything float32, powf from libm */
static float angle_cos(const float *a, const float *b, const float *c, const float *dim)
{
    float ba[3], bc[3], ac[3];
    for (int t = 0; t < 3; ++t) {
        ba[t] = fabsf(a[t] - b[t]);
        bc[t] = fabsf(c[t] - b[t]);
        ac[t] = fabsf(c[t] - a[t]);
    }
    for (int t = 0; t < 3; ++t) {
        float half = 0.5f * dim[t];
        if (ba[t] > half) ba[t] = ba[t] - dim[t];
        if (bc[t] > half) bc[t] = bc[t] - dim[t];
        if (ac[t] > half) ac[t] = ac[t] - dim[t];
    }
    float dba = sqrtf((ba[0] * ba[0] + ba[1] * ba[1]) + ba[2] * ba[2]);
    float dbc = sqrtf((bc[0] * bc[0] + bc[1] * bc[1]) + bc[2] * bc[2]);
    float dac = sqrtf((ac[0] * ac[0] + ac[1] * ac[1]) + ac[2] * ac[2]);
    float num = (powf(dac, 2.0f) - powf(dbc, 2.0f)) - powf(dba, 2.0f);
    float den = (-2.0f * dbc) * dba;
    return num / den;
}

typedef struct { double d; int idx; } cand_t;

static int cand_cmp(const void *pa, const void *pb)
{
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->d < b->d) return -1;
    if (a->d > b->d) return 1;
    return (a->idx > b->idx) - (a->idx < b->idx); /* stable: ascending atom index */
}

/*
 * Per-frame neighbour lists + RAD shells for the centres listed in `centres`
 * (slots into heavy_idx; pass n_centres = n_heavy and centres = NULL for all).
 *
 * pos        float32 [n_atoms][3]
 * dims       float32 [6]
 * heavy_idx  int32 [n_heavy]   atom indices with 2 <= mass <= 999, ascending
 * excl_ptr   int32 [n_heavy+1] CSR into excl: atoms bonded to heavy atom h
 * is_water   uint8 [n_atoms]
 * outputs are indexed by centre ordinal c (0..n_centres-1):
 *   nbr_total[c]      number of candidates with d <= 10
 *   nbr_idx/nbr_dist  first min(total,25) sorted candidates
 *   shell_n/shell_idx RAD shell (atom indices, distance order)
 *   gate[c]           1 if any of the first 20 sorted candidates is water
 *   status[c]         ST_*
 */
int we_oracle_rad(int n_atoms, const float *pos, const float *dims,
                  int n_heavy, const int *heavy_idx,
                  const int *excl_ptr, const int *excl,
                  const unsigned char *is_water,
                  int n_centres, const int *centres,
                  int *nbr_total, int *nbr_idx, double *nbr_dist,
                  int *shell_n, int *shell_idx,
                  unsigned char *gate, int *status)
{
    box_t b;
    make_box(dims, &b);
    const float *P = pos;
    float *wrapped = NULL;
    if (b.triclinic) {
        wrapped = (float *)malloc((size_t)n_atoms * 3 * sizeof(float));
        if (!wrapped) return -1;
        for (int i = 0; i < n_atoms; ++i) wrap_triclinic(pos + 3 * i, b.m, wrapped + 3 * i);
        P = wrapped;
    }
    /* centres are independent (the reference maps frames over cores; within a frame its per-centre
     * work is a pure function of the frame): OpenMP over centres, OMP_NUM_THREADS=1 for a serial run */
    int failed = 0;
#pragma omp parallel
    {
    cand_t *cand = (cand_t *)malloc((size_t)n_heavy * sizeof(cand_t));
    if (!cand) {
#pragma omp atomic write
        failed = 1;
    }
#pragma omp for schedule(dynamic, 4)
    for (int c = 0; c < n_centres; ++c) {
        if (!cand) continue;
        int hs = centres ? centres[c] : c;
        int i = heavy_idx[hs];
        const float *pi = pos + 3 * i;   /* raw coordinates for angles / duplicate test */
        int n = 0, dup = 0;
        for (int h = 0; h < n_heavy; ++h) {
            int j = heavy_idx[h];
            if (j == i) continue;
            int bonded = 0;
            for (int e = excl_ptr[hs]; e < excl_ptr[hs + 1]; ++e)
                if (excl[e] == j) { bonded = 1; break; }
            if (bonded) continue;
            const float *pj = pos + 3 * j;
            if (pj[0] == pi[0] && pj[1] == pi[1] && pj[2] == pi[2]) dup = 1;
            double d = pbc_distance(P + 3 * i, P + 3 * j, &b);
            if (d <= WE_CUTOFF) { cand[n].d = d; cand[n].idx = j; ++n; }
        }
        nbr_total[c] = n;
        shell_n[c] = 0;
        gate[c] = 0;
        if (dup) { status[c] = ST_DUPLICATE; continue; }
        if (n == 0) { status[c] = ST_NO_NEIGHBOURS; continue; }
        status[c] = ST_OK;
        qsort(cand, (size_t)n, sizeof(cand_t), cand_cmp);
        int lim = n < WE_MAX_NBR ? n : WE_MAX_NBR;
        for (int k = 0; k < lim; ++k) {
            nbr_idx[c * WE_MAX_NBR + k] = cand[k].idx;
            nbr_dist[c * WE_MAX_NBR + k] = cand[k].d;
        }
        int glim = n < WE_GATE_NBR ? n : WE_GATE_NBR;
        for (int k = 0; k < glim; ++k)
            if (is_water[cand[k].idx]) { gate[c] = 1; break; }
        /* analysis/RAD.py:39-71 */
        int ns = 0;
        for (int cj = 0; cj < lim; ++cj) {
            int j = cand[cj].idx;
            double rij = cand[cj].d;
            int blocked = 0;
            for (int ck = 0; ck < cj; ++ck) {
                int k = cand[ck].idx;
                double rik = cand[ck].d;
                float cs = angle_cos(pos + 3 * j, pi, pos + 3 * k, dims);
                if (isnan(cs)) break;
                double lhs = pow(1.0 / rij, 2.0);
                double rhs = pow(1.0 / rik, 2.0) * (double)cs;
                if (lhs < rhs) { blocked = 1; break; }
            }
            if (!blocked) shell_idx[c * WE_MAX_NBR + ns++] = j;
        }
        shell_n[c] = ns;
    }
    free(cand);
    }
    free(wrapped);
    return failed ? -1 : 0;
}

/*
 * Hydrogen-bond acceptor of each donor hydrogen (analysis/HB.py:96-139).
 *   don_x[d]   heavy atom X the donor hydrogen is bonded to (atom index)
 *   don_h[d]   donor hydrogen D (atom index)
 *   shl_n[d], shl_idx[d][25]  RAD shell of X (atom indices)
 *   charges    float64 [n_atoms]
 * acceptor[d] = atom index, or -1 with status[d] != 0
 */
int we_oracle_hb(int n_atoms, const float *pos, const float *dims,
                 const double *charges,
                 int n_don, const int *don_x, const int *don_h,
                 const int *shl_n, const int *shl_idx,
                 int *acceptor, int *status)
{
    box_t b;
    make_box(dims, &b);
    (void)n_atoms;
    for (int d = 0; d < n_don; ++d) {
        int X = don_x[d], D = don_h[d];
        const float *pd = pos + 3 * D;
        float wd[3];
        if (b.triclinic) wrap_triclinic(pd, b.m, wd);
        cand_t cand[WE_MAX_NBR];
        int n = 0, dup = 0;
        /* neighbours = select "index <shell>"  -> ascending atom index */
        int order[WE_MAX_NBR];
        int ns = shl_n[d];
        for (int k = 0; k < ns; ++k) order[k] = shl_idx[d * WE_MAX_NBR + k];
        for (int a = 1; a < ns; ++a) { /* insertion sort by index */
            int v = order[a], p = a - 1;
            while (p >= 0 && order[p] > v) { order[p + 1] = order[p]; --p; }
            order[p + 1] = v;
        }
        for (int k = 0; k < ns; ++k) {
            int A = order[k];
            const float *pa = pos + 3 * A;
            if (pa[0] == pd[0] && pa[1] == pd[1] && pa[2] == pd[2]) dup = 1;
            double dist;
            if (b.triclinic) {
                float wa[3];
                wrap_triclinic(pa, b.m, wa);
                dist = pbc_distance(wd, wa, &b);
            } else {
                dist = pbc_distance(pd, pa, &b);
            }
            if (dist <= WE_CUTOFF) { cand[n].d = dist; cand[n].idx = A; ++n; }
        }
        acceptor[d] = -1;
        if (dup) { status[d] = ST_DUPLICATE; continue; }
        if (n == 0) { status[d] = ST_NO_NEIGHBOURS; continue; }
        status[d] = ST_OK;
        qsort(cand, (size_t)n, sizeof(cand_t), cand_cmp);
        double cur = 99.0;
        int cur_acc = cand[0].idx;
        for (int k = 0; k < n; ++k) {
            int A = cand[k].idx;
            double rc = (charges[D] * charges[A]) / pow(cand[k].d, 2.0);
            float cs = angle_cos(pos + 3 * X, pd, pos + 3 * A, dims);
            /* float(degrees(arccos(cs))) > 90 in float32  <=>  -1 <= cs <= -2^-24
             * (checked against NumPy in tests/test_oracle_numerics.py) */
            int angle_ok = (cs >= -1.0f) && (cs <= -5.9604644775390625e-08f);
            if (rc < cur && angle_ok) { cur = rc; cur_acc = A; }
        }
        acceptor[d] = cur_acc;
    }
    return 0;
}

/* exported helpers so tests can pin the scalar arithmetic directly */
float we_oracle_angle_cos(const float *a, const float *b, const float *c, const float *dim)
{
    return angle_cos(a, b, c, dim);
}

double we_oracle_distance(const float *r, const float *c, const float *dims)
{
    box_t b;
    make_box(dims, &b);
    if (b.triclinic) {
        float wr[3], wc[3];
        wrap_triclinic(r, b.m, wr);
        wrap_triclinic(c, b.m, wc);
        return pbc_distance(wr, wc, &b);
    }
    return pbc_distance(r, c, &b);
}
