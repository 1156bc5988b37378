/*
 * oracle.c — CPU restatement of the squidpy spatial-statistics hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in squidpy_b200/ (the product) may import, link or call this
 * file; it is the checker for tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference
 * legs of bench.py.  Every function cites the reference file:line (relative to /root/reference/) whose
 * algorithm it restates.  Written from the algorithm descriptions, not copied from the reference
 * (which is Python/numba; this is C).
 *
 * Build: see oracle/Makefile  (gcc -O3 -ffp-contract=off -fopenmp -shared -fPIC).
 * -ffp-contract=off is REQUIRED: every fused multiply-add below is written explicitly with fmaf()
 * where the reference's JIT emits one, and nowhere else.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------------
 * numpy Generator(PCG64) replay.
 *   numpy/random/src/pcg64/pcg64.h (numpy 2.3.5, third-party dependency of the reference, not in
 *   /root/reference): PCG64 = pcg_setseq_128 + XSL-RR 128/64; next_uint32 hands out the low half of a
 *   64-bit draw first and buffers the high half; random_interval() = masked rejection sampling;
 *   Generator.shuffle = descending Fisher-Yates.  Reference call sites: src/squidpy/_utils.py:240-241
 *   (spawn_generators), src/squidpy/gr/_nhood.py:533-538 (rng.shuffle), src/squidpy/gr/_utils.py:208-212
 *   (_shuffle_group), src/squidpy/gr/_ppatterns.py:271 (rng.permutation).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    u128 state, inc;
    int has_uint32;
    uint32_t uinteger;
} orc_pcg64;

#define PCG_MULT ((((u128)0x2360ED051FC65DA4ULL) << 64) | (u128)0x4385DF649FCCF645ULL)

static inline uint64_t pcg_next64(orc_pcg64 *g) {
    g->state = g->state * PCG_MULT + g->inc; /* advance first ... */
    uint64_t hi = (uint64_t)(g->state >> 64), lo = (uint64_t)g->state;
    uint64_t x = hi ^ lo;
    unsigned rot = (unsigned)(hi >> 58);
    return (x >> rot) | (x << ((-rot) & 63)); /* ... then XSL-RR output */
}

static inline uint32_t pcg_next32(orc_pcg64 *g) {
    if (g->has_uint32) {
        g->has_uint32 = 0;
        return g->uinteger;
    }
    uint64_t v = pcg_next64(g);
    g->has_uint32 = 1;
    g->uinteger = (uint32_t)(v >> 32);
    return (uint32_t)v;
}

static inline uint64_t pcg_interval(orc_pcg64 *g, uint64_t max) {
    if (max == 0) return 0;
    uint64_t mask = max, v;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    mask |= mask >> 32;
    if (max <= 0xffffffffULL) {
        while ((v = (pcg_next32(g) & mask)) > max) {
        }
    } else {
        while ((v = (pcg_next64(g) & mask)) > max) {
        }
    }
    return v;
}

/* st[0..3] = state_hi, state_lo, inc_hi, inc_lo ; st[4] = has_uint32 ; st[5] = uinteger   (in/out) */
static void load_state(orc_pcg64 *g, const uint64_t *st) {
    g->state = ((u128)st[0] << 64) | st[1];
    g->inc = ((u128)st[2] << 64) | st[3];
    g->has_uint32 = (int)st[4];
    g->uinteger = (uint32_t)st[5];
}
static void store_state(const orc_pcg64 *g, uint64_t *st) {
    st[0] = (uint64_t)(g->state >> 64);
    st[1] = (uint64_t)g->state;
    st[2] = (uint64_t)(g->inc >> 64);
    st[3] = (uint64_t)g->inc;
    st[4] = (uint64_t)g->has_uint32;
    st[5] = g->uinteger;
}

/* Generator.shuffle on a 1-D uint32 array (numpy _generator.pyx _shuffle_raw: for i in n-1..1). */
void orc_shuffle_u32(uint64_t *st, uint32_t *a, int64_t n) {
    orc_pcg64 g;
    load_state(&g, st);
    for (int64_t i = n - 1; i >= 1; --i) {
        int64_t j = (int64_t)pcg_interval(&g, (uint64_t)i);
        uint32_t t = a[i];
        a[i] = a[j];
        a[j] = t;
    }
    store_state(&g, st);
}

/* Generator.permutation(n): shuffle of arange(n) (int64). */
void orc_permutation_i64(uint64_t *st, int64_t *a, int64_t n) {
    orc_pcg64 g;
    load_state(&g, st);
    for (int64_t i = 0; i < n; ++i) a[i] = i;
    for (int64_t i = n - 1; i >= 1; --i) {
        int64_t j = (int64_t)pcg_interval(&g, (uint64_t)i);
        int64_t t = a[i];
        a[i] = a[j];
        a[j] = t;
    }
    store_state(&g, st);
}

/* raw stream dump for unit tests of the device RNG: n 64-bit outputs */
void orc_pcg64_raw(uint64_t *st, uint64_t *out, int64_t n) {
    orc_pcg64 g;
    load_state(&g, st);
    for (int64_t i = 0; i < n; ++i) out[i] = pcg_next64(&g);
    store_state(&g, st);
}

/* ------------------------------------------------------------------------------------------------
 * nhood_enrichment count kernel — src/squidpy/gr/_nhood.py:54-141 (_nenrich_{n_cls}_{parallel}).
 * Pass 1 (:79-85): res[i, clustering[c]] += 1 for every stored CSR entry (row i -> col c) into an
 * (N, n_cls) uint32 scratch.  Pass 2 (:117-132): g[clustering[row]] += res[row].  Output uint32
 * (n_cls, n_cls), row = source cluster.  Edge weights are ignored; entries are counted as stored.
 * `scratch` must hold n*n_cls uint32 (caller-owned so the timing loop does not malloc).
 * ---------------------------------------------------------------------------------------------- */
void orc_nhood_count(const uint32_t *indptr, const uint32_t *indices, const uint32_t *clustering, int64_t n,
                     int n_cls, uint32_t *scratch, uint32_t *out) {
    memset(scratch, 0, (size_t)n * n_cls * sizeof(uint32_t));
    for (int64_t i = 0; i < n; ++i) {
        uint32_t *row = scratch + i * n_cls;
        for (uint32_t e = indptr[i]; e < indptr[i + 1]; ++e) row[clustering[indices[e]]] += 1;
    }
    memset(out, 0, (size_t)n_cls * n_cls * sizeof(uint32_t));
    for (int64_t i = 0; i < n; ++i) {
        uint32_t *g = out + (size_t)clustering[i] * n_cls;
        const uint32_t *row = scratch + i * n_cls;
        for (int c = 0; c < n_cls; ++c) g[c] += row[c];
    }
}

/* ------------------------------------------------------------------------------------------------
 * _nhood_enrichment_helper — src/squidpy/gr/_nhood.py:516-547, with _shuffle_group
 * (src/squidpy/gr/_utils.py:185-213) when n_groups > 0.
 * For every permutation p: shuffled = base.copy(); rng_p.shuffle(shuffled)  (or, per library category in
 * category order, shuffle that subset in place with the SAME generator); perms[p] = count(shuffled).
 * states: P x 6 uint64 (see load_state).  group_idx: concatenated member indices of every library
 * category (category order, ascending index within a category); group_ptr: n_groups+1 offsets.
 * out: P x n_cls x n_cls uint32.  OpenMP over permutations == the reference's joblib fan-out over
 * contiguous permutation chunks (src/squidpy/_utils.py:225-231), one numba thread per worker (:74-76).
 * ---------------------------------------------------------------------------------------------- */
void orc_nhood_perms(const uint32_t *indptr, const uint32_t *indices, const uint32_t *base, int64_t n, int n_cls,
                     const uint64_t *states, int64_t n_perms, const int64_t *group_idx, const int64_t *group_ptr,
                     int n_groups, uint32_t *out, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
#pragma omp parallel
    {
        uint32_t *scratch = (uint32_t *)malloc((size_t)n * n_cls * sizeof(uint32_t));
        uint32_t *lab = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
        uint32_t *sub = (uint32_t *)malloc((size_t)n * sizeof(uint32_t));
#pragma omp for schedule(static)
        for (int64_t p = 0; p < n_perms; ++p) {
            uint64_t st[6];
            memcpy(st, states + p * 6, sizeof(st));
            memcpy(lab, base, (size_t)n * sizeof(uint32_t));
            if (n_groups <= 0) {
                orc_shuffle_u32(st, lab, n);
            } else {
                for (int gidx = 0; gidx < n_groups; ++gidx) {
                    int64_t s = group_ptr[gidx], m = group_ptr[gidx + 1] - s;
                    for (int64_t t = 0; t < m; ++t) sub[t] = base[group_idx[s + t]];
                    orc_shuffle_u32(st, sub, m);
                    for (int64_t t = 0; t < m; ++t) lab[group_idx[s + t]] = sub[t];
                }
            }
            orc_nhood_count(indptr, indices, lab, n, n_cls, scratch, out + (size_t)p * n_cls * n_cls);
        }
        free(scratch);
        free(lab);
        free(sub);
    }
}

/* same, but also returns the shuffled label vectors (P x n) for tests of the device shuffle */
void orc_shuffle_labels(const uint32_t *base, int64_t n, const uint64_t *states, int64_t n_perms,
                        const int64_t *group_idx, const int64_t *group_ptr, int n_groups, uint32_t *out_labels) {
#pragma omp parallel
    {
        uint32_t *sub = (uint32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(uint32_t));
#pragma omp for schedule(static)
        for (int64_t p = 0; p < n_perms; ++p) {
            uint64_t st[6];
            memcpy(st, states + p * 6, sizeof(st));
            uint32_t *lab = out_labels + (size_t)p * n;
            memcpy(lab, base, (size_t)n * sizeof(uint32_t));
            if (n_groups <= 0) {
                orc_shuffle_u32(st, lab, n);
            } else {
                for (int gidx = 0; gidx < n_groups; ++gidx) {
                    int64_t s = group_ptr[gidx], m = group_ptr[gidx + 1] - s;
                    for (int64_t t = 0; t < m; ++t) sub[t] = base[group_idx[s + t]];
                    orc_shuffle_u32(st, sub, m);
                    for (int64_t t = 0; t < m; ++t) lab[group_idx[s + t]] = sub[t];
                }
            }
        }
        free(sub);
    }
}

/* ------------------------------------------------------------------------------------------------
 * _occur_count — src/squidpy/gr/_ppatterns.py:283-310.
 * counts[a,b,r] = #{ordered (i,j), i != j : lab_i = a, lab_j = b, d2_ij <= thr[r]} (cumulative in r),
 * float32 arithmetic.  The reference is numba fastmath: on an FMA-capable x86-64 host LLVM emits
 *   d2 = fma(dy, dy, dx*dx)      (vmulss dx,dx ; vfmadd231ss dy,dy)    [verified with inspect_asm]
 * which is what `use_fma != 0` reproduces; use_fma == 0 gives the uncontracted dx*dx + dy*dy.
 * Per-point int32 partial rows (n, L*k*k) then a column sum, as in the reference (:289, :305-306);
 * `compact != 0` keeps one partial row per thread instead (same integers, far less memory).
 * ---------------------------------------------------------------------------------------------- */
int orc_cooc_counts(const float *x, const float *y, int64_t n, const int32_t *labs, int k, const float *thr, int L,
                    int use_fma, int compact, int64_t *out, int n_threads) {
    const int64_t width = (int64_t)L * k * k;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
    int nt = omp_get_max_threads();
#else
    int nt = 1;
#endif
    int64_t rows = compact ? nt : n;
    int32_t *local = (int32_t *)calloc((size_t)rows * width, sizeof(int32_t));
    if (!local) return -1;
    memset(out, 0, (size_t)width * sizeof(int64_t));
#pragma omp parallel
    {
#ifdef _OPENMP
        int tid = omp_get_thread_num();
#else
        int tid = 0;
#endif
        /* compact mode accumulates in int64 per thread to stay exact for any n */
        int64_t *acc = compact ? (int64_t *)calloc((size_t)width, sizeof(int64_t)) : NULL;
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < n; ++i) {
            int32_t *row = compact ? NULL : local + i * width;
            for (int64_t j = 0; j < n; ++j) {
                if (i == j) continue;
                float dx = x[i] - x[j], dy = y[i] - y[j];
                float d2 = use_fma ? fmaf(dy, dy, dx * dx) : (dx * dx + dy * dy);
                int64_t base = ((int64_t)labs[i] * k + labs[j]) * L;
                for (int r = 0; r < L; ++r) {
                    if (d2 <= thr[r]) {
                        if (compact)
                            acc[base + r] += 1;
                        else
                            row[base + r] += 1;
                    }
                }
            }
        }
        if (compact) {
#pragma omp critical
            for (int64_t c = 0; c < width; ++c) out[c] += acc[c];
            free(acc);
        }
        (void)tid;
    }
    if (!compact) {
        for (int64_t i = 0; i < n; ++i)
            for (int64_t c = 0; c < width; ++c) out[c] += local[i * width + c];
    }
    free(local);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Ripley L pair counting — src/squidpy/gr/_ripley.py:212-227 calls
 * sklearn.neighbors.KDTree(points).two_point_correlation(points, support, dualtree=True)
 * (scikit-learn 1.9.0, third-party, not under /root/reference).  Its published semantics
 * (sklearn/neighbors/_binary_tree.pxi.tp:_two_point_dual + metrics/_dist_metrics.pxd.tp:euclidean_dist):
 *   count[s] = #{ordered (i,j), INCLUDING i == j : sqrt(dx*dx + dy*dy) <= r[s]}   in float64.
 * Brute-force restatement (the tree only prunes; the integers are the same).
 * ---------------------------------------------------------------------------------------------- */
void orc_pair_counts_f64(const double *pts, int64_t m, const double *r, int S, int64_t *out, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    memset(out, 0, (size_t)S * sizeof(int64_t));
#pragma omp parallel
    {
        int64_t *acc = (int64_t *)calloc((size_t)S, sizeof(int64_t));
#pragma omp for schedule(dynamic, 64)
        for (int64_t i = 0; i < m; ++i) {
            for (int64_t j = 0; j < m; ++j) {
                double dx = pts[2 * i] - pts[2 * j], dy = pts[2 * i + 1] - pts[2 * j + 1];
                double d = 0.0;
                d += dx * dx;
                d += dy * dy;
                d = sqrt(d);
                for (int s = S - 1; s >= 0 && d <= r[s]; --s) acc[s] += 1;
            }
        }
#pragma omp critical
        for (int s = 0; s < S; ++s) out[s] += acc[s];
        free(acc);
    }
}

/* ------------------------------------------------------------------------------------------------
 * Moran's I / Geary's C — the reference calls scanpy.metrics.morans_i / gearys_c
 * (src/squidpy/gr/_ppatterns.py:14,200,205,216,267,272); scanpy (>=1.9.3, unpinned, pyproject.toml:68)
 * is third-party and absent from /root/reference, so this restates its published algorithm
 * (scanpy/metrics/_morans_i.py, _gearys_c.py): float64 throughout; W used as given (not symmetrised);
 *   S0 = sum(W.data)
 *   I  = N/S0 * sum_i z_i * sum_j w_ij z_j / sum_i z_i^2,          z = x - mean(x)
 *   C  = (N-1) * sum_ij w_ij (x_i - x_j)^2 / (2 * S0 * sum_i (x_i - mean)^2)
 * constant feature -> NaN.  X is features x obs; dense row-major (xd) or CSR by feature (densified per
 * feature exactly like scanpy's _morans_i_mtx_csr).  `row_perm` (or NULL) applies
 * g[idx,:] of _score_helper (_ppatterns.py:258-280): row r of the permuted W is row row_perm[r] of W.
 * PARITY UNPINNED at this boundary: no reference test pins I or C numerically.
 * ---------------------------------------------------------------------------------------------- */
static double autocorr_one(int mode, const int32_t *wp, const int32_t *wi, const double *wd, int64_t n, double s0,
                           const double *x, const int64_t *row_perm) {
    double mean = 0.0;
    for (int64_t i = 0; i < n; ++i) mean += x[i];
    mean /= (double)n;
    double z2 = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double z = x[i] - mean;
        z2 += z * z;
    }
    if (z2 == 0.0) return NAN;
    double total = 0.0;
    if (mode == 0) { /* moran */
        for (int64_t r = 0; r < n; ++r) {
            int64_t src = row_perm ? row_perm[r] : r;
            double acc = 0.0;
            for (int32_t e = wp[src]; e < wp[src + 1]; ++e) acc += wd[e] * (x[wi[e]] - mean);
            total += acc * (x[r] - mean);
        }
        return (double)n / s0 * total / z2;
    }
    for (int64_t r = 0; r < n; ++r) {
        int64_t src = row_perm ? row_perm[r] : r;
        double acc = 0.0;
        for (int32_t e = wp[src]; e < wp[src + 1]; ++e) {
            double d = x[r] - x[wi[e]];
            acc += wd[e] * (d * d);
        }
        total += acc;
    }
    return ((double)(n - 1) * total) / (2.0 * s0 * z2);
}

void orc_autocorr_dense(int mode, const int32_t *wp, const int32_t *wi, const double *wd, int64_t n, const double *xd,
                        int64_t n_feat, const int64_t *row_perm, double *out, int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    double s0 = 0.0;
    for (int32_t e = 0; e < wp[n]; ++e) s0 += wd[e];
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t g = 0; g < n_feat; ++g) out[g] = autocorr_one(mode, wp, wi, wd, n, s0, xd + g * n, row_perm);
}

void orc_autocorr_csr(int mode, const int32_t *wp, const int32_t *wi, const double *wd, int64_t n, const int64_t *xp,
                      const int32_t *xi, const double *xv, int64_t n_feat, const int64_t *row_perm, double *out,
                      int n_threads) {
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    double s0 = 0.0;
    for (int32_t e = 0; e < wp[n]; ++e) s0 += wd[e];
#pragma omp parallel
    {
        double *x = (double *)malloc((size_t)n * sizeof(double));
#pragma omp for schedule(dynamic, 4)
        for (int64_t g = 0; g < n_feat; ++g) {
            memset(x, 0, (size_t)n * sizeof(double));
            for (int64_t e = xp[g]; e < xp[g + 1]; ++e) x[xi[e]] = xv[e];
            out[g] = autocorr_one(mode, wp, wi, wd, n, s0, x, row_perm);
        }
        free(x);
    }
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
