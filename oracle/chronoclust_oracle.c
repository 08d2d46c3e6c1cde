/*
 * chronoclust_oracle.c -- CPU restatement of ChronoClust's per-timepoint clustering hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under chronoclust_b200/ may import, link or call this file;
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it,
 * and only as the checker.  It is a scalar, single-threaded, plain-C restatement written from the
 * behaviour of the reference (citations are relative to /root/reference/), *not* a copy of it:
 *
 *   online phase   chronoclust/clustering/hddstream.py:166-245, 247-286, 288-462, 512-549
 *   MC maths       chronoclust/objects/microcluster.py:89-153, 167-197, 213-256
 *                  chronoclust/utilities/mc_functions.py:14-77
 *   offline phase  chronoclust/clustering/hddstream.py:464-510
 *                  chronoclust/clustering/predecon.py:49-120, 136-267
 *                  chronoclust/objects/predecon_mc.py:50-80
 *                  chronoclust/utilities/predeconmc_functions.py:4-62
 *
 * Numerics rules: every sum over dimensions is a sequential left-to-right
 * fp64 accumulation starting from 0.0; no FMA contraction (build with -ffp-contract=off); IEEE
 * division.  The Euclidean neighbourhood test of the offline phase goes through BLAS dnrm2
 * (predeconmc_functions.py:16-17); the caller may inject the very function numba binds
 * (scipy.linalg.cython_blas dnrm2) with cco_set_dnrm2(); the built-in default restates the
 * x86-64 OpenBLAS kernel as "sum of squares and square root in x87 extended precision".
 *
 * Parity status: PINNED -- checked bit-for-bit against the live reference (tests/golden/, made by
 * tests/golden/make_golden.py) and against the reference's own golden result.csv.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef double (*cco_dnrm2_fn)(int *n, double *x, int *incx);

typedef struct {
    double *cf1, *cf2, *cen, *pref; /* [D] each; pref holds k or 1.0 literally */
    double w;
    int64_t id;              /* the single element of Microcluster.id */
    int64_t prev_outlier_id; /* microcluster.py:83-84; unique per creation -> used as uid */
} cco_mc;

typedef struct {
    cco_mc **v;
    int64_t n, cap;
} cco_list;

typedef struct {
    int D;
    /* constants derived by the Python side exactly as hddstream.py:44-52 does */
    double eps2, upsilon_eps, upsilon_eps2, delta, delta2, beta, k, lambda_;
    /* dataset dependent (hddstream.py:89-128), set per timepoint by the caller */
    double mu, omicron;
    int64_t pi;
    cco_list pcore, outlier;
    int64_t pcore_last_id, outlier_last_id;
    /* offline result of the last timepoint */
    int64_t n_clusters;
    int64_t *cl_off;     /* [n_clusters+1] offsets into cl_members */
    int64_t *cl_members; /* pcore ids in claim order */
    double *cl_cf1, *cl_cf2, *cl_cen, *cl_pref, *cl_w; /* [n_clusters][D], w: [n_clusters] */
    /* offline intermediates kept for white-box tests */
    int64_t off_m;
    uint8_t *off_core, *off_nbr, *off_wnbr; /* [m], [m*m], [m*m] */
    double *off_w;                            /* [m][D] subspace preference vectors */
    cco_dnrm2_fn dnrm2;
    int64_t n_dist_pairs; /* (point x MC) distance evaluations, for the CPU baseline's flop count */
} cco_state;

/* ------------------------------------------------------------------------------------------ */
static cco_mc *mc_new(int D) {
    cco_mc *m = (cco_mc *)calloc(1, sizeof(cco_mc));
    m->cf1 = (double *)calloc((size_t)4 * D, sizeof(double));
    m->cf2 = m->cf1 + D;
    m->cen = m->cf2 + D;
    m->pref = m->cen + D;
    return m;
}
static void mc_free(cco_mc *m) {
    if (m) {
        free(m->cf1);
        free(m);
    }
}
static void list_push(cco_list *l, cco_mc *m) {
    if (l->n == l->cap) {
        l->cap = l->cap ? l->cap * 2 : 64;
        l->v = (cco_mc **)realloc(l->v, (size_t)l->cap * sizeof(cco_mc *));
    }
    l->v[l->n++] = m;
}
static void list_remove_at(cco_list *l, int64_t i) {
    memmove(l->v + i, l->v + i + 1, (size_t)(l->n - i - 1) * sizeof(cco_mc *));
    l->n--;
}

/* mc_functions.py:35-43 -- sum_d ((p_d - c_d)^2) / pref_d, sequential. */
static double projected_distance(const cco_mc *m, const double *p, int D) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) {
        double t = p[d] - m->cen[d];
        t = t * t;
        t = t / m->pref[d];
        s = s + t;
    }
    return s;
}

/* Tentative clone + add + preference recompute: microcluster.py:213-233 (get_copy_with_new_point),
 * mc_functions.py:24-29 (update_cf), :14-22 (calculate_squared_variance), microcluster.py:89-115. */
static void tentative(const cco_mc *m, const double *p, int D, double delta2, double k, double *cf1,
                      double *cf2, double *w_out, double *pref) {
    double w = m->w + 1.0;
    for (int d = 0; d < D; ++d) {
        cf1[d] = m->cf1[d] + p[d];
        cf2[d] = m->cf2[d] + p[d] * p[d];
        double a = cf2[d] / w;
        double b = cf1[d] / w;
        b = b * b;
        double var = a - b;
        pref[d] = (var <= delta2) ? k : 1.0;
    }
    *w_out = w;
}

/* mc_functions.py:45-56 */
static double projected_radius2(const double *cf1, const double *cf2, const double *pref, double w, int D) {
    double s = 0.0;
    for (int d = 0; d < D; ++d) {
        double a = cf2[d] / w;
        double b = cf1[d] / w;
        b = b * b;
        double t = (a - b) / pref[d];
        s = s + t;
    }
    return s;
}

static int64_t count_ne1(const double *pref, int D) {
    int64_t c = 0;
    for (int d = 0; d < D; ++d) c += (pref[d] != 1.0);
    return c;
}
static int64_t count_gt1(const double *pref, int D) {
    int64_t c = 0;
    for (int d = 0; d < D; ++d) c += (pref[d] > 1.0);
    return c;
}

/* add_new_point + update_preferred_dimensions (microcluster.py:117-153, 89-115) */
static void commit(cco_mc *m, const double *cf1, const double *cf2, double w, const double *pref, int D) {
    for (int d = 0; d < D; ++d) {
        m->cf1[d] = cf1[d];
        m->cf2[d] = cf2[d];
        m->cen[d] = cf1[d] / w; /* calculate_centroid, mc_functions.py:31-33 */
        m->pref[d] = pref[d];
    }
    m->w = w;
}

/* ------------------------------------------------------------------------------------------ */
cco_state *cco_create(int D, double eps2, double upsilon_eps, double upsilon_eps2, double delta, double delta2,
                      double beta, double k, double lambda_) {
    cco_state *s = (cco_state *)calloc(1, sizeof(cco_state));
    s->D = D;
    s->eps2 = eps2;
    s->upsilon_eps = upsilon_eps;
    s->upsilon_eps2 = upsilon_eps2;
    s->delta = delta;
    s->delta2 = delta2;
    s->beta = beta;
    s->k = k;
    s->lambda_ = lambda_;
    return s;
}

static void free_offline(cco_state *s) {
    free(s->cl_off);
    free(s->cl_members);
    free(s->cl_cf1);
    free(s->cl_w);
    free(s->off_core);
    free(s->off_nbr);
    free(s->off_wnbr);
    free(s->off_w);
    s->cl_off = s->cl_members = NULL;
    s->cl_cf1 = s->cl_cf2 = s->cl_cen = s->cl_pref = s->cl_w = NULL;
    s->off_core = s->off_nbr = s->off_wnbr = NULL;
    s->off_w = NULL;
    s->n_clusters = 0;
    s->off_m = 0;
}

void cco_destroy(cco_state *s) {
    if (!s) return;
    for (int64_t i = 0; i < s->pcore.n; ++i) mc_free(s->pcore.v[i]);
    for (int64_t i = 0; i < s->outlier.n; ++i) mc_free(s->outlier.v[i]);
    free(s->pcore.v);
    free(s->outlier.v);
    free_offline(s);
    free(s);
}

void cco_set_dnrm2(cco_state *s, cco_dnrm2_fn fn) { s->dnrm2 = fn; }

/* Timepoint start: hddstream.py:199-213.  `decay` != 0 iff t != last_data_timestamp; the factor
 * 2 ** (-lambda * interval) is evaluated by the caller in Python (hddstream.py:283). */
void cco_begin_timepoint(cco_state *s, double mu, double omicron, int64_t pi, int decay, double decay_factor) {
    const int D = s->D;
    s->mu = mu;
    s->omicron = omicron;
    s->pi = pi;
    if (!decay) return;
    /* _decay_clusters_weight, hddstream.py:247-286: centroid and preference vector untouched */
    for (int pass = 0; pass < 2; ++pass) {
        cco_list *l = pass ? &s->outlier : &s->pcore;
        for (int64_t i = 0; i < l->n; ++i) {
            cco_mc *m = l->v[i];
            for (int d = 0; d < D; ++d) {
                m->cf1[d] *= decay_factor;
                m->cf2[d] *= decay_factor;
            }
            m->w *= decay_factor;
        }
    }
    /* _downgrade_potential_microclusters, hddstream.py:516-537: Python's list iterator keeps an
     * index, so removing the element just visited makes the next one slide into its slot unseen. */
    const double bm = s->beta * s->mu;
    for (int64_t i = 0; i < s->pcore.n; ++i) {
        cco_mc *m = s->pcore.v[i];
        int w_bad = m->w < bm;
        int p_bad = count_gt1(m->pref, D) > s->pi;
        if (w_bad || p_bad) {
            m->id = m->prev_outlier_id;
            list_remove_at(&s->pcore, i);
            list_push(&s->outlier, m);
            /* i is NOT decremented: skip-next-after-removal */
        }
    }
    /* _downgrade_outlier_microclusters, hddstream.py:539-549 (same iteration quirk) */
    for (int64_t i = 0; i < s->outlier.n; ++i) {
        cco_mc *m = s->outlier.v[i];
        if (m->w <= s->omicron) {
            list_remove_at(&s->outlier, i);
            mc_free(m);
        }
    }
}

/* The ordered per-point loop: hddstream.py:220-237.  X is row-major with leading dimension ld
 * (in doubles).  assign_uid[r] receives the prev_outlier_id (unique per MC) of the MC that took
 * row r; stage[r] = 0 pcore absorb, 1 outlier absorb, 2 outlier absorb + upgrade, 3 new outlier. */
void cco_ingest(cco_state *s, const double *X, int64_t N, int64_t ld, int64_t *assign_uid, uint8_t *stage) {
    const int D = s->D;
    double *cf1 = (double *)malloc((size_t)4 * D * sizeof(double));
    double *cf2 = cf1 + D, *pref = cf2 + D, *tmp = pref + D;
    (void)tmp;
    for (int64_t r = 0; r < N; ++r) {
        const double *p = X + r * ld;
        int done = 0;
        /* ---- _add_to_pcore, hddstream.py:288-343 ---- */
        {
            int64_t best = -1;
            double best_d = 0.0, w2;
            for (int64_t j = 0; j < s->pcore.n; ++j) {
                cco_mc *m = s->pcore.v[j];
                tentative(m, p, D, s->delta2, s->k, cf1, cf2, &w2, pref);
                if (count_ne1(pref, D) <= s->pi) {
                    double dist = projected_distance(m, p, D);
                    s->n_dist_pairs++;
                    if (best < 0 || dist < best_d) {
                        best = j;
                        best_d = dist;
                    }
                }
            }
            if (best >= 0) {
                cco_mc *m = s->pcore.v[best];
                tentative(m, p, D, s->delta2, s->k, cf1, cf2, &w2, pref);
                if (projected_radius2(cf1, cf2, pref, w2, D) <= s->eps2) {
                    commit(m, cf1, cf2, w2, pref, D);
                    assign_uid[r] = m->prev_outlier_id;
                    if (stage) stage[r] = 0;
                    done = 1;
                }
            }
        }
        if (done) continue;
        /* ---- _add_to_outlier, hddstream.py:345-395 ---- */
        {
            int64_t best = -1;
            double best_d = 0.0, w2;
            for (int64_t j = 0; j < s->outlier.n; ++j) {
                double dist = projected_distance(s->outlier.v[j], p, D);
                if (best < 0 || dist < best_d) {
                    best = j;
                    best_d = dist;
                }
            }
            s->n_dist_pairs += s->outlier.n;
            if (best >= 0) {
                cco_mc *m = s->outlier.v[best];
                tentative(m, p, D, s->delta2, s->k, cf1, cf2, &w2, pref);
                if (projected_radius2(cf1, cf2, pref, w2, D) <= s->eps2) {
                    commit(m, cf1, cf2, w2, pref, D);
                    assign_uid[r] = m->prev_outlier_id;
                    if (stage) stage[r] = 1;
                    /* _upgrade_outlier_microcluster, hddstream.py:397-430; prev_pcore_id is never
                     * set anywhere, so a fresh pcore id is always minted. */
                    if (m->w >= s->beta * s->mu && count_gt1(m->pref, D) <= s->pi) {
                        m->id = s->pcore_last_id++;
                        list_remove_at(&s->outlier, best);
                        list_push(&s->pcore, m);
                        if (stage) stage[r] = 2;
                    }
                    done = 1;
                }
            }
        }
        if (done) continue;
        /* ---- _create_new_outlier_cluster, hddstream.py:434-462 ---- */
        {
            cco_mc *m = mc_new(D);
            double w2;
            m->w = 0.0;
            tentative(m, p, D, s->delta2, s->k, cf1, cf2, &w2, pref);
            commit(m, cf1, cf2, w2, pref, D);
            m->id = s->outlier_last_id;
            m->prev_outlier_id = s->outlier_last_id;
            s->outlier_last_id++;
            list_push(&s->outlier, m);
            assign_uid[r] = m->prev_outlier_id;
            if (stage) stage[r] = 3;
        }
    }
    free(cf1);
}

/* ------------------------------------------------------------------------------------------ */
/* Offline phase.                                                                             */

/* Default Euclidean norm: restates OpenBLAS' x86-64 dnrm2 (x87 kernel): squares and their sum are
 * kept in 80-bit extended precision, square root in extended precision, one final rounding. */
static double default_nrm2(const double *x, int n) {
    long double s = 0.0L;
    for (int i = 0; i < n; ++i) s += (long double)x[i] * (long double)x[i];
    return (double)sqrtl(s);
}

static double euclid(cco_state *s, const double *a, const double *b, double *scratch) {
    int D = s->D;
    for (int d = 0; d < D; ++d) scratch[d] = a[d] - b[d]; /* predeconmc_functions.py:16 */
    if (s->dnrm2) {
        int n = D, inc = 1;
        return s->dnrm2(&n, scratch, &inc);
    }
    return default_nrm2(scratch, D);
}

/* hddstream.py:464-510 + predecon.py:49-120,136-267.  Returns the number of clusters. */
int64_t cco_offline(cco_state *s) {
    const int D = s->D;
    const int64_t M = s->pcore.n;
    free_offline(s);
    s->off_m = M;
    s->off_core = (uint8_t *)calloc((size_t)(M ? M : 1), 1);
    s->off_nbr = (uint8_t *)calloc((size_t)(M ? M * M : 1), 1);
    s->off_wnbr = (uint8_t *)calloc((size_t)(M ? M * M : 1), 1);
    s->off_w = (double *)calloc((size_t)(M ? M * D : 1), sizeof(double));
    double *scratch = (double *)malloc((size_t)D * sizeof(double));
    uint8_t *cls = (uint8_t *)calloc((size_t)(M ? M : 1), 1); /* 0 = 'u', 1 = 'c', 2 = 'n' */
    int64_t *pdim = (int64_t *)calloc((size_t)(M ? M : 1), sizeof(int64_t));
    const double E = s->upsilon_eps, E2 = s->upsilon_eps2;

    /* core flags: Microcluster.is_core, microcluster.py:235-256 -> mc_functions.py:64-77 */
    for (int64_t i = 0; i < M; ++i) {
        cco_mc *m = s->pcore.v[i];
        double r2 = projected_radius2(m->cf1, m->cf2, m->pref, m->w, D);
        s->off_core[i] = (r2 <= s->eps2) && (m->w >= s->mu) && (count_gt1(m->pref, D) <= s->pi);
    }
    /* N(p) and w_p: predecon.py:149-152, 161-188, 190-217 */
    for (int64_t p = 0; p < M; ++p) {
        const double *cp = s->pcore.v[p]->cen;
        int64_t cnt = 0;
        for (int64_t q = 0; q < M; ++q) {
            double dist = euclid(s, s->pcore.v[q]->cen, cp, scratch);
            if (dist <= E) {
                s->off_nbr[p * M + q] = 1;
                cnt++;
            }
        }
        for (int d = 0; d < D; ++d) {
            double sum = 0.0;
            for (int64_t q = 0; q < M; ++q)
                if (s->off_nbr[p * M + q]) {
                    double t = cp[d] - s->pcore.v[q]->cen[d];
                    sum = sum + t * t;
                }
            double var = sum / (double)cnt;
            s->off_w[p * D + d] = (var <= s->delta) ? s->k : 1.0; /* delta, not delta^2: predecon.py:213 */
        }
        pdim[p] = count_gt1(s->off_w + p * D, D);
    }
    /* WN(p): predecon.py:155-159, 219-239 */
    for (int64_t p = 0; p < M; ++p) {
        const double *cp = s->pcore.v[p]->cen;
        for (int64_t q = 0; q < M; ++q) {
            if (!s->off_nbr[p * M + q]) continue;
            const double *cq = s->pcore.v[q]->cen;
            double dpq = 0.0, dqp = 0.0;
            for (int d = 0; d < D; ++d) {
                double t = cp[d] - cq[d];
                dpq = dpq + s->off_w[p * D + d] * (t * t);
            }
            for (int d = 0; d < D; ++d) {
                double t = cq[d] - cp[d];
                dqp = dqp + s->off_w[q * D + d] * (t * t);
            }
            double dist = dpq;
            if (dqp > dpq) dist = dqp; /* Python max(a, b) */
            if (dist <= E2) s->off_wnbr[p * M + q] = 1;
        }
    }
    /* cluster growth: predecon.py:62-87 (run) and :89-120 (_expand) */
    int64_t *queue = (int64_t *)malloc((size_t)(2 * M + 2) * sizeof(int64_t));
    s->cl_off = (int64_t *)calloc((size_t)(M + 2), sizeof(int64_t));
    s->cl_members = (int64_t *)calloc((size_t)(M + 1), sizeof(int64_t));
    s->cl_cf1 = (double *)calloc((size_t)((M + 1) * 4 * D), sizeof(double));
    s->cl_cf2 = s->cl_cf1 + (M + 1) * D;
    s->cl_cen = s->cl_cf2 + (M + 1) * D;
    s->cl_pref = s->cl_cen + (M + 1) * D;
    s->cl_w = (double *)calloc((size_t)(M + 1), sizeof(double));
    int64_t nc = 0, nmem = 0;
    for (int64_t sd = 0; sd < M; ++sd) {
        if (cls[sd] != 0) continue;
        if (!s->off_core[sd]) {
            cls[sd] = 2;
            continue;
        }
        double *k1 = s->cl_cf1 + nc * D, *k2 = s->cl_cf2 + nc * D;
        double kw = 0.0;
        int64_t start = nmem;
        for (int d = 0; d < D; ++d) k1[d] = k2[d] = 0.0;
        int64_t qh = 0, qt = 0;
        for (int64_t x = 0; x < M; ++x)
            if (s->off_wnbr[sd * M + x]) queue[qt++] = x;
        while (qh < qt) {
            int64_t q = queue[qh++];
            if (!s->off_core[q]) continue; /* _find_directly_reachable_points, predecon.py:242-267 */
            for (int64_t x = 0; x < M; ++x) {
                if (pdim[x] > s->pi || !s->off_wnbr[q * M + x]) continue;
                if (cls[x] == 0) queue[qt++] = x; /* every MC is enqueued by a claim at most once */
                if (cls[x] == 0 || cls[x] == 2) {
                    cls[x] = 1;
                    cco_mc *m = s->pcore.v[x]; /* merge_mc, predecon_mc.py:50-68 */
                    for (int d = 0; d < D; ++d) {
                        k1[d] += m->cf1[d];
                        k2[d] += m->cf2[d];
                    }
                    kw += m->w;
                    s->cl_members[nmem++] = m->id;
                }
            }
        }
        if (kw > 0.0) { /* predecon.py:80-84 */
            for (int d = 0; d < D; ++d) {
                s->cl_cen[nc * D + d] = k1[d] / kw;
                double a = k2[d] / kw, b = k1[d] / kw;
                b = b * b;
                s->cl_pref[nc * D + d] = ((a - b) <= s->delta2) ? s->k : 1.0;
            }
            s->cl_w[nc] = kw;
            s->cl_off[nc] = start;
            nc++;
            s->cl_off[nc] = nmem;
        } else {
            nmem = start;
        }
    }
    s->n_clusters = nc;
    free(queue);
    free(scratch);
    free(cls);
    free(pdim);
    return nc;
}

/* ------------------------------------------------------------------------------------------ */
/* Accessors for ctypes.                                                                      */
int64_t cco_count(const cco_state *s, int which) { return which ? s->outlier.n : s->pcore.n; }
int64_t cco_dist_pairs(const cco_state *s) { return s->n_dist_pairs; }
int64_t cco_last_id(const cco_state *s, int which) { return which ? s->outlier_last_id : s->pcore_last_id; }

/* Copies one list out: ids/uids [n], w [n], cf1/cf2/cen/pref [n][D]. */
void cco_export(const cco_state *s, int which, int64_t *ids, int64_t *uids, double *w, double *cf1, double *cf2,
                double *cen, double *pref) {
    const cco_list *l = which ? &s->outlier : &s->pcore;
    const int D = s->D;
    for (int64_t i = 0; i < l->n; ++i) {
        const cco_mc *m = l->v[i];
        ids[i] = m->id;
        uids[i] = m->prev_outlier_id;
        w[i] = m->w;
        memcpy(cf1 + i * D, m->cf1, (size_t)D * sizeof(double));
        memcpy(cf2 + i * D, m->cf2, (size_t)D * sizeof(double));
        memcpy(cen + i * D, m->cen, (size_t)D * sizeof(double));
        memcpy(pref + i * D, m->pref, (size_t)D * sizeof(double));
    }
}

/* Appends an MC to a list verbatim (used to start white-box tests from an arbitrary state). */
void cco_import_mc(cco_state *s, int which, int64_t id, int64_t uid, double w, const double *cf1, const double *cf2,
                   const double *cen, const double *pref) {
    const int D = s->D;
    cco_mc *m = mc_new(D);
    m->id = id;
    m->prev_outlier_id = uid;
    m->w = w;
    memcpy(m->cf1, cf1, (size_t)D * sizeof(double));
    memcpy(m->cf2, cf2, (size_t)D * sizeof(double));
    memcpy(m->cen, cen, (size_t)D * sizeof(double));
    memcpy(m->pref, pref, (size_t)D * sizeof(double));
    list_push(which ? &s->outlier : &s->pcore, m);
}
void cco_set_counters(cco_state *s, int64_t pcore_last_id, int64_t outlier_last_id) {
    s->pcore_last_id = pcore_last_id;
    s->outlier_last_id = outlier_last_id;
}
void cco_set_thresholds(cco_state *s, double mu, double omicron, int64_t pi) {
    s->mu = mu;
    s->omicron = omicron;
    s->pi = pi;
}

int64_t cco_n_clusters(const cco_state *s) { return s->n_clusters; }
int64_t cco_n_members(const cco_state *s) { return s->n_clusters ? s->cl_off[s->n_clusters] : 0; }
void cco_export_clusters(const cco_state *s, int64_t *off, int64_t *members, double *w, double *cf1, double *cf2,
                         double *cen, double *pref) {
    const int D = s->D;
    const int64_t nc = s->n_clusters;
    for (int64_t c = 0; c <= nc; ++c) off[c] = nc ? s->cl_off[c] : 0;
    if (!nc) return;
    memcpy(members, s->cl_members, (size_t)s->cl_off[nc] * sizeof(int64_t));
    memcpy(w, s->cl_w, (size_t)nc * sizeof(double));
    memcpy(cf1, s->cl_cf1, (size_t)nc * D * sizeof(double));
    memcpy(cf2, s->cl_cf2, (size_t)nc * D * sizeof(double));
    memcpy(cen, s->cl_cen, (size_t)nc * D * sizeof(double));
    memcpy(pref, s->cl_pref, (size_t)nc * D * sizeof(double));
}
/* white-box offline intermediates: core [m], nbr/wnbr [m*m] bytes, w [m][D] */
int64_t cco_offline_m(const cco_state *s) { return s->off_m; }
void cco_export_offline(const cco_state *s, uint8_t *core, uint8_t *nbr, uint8_t *wnbr, double *w) {
    const int64_t M = s->off_m;
    if (!M) return;
    memcpy(core, s->off_core, (size_t)M);
    memcpy(nbr, s->off_nbr, (size_t)(M * M));
    memcpy(wnbr, s->off_wnbr, (size_t)(M * M));
    memcpy(w, s->off_w, (size_t)(M * s->D) * sizeof(double));
}

/* Stand-alone kernels for the reference's known-answer tests (unittest_microcluster.py,
 * unittest_predecon.py). */
double cco_kat_projected_distance(const double *cen, const double *pref, const double *p, int D) {
    cco_mc m;
    m.cen = (double *)cen;
    m.pref = (double *)pref;
    return projected_distance(&m, p, D);
}
double cco_kat_radius2(const double *cf1, const double *cf2, const double *pref, double w, int D) {
    return projected_radius2(cf1, cf2, pref, w, D);
}
