/*
 * bsv_proto.c -- CPU model of the block-speculative versioned commit ("BSV") that the CUDA engine
 * (chronoclust_b200/csrc/engine.cuh) implements.  TEST INFRASTRUCTURE ONLY (it includes the oracle).
 *
 * The reference's online phase (clustering/hddstream.py:220-237) is a strictly ordered loop: every cell's
 * decision depends on the exact fp64 state left by all earlier cells.  BSV keeps those semantics bit for bit
 * while exposing parallelism, one block of B consecutive cells at a time:
 *
 *   S   speculate every cell's decision from the SNAPSHOT state at the block start.  A decision is a
 *       target key:  j            (absorbed by pcore MC j,                       hddstream.py:288-343)
 *                    Mp + o       (absorbed by outlier MC o of the snapshot,     hddstream.py:345-395)
 *                    Mp + Mo0 + c (absorbed by the MC that cell c < i of this block created)
 *                    Mp + Mo0 + i (creates a new outlier MC,                     hddstream.py:434-462)
 *   C   per target key, replay its cells in input order: CF1 += x, CF2 += x*x, W += 1 -- sequential fp64
 *       adds exactly as mc_functions.py:24-29 would perform them -- and keep the state after every cell
 *       (its "version").  Creation is an absorb into the all-zero MC.
 *   D   per version: centroid, variance, preference vector, projected radius^2 (cell-parallel).
 *   V   per cell, recompute the decision EXACTLY against, for every MC, the version it had just before
 *       that cell (all pcore MCs; for the outlier list: the nearest still-unmodified snapshot MC from a
 *       top-K list, every MC modified earlier in the block, every MC created earlier in the block).
 *   M   the first cell whose exact decision differs from the speculated one bounds the exact prefix:
 *       every cell before it saw exact versions, so by induction their decisions -- and that cell's own
 *       recomputed decision -- are the reference's.  The recomputed decisions become the next speculation
 *       (fixed-point refinement); an outlier->pcore upgrade (hddstream.py:397-430) ends the block.
 *   commit the exact prefix: last versions -> MC lists, new MCs appended in creation order, upgrade applied.
 *
 * This file runs that algorithm with plain loops on the oracle's state so that the scheme (not the CUDA
 * code) can be validated on the CPU against the sequential oracle, bit for bit.
 */
#include "../../oracle/chronoclust_oracle.c"

#include <stdio.h>

typedef struct {
    int64_t blocks, iters, cells, mismatches, unknown_cuts, iter_cuts, upgrades, max_iters, rejects, tk_miss;
} bsv_stats;

typedef struct {
    const double *cf1, *cf2, *cen, *pref;
    double w;
} bsv_view;

#define KEY_REJ (-2)
#define KEY_UNKNOWN (-1)
#define KEY_NEED (-3) /* rejected by the pcore stage but no top-K list of the snapshot outlier list yet */

static int64_t cmp_i64pair(const void *a, const void *b) {
    const int64_t *x = (const int64_t *)a, *y = (const int64_t *)b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[1] != y[1]) return x[1] < y[1] ? -1 : 1;
    return 0;
}
static int cmp_pair(const void *a, const void *b) { return (int)cmp_i64pair(a, b); }

/* tentative absorb of p into a raw state (microcluster.py:213-233); returns projected radius^2 */
static double tent_view(const bsv_view *v, const double *p, int D, double delta2, double k, double *cf1, double *cf2,
                        double *w_out, double *pref) {
    cco_mc m;
    m.cf1 = (double *)v->cf1;
    m.cf2 = (double *)v->cf2;
    m.w = v->w;
    tentative(&m, p, D, delta2, k, cf1, cf2, w_out, pref);
    return projected_radius2(cf1, cf2, pref, *w_out, D);
}
static double dist_view(const bsv_view *v, const double *p, int D) {
    cco_mc m;
    m.cen = (double *)v->cen;
    m.pref = (double *)v->pref;
    return projected_distance(&m, p, D);
}
static bsv_view view_mc(const cco_mc *m) {
    bsv_view v = {m->cf1, m->cf2, m->cen, m->pref, m->w};
    return v;
}

/* latest element of the ascending list l[0..n) that is < i, or -1 */
static int64_t latest_before(const int64_t *l, int64_t n, int64_t i) {
    int64_t lo = 0, hi = n; /* first index with l[idx] >= i */
    while (lo < hi) {
        int64_t mid = (lo + hi) / 2;
        if (l[mid] < i) lo = mid + 1;
        else hi = mid;
    }
    return lo ? l[lo - 1] : -1;
}

/* One block starting at row p0; returns the number of committed cells (>= 1).
 * theta = contest * eps2: a pcore candidate whose snapshot distance exceeds theta (or that the snapshot
 * radius test rejects) is CONTESTED: its radius test is evaluated exactly, in order, inside the chain of its
 * candidate MC, so that one wrong guess cannot poison the rest of that chain. */
static int64_t bsv_block(cco_state *s, const double *X, int64_t ld, int64_t p0, int64_t B, int itmax, int K,
                         int64_t rmax, double contest, int64_t *assign_uid, uint8_t *stage, bsv_stats *st) {
    const int D = s->D;
    const int64_t Mp = s->pcore.n, Mo0 = s->outlier.n;
    const int64_t KNEW = Mp + Mo0; /* key of the MC created by cell i = KNEW + i */
    const double theta = contest * s->eps2;
    int64_t *pcand = (int64_t *)malloc((size_t)B * 8), *ospec = (int64_t *)malloc((size_t)B * 8);
    int64_t *dec = (int64_t *)malloc((size_t)B * 8), *eff = (int64_t *)malloc((size_t)B * 8);
    uint8_t *pflag = (uint8_t *)calloc((size_t)B, 1), *prej = (uint8_t *)calloc((size_t)B, 1);
    uint8_t *upf = (uint8_t *)calloc((size_t)B, 1);
    int64_t *tkpos = (int64_t *)malloc((size_t)B * 8);
    double *vcf1 = (double *)malloc((size_t)B * D * 8 * 4), *vcf2 = vcf1 + B * D, *vcen = vcf2 + B * D,
           *vpref = vcen + B * D;
    double *vw = (double *)malloc((size_t)B * 8 * 2), *vr2 = vw + B;
    double *t1 = (double *)malloc((size_t)D * 8 * 3), *t2 = t1 + D, *tp = t2 + D;
    double *zero = (double *)calloc((size_t)D, 8);
    double *tkd = (double *)malloc((size_t)(rmax ? rmax : 1) * K * 8);
    int64_t *tki = (int64_t *)malloc((size_t)(rmax ? rmax : 1) * K * 8);
    double tw;
#define BSV_TOPK(i)                                                                        \
    do {                                                                                   \
        const double *p_ = X + (p0 + (i)) * ld;                                            \
        double *bd_ = tkd + tkpos[i] * K;                                                  \
        int64_t *bi_ = tki + tkpos[i] * K;                                                 \
        for (int e = 0; e < K; ++e) bd_[e] = INFINITY, bi_[e] = -1;                        \
        for (int64_t o = 0; o < Mo0; ++o) {                                                \
            double d_ = projected_distance(s->outlier.v[o], p_, D);                        \
            if (!(d_ < bd_[K - 1])) continue;                                              \
            int e = K - 1;                                                                 \
            while (e > 0 && d_ < bd_[e - 1]) bd_[e] = bd_[e - 1], bi_[e] = bi_[e - 1], --e; \
            bd_[e] = d_, bi_[e] = o;                                                       \
        }                                                                                  \
    } while (0)

    /* ---- S: pcore candidate against the snapshot, SAFE / CONTESTED, top-K + outlier decision for the rest */
    int64_t nrej = 0, Beff = B;
    for (int64_t i = 0; i < B; ++i) {
        const double *p = X + (p0 + i) * ld;
        int64_t best = -1;
        double bd = 0.0;
        for (int64_t j = 0; j < Mp; ++j) {
            cco_mc *m = s->pcore.v[j];
            tentative(m, p, D, s->delta2, s->k, t1, t2, &tw, tp);
            if (count_ne1(tp, D) <= s->pi) {
                double d = projected_distance(m, p, D);
                if (best < 0 || d < bd) best = j, bd = d;
            }
        }
        pcand[i] = best;
        pflag[i] = 1;
        ospec[i] = KEY_REJ;
        tkpos[i] = -1;
        if (best >= 0) {
            bsv_view v = view_mc(s->pcore.v[best]);
            double r2s = tent_view(&v, p, D, s->delta2, s->k, t1, t2, &tw, tp);
            if (bd <= theta && r2s <= s->eps2) pflag[i] = 0;
        }
        if (pflag[i]) {
            if (nrej == rmax) { /* bound the outlier-stage work of one block */
                Beff = i;
                break;
            }
            tkpos[i] = nrej++;
            BSV_TOPK(i);
            const int64_t *bi = tki + tkpos[i] * K;
            ospec[i] = KNEW + i;
            if (bi[0] >= 0) {
                bsv_view v = view_mc(s->outlier.v[bi[0]]);
                if (tent_view(&v, p, D, s->delta2, s->k, t1, t2, &tw, tp) <= s->eps2) ospec[i] = Mp + bi[0];
            }
        }
    }
    st->rejects += nrej;

    int64_t *pcnt = (int64_t *)malloc((size_t)(Mp + 1) * 8), *poff = (int64_t *)malloc((size_t)(Mp + 2) * 8);
    int64_t *plist = (int64_t *)malloc((size_t)B * 8), *pacc = (int64_t *)malloc((size_t)B * 8);
    int64_t *nacc = (int64_t *)malloc((size_t)(Mp + 1) * 8);
    int64_t *opair = (int64_t *)malloc((size_t)B * 16);
    int64_t *hkey = (int64_t *)malloc((size_t)B * 8), *hoff = (int64_t *)malloc((size_t)(B + 1) * 8);
    int64_t *omem = (int64_t *)malloc((size_t)B * 8);
    int64_t *firstmember = (int64_t *)malloc((size_t)(Mo0 ? Mo0 : 1) * 8);
    int64_t m_commit = 0, nh = 0;
    int upgrade_at_end = 0;

    for (int it = 0;; ++it) {
        st->iters++;
        /* ---- L_P: ordered candidate lists of the pcore keys */
        for (int64_t j = 0; j <= Mp; ++j) pcnt[j] = 0;
        for (int64_t i = 0; i < Beff; ++i)
            if (pcand[i] >= 0) pcnt[pcand[i]]++;
        poff[0] = 0;
        for (int64_t j = 0; j < Mp; ++j) poff[j + 1] = poff[j] + pcnt[j], pcnt[j] = 0;
        for (int64_t i = 0; i < Beff; ++i)
            if (pcand[i] >= 0) plist[poff[pcand[i]] + pcnt[pcand[i]]++] = i;
        /* ---- C_P: pcore chains; SAFE members only add, CONTESTED members take the exact radius test */
        for (int64_t i = 0; i < Beff; ++i) prej[i] = pcand[i] < 0;
        for (int64_t j = 0; j < Mp; ++j) {
            const double *c1 = s->pcore.v[j]->cf1, *c2 = s->pcore.v[j]->cf2;
            double w = s->pcore.v[j]->w;
            nacc[j] = 0;
            for (int64_t e = poff[j]; e < poff[j + 1]; ++e) {
                const int64_t i = plist[e];
                const double *p = X + (p0 + i) * ld;
                if (pflag[i]) {
                    bsv_view v = {c1, c2, 0, 0, w};
                    if (!(tent_view(&v, p, D, s->delta2, s->k, t1, t2, &tw, tp) <= s->eps2)) {
                        prej[i] = 1;
                        continue;
                    }
                }
                for (int d = 0; d < D; ++d) {
                    vcf1[i * D + d] = c1[d] + p[d];
                    vcf2[i * D + d] = c2[d] + p[d] * p[d];
                }
                vw[i] = w + 1.0;
                c1 = vcf1 + i * D, c2 = vcf2 + i * D, w = vw[i];
                pacc[poff[j] + nacc[j]++] = i;
            }
        }
        /* ---- L_O: outlier-side keys of the pcore-rejected cells, sorted by (key, cell) */
        int64_t no = 0;
        for (int64_t i = 0; i < Beff; ++i)
            if (prej[i]) opair[2 * no] = ospec[i], opair[2 * no + 1] = i, no++;
        qsort(opair, (size_t)no, 16, cmp_pair);
        nh = 0;
        for (int64_t o = 0; o < Mo0; ++o) firstmember[o] = INT64_MAX;
        for (int64_t e = 0; e < no; ++e) {
            if (e == 0 || opair[2 * e] != opair[2 * e - 2]) {
                hkey[nh] = opair[2 * e], hoff[nh] = e, nh++;
                if (opair[2 * e] < KNEW) firstmember[opair[2 * e] - Mp] = opair[2 * e + 1];
            }
            omem[e] = opair[2 * e + 1];
        }
        hoff[nh] = no;
        /* ---- C_O: outlier-side chains (modified snapshot MCs, MCs created in this block) */
        for (int64_t h = 0; h < nh; ++h) {
            const double *c1 = zero, *c2 = zero;
            double w = 0.0;
            if (hkey[h] < KNEW) {
                cco_mc *m = s->outlier.v[hkey[h] - Mp];
                c1 = m->cf1, c2 = m->cf2, w = m->w;
            }
            for (int64_t e = hoff[h]; e < hoff[h + 1]; ++e) {
                const int64_t i = omem[e];
                const double *p = X + (p0 + i) * ld;
                for (int d = 0; d < D; ++d) {
                    vcf1[i * D + d] = c1[d] + p[d];
                    vcf2[i * D + d] = c2[d] + p[d] * p[d];
                }
                vw[i] = w + 1.0;
                c1 = vcf1 + i * D, c2 = vcf2 + i * D, w = vw[i];
            }
        }
        /* ---- D: derive centroid / preference vector / radius^2 of every version */
        for (int64_t i = 0; i < Beff; ++i) {
            double r2 = 0.0;
            for (int d = 0; d < D; ++d) {
                double a = vcf2[i * D + d] / vw[i], b = vcf1[i * D + d] / vw[i];
                vcen[i * D + d] = b;
                b = b * b;
                double var = a - b;
                vpref[i * D + d] = (var <= s->delta2) ? s->k : 1.0;
                r2 = r2 + var / vpref[i * D + d];
            }
            vr2[i] = r2;
        }
        /* ---- V: exact decision of every cell given the versions before it */
        for (int64_t i = 0; i < Beff; ++i) {
            const double *p = X + (p0 + i) * ld;
            upf[i] = 0;
            eff[i] = prej[i] ? ospec[i] : pcand[i];
            int64_t best = -1;
            double bd = 0.0;
            bsv_view bv = {0, 0, 0, 0, 0};
            for (int64_t j = 0; j < Mp; ++j) {
                bsv_view v = view_mc(s->pcore.v[j]);
                int64_t lat = latest_before(pacc + poff[j], nacc[j], i);
                if (lat >= 0) {
                    bsv_view vv = {vcf1 + lat * D, vcf2 + lat * D, vcen + lat * D, vpref + lat * D, vw[lat]};
                    v = vv;
                }
                tent_view(&v, p, D, s->delta2, s->k, t1, t2, &tw, tp);
                if (count_ne1(tp, D) <= s->pi) {
                    double d = dist_view(&v, p, D);
                    if (best < 0 || d < bd) best = j, bd = d, bv = v;
                }
            }
            if (best >= 0 && tent_view(&bv, p, D, s->delta2, s->k, t1, t2, &tw, tp) <= s->eps2) {
                dec[i] = best;
                continue;
            }
            /* outlier stage.  (1) nearest snapshot MC not modified before i, from the top-K list */
            int64_t obest = -1; /* key */
            double obd = 0.0;
            bsv_view ov = {0, 0, 0, 0, 0};
            int unknown = 0, bounded = 0;
            double bound = 0.0; /* every snapshot MC outside a full, all-stale list is at least this far away */
            if (Mo0 > 0) {
                if (tkpos[i] < 0) {
                    unknown = 2;
                } else {
                    const double *td = tkd + tkpos[i] * K;
                    const int64_t *ti = tki + tkpos[i] * K;
                    int e = 0;
                    for (; e < K; ++e) {
                        if (ti[e] < 0) break;
                        if (firstmember[ti[e]] >= i) {
                            obest = Mp + ti[e], obd = td[e], ov = view_mc(s->outlier.v[ti[e]]);
                            break;
                        }
                    }
                    if (e == K) bounded = 1, bound = td[K - 1];
                }
            }
            if (unknown) {
                dec[i] = KEY_NEED;
                continue;
            }
            /* (2) every MC modified or created earlier in the block, at its version just before i */
            for (int64_t h = 0; h < nh; ++h) {
                if (omem[hoff[h]] >= i) continue;
                int64_t lat = latest_before(omem + hoff[h], hoff[h + 1] - hoff[h], i);
                bsv_view v = {vcf1 + lat * D, vcf2 + lat * D, vcen + lat * D, vpref + lat * D, vw[lat]};
                double d = dist_view(&v, p, D);
                if (obest < 0 || d < obd || (d == obd && hkey[h] < obest)) obest = hkey[h], obd = d, ov = v;
            }
            /* all listed snapshot candidates were stale: the nearest clean MC is unknown, but it cannot be nearer
             * than the last list entry -- the decision stands iff a modified / created MC beats that bound */
            if (bounded && !(obest >= 0 && obd < bound)) {
                dec[i] = KEY_UNKNOWN;
                continue;
            }
            dec[i] = KNEW + i;
            if (obest >= 0 && tent_view(&ov, p, D, s->delta2, s->k, t1, t2, &tw, tp) <= s->eps2) {
                dec[i] = obest;
                if (tw >= s->beta * s->mu && count_gt1(tp, D) <= s->pi) upf[i] = 1;
            }
        }
        /* ---- M: exact prefix, upgrade, refinement */
        int64_t m0 = Beff;
        for (int64_t i = 0; i < Beff; ++i)
            if (dec[i] != eff[i]) {
                m0 = i;
                break;
            }
        int64_t up = -1;
        for (int64_t i = 0; i < m0; ++i)
            if (upf[i]) {
                up = i;
                break;
            }
        if (it + 1 > st->max_iters) st->max_iters = it + 1;
        if (up >= 0) {
            m_commit = up + 1;
            upgrade_at_end = 1;
            st->upgrades++;
            break;
        }
        if (m0 == Beff) {
            m_commit = Beff;
            break;
        }
        st->mismatches++;
        if (dec[m0] == KEY_UNKNOWN || (dec[m0] == KEY_NEED && nrej == rmax)) {
            m_commit = m0;
            st->unknown_cuts++;
            break;
        }
        if (it + 1 >= itmax) {
            m_commit = m0;
            st->iter_cuts++;
            break;
        }
        for (int64_t i = m0; i < Beff; ++i) {
            if (dec[i] == eff[i]) continue;
            if (dec[i] == KEY_UNKNOWN || (dec[i] == KEY_NEED && nrej == rmax)) {
                Beff = i;
                break;
            }
            if (dec[i] == KEY_NEED) { /* fetch its top-K list now; provisional outlier decision: create */
                tkpos[i] = nrej++;
                BSV_TOPK(i);
                st->tk_miss++;
                ospec[i] = KNEW + i;
                pflag[i] = 1;
            } else if (dec[i] < Mp) {
                /* another pcore MC is nearest: test it exactly in its chain, which needs a fallback outlier
                 * decision (hence a top-K list) should that chain reject the cell */
                if (pcand[i] != dec[i] && !pflag[i] && (Mo0 == 0 || nrej < rmax)) {
                    if (Mo0 > 0) {
                        tkpos[i] = nrej++;
                        BSV_TOPK(i);
                    }
                    ospec[i] = KNEW + i;
                    pflag[i] = 1;
                }
                pcand[i] = dec[i];
            } else {
                ospec[i] = dec[i];
                pflag[i] = 1;
            }
        }
    }
    if (m_commit < 1) {
        fprintf(stderr, "bsv: no progress at row %lld\n", (long long)p0);
        abort();
    }
    /* ---- commit [0, m_commit): versions -> lists, creations appended in order, upgrade applied */
    const int64_t m = m_commit;
    int64_t *newslot = (int64_t *)malloc((size_t)m * 8);
    for (int64_t i = 0; i < m; ++i) {
        newslot[i] = -1;
        if (eff[i] == KNEW + i) { /* creation */
            cco_mc *mc = mc_new(D);
            mc->id = mc->prev_outlier_id = s->outlier_last_id++;
            newslot[i] = s->outlier.n;
            list_push(&s->outlier, mc);
        }
    }
    for (int64_t i = 0; i < m; ++i) { /* ascending i: the last version of every key wins */
        cco_mc *mc;
        const int64_t key = eff[i];
        if (key < Mp) mc = s->pcore.v[key], stage[p0 + i] = 0;
        else if (key < KNEW) mc = s->outlier.v[key - Mp], stage[p0 + i] = 1;
        else mc = s->outlier.v[newslot[key - KNEW]], stage[p0 + i] = (key == KNEW + i) ? 3 : 1;
        memcpy(mc->cf1, vcf1 + i * D, (size_t)D * 8);
        memcpy(mc->cf2, vcf2 + i * D, (size_t)D * 8);
        memcpy(mc->cen, vcen + i * D, (size_t)D * 8);
        memcpy(mc->pref, vpref + i * D, (size_t)D * 8);
        mc->w = vw[i];
        assign_uid[p0 + i] = mc->prev_outlier_id;
    }
    if (upgrade_at_end) {
        const int64_t key = eff[m - 1];
        const int64_t slot = key < KNEW ? key - Mp : newslot[key - KNEW];
        cco_mc *mc = s->outlier.v[slot];
        mc->id = s->pcore_last_id++;
        list_remove_at(&s->outlier, slot);
        list_push(&s->pcore, mc);
        stage[p0 + m - 1] = 2;
    }
    st->blocks++;
    st->cells += m;
    free(newslot);
    free(pcand), free(ospec), free(dec), free(eff), free(pflag), free(prej), free(upf), free(tkpos), free(vcf1), free(vw);
    free(t1), free(zero), free(tkd), free(tki);
    free(pcnt), free(poff), free(plist), free(pacc), free(nacc), free(opair), free(hkey), free(hoff), free(omem);
    free(firstmember);
    return m;
}

/* Drop-in for cco_ingest that runs the block-speculative scheme.  Block length adapts: doubles after a
 * block that committed completely, halves after one that was cut. */
void cco_ingest_bsv(cco_state *s, const double *X, int64_t N, int64_t ld, int64_t *assign_uid, uint8_t *stage,
                    int64_t Bmin, int64_t Bmax, int itmax, int K, int64_t rmax, double contest, bsv_stats *st) {
    int64_t pos = 0, B = Bmin;
    while (pos < N) {
        const int64_t b = B < N - pos ? B : N - pos;
        const int64_t m = bsv_block(s, X, ld, pos, b, itmax, K, rmax, contest, assign_uid, stage, st);
        pos += m;
        if (m == b) B = B * 2 < Bmax ? B * 2 : Bmax;
        else B = B / 2 > Bmin ? B / 2 : Bmin;
    }
}
