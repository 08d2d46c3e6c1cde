// offline.cuh -- KERNEL 4: the PreDeCon-style offline phase over the potential microclusters.
//
// Replaces HDDStream.offline_clustering (clustering/hddstream.py:464-510) and PreDeCon.run
// (clustering/predecon.py:49-120, 136-267) with the maths of utilities/predeconmc_functions.py:4-62,
// objects/predecon_mc.py:50-80 and Microcluster.is_core (utilities/mc_functions.py:64-77).
//
//   4a k_off_core        core flags                                   (hddstream.py:483-496)
//   4b k_off_neighbours  eps-neighbourhood bit rows, row-sharded       (predecon.py:161-188)
//   4c k_off_subspace    subspace preference vector per row            (predecon.py:190-217)
//   4d k_off_weighted    preference-weighted neighbourhood bit rows    (predecon.py:155-159, 219-239)
//   4e k_off_clusters    ordered cluster growth (FIFO expansion)       (predecon.py:62-120, 242-267)
//   4f k_off_cluster_cf  merged-cluster statistics in claim order      (predecon_mc.py:50-68, predecon.py:80)
//
// 4b-4d work on row ranges [r0, r1) so that G GPUs can each take M/G rows; the host all-gathers the
// subspace masks (before 4d) and the weighted-neighbour rows (before 4e) over NCCL.
//
// The Euclidean test of 4b is `dnrm2(c_q - c_p) <= E` in the reference (OpenBLAS, third party).  The
// device evaluates s = sum_d (c_q[d] - c_p[d])^2 in fp64 (index order, fused multiply-add) and decides every pair
// whose s is outside a relative guard band of 2^-40 around E^2 -- far wider than the rounding difference between any
// two summation schemes, dnrm2's included -- and reports the (astronomically rare) pairs inside the band to the
// host, which settles them with the very dnrm2 the reference calls.
#pragma once
#include "common.cuh"

namespace ccb {

constexpr double OFF_GUARD = 9.094947017729282e-13; // 2^-40

// ---- 4a ------------------------------------------------------------------------------------------
__global__ void k_off_core(const double *cf1, const double *cf2, const double *w, const uint64_t *mask, int M, int D,
                           double k, double wsel, int div_mode, int cnt_gt1, double eps2, double mu, int64_t pi,
                           uint8_t *core) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= M) return;
    const double wj = w[j];
    const uint64_t m = mask[j];
    double s = 0.0; // calculate_projected_radius_squared, mc_functions.py:45-56
    for (int d = 0; d < D; ++d) {
        const double a = ddiv(cf2[(size_t)j * D + d], wj);
        const double b = ddiv(cf1[(size_t)j * D + d], wj);
        double t = dsub(a, dmul(b, b));
        if ((m >> d) & 1ull) t = div_mode ? ddiv(t, k) : dmul(t, wsel);
        s = dadd(s, t);
    }
    const int pd = cnt_gt1 ? popc64(m) : 0;
    core[j] = (s <= eps2) && (wj >= mu) && ((int64_t)pd <= pi);
}

// ---- 4b ------------------------------------------------------------------------------------------
// FP64-pipe kernel (3 D flops per ORDERED pair).  One thread owns PPT rows with their coordinates in
// registers; the columns stream through shared memory in double-buffered tiles (1-D bulk TMA) and every lane reads
// the same coordinates (broadcast LDS.128); JU columns x PPT rows advance together for ILP.  The test only has to be
// right OUTSIDE the guard band (pairs inside it go to the host's dnrm2), so the sum of squares may be contracted:
// x = c - p; s = fma(x, x, s) -- 2 instructions for the 3 credited flops; |s_fma - s_seq| <= D 2^-52 s << 2^-40 s.
constexpr int OFFN_THREADS = 128;
constexpr int OFFN_JU = 4;
template <int DP>
struct OffCfg {
    static constexpr int PPT = DP <= 40 ? 2 : 1;
    static constexpr int TM = ((2048 / DP) < 32 ? 32 : (2048 / DP)) / 32 * 32; // MCs per tile, multiple of 32
    static constexpr int ROWS = OFFN_THREADS * PPT;
};

template <int DP>
__global__ void __launch_bounds__(OFFN_THREADS)
    k_off_neighbours(const double *__restrict__ cen, int M, int D, int r0, int r1, double E2, uint32_t *__restrict__ nbr,
                     int32_t *__restrict__ cnt, int32_t *__restrict__ border, int border_cap, int32_t *n_border) {
    constexpr int TM = OffCfg<DP>::TM, PPT = OffCfg<DP>::PPT, JU = OFFN_JU;
    __shared__ __align__(128) double tile[2][TM * DP];
    __shared__ __align__(8) uint64_t bar[2];
    const int words = (M + 31) / 32;
    int row[PPT];
    bool live[PPT];
    double p[PPT][DP];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        row[u] = r0 + blockIdx.x * OffCfg<DP>::ROWS + u * OFFN_THREADS + threadIdx.x;
        live[u] = row[u] < r1;
#pragma unroll
        for (int d = 0; d < DP; ++d) p[u][d] = (live[u] && d < D) ? cen[(size_t)row[u] * D + d] : 0.0;
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    // column slab of this CTA (blockIdx.y): tiles [t_lo, t_hi); the row counts are accumulated with atomicAdd
    const int ntiles_all = (M + TM - 1) / TM;
    const int tps = (ntiles_all + gridDim.y - 1) / gridDim.y;
    const int t_lo = blockIdx.y * tps, t_hi = min(ntiles_all, t_lo + tps);
    auto issue = [&](int t) {
        const int q = t - t_lo;
        const int jt = t * TM;
        const int n = min(TM, M - jt);
        uint32_t bytes = (uint32_t)((size_t)n * D * sizeof(double));
        if (bytes & 15u) { // odd element count: the last double travels by a plain store (ordered by the arrive)
            bytes -= 8u;
            tile[q & 1][(size_t)n * D - 1] = cen[(size_t)jt * D + (size_t)n * D - 1];
        }
        mbar_expect_tx(&bar[q & 1], bytes);
        if (bytes) tma_load_1d(&tile[q & 1][0], cen + (size_t)jt * D, bytes, &bar[q & 1]);
    };
    if (threadIdx.x == 0 && t_lo < t_hi) issue(t_lo);
    const double guard = dmul(E2, OFF_GUARD);
    const bool full = D == DP; // compile-time offsets and 16-byte loads (DP is a multiple of 4)
    int count[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) count[u] = 0;
    for (int t = t_lo; t < t_hi; ++t) {
        const int q = t - t_lo;
        if (threadIdx.x == 0 && t + 1 < t_hi) issue(t + 1);
        mbar_wait(&bar[q & 1], (q >> 1) & 1);
        const double *tl = tile[q & 1];
        const int jt = t * TM;
        const int n = min(TM, M - jt);
        for (int wq = 0; wq < n; wq += 32) {
            uint32_t bits[PPT];
#pragma unroll
            for (int u = 0; u < PPT; ++u) bits[u] = 0u;
            for (int b0 = 0; b0 < 32 && wq + b0 < n; b0 += JU) {
                double s[PPT][JU];
#pragma unroll
                for (int u = 0; u < PPT; ++u)
#pragma unroll
                    for (int v = 0; v < JU; ++v) s[u][v] = 0.0;
                if (full) {
#pragma unroll
                    for (int d = 0; d < DP; d += 2) {
#pragma unroll
                        for (int v = 0; v < JU; ++v) { // columns past n read stale shared memory; masked below
                            const double2 c = *reinterpret_cast<const double2 *>(tl + (size_t)(wq + b0 + v) * DP + d);
#pragma unroll
                            for (int u = 0; u < PPT; ++u) {
                                const double x0 = dsub(c.x, p[u][d]); // predeconmc_functions.py:16 (a - b, a = the other MC)
                                s[u][v] = __fma_rn(x0, x0, s[u][v]);
                                const double x1 = dsub(c.y, p[u][d + 1]);
                                s[u][v] = __fma_rn(x1, x1, s[u][v]);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int d = 0; d < DP; ++d) {
                        if (d < D) {
#pragma unroll
                            for (int v = 0; v < JU; ++v) {
                                const double c = tl[(size_t)(wq + b0 + v) * D + d];
#pragma unroll
                                for (int u = 0; u < PPT; ++u) {
                                    const double x = dsub(c, p[u][d]);
                                    s[u][v] = __fma_rn(x, x, s[u][v]);
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int v = 0; v < JU; ++v) {
                    const int b = b0 + v;
                    if (wq + b < n) {
#pragma unroll
                        for (int u = 0; u < PPT; ++u) {
                            if (!live[u]) continue;
                            if (fabs(dsub(s[u][v], E2)) <= guard) {
                                const int slot = atomicAdd(n_border, 1);
                                if (slot < border_cap) {
                                    border[2 * slot] = row[u];
                                    border[2 * slot + 1] = jt + wq + b;
                                }
                            } else if (s[u][v] < E2) {
                                bits[u] |= 1u << b;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < PPT; ++u)
                if (live[u]) {
                    nbr[(size_t)(row[u] - r0) * words + ((jt + wq) >> 5)] = bits[u];
                    count[u] += __popc(bits[u]);
                }
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u)
        if (live[u] && count[u]) atomicAdd(&cnt[row[u] - r0], count[u]); // cnt is zeroed by the launcher
}

// settles borderline pairs decided on the host: sets the bit and bumps the row count
__global__ void k_off_patch(uint32_t *nbr, int32_t *cnt, const int32_t *pairs, const uint8_t *decision, int n, int r0,
                            int words) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !decision[i]) return;
    const int row = pairs[2 * i], col = pairs[2 * i + 1];
    atomicOr(&nbr[(size_t)(row - r0) * words + (col >> 5)], 1u << (col & 31));
    atomicAdd(&cnt[row - r0], 1);
}

// ---- 4c: one thread per (row, dim); neighbours visited in index order (the sum order of np.sum) ----
__global__ void k_off_subspace(const double *__restrict__ cen, int M, int D, int r0, int r1,
                               const uint32_t *__restrict__ nbr, const int32_t *__restrict__ cnt, double delta,
                               uint64_t *submask) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int row = r0 + (int)(t / D), d = (int)(t % D);
    if (row >= r1) return;
    const int words = (M + 31) / 32;
    const uint32_t *nr = nbr + (size_t)(row - r0) * words;
    const double cp = cen[(size_t)row * D + d];
    double sum = 0.0;
    for (int wq = 0; wq < words; ++wq) {
        uint32_t bits = nr[wq];
        while (bits) {
            const int q = (wq << 5) + __ffs(bits) - 1;
            bits &= bits - 1;
            const double x = dsub(cp, cen[(size_t)q * D + d]); // predeconmc_functions.py:36 (point - neighbours)
            sum = dadd(sum, dmul(x, x));
        }
    }
    const double var = ddiv(sum, (double)cnt[row - r0]);
    if (var <= delta) atomicOr(reinterpret_cast<unsigned long long *>(&submask[row - r0]), 1ull << d); // delta, not delta^2
}

// ---- 4d: one thread per (row, word of the neighbour row) -----------------------------------------------
__global__ void k_off_weighted(const double *__restrict__ cen, int M, int D, int r0, int r1,
                               const uint32_t *__restrict__ nbr, const uint64_t *__restrict__ submask_all, double k,
                               double E2, uint32_t *__restrict__ wnbr) {
    const int words = (M + 31) / 32;
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int row = r0 + (int)(t / words), wq = (int)(t % words);
    if (row >= r1) return;
    uint32_t bits = nbr[(size_t)(row - r0) * words + wq];
    uint32_t out = 0u;
    const double *cp = cen + (size_t)row * D;
    const uint64_t mp = submask_all[row];
    while (bits) {
        const int b = __ffs(bits) - 1;
        bits &= bits - 1;
        const int q = (wq << 5) + b;
        const double *cq = cen + (size_t)q * D;
        const uint64_t mq = submask_all[q];
        double dpq = 0.0, dqp = 0.0; // calculate_weighted_dist_squared, predeconmc_functions.py:44-62
        for (int d = 0; d < D; ++d) {
            const double x = dsub(cp[d], cq[d]);
            const double sq = dmul(x, x);
            dpq = dadd(dpq, ((mp >> d) & 1ull) ? dmul(k, sq) : sq);
            const double y = dsub(cq[d], cp[d]);
            const double sy = dmul(y, y);
            dqp = dadd(dqp, ((mq >> d) & 1ull) ? dmul(k, sy) : sy);
        }
        double dist = dpq;
        if (dqp > dpq) dist = dqp; // Python max(a, b)
        if (dist <= E2) out |= 1u << b;
    }
    wnbr[(size_t)(row - r0) * words + wq] = out;
}

// ---- 4e: ordered cluster growth ------------------------------------------------------------------------
// One CTA reproduces PreDeCon.run / _expand literally: seeds in list order; the queue starts as a copy of
// WN(seed); every popped core MC q claims, in index order, each x in WN(q) with pdim(x) <= pi that is
// unclassified (-> also enqueued) or noise.  Rows are scanned by the whole CTA with an ordered
// (prefix-sum) compaction so queue and claim order equal the reference's.
constexpr int OFFC_THREADS = 1024;

__device__ __forceinline__ int block_excl_scan(int v, int *s_warp, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int wv = s_warp[lane];
        int winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += n;
        }
        s_warp[lane] = winc - wv;
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    const int r = s_warp[warp] + inc - v;
    return r;
}

__global__ void __launch_bounds__(OFFC_THREADS, 1)
    k_off_clusters(int M, const uint32_t *__restrict__ wnbr, const uint8_t *__restrict__ core,
                   const uint64_t *__restrict__ submask, int cnt_gt1, int64_t pi, uint8_t *cls /*[M] zeroed*/,
                   int32_t *queue /*[2M+2]*/, int32_t *label, int32_t *order, int32_t *cl_off, int32_t *n_cl) {
    __shared__ int s_warp[32];
    __shared__ int s_tot[2];
    __shared__ int s_qt, s_nmem;
    const int tid = threadIdx.x;
    const int words = (M + 31) / 32;
    for (int i = tid; i < M; i += OFFC_THREADS) label[i] = -1;
    if (tid == 0) {
        s_nmem = 0;
        cl_off[0] = 0;
    }
    __syncthreads();
    int ncl = 0;
    for (int sd = 0; sd < M; ++sd) {
        const int c0 = cls[sd];
        if (c0 != 0) continue;
        if (!core[sd]) {
            __syncthreads();
            if (tid == 0) cls[sd] = 2;
            __syncthreads();
            continue;
        }
        // initial queue: a copy of WN(seed), in index order, no filtering (predecon.py:103)
        int qt = 0;
        for (int base = 0; base < words; base += OFFC_THREADS) {
            const int wq = base + tid;
            const uint32_t bits = wq < words ? wnbr[(size_t)sd * words + wq] : 0u;
            int tot;
            int off = block_excl_scan(__popc(bits), s_warp, &s_tot[0]);
            tot = s_tot[0];
            uint32_t bb = bits;
            while (bb) {
                const int b = __ffs(bb) - 1;
                bb &= bb - 1;
                queue[qt + off++] = (wq << 5) + b;
            }
            qt += tot;
            __syncthreads();
        }
        int qh = 0;
        while (qh < qt) {
            const int q = queue[qh++];
            if (!core[q]) continue; // _find_directly_reachable_points: point_is_core
            for (int base = 0; base < words; base += OFFC_THREADS) {
                const int wq = base + tid;
                uint32_t bits = wq < words ? wnbr[(size_t)q * words + wq] : 0u;
                uint32_t claim = 0u, enq = 0u;
                uint32_t bb = bits;
                while (bb) {
                    const int b = __ffs(bb) - 1;
                    bb &= bb - 1;
                    const int x = (wq << 5) + b;
                    const int pd = cnt_gt1 ? popc64(submask[x]) : 0;
                    if ((int64_t)pd > pi) continue;
                    const int cx = cls[x];
                    if (cx == 0) enq |= 1u << b;
                    if (cx == 0 || cx == 2) claim |= 1u << b;
                }
                const int qoff = block_excl_scan(__popc(enq), s_warp, &s_tot[0]);
                const int qtot = s_tot[0];
                __syncthreads();
                const int moff = block_excl_scan(__popc(claim), s_warp, &s_tot[1]);
                const int mtot = s_tot[1];
                const int nmem = s_nmem;
                int o1 = qt + qoff, o2 = nmem + moff;
                bb = enq;
                while (bb) {
                    const int b = __ffs(bb) - 1;
                    bb &= bb - 1;
                    queue[o1++] = (wq << 5) + b;
                }
                bb = claim;
                while (bb) {
                    const int b = __ffs(bb) - 1;
                    bb &= bb - 1;
                    const int x = (wq << 5) + b;
                    order[o2++] = x;
                    label[x] = ncl;
                    cls[x] = 1;
                }
                qt += qtot;
                __syncthreads();
                if (tid == 0) s_nmem = nmem + mtot;
                __syncthreads();
            }
        }
        __syncthreads();
        ncl += 1; // emitted even if empty; the host drops clusters whose weight is not > 0 (predecon.py:83)
        if (tid == 0) cl_off[ncl] = s_nmem;
        __syncthreads();
    }
    if (tid == 0) *n_cl = ncl;
}

// ---- 4f: merged cluster statistics: one CTA per cluster, thread per dim, members in claim order ----------
__global__ void k_off_cluster_cf(const double *cf1, const double *cf2, const double *w, int D, const int32_t *order,
                                 const int32_t *cl_off, double delta2, double *o_cf1, double *o_cf2, double *o_cen,
                                 uint64_t *o_mask, double *o_w) {
    const int c = blockIdx.x, d = threadIdx.x;
    const int b = cl_off[c], e = cl_off[c + 1];
    if (d == 0) o_mask[c] = 0ull;
    __syncthreads();
    double k1 = 0.0, k2 = 0.0, kw = 0.0;
    for (int i = b; i < e; ++i) { // merge_mc, predecon_mc.py:50-68
        const int x = order[i];
        if (d < D) {
            k1 = dadd(k1, cf1[(size_t)x * D + d]);
            k2 = dadd(k2, cf2[(size_t)x * D + d]);
        }
        kw = dadd(kw, w[x]);
    }
    if (d < D) {
        o_cf1[(size_t)c * D + d] = k1;
        o_cf2[(size_t)c * D + d] = k2;
        const double cen = ddiv(k1, kw);
        o_cen[(size_t)c * D + d] = cen;
        const double var = dsub(ddiv(k2, kw), dmul(cen, cen)); // update_preferred_dimensions(delta^2, k), predecon.py:80
        if (var <= delta2) atomicOr(reinterpret_cast<unsigned long long *>(&o_mask[c]), 1ull << d);
    }
    if (d == 0) o_w[c] = kw;
}

} // namespace ccb
