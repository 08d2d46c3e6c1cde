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

// ---- 4d: one WARP per row; lane l takes words l, l + 32, ... of the neighbour row ------------------------------------
// (a thread per (row, word) launched M * M / 32 threads, nearly all of them for empty words: 2.4 M CTAs at M = 1e5)
constexpr int OFFW_THREADS = 128;
__global__ void __launch_bounds__(OFFW_THREADS)
    k_off_weighted(const double *__restrict__ cen, int M, int D, int r0, int r1, const uint32_t *__restrict__ nbr,
                   const uint64_t *__restrict__ submask_all, double k, double E2, uint32_t *__restrict__ wnbr) {
    const int words = (M + 31) / 32;
    const int lane = threadIdx.x & 31;
    const int row = r0 + blockIdx.x * (OFFW_THREADS / 32) + (threadIdx.x >> 5);
    if (row >= r1) return;
    const double *cp = cen + (size_t)row * D;
    const uint64_t mp = submask_all[row];
    for (int wq = lane; wq < words; wq += 32) {
        uint32_t bits = nbr[(size_t)(row - r0) * words + wq];
        uint32_t out = 0u;
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
    __shared__ int s_nmem;
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

// ---- 4e, large M: the same ordered growth on a CSR of the weighted neighbourhoods -----------------------------------
// k_off_clusters costs one scan of a whole bit row (M / 32 words) per seed and per popped core MC: 2 M row scans,
// 1.4 s at M = 1e5 (config C4), 95 % of the offline phase.  Three observations remove almost all of it without touching
// the literal rule of PreDeCon.run / _expand:
//  (1) an ISOLATED MC (WN(p) = {p}) can neither reach nor be reached by anything (WN is symmetric): core -> it opens
//      its own cluster {p} (empty if pdim(p) > pi), not core -> noise.  Decided in parallel.
//  (2) seeds are visited in increasing index order and every seed emits exactly one cluster, so the index of a cluster
//      is the RANK of its seed among all seeds: clusters grown serially and clusters of isolated MCs are merged by a
//      prefix sum over seed flags afterwards.
//  (3) the remaining MCs are walked by ONE CTA exactly like k_off_clusters, but over CSR lists (cost = entries of the
//      popped lists, not M / 32 words per pop), skipping classified / isolated seeds 1024 at a time.
constexpr int OFFG_THREADS = 1024;

__device__ __forceinline__ int offc_block_min(int v, int *s_warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = v;
    __syncthreads();
    v = s_warp[threadIdx.x & 31];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    return v;
}

// warp per row of [r0, r1): |WN(row)|, isolated flag, CSR entries the row will need (0 for isolated rows).  wnbr holds the
// rows from r0 on (a rank's own rows in the sharded path); iso / nnz are indexed from r0, too.
__global__ void k_offc_rowinfo(const uint32_t *__restrict__ wnbr, int r0, int r1, int words, uint8_t *iso, int32_t *nnz) {
    const int lane = threadIdx.x & 31;
    const int row = r0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= r1) return;
    const uint32_t *r = wnbr + (size_t)(row - r0) * words;
    int c = 0;
    for (int w = lane; w < words; w += 32) c += __popc(r[w]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) {
        const bool self = (r[row >> 5] >> (row & 31)) & 1u;
        const bool is = (c == 1 && self);
        iso[row - r0] = is;
        nnz[row - r0] = is ? 0 : c;
    }
}

// single-CTA exclusive scan of n int32 values into n + 1 outputs (T = int64_t offsets or int32_t ranks); every thread owns
// 8 consecutive values per step, so M = 1e5 takes 13 steps of three block barriers instead of 98
template <typename T>
__global__ void __launch_bounds__(OFFG_THREADS, 1) k_offc_scan(const int32_t *__restrict__ in, int n, T *__restrict__ out) {
    constexpr int E = 8;
    __shared__ int s_warp[32];
    __shared__ int s_tot;
    T carry = 0;
    for (int b = 0; b < n; b += OFFG_THREADS * E) {
        const int i0 = b + threadIdx.x * E;
        int v[E], sum = 0;
#pragma unroll
        for (int u = 0; u < E; ++u) {
            v[u] = i0 + u < n ? in[i0 + u] : 0;
            sum += v[u];
        }
        const int ex = block_excl_scan(sum, s_warp, &s_tot);
        T run = carry + (T)ex;
#pragma unroll
        for (int u = 0; u < E; ++u) {
            if (i0 + u < n) out[i0 + u] = run;
            run += (T)v[u];
        }
        carry += (T)s_tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

// warp per non-isolated row of [r0, r1): column indices of its set bits, ascending.  wnbr holds the rows from r0 on; iso and
// off are indexed by the GLOBAL row, col is the global CSR (a rank fills the segment of its own rows).
__global__ void k_offc_fill(const uint32_t *__restrict__ wnbr, int r0, int r1, int words, const uint8_t *__restrict__ iso,
                            const int64_t *__restrict__ off, int32_t *__restrict__ col) {
    const int lane = threadIdx.x & 31;
    const int row = r0 + blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= r1 || iso[row]) return;
    const uint32_t *r = wnbr + (size_t)(row - r0) * words;
    int64_t o = off[row];
    for (int w0 = 0; w0 < words; w0 += 32) {
        const int w = w0 + lane;
        uint32_t bits = w < words ? r[w] : 0u;
        const int c = __popc(bits);
        int inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        int64_t my = o + inc - c;
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            col[my++] = (w << 5) + b;
        }
        o += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// The same ordered growth by ONE WARP over the compacted list of NON-ISOLATED microclusters: when the lists are short (a
// sparse weighted-neighbour graph: config C4 has ~2 entries per non-isolated row and 98 % isolated rows) the 1024-thread
// version below spends its time in block barriers and in scanning all M flags for the next seed; a warp needs no barrier
// and only visits the candidates (cand [ncand], ascending; isolated microclusters never take part in the growth).  Seeds
// are scanned 32 candidates at a time, lists are compacted with ballots in index order.  Identical output; the launcher
// picks by the mean list length.
__global__ void k_offc_candflag(int M, const uint8_t *__restrict__ iso, int32_t *__restrict__ flag) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) flag[i] = iso[i] ? 0 : 1;
}
__global__ void k_offc_candscatter(int M, const int32_t *__restrict__ flag, const int32_t *__restrict__ rank,
                                   int32_t *__restrict__ cand) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M && flag[i]) cand[rank[i]] = i;
}
__global__ void __launch_bounds__(32, 1)
    k_offc_grow_warp(const int32_t *__restrict__ cand, const int32_t *__restrict__ ncand_p, const int64_t *__restrict__ off,
                     const int32_t *__restrict__ col, const uint8_t *__restrict__ core, const uint64_t *__restrict__ submask,
                     int cnt_gt1, int64_t pi, uint8_t *cls /*[M] zeroed*/, int32_t *queue /*[2M+2]*/, int32_t *order_s,
                     int32_t *cl_off_s, int32_t *seed_of, int32_t *n_cl_s) {
    const int lane = threadIdx.x;
    const unsigned lt = (1u << lane) - 1u;
    const int ncand = *ncand_p;
    int ncl = 0, nmem = 0;
    if (lane == 0) cl_off_s[0] = 0;
    int base = 0;
    while (base < ncand) {
        const int ci = base + lane;
        const int i = ci < ncand ? cand[ci] : -1;
        int c = 1;
        bool isc = false;
        if (i >= 0) {
            c = cls[i];
            isc = core[i] != 0;
        }
        const unsigned seeds = __ballot_sync(0xffffffffu, c == 0 && isc);
        const int fl = seeds ? __ffs(seeds) - 1 : 32; // first unclassified core MC of this window
        if (c == 0 && !isc && lane < fl) cls[i] = 2;   // an unclassified non-core MC reached by the seed loop becomes noise
        if (!seeds) {
            base += 32;
            continue;
        }
        const int fc = __shfl_sync(0xffffffffu, i, fl);
        // ---- expand(fc): the queue starts as a copy of WN(seed), unfiltered (predecon.py:103)
        const int64_t o0 = off[fc];
        const int len0 = (int)(off[fc + 1] - o0);
        for (int t = lane; t < len0; t += 32) queue[t] = col[o0 + t];
        int qt = len0, qh = 0;
        __syncwarp();
        while (qh < qt) {
            const int q = queue[qh++];
            if (!core[q]) continue; // _find_directly_reachable_points: point_is_core
            const int64_t o = off[q];
            const int len = (int)(off[q + 1] - o);
            for (int c0 = 0; c0 < len; c0 += 32) {
                const int t = c0 + lane;
                int x = -1;
                bool enq = false, claim = false;
                if (t < len) {
                    x = col[o + t];
                    const int pd = cnt_gt1 ? popc64(submask[x]) : 0;
                    if ((int64_t)pd <= pi) {
                        const int cx = cls[x];
                        enq = cx == 0;
                        claim = cx == 0 || cx == 2;
                    }
                }
                const unsigned be = __ballot_sync(0xffffffffu, enq), bc = __ballot_sync(0xffffffffu, claim);
                if (enq) queue[qt + __popc(be & lt)] = x;
                if (claim) {
                    order_s[nmem + __popc(bc & lt)] = x;
                    cls[x] = 1;
                }
                qt += __popc(be);
                nmem += __popc(bc);
                __syncwarp(); // the next chunk / pop reads cls and the queue written above
            }
        }
        if (lane == 0) {
            seed_of[ncl] = fc;
            cl_off_s[ncl + 1] = nmem;
        }
        ncl += 1; // emitted even if empty; the host drops clusters whose weight is not > 0 (predecon.py:83)
        base = base + fl + 1;
        __syncwarp();
    }
    if (lane == 0) *n_cl_s = ncl;
}

// The one-warp walk with its whole working set in SHARED memory: the walk is a chain of dependent loads (candidate ->
// class -> core flag -> list bounds -> list -> class of every entry ...), ~0.7 us each from L2, ~4 us per seed; config C4
// has ~850 seeds and that was 3 of the 3.5 ms of the whole growth.  When the non-isolated part of the graph is small
// (22 bytes per CSR entry fit the CTA's shared memory: nnz <= ~9 000) the CTA first copies it -- candidates, local CSR
// (neighbour ids remapped to candidate positions through `rank`), core / pdim flags -- and warp 0 then walks it at
// shared-memory latency.  Outputs as k_offc_grow_warp (global ids).
__global__ void __launch_bounds__(1024, 1)
    k_offc_grow_smem(const int32_t *__restrict__ cand, const int32_t *__restrict__ rank, int M, const int64_t *__restrict__ off,
                     const int32_t *__restrict__ col, const uint8_t *__restrict__ core, const uint64_t *__restrict__ submask,
                     int cnt_gt1, int64_t pi, int32_t *order_s, int32_t *cl_off_s, int32_t *seed_of, int32_t *n_cl_s,
                     int cap_nodes, int cap_nnz) {
    extern __shared__ __align__(16) unsigned char offc_smem[];
    const int ncand = rank[M];
    int32_t *gid = reinterpret_cast<int32_t *>(offc_smem); // [cap_nodes]
    int32_t *loff = gid + cap_nodes;                       // [cap_nodes + 1]
    int32_t *queue = loff + cap_nodes + 1;                 // [2 cap_nodes + 2]
    int32_t *lcol = queue + 2 * cap_nodes + 2;             // [cap_nnz]
    uint8_t *flags = reinterpret_cast<uint8_t *>(lcol + cap_nnz); // [cap_nodes] bit 0 core, bit 1 pdim <= pi
    uint8_t *cls = flags + cap_nodes;                             // [cap_nodes] 0 unclassified, 1 classified, 2 noise
    const int tid = threadIdx.x;
    if (ncand > cap_nodes) { // cannot happen (the launcher sizes by nnz >= ncand); leave a marker instead of corrupting memory
        if (tid == 0) *n_cl_s = -1;
        return;
    }
    for (int c = tid; c < ncand; c += blockDim.x) {
        const int g = cand[c];
        gid[c] = g;
        const int pd = cnt_gt1 ? popc64(submask[g]) : 0;
        flags[c] = (uint8_t)((core[g] ? 1 : 0) | (((int64_t)pd <= pi) ? 2 : 0));
        cls[c] = 0;
    }
    __syncthreads();
    // local CSR: the lists of the candidates are contiguous in the global CSR in candidate order (isolated rows own none)
    const int64_t o_first = ncand ? off[gid[0]] : 0;
    if (ncand == 0) {
        if (tid == 0) {
            cl_off_s[0] = 0;
            *n_cl_s = 0;
        }
        return;
    }
    for (int c = tid; c <= ncand; c += blockDim.x)
        loff[c] = c < ncand ? (int)(off[gid[c]] - o_first) : (int)(off[gid[ncand - 1] + 1] - o_first);
    __syncthreads();
    const int nnz = ncand ? loff[ncand] : 0;
    if (nnz > cap_nnz) {
        if (tid == 0) *n_cl_s = -1;
        return;
    }
    for (int t = tid; t < nnz; t += blockDim.x) lcol[t] = rank[col[o_first + t]];
    __syncthreads();
    if (tid >= 32) return;
    const int lane = tid;
    const unsigned lt = (1u << lane) - 1u;
    int ncl = 0, nmem = 0;
    if (lane == 0) cl_off_s[0] = 0;
    int base = 0;
    while (base < ncand) {
        const int ci = base + lane;
        int c = 1;
        bool isc = false;
        if (ci < ncand) {
            c = cls[ci];
            isc = flags[ci] & 1;
        }
        const unsigned seeds = __ballot_sync(0xffffffffu, c == 0 && isc);
        const int fl = seeds ? __ffs(seeds) - 1 : 32;
        if (c == 0 && !isc && lane < fl) cls[ci] = 2; // an unclassified non-core MC reached by the seed loop becomes noise
        if (!seeds) {
            base += 32;
            continue;
        }
        const int fc = base + fl; // candidate position of the seed
        const int o0 = loff[fc], len0 = loff[fc + 1] - o0;
        for (int t = lane; t < len0; t += 32) queue[t] = lcol[o0 + t];
        int qt = len0, qh = 0;
        __syncwarp();
        while (qh < qt) {
            const int q = queue[qh++];
            if (!(flags[q] & 1)) continue; // _find_directly_reachable_points: point_is_core
            const int o = loff[q], len = loff[q + 1] - o;
            for (int c0 = 0; c0 < len; c0 += 32) {
                const int t = c0 + lane;
                int x = -1;
                bool enq = false, claim = false;
                if (t < len) {
                    x = lcol[o + t];
                    if (flags[x] & 2) {
                        const int cx = cls[x];
                        enq = cx == 0;
                        claim = cx == 0 || cx == 2;
                    }
                }
                const unsigned be = __ballot_sync(0xffffffffu, enq), bc = __ballot_sync(0xffffffffu, claim);
                if (enq) queue[qt + __popc(be & lt)] = x;
                if (claim) {
                    order_s[nmem + __popc(bc & lt)] = gid[x];
                    cls[x] = 1;
                }
                qt += __popc(be);
                nmem += __popc(bc);
                __syncwarp();
            }
        }
        if (lane == 0) {
            seed_of[ncl] = gid[fc];
            cl_off_s[ncl + 1] = nmem;
        }
        ncl += 1;
        base = fc + 1;
        __syncwarp();
    }
    if (lane == 0) *n_cl_s = ncl;
}

// the ordered growth over the non-isolated MCs (PreDeCon.run / _expand, predecon.py:62-120, 242-267)
__global__ void __launch_bounds__(OFFG_THREADS, 1)
    k_offc_grow(int M, const int64_t *__restrict__ off, const int32_t *__restrict__ col, const uint8_t *__restrict__ core,
                const uint8_t *__restrict__ iso, const uint64_t *__restrict__ submask, int cnt_gt1, int64_t pi,
                uint8_t *cls /*[M] zeroed*/, int32_t *queue /*[2M+2]*/, int32_t *order_s, int32_t *cl_off_s, int32_t *seed_of,
                int32_t *n_cl_s) {
    __shared__ int s_warp[32];
    __shared__ int s_tot[2][2];
    __shared__ int s_scan[2];
    const int tid = threadIdx.x;
    const unsigned lt = (1u << (tid & 31)) - 1u;
    int ncl = 0, nmem = 0, par = 0;
    if (tid == 0) cl_off_s[0] = 0;
    int base = 0;
    while (base < M) {
        const int i = base + tid;
        int c = 1;
        bool isc = false;
        if (i < M && !iso[i]) {
            c = cls[i];
            isc = core[i] != 0;
        }
        const int fc = offc_block_min((c == 0 && isc) ? i : INT_MAX, s_warp);
        if (c == 0 && !isc && i < fc) cls[i] = 2; // an unclassified non-core MC reached by the seed loop becomes noise
        __syncthreads();
        if (fc == INT_MAX) {
            base += OFFG_THREADS;
            continue;
        }
        // ---- expand(fc): the queue starts as a copy of WN(seed), unfiltered (predecon.py:103)
        const int64_t o0 = off[fc];
        const int len0 = (int)(off[fc + 1] - o0);
        for (int t = tid; t < len0; t += OFFG_THREADS) queue[t] = col[o0 + t];
        int qt = len0, qh = 0;
        __syncthreads();
        while (qh < qt) {
            const int q = queue[qh++];
            if (!core[q]) continue; // _find_directly_reachable_points: point_is_core
            const int64_t o = off[q];
            const int len = (int)(off[q + 1] - o);
            for (int c0 = 0; c0 < len; c0 += OFFG_THREADS) {
                const int t = c0 + tid;
                int x = -1;
                bool enq = false, claim = false;
                if (t < len) {
                    x = col[o + t];
                    const int pd = cnt_gt1 ? popc64(submask[x]) : 0;
                    if ((int64_t)pd <= pi) {
                        const int cx = cls[x];
                        enq = cx == 0;
                        claim = cx == 0 || cx == 2;
                    }
                }
                par ^= 1;
                if (len - c0 <= 32) { // short list: warp 0 compacts with ballots, in index order
                    if (tid < 32) {
                        const unsigned be = __ballot_sync(0xffffffffu, enq), bc = __ballot_sync(0xffffffffu, claim);
                        if (enq) queue[qt + __popc(be & lt)] = x;
                        if (claim) {
                            order_s[nmem + __popc(bc & lt)] = x;
                            cls[x] = 1;
                        }
                        if (tid == 0) {
                            s_tot[par][0] = __popc(be);
                            s_tot[par][1] = __popc(bc);
                        }
                    }
                    __syncthreads();
                } else {
                    const int qoff = block_excl_scan(enq ? 1 : 0, s_warp, &s_scan[0]);
                    const int qtot = s_scan[0];
                    __syncthreads();
                    const int moff = block_excl_scan(claim ? 1 : 0, s_warp, &s_scan[1]);
                    const int mtot = s_scan[1];
                    if (enq) queue[qt + qoff] = x;
                    if (claim) {
                        order_s[nmem + moff] = x;
                        cls[x] = 1;
                    }
                    if (tid == 0) {
                        s_tot[par][0] = qtot;
                        s_tot[par][1] = mtot;
                    }
                    __syncthreads();
                }
                qt += s_tot[par][0];
                nmem += s_tot[par][1];
            }
        }
        if (tid == 0) {
            seed_of[ncl] = fc;
            cl_off_s[ncl + 1] = nmem;
        }
        ncl += 1; // emitted even if empty; the host drops clusters whose weight is not > 0 (predecon.py:83)
        base = fc + 1;
        __syncthreads();
    }
    if (tid == 0) *n_cl_s = ncl;
}

// seeds of both kinds and the size of the cluster each one emits
__global__ void k_offc_seeds(int M, const uint8_t *__restrict__ iso, const uint8_t *__restrict__ core,
                             const uint64_t *__restrict__ submask, int cnt_gt1, int64_t pi, const int32_t *__restrict__ seed_of,
                             const int32_t *__restrict__ cl_off_s, const int32_t *__restrict__ n_cl_s, int32_t *seedflag,
                             int32_t *size_by_node, int32_t *label) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) {
        label[i] = -1;
        const bool sd = iso[i] && core[i];
        seedflag[i] = sd;
        const int pd = cnt_gt1 ? popc64(submask[i]) : 0;
        size_by_node[i] = (sd && (int64_t)pd <= pi) ? 1 : 0;
    }
}
__global__ void k_offc_seeds_serial(const int32_t *__restrict__ seed_of, const int32_t *__restrict__ cl_off_s,
                                    const int32_t *__restrict__ n_cl_s, int32_t *seedflag, int32_t *size_by_node) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= *n_cl_s) return;
    const int sd = seed_of[c];
    seedflag[sd] = 1;
    size_by_node[sd] = cl_off_s[c + 1] - cl_off_s[c];
}
// size of cluster rank[i] for every seed i
__global__ void k_offc_sizes(int M, const int32_t *__restrict__ seedflag, const int32_t *__restrict__ rank,
                             const int32_t *__restrict__ size_by_node, int32_t *size_by_cluster, int32_t *n_cl) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *n_cl = rank[M];
    if (i < M && seedflag[i]) size_by_cluster[rank[i]] = size_by_node[i];
}
// members of every cluster in claim order, cluster label of every claimed MC
__global__ void k_offc_scatter_iso(int M, const uint8_t *__restrict__ iso, const int32_t *__restrict__ seedflag,
                                   const int32_t *__restrict__ rank, const int32_t *__restrict__ size_by_node,
                                   const int32_t *__restrict__ cl_off, int32_t *order, int32_t *label) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M || !iso[i] || !seedflag[i] || !size_by_node[i]) return;
    const int r = rank[i];
    order[cl_off[r]] = i;
    label[i] = r;
}
__global__ void k_offc_scatter_serial(const int32_t *__restrict__ seed_of, const int32_t *__restrict__ cl_off_s,
                                      const int32_t *__restrict__ n_cl_s, const int32_t *__restrict__ order_s,
                                      const int32_t *__restrict__ rank, const int32_t *__restrict__ cl_off, int32_t *order,
                                      int32_t *label) {
    const int lane = threadIdx.x & 31;
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); c < *n_cl_s; c += nw) {
        const int r = rank[seed_of[c]];
        const int b = cl_off_s[c], n = cl_off_s[c + 1] - b, dst = cl_off[r];
        for (int t = lane; t < n; t += 32) {
            const int x = order_s[b + t];
            order[dst + t] = x;
            label[x] = r;
        }
    }
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
