// nearest.cuh -- KERNEL 1: tiled fp64 cell x microcluster preference-weighted projected distance with
// a top-K (K = 1: argmin) selection.
//
// Replaces: Microcluster.get_projected_dist_to_point (objects/microcluster.py:167-181 ->
// utilities/mc_functions.py:35-43) evaluated for every MC of a list, and the strict-< first-wins
// argmin of the scans (clustering/hddstream.py:311-328, 371-375).
//
// Layout.  Microclusters are packed as cw[M][DP] double2 = (centroid_d, weight_d), row-major, DP = D
// rounded up to a multiple of 4 with (0.0, 1.0) padding (adds exact zeros).  weight_d is 1/pref_d
// when k is a power of two (x * 2^-m is bit-identical to x / 2^m for every x, subnormals included)
// and pref_d itself otherwise (true IEEE division).  A dead slot carries NaN in its first centroid
// word: its distance is NaN and never wins a strict-< comparison.
//
// Mapping.  One thread owns PPT cells whose DP coordinates live in registers for the whole kernel;
// the MC axis is cut into slabs (blockIdx.y) and each slab is streamed through shared memory in tiles
// of TM microclusters by 1-D bulk TMA copies (cp.async.bulk + mbarrier, double buffered).  Every lane
// of a warp reads the same (c, w) pair -> one broadcast LDS.128 per (MC, dim).  The sum over
// dimensions is a sequential dependent chain per (cell, MC) as parity requires; instruction-level
// parallelism comes from JU microclusters x PPT cells advancing together.
//
// Roofline: FP64 pipe.  Algorithmic work = 4*D flops per evaluated pair (sub, mul, div, add); bytes =
// 8*D per cell read + 12 B per cell written.
#pragma once
#include "common.cuh"

namespace ccb {

constexpr int NEAREST_THREADS = 128;
constexpr int NEAREST_JU = 4;

// K > 1 is the in-engine use (a few hundred cells x 1e4 microclusters per launch): there the kernel is bound by the
// instruction stream of a single warp per work item, so a thread owns ONE cell and the split makes more, smaller items
template <int DP, int K = 1>
struct NearestCfg {
    static constexpr int PPT = (K == 1 && DP <= 40) ? 2 : 1;
    // tile of TM microclusters, ~12-16 KB per buffer, TM a multiple of JU
    static constexpr int TM = ((1024 / DP) < 8 ? 8 : (1024 / DP)) / NEAREST_JU * NEAREST_JU;
    static constexpr int CELLS = NEAREST_THREADS * PPT;
    // SPLIT layout (K > 1 and few rows: the steady state of the engine, a few hundred cells against 1e4 microclusters): a
    // work item is 32 cells (lane = cell, the same cells in every warp) and the CTA's warps share the tile's microclusters
    // -- four times the warps on the same work, each with a quarter of the instruction stream
    static constexpr int TMS = (TM / (4 * (NEAREST_THREADS / 32))) * (4 * (NEAREST_THREADS / 32)) > 0
                                   ? (TM / (4 * (NEAREST_THREADS / 32))) * (4 * (NEAREST_THREADS / 32))
                                   : TM;
    static constexpr bool CAN_SPLIT = K > 1 && TM >= 4 * (NEAREST_THREADS / 32);
};
constexpr int NEAREST_SPLIT_MAX_ROWS = 1024; // rows of one launch up to which the SPLIT layout is used

// guards of the fused weight step (nearest_item): a non-zero magnitude below 2^-400, or a weight below 2^-64 / not a
// power of two (never produced by k_pack_cw / the commit kernels, checked anyway), forces the unfused sequence
__device__ __forceinline__ bool tiny_nonzero(double v) {
    const uint64_t b = (uint64_t)__double_as_longlong(v) & 0x7fffffffffffffffull;
    return b != 0ull && b < 0x26f0000000000000ull; // biased exponent 623 = 2^-400
}
__device__ __forceinline__ bool small_weight(double w) {
    const uint64_t b = (uint64_t)__double_as_longlong(w);
    return (b & 0x000fffffffffffffull) != 0ull || b < 0x3bf0000000000000ull || b > 0x3ff0000000000000ull; // [2^-64, 1]
}

template <int K>
__device__ __forceinline__ void topk_insert(double (&bd)[K], int (&bi)[K], double d, int j) {
    // strict <: an equal distance never displaces an earlier (smaller-index) entry
    if constexpr (K == 1) {
        if (!(d < bd[0])) return;
        bd[0] = d;
        bi[0] = j;
        return;
    } else {
    // K > 1 (in-engine lists): BRANCH-FREE insertion into the ascending list.  lt[s] = d < bd[s] is monotone in s, so the
    // new entry s is the old entry s - 1 where lt[s - 1], the candidate where only lt[s], itself otherwise.  A list that
    // starts empty takes a candidate nearly every time in its first hundred microclusters, in some lane of the warp nearly
    // always: with data-dependent branches the warp walked all K compare-and-swap steps divergently for every candidate
    // (3 000 cycles per group of four, measured), with selects it is K compares + 3 K selects.
    bool lt[K];
#pragma unroll
    for (int s = 0; s < K; ++s) lt[s] = d < bd[s];
#pragma unroll
    for (int s = K - 1; s > 0; --s) {
        bd[s] = lt[s - 1] ? bd[s - 1] : (lt[s] ? d : bd[s]);
        bi[s] = lt[s - 1] ? bi[s - 1] : (lt[s] ? j : bi[s]);
    }
    bd[0] = lt[0] ? d : bd[0];
    bi[0] = lt[0] ? j : bi[0];
    }
}

// Device-side work split of the block-speculative engine's calls: R rows x M microclusters are cut into
// row groups of `cells` rows and nslab slabs of slab_mcs microclusters so that about `target` CTAs have work,
// whatever R and M turn out to be on the device (the host only knows upper bounds).
__device__ __forceinline__ void dyn_split(int R, int M, int target, int max_slabs, int tm, int cells, int &groups,
                                          int &slab_mcs, int &nslab) {
    groups = (R + cells - 1) / cells;
    const int tiles = (M + tm - 1) / tm;
    int want = groups > 0 ? (target + groups - 1) / groups : 1;
    want = min(want, min(max_slabs, tiles));
    want = max(want, 1);
    slab_mcs = max(1, (tiles + want - 1) / want) * tm;
    nslab = M > 0 ? (M + slab_mcs - 1) / slab_mcs : 0;
}

// One (row group, slab) work item: cells [cell0, cell0 + CELLS) of the row list against MCs [j0, j1).
// seq counts the tiles this CTA has streamed so far (mbarrier buffer / parity bookkeeping across items).
template <int DP, int K, bool DIV, bool SPLIT = false>
__device__ __forceinline__ void nearest_item(const double *__restrict__ X, const int32_t *__restrict__ rows, int64_t row_off,
                                             int64_t nrows, int64_t ld, int D, const double2 *__restrict__ cw, int64_t cell0,
                                             int j0, int j1, double *__restrict__ out_dist, int32_t *__restrict__ out_idx,
                                             int64_t out_off, int out_stride, int out_slab, double2 (*tile)[NearestCfg<DP, K>::TM * DP],
                                             uint64_t *bar, uint32_t &seq) {
    using Cfg = NearestCfg<DP, K>;
    constexpr int PPT = Cfg::PPT, TM = SPLIT ? Cfg::TMS : Cfg::TM, JU = NEAREST_JU;
    constexpr int NWARP = NEAREST_THREADS / 32, QS = TM / NWARP; // SPLIT: microclusters of a tile per warp
    static_assert(!SPLIT || (PPT == 1 && QS % JU == 0 && QS > 0), "split layout");
    // ---- this thread's cells -> registers (row-contiguous 16-byte loads; every fetched sector is used)
    double p[PPT][DP];
    int64_t cell[PPT];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
        cell[u] = cell0 + (SPLIT ? (threadIdx.x & 31) : threadIdx.x) + u * NEAREST_THREADS;
        const bool live = cell[u] < nrows;
        const int64_t r = live ? (rows ? (int64_t)rows[row_off + cell[u]] : row_off + cell[u]) : 0;
        const double *xr = X + r * ld;
        if (live && ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0)) {
#pragma unroll
            for (int d = 0; d < DP; d += 2) {
                if (d + 1 < D) {
                    double2 v = *reinterpret_cast<const double2 *>(xr + d);
                    p[u][d] = v.x;
                    p[u][d + 1] = v.y;
                } else {
                    p[u][d] = (d < D) ? xr[d] : 0.0;
                    p[u][d + 1] = 0.0;
                }
            }
        } else {
#pragma unroll
            for (int d = 0; d < DP; ++d) p[u][d] = (live && d < D) ? xr[d] : 0.0;
        }
    }
    if (K > 1) { CCB_TS(24); }
    double bd[PPT][K];
    int bi[PPT][K];
#pragma unroll
    for (int u = 0; u < PPT; ++u)
#pragma unroll
        for (int s = 0; s < K; ++s) {
            bd[u][s] = __longlong_as_double(0x7ff0000000000000LL);
            bi[u][s] = -1;
        }

    const int ntiles = (j1 - j0 + TM - 1) / TM;
    auto issue = [&](int t) {
        const int jt = j0 + t * TM;
        const int n = min(TM, j1 - jt);
        const uint32_t bytes = (uint32_t)n * DP * (uint32_t)sizeof(double2);
        const uint32_t q = seq + (uint32_t)t;
        mbar_expect_tx(&bar[q & 1], bytes);
        tma_load_1d(&tile[q & 1][0], cw + (size_t)jt * DP, bytes, &bar[q & 1]);
    };
    if (threadIdx.x == 0 && ntiles > 0) issue(0);

    // FUSED weight step (power-of-two weights only): acc = fma(t^2, w, acc) instead of acc + (t^2 * w).  The product
    // of a double and 2^-m is exact unless it lands in the subnormal range, so the two are the same IEEE result
    // whenever every non-zero (p - c)^2 is far above 2^-1022 / w.  That holds when every coordinate involved is 0 or
    // at least 2^-400 in magnitude: p and c are then multiples of 2^-452, a non-zero difference is >= 2^-452, its
    // square >= 2^-904.  The kernel checks that itself -- this warp's cells here, every tile below (each warp scans
    // the whole tile: ~DP * TM / 32 integer compares per lane against ~3 * DP * TM * PPT fp64 instructions) -- and
    // falls back to the unfused sequence for a tile that fails, so the result is bit-exact unconditionally.
    bool tiny_p = false;
    if (!DIV) {
#pragma unroll
        for (int u = 0; u < PPT; ++u)
#pragma unroll
            for (int d = 0; d < DP; ++d) tiny_p |= tiny_nonzero(p[u][d]);
        tiny_p = __any_sync(0xffffffffu, tiny_p);
    }

    for (int t = 0; t < ntiles; ++t) {
        if (threadIdx.x == 0 && t + 1 < ntiles) issue(t + 1); // that buffer was released by the barrier below
        const uint32_t q = seq + (uint32_t)t;
        mbar_wait(&bar[q & 1], (q >> 1) & 1);
        if (K > 1 && t == 0) { CCB_TS(25); }
        const double2 *tl = tile[q & 1];
        const int jt = j0 + t * TM;
        const int n = min(TM, j1 - jt);
        bool fuse = false;
        if (!DIV) {
            bool tiny_c = tiny_p;
            if (!tiny_c) {
                const int lane = threadIdx.x & 31;
                for (int e = lane; e < n * DP; e += 32) tiny_c |= tiny_nonzero(tl[e].x) | small_weight(tl[e].y);
                tiny_c = __any_sync(0xffffffffu, tiny_c);
            }
            fuse = !tiny_c;
        }
        const int jlo = SPLIT ? (int)(threadIdx.x >> 5) * QS : 0, jhi = SPLIT ? min(n, jlo + QS) : n;
        for (int jj = jlo; jj < jhi; jj += JU) {
            double acc[PPT][JU];
#pragma unroll
            for (int u = 0; u < PPT; ++u)
#pragma unroll
                for (int v = 0; v < JU; ++v) acc[u][v] = 0.0;
            if (fuse) {
#pragma unroll
                for (int d = 0; d < DP; ++d) {
#pragma unroll
                    for (int v = 0; v < JU; ++v) {
                        const double2 c = tl[(jj + v) * DP + d]; // broadcast LDS.128 (garbage past n is masked below)
#pragma unroll
                        for (int u = 0; u < PPT; ++u) {
                            double tt = dsub(p[u][d], c.x);
                            tt = dmul(tt, tt);
                            acc[u][v] = __fma_rn(tt, c.y, acc[u][v]); // exact product (see above): == acc + tt * w
                        }
                    }
                }
            } else {
#pragma unroll
                for (int d = 0; d < DP; ++d) {
#pragma unroll
                    for (int v = 0; v < JU; ++v) {
                        const double2 c = tl[(jj + v) * DP + d];
#pragma unroll
                        for (int u = 0; u < PPT; ++u) {
                            double tt = dsub(p[u][d], c.x);
                            tt = dmul(tt, tt);
                            tt = DIV ? ddiv(tt, c.y) : dmul(tt, c.y);
                            acc[u][v] = dadd(acc[u][v], tt);
                        }
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < JU; ++v) {
                if (jj + v < jhi) {
                    if (K == 1) {
#pragma unroll
                        for (int u = 0; u < PPT; ++u) topk_insert<K>(bd[u], bi[u], acc[u][v], jt + jj + v);
                    } else { // (warp-uniform branch: skipped once every lane's list has settled below the candidate)
                        bool any = false;
#pragma unroll
                        for (int u = 0; u < PPT; ++u) any |= acc[u][v] < bd[u][K - 1];
                        if (__any_sync(0xffffffffu, any)) {
#pragma unroll
                            for (int u = 0; u < PPT; ++u) topk_insert<K>(bd[u], bi[u], acc[u][v], jt + jj + v);
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    seq += (uint32_t)ntiles;
    if (K > 1) { CCB_TS(26); CCB_TS_ANY(27); }

    if constexpr (SPLIT) {
        // the warps' lists of the same 32 cells -> shared memory (the tile buffers are free: the loop above ended with a
        // barrier), then warp 0 merges the four ascending lists of every cell.  The warps cover ascending index ranges of
        // every tile, but tiles interleave them, so ties are broken by the index explicitly (better()).
        double *sd = reinterpret_cast<double *>(&tile[0][0]);
        int *si = reinterpret_cast<int *>(&tile[1][0]);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int s = 0; s < K; ++s) {
            sd[(warp * 32 + lane) * K + s] = bd[0][s];
            si[(warp * 32 + lane) * K + s] = bi[0][s];
        }
        __syncthreads();
        if (warp == 0 && cell[0] < nrows) {
            int pw[NWARP], hi_[NWARP];
            double hd_[NWARP];
#pragma unroll
            for (int w = 0; w < NWARP; ++w) {
                pw[w] = 0;
                hd_[w] = sd[(w * 32 + lane) * K];
                hi_[w] = si[(w * 32 + lane) * K];
            }
            const size_t o = ((size_t)(out_off + cell[0]) * out_stride + out_slab) * K;
            for (int s = 0; s < K; ++s) {
                double bdv = 0.0;
                int biv = -1, bw = 0;
#pragma unroll
                for (int w = 0; w < NWARP; ++w)
                    if (better(hd_[w], hi_[w], bdv, biv)) {
                        bdv = hd_[w];
                        biv = hi_[w];
                        bw = w;
                    }
                out_dist[o + s] = biv >= 0 ? bdv : __longlong_as_double(0x7ff0000000000000LL);
                out_idx[o + s] = biv;
#pragma unroll
                for (int w = 0; w < NWARP; ++w)
                    if (w == bw && biv >= 0) {
                        pw[w] += 1;
                        hi_[w] = -1;
                        if (pw[w] < K) {
                            hd_[w] = sd[(w * 32 + lane) * K + pw[w]];
                            hi_[w] = si[(w * 32 + lane) * K + pw[w]];
                        }
                    }
            }
        }
        __syncthreads(); // the next item's first tile lands in the same memory
    } else {
#pragma unroll
        for (int u = 0; u < PPT; ++u) {
            if (cell[u] < nrows) {
                const size_t o = ((size_t)(out_off + cell[u]) * out_stride + out_slab) * K;
#pragma unroll
                for (int s = 0; s < K; ++s) {
                    out_dist[o + s] = bd[u][s];
                    out_idx[o + s] = bi[u][s];
                }
            }
        }
    }
}

// rows == nullptr: cell r is row r of X.  nrows_dev (optional) overrides nrows with a device-side count.
// out_dist/out_idx: [nrows][nslab][K]; unused entries hold (+inf, -1).
// STATIC split (range_dev == nullptr): grid = (row groups, slabs) chosen by the host.
// DYNAMIC split (range_dev != nullptr, block-speculative engine): 1-D grid; rows[range_dev[0] .. range_dev[1])
// against the first min(M, *M_dev) microclusters; every CTA derives the split with dyn_split and loops over its
// work items; results land at [absolute list position][dyn_max_slabs][K].
template <int DP, int K, bool DIV>
__global__ void __launch_bounds__(NEAREST_THREADS)
    k_nearest(const double *__restrict__ X, const int32_t *__restrict__ rows, const int32_t *__restrict__ nrows_dev,
              int64_t row_off, int64_t nrows, int64_t ld, int D, const double2 *__restrict__ cw, int M, int slab_mcs,
              double *__restrict__ out_dist, int32_t *__restrict__ out_idx, const int32_t *__restrict__ range_dev,
              const int32_t *__restrict__ M_dev, int dyn_max_slabs, const XRef *__restrict__ xref) {
    using Cfg = NearestCfg<DP, K>;
    __shared__ __align__(128) double2 tile[2][Cfg::TM * DP];
    __shared__ __align__(8) uint64_t bar[2];

    if (xref) { // graph launches: the input array of this call
        X = xref->X;
        ld = xref->ld;
    }
    CCB_TS(3);
    CCB_PDL();
    if (nrows_dev) nrows = (int64_t)(*nrows_dev) - row_off;
    if (M_dev) M = min(M, *M_dev);
    if (range_dev) {
        row_off = range_dev[0];
        nrows = (int64_t)range_dev[1] - row_off;
        if (nrows <= 0) return;
    } else if ((int64_t)blockIdx.x * Cfg::CELLS >= nrows) {
        return;
    }
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    uint32_t seq = 0;
    if (range_dev) {
        int groups, smcs, nslab;
        if (Cfg::CAN_SPLIT && nrows <= NEAREST_SPLIT_MAX_ROWS) {
            dyn_split((int)nrows, M, gridDim.x, dyn_max_slabs, Cfg::TMS, 32, groups, smcs, nslab);
            const int items = groups * nslab;
            for (int it = blockIdx.x; it < items; it += gridDim.x) {
                const int g = it / nslab, sl = it - g * nslab;
                const int j0 = sl * smcs;
                nearest_item<DP, K, DIV, Cfg::CAN_SPLIT>(X, rows, row_off, nrows, ld, D, cw, (int64_t)g * 32, j0, min(M, j0 + smcs),
                                                         out_dist, out_idx, row_off, dyn_max_slabs, sl, tile, bar, seq);
            }
            return;
        }
        dyn_split((int)nrows, M, gridDim.x, dyn_max_slabs, Cfg::TM, Cfg::CELLS, groups, smcs, nslab);
        const int items = groups * nslab;
        for (int it = blockIdx.x; it < items; it += gridDim.x) {
            const int g = it / nslab, sl = it - g * nslab;
            const int j0 = sl * smcs;
            nearest_item<DP, K, DIV>(X, rows, row_off, nrows, ld, D, cw, (int64_t)g * Cfg::CELLS, j0, min(M, j0 + smcs),
                                     out_dist, out_idx, row_off, dyn_max_slabs, sl, tile, bar, seq);
        }
    } else {
        const int j0 = blockIdx.y * slab_mcs;
        nearest_item<DP, K, DIV>(X, rows, row_off, nrows, ld, D, cw, (int64_t)blockIdx.x * Cfg::CELLS, j0,
                                 min(M, j0 + slab_mcs), out_dist, out_idx, 0, gridDim.y, blockIdx.y, tile, bar, seq);
    }
}

// Merge for the DYNAMIC split: one warp per row; lane l walks the (already ascending) lists of slabs l, l + 32, ...
// and the warp extracts the K lexicographically smallest (dist, idx) entries -- ties go to the smaller list
// position, exactly like a scan in list order with strict <.
template <int K>
__global__ void __launch_bounds__(128) k_topk_merge_dyn(const double *__restrict__ in_dist, const int32_t *__restrict__ in_idx,
                                                        const int32_t *__restrict__ range_dev, const int32_t *__restrict__ M_dev,
                                                        int M, int target, int max_slabs, int tm, int cells, int tm_split,
                                                        double *__restrict__ out_dist, int32_t *__restrict__ out_idx) {
    constexpr int QMAX = 8; // max_slabs <= 256
    CCB_TS(4);
    CCB_PDL();
    const int lane = threadIdx.x & 31;
    const int R = range_dev[1] - range_dev[0];
    if (M_dev) M = min(M, *M_dev);
    int groups, smcs, nslab;
    if (tm_split > 0 && R <= NEAREST_SPLIT_MAX_ROWS) dyn_split(R, M, target, max_slabs, tm_split, 32, groups, smcs, nslab); // (as k_nearest)
    else dyn_split(R, M, target, max_slabs, tm, cells, groups, smcs, nslab);
    const int nw = gridDim.x * (blockDim.x >> 5);
    for (int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < R; r += nw) {
        const size_t row = (size_t)range_dev[0] + r;
        // every lane keeps the head of each of its lists AND the entry behind it in registers: a list that wins a round
        // moves on to an entry that is already there, and the load of the one after that has a whole round (usually many)
        // to arrive -- no dependent global load inside the K extraction rounds
        int hd[QMAX], ci[QMAX], ni[QMAX];
        double cd[QMAX], nd[QMAX];
#pragma unroll
        for (int q = 0; q < QMAX; ++q) {
            const int sl = lane + 32 * q;
            hd[q] = 0;
            ci[q] = ni[q] = -1;
            cd[q] = nd[q] = 0.0;
            if (sl < nslab) {
                const size_t o = (row * max_slabs + sl) * K;
                ci[q] = in_idx[o];
                cd[q] = in_dist[o];
                if (K > 1) {
                    ni[q] = in_idx[o + 1];
                    nd[q] = in_dist[o + 1];
                }
            }
        }
        for (int s = 0; s < K; ++s) {
            double bdv = 0.0;
            int biv = -1, bq = 0;
#pragma unroll
            for (int q = 0; q < QMAX; ++q) {
                if (better(cd[q], ci[q], bdv, biv)) { // (an exhausted or empty list carries index -1: never better)
                    bdv = cd[q];
                    biv = ci[q];
                    bq = q;
                }
            }
            double wd = bdv;
            int wi = biv;
            warp_argmin(wd, wi);
            if (wi >= 0 && wi == biv) {
#pragma unroll
                for (int q = 0; q < QMAX; ++q)
                    if (q == bq) {
                        hd[q] += 1;
                        cd[q] = nd[q];
                        ci[q] = ni[q];
                        ni[q] = -1;
                        if (hd[q] + 1 < K) {
                            const size_t o = (row * max_slabs + lane + 32 * q) * K + hd[q] + 1;
                            ni[q] = in_idx[o];
                            nd[q] = in_dist[o];
                        }
                    }
            }
            if (lane == 0) {
                out_dist[row * K + s] = wi >= 0 ? wd : __longlong_as_double(0x7ff0000000000000LL);
                out_idx[row * K + s] = wi;
            }
        }
    }
}

// Merges the per-slab top-K lists of every cell into one ascending list of K (dist, idx) entries.
// Slabs cover ascending index ranges, so scanning them in order with strict < keeps ties on the
// smaller index.
template <int K>
__global__ void k_topk_merge(const double *__restrict__ in_dist, const int32_t *__restrict__ in_idx,
                             const int32_t *__restrict__ nrows_dev, int64_t row_off, int64_t nrows, int nslab,
                             double *__restrict__ out_dist, int32_t *__restrict__ out_idx,
                             const int32_t *__restrict__ range_dev) {
    if (nrows_dev) nrows = (int64_t)(*nrows_dev) - row_off;
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (range_dev) {
        nrows = (int64_t)range_dev[1] - range_dev[0];
        if (c >= nrows) return;
        c += range_dev[0];
    } else if (c >= nrows) {
        return;
    }
    double bd[K];
    int bi[K];
#pragma unroll
    for (int s = 0; s < K; ++s) {
        bd[s] = __longlong_as_double(0x7ff0000000000000LL);
        bi[s] = -1;
    }
    for (int e = 0; e < nslab * K; ++e) {
        const int j = in_idx[(size_t)c * nslab * K + e];
        if (j >= 0) topk_insert<K>(bd, bi, in_dist[(size_t)c * nslab * K + e], j);
    }
#pragma unroll
    for (int s = 0; s < K; ++s) {
        out_dist[(size_t)c * K + s] = bd[s];
        out_idx[(size_t)c * K + s] = bi[s];
    }
}

// ---- association scan ----------------------------------------------------------------------------------
// TrackByHistoricalAssociation.track_cluster_history (tracking/cluster_tracker.py:127-144): for every CURRENT pcore MC q
// the previous-timepoint pcore MC whose centroid p_j minimises q.get_projected_dist_to_point(p_j)
// (microcluster.py:167-181 -> mc_functions.py:35-43: sum_d ((p_jd - c_qd)^2) / pref_qd, the QUERY's centroid and
// preference vector), scanning j in list order, first strictly smaller wins.  Same arithmetic as kernel 1 with the roles
// swapped: here the thread-resident side carries the weights.  One thread owns one query (c_q, w_q in registers), the
// scanned centroids stream through shared memory in double-buffered tiles (1-D bulk TMA), every lane reads the same
// coordinates (broadcast LDS), 4 scanned MCs advance together; the fused weight step is guarded exactly like kernel 1's.
constexpr int ASSOC_THREADS = 128;
constexpr int ASSOC_JU = 4;
template <int DP>
struct AssocCfg {
    static constexpr int TM = ((2048 / DP) < 32 ? 32 : (2048 / DP)) / ASSOC_JU * ASSOC_JU;
};

template <int DP, bool DIV>
__global__ void __launch_bounds__(ASSOC_THREADS)
    k_assoc(const double *__restrict__ qcen, const uint64_t *__restrict__ qmask, int Q, const double *__restrict__ pcen, int P,
            int D, double k, double wsel, int32_t *__restrict__ best, double *__restrict__ bdist,
            double *__restrict__ bdist2) {
    constexpr int TM = AssocCfg<DP>::TM, JU = ASSOC_JU;
    __shared__ __align__(128) double tile[2][TM * DP];
    __shared__ __align__(8) uint64_t bar[2];
    const int q = blockIdx.x * ASSOC_THREADS + threadIdx.x;
    const bool live = q < Q;
    double c[DP], w[DP];
    bool tiny_q = false;
    {
        const uint64_t m = live ? qmask[q] : 0ull;
#pragma unroll
        for (int d = 0; d < DP; ++d) {
            c[d] = (live && d < D) ? qcen[(size_t)q * D + d] : 0.0;
            w[d] = (d < D && ((m >> d) & 1ull)) ? (DIV ? k : wsel) : 1.0;
            tiny_q |= tiny_nonzero(c[d]) | (!DIV && small_weight(w[d]));
        }
    }
    tiny_q = __any_sync(0xffffffffu, tiny_q);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int ntiles = (P + TM - 1) / TM;
    auto issue = [&](int t) {
        const int jt = t * TM;
        const int n = min(TM, P - jt);
        uint32_t bytes = (uint32_t)((size_t)n * D * sizeof(double));
        if (bytes & 15u) { // odd element count: the last double travels by a plain store (ordered by the arrive)
            bytes -= 8u;
            tile[t & 1][(size_t)n * D - 1] = pcen[(size_t)jt * D + (size_t)n * D - 1];
        }
        mbar_expect_tx(&bar[t & 1], bytes);
        if (bytes) tma_load_1d(&tile[t & 1][0], pcen + (size_t)jt * D, bytes, &bar[t & 1]);
    };
    if (threadIdx.x == 0 && ntiles > 0) issue(0);
    int bi = -1;
    double bd = 0.0, bd2 = __longlong_as_double(0x7ff0000000000000LL); // runner-up distance (+inf: none)
    for (int t = 0; t < ntiles; ++t) {
        if (threadIdx.x == 0 && t + 1 < ntiles) issue(t + 1);
        mbar_wait(&bar[t & 1], (t >> 1) & 1);
        const double *tl = tile[t & 1];
        const int jt = t * TM;
        const int n = min(TM, P - jt);
        bool fuse = false;
        if (!DIV) {
            bool tiny = tiny_q;
            if (!tiny) {
                for (int e = threadIdx.x & 31; e < n * D; e += 32) tiny |= tiny_nonzero(tl[e]);
                tiny = __any_sync(0xffffffffu, tiny);
            }
            fuse = !tiny;
        }
        for (int j0 = 0; j0 < n; j0 += JU) {
            double acc[JU];
#pragma unroll
            for (int v = 0; v < JU; ++v) acc[v] = 0.0;
#pragma unroll
            for (int d = 0; d < DP; ++d) {
                if (d < D) {
#pragma unroll
                    for (int v = 0; v < JU; ++v) {
                        const double pv = tl[(size_t)(j0 + v) * D + d]; // rows past n: stale shared memory, masked below
                        double tt = dsub(pv, c[d]);
                        tt = dmul(tt, tt);
                        if (fuse) acc[v] = __fma_rn(tt, w[d], acc[v]); // exact product (see nearest_item)
                        else acc[v] = dadd(acc[v], DIV ? ddiv(tt, w[d]) : dmul(tt, w[d]));
                    }
                }
            }
#pragma unroll
            for (int v = 0; v < JU; ++v) {
                if (j0 + v >= n) continue;
                if (bi < 0 || acc[v] < bd) { // strict <: the first scanned MC keeps a tie
                    if (bi >= 0) bd2 = bd;
                    bi = jt + j0 + v;
                    bd = acc[v];
                } else if (acc[v] < bd2) {
                    bd2 = acc[v];
                }
            }
        }
        __syncthreads();
    }
    if (live) {
        best[q] = bi;
        bdist[q] = bd;
        if (bdist2) bdist2[q] = bd2;
    }
}

// Builds the packed (centroid, weight) rows from separate centroid / preference-mask arrays.
__global__ void k_pack_cw(const double *__restrict__ cen, const uint64_t *__restrict__ mask, int64_t M, int D, int DP,
                          double wsel, double2 *__restrict__ cw) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M * DP) return;
    const int64_t j = i / DP;
    const int d = (int)(i % DP);
    double2 v;
    v.x = d < D ? cen[j * D + d] : 0.0;
    v.y = (d < D && ((mask[j] >> d) & 1ull)) ? wsel : 1.0;
    cw[i] = v;
}

} // namespace ccb
