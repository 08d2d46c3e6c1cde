// online.cuh -- KERNEL 2 (exact ordered commit) and KERNEL 3 (decay / downgrade / compaction).
//
// Replaces the ordered per-cell loop of HDDStream.online_microcluster_maintenance
// (clustering/hddstream.py:220-237): _add_to_pcore :288-343, _add_to_outlier :345-395,
// _upgrade_outlier_microcluster :397-430, _create_new_outlier_cluster :434-462, together with the
// Microcluster maths they call (objects/microcluster.py:89-153, 213-233; utilities/mc_functions.py:14-56),
// and the timepoint-start maintenance (hddstream.py:247-286 decay, :512-549 downgrade / delete).
//
// Kernel 2a  k_pcore_stage   one CTA walks a chunk of cells in input order against ALL potential
//            microclusters, whose state lives in shared memory.  Cells are taken in micro-batches
//            ("waves") of up to 32: (A) every (cell, MC) distance from the state at the start of the
//            wave, (B) warp-shuffle argmin per cell, (C) one warp per candidate MC replays its cells in
//            order -- tentative CF update, variance, preference vector, radius test, commit -- keeping a
//            version of the MC after every accepted cell, (D) every cell re-evaluates only the MCs that
//            were modified earlier in the wave, against the version they had at that cell's turn, and
//            the argmin is repeated, (E) the wave is committed up to the first cell whose argmin changed.
//            A wave of width 1 is the plain sequential algorithm; wider waves give identical results by
//            construction (the first cell of a wave is always exact, so progress is guaranteed).
// Kernel 2b  k_resolve       one CTA walks the rejected cells in order against the outlier list:
//            nearest unmodified MC from kernel 1's snapshot top-K, exact distances to every MC modified or
//            created since the snapshot, radius test, commit / upgrade / create.
// Kernel 3   k_maint_plan + k_maint_gather: fused decay x 2^(-lambda dt), downgrade with the reference's
//            skip-next-after-removal iteration, outlier deletion, order-preserving compaction.
#pragma once
#include "common.cuh"

namespace ccb {

// Device-resident control block shared by the host and the single-CTA kernels.
struct Ctl {
    int32_t n_pcore, n_outlier; // physical list lengths (outlier list includes tombstones)
    int32_t n_outlier_alive, pad0;
    int64_t pcore_last_id, outlier_last_id;
    int64_t pos_end;    // kernel 2a: first cell not yet processed
    int32_t n_rej;      // kernel 2a: rejects collected in this chunk
    int32_t res_done;   // kernel 2b: rejects fully processed so far in this chunk
    int32_t res_reason; // kernel 2b: see RES_*
    int32_t n_dirty;
    double max_w_outlier;
    int64_t waves, rollbacks, pcore_pairs, upgrades, created;
    int32_t new_n_pcore, new_n_outlier; // kernel 3 plan output
    int64_t downgraded, deleted;
    int64_t phase_cycles[8]; // kernel 2a, thread 0: cycles between barriers per phase (diagnostics)
};
enum { RES_DONE = 0, RES_UPGRADE = 1, RES_CUT = 2, RES_OCAP = 3, RES_PCAP = 4 };

struct Store { // one ordered MC list, physical order == list order
    double *cf1, *cf2, *cen; // [cap][D]
    double *w;               // [cap]
    uint64_t *mask;          // [cap]  bit d <=> preferred_dimension_vector[d] == k
    int64_t *id;             // [cap]
    int32_t *uid;            // [cap]  prev_outlier_id (unique per creation)
    double2 *cw;             // [cap][DP] packed (centroid, weight) rows for kernel 1 (outlier list only)
    int32_t cap;
};

struct Num { // numeric parameters common to the ordered kernels
    double delta2, k, wsel, eps2, beta_mu;
    int64_t pi;
    int32_t D, DP, div_mode, pi_active, cnt_gt1; // cnt_gt1: k > 1 (count(pref > 1) == popc(mask)) else 0
};

// ---------------------------------------------------------------------------------------------------
// pieces of arithmetic shared by kernels 2a / 2b

// sum_d ((x_d - c_d)^2) / pref_d, d ascending (mc_functions.py:35-43); x at stride xs, c at stride 1
__device__ __forceinline__ double proj_dist(const double *x, int xs, const double *c, uint64_t mask, const Num &nm) {
    double acc = 0.0;
    for (int d = 0; d < nm.D; ++d) {
        double t = dsub(x[d * xs], c[d]);
        t = dmul(t, t);
        if ((mask >> d) & 1ull) t = nm.div_mode ? ddiv(t, nm.k) : dmul(t, nm.wsel);
        acc = dadd(acc, t);
    }
    return acc;
}

// feasibility gate of _add_to_pcore (hddstream.py:315-321): count(pref' != 1) <= pi on the tentative MC
__device__ __forceinline__ bool feasible(const double *x, int xs, const double *cf1, const double *cf2, double w,
                                         const Num &nm) {
    const double w1 = dadd(w, 1.0);
    int cnt = 0;
    for (int d = 0; d < nm.D; ++d) {
        const double xv = x[d * xs];
        const double a = ddiv(dadd(cf2[d], dmul(xv, xv)), w1);
        double b = ddiv(dadd(cf1[d], xv), w1);
        b = dmul(b, b);
        cnt += (dsub(a, b) <= nm.delta2);
    }
    return (int64_t)cnt <= nm.pi; // pi_active implies k != 1, so "!= 1" counts exactly the preferred bits
}

// One tentative absorb of cell x into an MC held by a warp (lane d owns dims d and d+32):
// get_copy_with_new_point + calculate_projected_radius_squared (microcluster.py:213-233, mc_functions.py:45-56).
// Returns the radius test; on return n1/n2/nc hold CF1', CF2', centroid', *wn = W', *nmask = pref'.
struct LaneMc {
    double cf1[2], cf2[2], cen[2];
};
__device__ __forceinline__ bool tentative_absorb(const LaneMc &m, double w, const double x[2], const Num &nm,
                                                 double *sumscr, LaneMc &o, double &wn, uint64_t &nmask) {
    const int lane = threadIdx.x & 31;
    wn = dadd(w, 1.0);
    uint32_t bits[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int d = lane + 32 * h;
        double term = 0.0;
        bool bit = false;
        if (d < nm.D) {
            o.cf1[h] = dadd(m.cf1[h], x[h]);
            o.cf2[h] = dadd(m.cf2[h], dmul(x[h], x[h]));
            const double a = ddiv(o.cf2[h], wn);
            const double c = ddiv(o.cf1[h], wn);
            o.cen[h] = c;
            const double var = dsub(a, dmul(c, c));
            bit = var <= nm.delta2;
            term = bit ? (nm.div_mode ? ddiv(var, nm.k) : dmul(var, nm.wsel)) : var;
            sumscr[d] = term;
        }
        if (h == 0 || nm.D > 32) bits[h] = __ballot_sync(0xffffffffu, bit);
    }
    nmask = (uint64_t)bits[0] | ((uint64_t)bits[1] << 32);
    __syncwarp();
    double s = 0.0; // sequential sum over d, redundantly in every lane (broadcast LDS)
    for (int d = 0; d < nm.D; ++d) s = dadd(s, sumscr[d]);
    __syncwarp();
    return s <= nm.eps2;
}

// ---------------------------------------------------------------------------------------------------
struct PcoreArgs {
    const double *X;
    int64_t ld, start, end;
    Store P;
    Num nm;
    Ctl *ctl;
    int32_t *assign;
    uint8_t *stage;
    int32_t *rej_list;
    int32_t rej_cap, wave, state_in_smem, dist_in_smem;
    double *dist_gmem; // [Mp][33] fallback
};

constexpr int PCORE_THREADS = 768;
constexpr int PCORE_WARPS = PCORE_THREADS / 32;
constexpr int XS = 33; // padded stride of the transposed wave tile and of the distance matrix

__host__ __device__ inline size_t pcore_smem_bytes(int D, int Mp, bool state, bool dist) {
    size_t dbl = 2 * (size_t)D * XS          // xT, double buffered
                 + 3 * (size_t)32 * D + 32;  // versions cf1, cf2, cen, w
    if (dist) dbl += (size_t)Mp * XS;
    if (state) dbl += 3 * (size_t)Mp * D + Mp;
    size_t b = dbl * 8 + (32 + (state ? (size_t)Mp : 0)) * 8; // v_mask, m_mask
    b += 6 * 32 * 4 + 64;
    return b;
}

// 8-byte asynchronous global->shared copy (LDGSTS); used to prefetch the next wave's cells transposed
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

template <int DP>
__device__ __forceinline__ double proj_dist_t(const double *x, const double *c, uint64_t mask, const Num &nm) {
    double acc = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
        if (d < nm.D) {
            double t = dsub(x[d * XS], c[d]);
            t = dmul(t, t);
            if ((mask >> d) & 1ull) t = nm.div_mode ? ddiv(t, nm.k) : dmul(t, nm.wsel);
            acc = dadd(acc, t);
        }
    }
    return acc;
}

// Tentative absorb held by one warp, lane d owns dims d and d+32; the D radius terms are summed in index
// order by every lane redundantly from warp-shuffle broadcasts (no shared memory, no barrier).
template <int DP>
__device__ __forceinline__ bool tentative_absorb_t(const LaneMc &m, double w, const double x[2], const Num &nm,
                                                   LaneMc &o, double &wn, uint64_t &nmask) {
    const int lane = threadIdx.x & 31;
    wn = dadd(w, 1.0);
    uint32_t bits[2] = {0u, 0u};
    double term[2] = {0.0, 0.0};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        if (h == 0 || DP > 32) {
            const int d = lane + 32 * h;
            // idle lanes (d >= D) carry CF = 1 so that their (discarded) quotients stay on the fast path of the
            // IEEE division; a zero dividend would drag the whole warp through the slow-path subroutine
            const bool act = d < nm.D;
            o.cf1[h] = dadd(m.cf1[h], x[h]);
            o.cf2[h] = dadd(m.cf2[h], dmul(x[h], x[h]));
            const double a = ddiv(o.cf2[h], wn);
            const double c = ddiv(o.cf1[h], wn);
            o.cen[h] = c;
            const double var = dsub(a, dmul(c, c));
            const bool bit = act && (var <= nm.delta2);
            term[h] = bit ? (nm.div_mode ? ddiv(var, nm.k) : dmul(var, nm.wsel)) : var;
            bits[h] = __ballot_sync(0xffffffffu, bit);
        }
    }
    nmask = (uint64_t)bits[0] | ((uint64_t)bits[1] << 32);
    double s = 0.0;
#pragma unroll
    for (int d = 0; d < DP; ++d) {
        if (d < nm.D) s = dadd(s, __shfl_sync(0xffffffffu, term[d >> 5], d & 31));
    }
    return s <= nm.eps2;
}

template <int DP>
__global__ void __launch_bounds__(PCORE_THREADS, 1) k_pcore_stage(PcoreArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Num nm = a.nm;
    const int D = nm.D;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Mp = a.ctl->n_pcore;

    double *sp = reinterpret_cast<double *>(smem_raw);
    double *xTb[2];
    xTb[0] = sp;
    sp += (size_t)D * XS;
    xTb[1] = sp;
    sp += (size_t)D * XS;
    double *v_cf1 = sp;
    sp += 32 * D;
    double *v_cf2 = sp;
    sp += 32 * D;
    double *v_cen = sp;
    sp += 32 * D;
    double *v_w = sp;
    sp += 32;
    double *dist = a.dist_gmem;
    if (a.dist_in_smem) {
        dist = sp;
        sp += (size_t)Mp * XS;
    }
    double *m_cf1 = a.P.cf1, *m_cf2 = a.P.cf2, *m_cen = a.P.cen, *m_w = a.P.w;
    if (a.state_in_smem) {
        m_cf1 = sp;
        sp += (size_t)Mp * D;
        m_cf2 = sp;
        sp += (size_t)Mp * D;
        m_cen = sp;
        sp += (size_t)Mp * D;
        m_w = sp;
        sp += Mp;
    }
    uint64_t *up = reinterpret_cast<uint64_t *>(sp);
    uint64_t *v_mask = up;
    up += 32;
    uint64_t *m_mask = a.P.mask;
    if (a.state_in_smem) {
        m_mask = up;
        up += Mp;
    }
    int *ip = reinterpret_cast<int *>(up);
    int *cand = ip;
    ip += 32;
    int *distinct = ip;
    ip += 32;
    unsigned *dmask = reinterpret_cast<unsigned *>(ip);
    ip += 32;
    unsigned *amask = reinterpret_cast<unsigned *>(ip);
    ip += 32;
    int *misc = ip; // [0] n_distinct, [1] m (commit length), [2] stop flag, [3] nrej (running), [4] rollback flag

    if (a.state_in_smem) {
        for (int i = tid; i < Mp * D; i += PCORE_THREADS) {
            m_cf1[i] = a.P.cf1[i];
            m_cf2[i] = a.P.cf2[i];
            m_cen[i] = a.P.cen[i];
        }
        for (int i = tid; i < Mp; i += PCORE_THREADS) {
            m_w[i] = a.P.w[i];
            m_mask[i] = a.P.mask[i];
        }
    }
    if (tid == 0) {
        misc[2] = 0;
        misc[3] = 0;
    }
    const double maxw0 = fmax(a.ctl->max_w_outlier, 0.0);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    int64_t pos = a.start;
    int64_t st_waves = 0, st_roll = 0, st_pairs = 0;

    // prefetch of a wave tile: cells [p0, p0 + 32) clipped to the chunk, transposed into xT[d][i]
    auto prefetch = [&](int buf, int64_t p0) {
        const int nb = (int)min((int64_t)a.wave, a.end - p0);
        for (int i = tid; i < nb * D; i += PCORE_THREADS) {
            const int r = i / D, d = i - r * D;
            cp_async8(&xTb[buf][d * XS + r], a.X + (p0 + r) * a.ld + d);
        }
    };
    long long pc_t = clock64();
    long long pc_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define CCB_PHASE(k)                     \
    do {                                 \
        if (tid == 0) {                  \
            const long long n_ = clock64(); \
            pc_acc[k] += n_ - pc_t;      \
            pc_t = n_;                   \
        }                                \
    } while (0)
    int buf = 0;
    if (pos < a.end) prefetch(0, pos);
    int64_t pf_pos = pos; // position the tile in xTb[buf] was fetched for
    cp_async_commit_wait_all();
    __syncthreads();

    while (pos < a.end && !misc[2]) {
        const int b = (int)min((int64_t)a.wave, a.end - pos);
        if (pf_pos != pos) { // the previous wave was cut short: the speculative tile is misaligned
            prefetch(buf, pos);
            pf_pos = pos;
            cp_async_commit_wait_all();
            __syncthreads();
        }
        const double *xT = xTb[buf];
        // speculative prefetch of the next wave (assumes this one commits completely)
        if (pos + b < a.end) prefetch(buf ^ 1, pos + b);
        asm volatile("cp.async.commit_group;" ::: "memory");
        // ---- A: speculative distances against the state at the start of the wave
        for (int pidx = tid; pidx < Mp * 32; pidx += PCORE_THREADS) {
            const int i = pidx & 31, j = pidx >> 5;
            if (i < b) {
                double dv;
                if (nm.pi_active && !feasible(xT + i, XS, m_cf1 + (size_t)j * D, m_cf2 + (size_t)j * D, m_w[j], nm))
                    dv = qnan; // infeasible: never a candidate
                else
                    dv = proj_dist_t<DP>(xT + i, m_cen + (size_t)j * D, m_mask[j], nm);
                dist[(size_t)j * XS + i] = dv;
            }
        }
        __syncthreads();
        CCB_PHASE(0);
        // ---- B + C0 (warp 0): argmin per cell (lane = cell, ascending MC index, strict <), then the distinct
        //      candidate MCs with the bitmask of their cells (ascending bit = input order)
        if (warp == 0) {
            double bd = 0.0;
            int bj = -1;
            if (lane < b) {
                for (int j = 0; j < Mp; ++j) {
                    const double v = dist[(size_t)j * XS + lane];
                    if (!(v != v) && (bj < 0 || v < bd)) {
                        bd = v;
                        bj = j;
                    }
                }
            }
            cand[lane] = bj;
            const unsigned grp = __match_any_sync(0xffffffffu, bj);
            const bool leader = (bj >= 0) && ((__ffs(grp) - 1) == lane);
            const unsigned lead = __ballot_sync(0xffffffffu, leader);
            if (leader) {
                const int r = __popc(lead & ((1u << lane) - 1u));
                distinct[r] = bj;
                dmask[r] = grp;
            }
            if (lane == 0) misc[0] = __popc(lead);
        }
        __syncthreads();
        CCB_PHASE(1);
        const int nd = misc[0];
        // ---- C: one warp per candidate MC replays its cells in order
        for (int e = warp; e < nd; e += PCORE_WARPS) {
            const int j = distinct[e];
            unsigned pts = dmask[e];
            LaneMc m;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int d = lane + 32 * h;
                m.cf1[h] = d < D ? m_cf1[(size_t)j * D + d] : 1.0;
                m.cf2[h] = d < D ? m_cf2[(size_t)j * D + d] : 1.0;
                m.cen[h] = 0.0;
            }
            double w = m_w[j];
            unsigned acc = 0u;
            while (pts) {
                const int i = __ffs(pts) - 1;
                pts &= pts - 1;
                double x[2];
                x[0] = lane < D ? xT[lane * XS + i] : 0.0;
                x[1] = (DP > 32 && lane + 32 < D) ? xT[(lane + 32) * XS + i] : 0.0;
                LaneMc o;
                double wn;
                uint64_t nmask;
                if (tentative_absorb_t<DP>(m, w, x, nm, o, wn, nmask)) {
                    m = o;
                    w = wn;
                    acc |= 1u << i;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int d = lane + 32 * h;
                        if ((h == 0 || DP > 32) && d < D) {
                            v_cf1[i * D + d] = o.cf1[h];
                            v_cf2[i * D + d] = o.cf2[h];
                            v_cen[i * D + d] = o.cen[h];
                        }
                    }
                    if (lane == 0) {
                        v_w[i] = wn;
                        v_mask[i] = nmask;
                    }
                }
            }
            if (lane == 0) amask[e] = acc;
        }
        __syncthreads();
        CCB_PHASE(2);
        // ---- D: re-evaluate (cell, MC) pairs whose MC changed earlier in the wave
        int npatch = 0;
        for (int pidx = tid; pidx < nd * 32; pidx += PCORE_THREADS) {
            const int i = pidx & 31, e = pidx >> 5;
            if (i < b && i > 0) {
                const unsigned prior = amask[e] & ((1u << i) - 1u);
                if (prior) {
                    const int v = 31 - __clz(prior);
                    const int j = distinct[e];
                    double dv;
                    if (nm.pi_active && !feasible(xT + i, XS, v_cf1 + v * D, v_cf2 + v * D, v_w[v], nm))
                        dv = qnan;
                    else
                        dv = proj_dist_t<DP>(xT + i, v_cen + v * D, v_mask[v], nm);
                    dist[(size_t)j * XS + i] = dv;
                    ++npatch;
                }
            }
        }
        st_pairs += npatch;
        __syncthreads();
        CCB_PHASE(3);
        // ---- D2 + E (warp 0): argmin again, first mismatch, reject budget, per-cell outputs
        if (warp == 0) {
            double bd = 0.0;
            int bj = -1;
            if (lane < b) {
                for (int j = 0; j < Mp; ++j) {
                    const double v = dist[(size_t)j * XS + lane];
                    if (!(v != v) && (bj < 0 || v < bd)) {
                        bd = v;
                        bj = j;
                    }
                }
            }
            const int mycand = cand[lane];
            const bool bad = (lane < b) && (bj != mycand);
            const unsigned badm = __ballot_sync(0xffffffffu, bad);
            int m = badm ? (__ffs(badm) - 1) : b; // cell m itself is not committed (its candidate was wrong)
            unsigned accb = 0u;
            for (int e = lane; e < nd; e += 32) accb |= amask[e];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) accb |= __shfl_xor_sync(0xffffffffu, accb, o);
            const unsigned low = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
            unsigned rej = ~accb & low;
            // reject budget: an upgrade (hddstream.py:413-418) needs W >= beta*mu, and after q rejects of
            // this chunk no outlier MC can weigh more than max_w_outlier + q; the chunk ends one reject
            // before that becomes possible, so an upgrade can only ever happen on a chunk's last cell.
            int nrej = misc[3];
            int stop = 0;
            unsigned walk = rej;
            while (walk) {
                const int i = __ffs(walk) - 1;
                walk &= walk - 1;
                ++nrej;
                if (nrej >= a.rej_cap || (double)nrej + maxw0 + 1.0 >= nm.beta_mu) {
                    m = i + 1;
                    stop = 1;
                    break;
                }
            }
            const unsigned low2 = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
            rej &= low2;
            if (lane < m) {
                const int64_t r = pos + lane;
                if ((rej >> lane) & 1u) {
                    a.assign[r] = -1;
                    a.rej_list[misc[3] + __popc(rej & ((1u << lane) - 1u))] = (int32_t)r;
                } else {
                    a.assign[r] = a.P.uid[mycand];
                    if (a.stage) a.stage[r] = 0;
                }
            }
            __syncwarp();
            if (lane == 0) {
                misc[1] = m;
                misc[2] = stop;
                misc[3] += __popc(rej);
                misc[4] = badm ? 1 : 0;
            }
        }
        __syncthreads();
        CCB_PHASE(4);
        const int m = misc[1];
        {
            const unsigned low = (m >= 32) ? 0xffffffffu : ((1u << m) - 1u);
            for (int idx = tid; idx < nd * D; idx += PCORE_THREADS) {
                const int e = idx / D, d = idx - e * D;
                const unsigned am = amask[e] & low;
                if (am) {
                    const int v = 31 - __clz(am), j = distinct[e];
                    m_cf1[(size_t)j * D + d] = v_cf1[v * D + d];
                    m_cf2[(size_t)j * D + d] = v_cf2[v * D + d];
                    m_cen[(size_t)j * D + d] = v_cen[v * D + d];
                    if (d == 0) {
                        m_w[j] = v_w[v];
                        m_mask[j] = v_mask[v];
                    }
                }
            }
        }
        st_waves += 1;
        st_roll += misc[4];
        if (tid == 0) st_pairs += (int64_t)b * Mp;
        pf_pos = pos + b; // what the speculative prefetch targeted
        pos += m;
        buf ^= 1;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        CCB_PHASE(5);
    }
#undef CCB_PHASE
    if (tid == 0)
        for (int k = 0; k < 8; ++k) a.ctl->phase_cycles[k] += pc_acc[k];

    if (a.state_in_smem) {
        for (int i = tid; i < Mp * D; i += PCORE_THREADS) {
            a.P.cf1[i] = m_cf1[i];
            a.P.cf2[i] = m_cf2[i];
            a.P.cen[i] = m_cen[i];
        }
        for (int i = tid; i < Mp; i += PCORE_THREADS) {
            a.P.w[i] = m_w[i];
            a.P.mask[i] = m_mask[i];
        }
    }
    // per-thread patch counts -> one atomic per warp
    int64_t pairs = st_pairs;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    if (lane == 0) atomicAdd(reinterpret_cast<unsigned long long *>(&a.ctl->pcore_pairs), (unsigned long long)pairs);
    if (tid == 0) {
        a.ctl->pos_end = pos;
        a.ctl->n_rej = misc[3];
        a.ctl->res_done = 0; // the outlier stage of this chunk starts from a fresh snapshot
        a.ctl->res_reason = RES_DONE;
        a.ctl->n_dirty = 0;
        a.ctl->waves += st_waves;
        a.ctl->rollbacks += st_roll;
    }
}

// ---------------------------------------------------------------------------------------------------
// Kernel 2b: ordered resolution of the rejected cells against the outlier list.
struct ResolveArgs {
    const double *X;
    int64_t ld;
    Store O, P;
    Num nm;
    Ctl *ctl;
    const int32_t *rej_list;
    int32_t q_snap;   // index into rej_list of the first cell covered by the snapshot top-K arrays
    int32_t mo_snap;  // outlier list length when the snapshot was taken
    int32_t topk;     // K
    const double *tk_dist; // [n_rej - q_snap][K]
    const int32_t *tk_idx;
    uint8_t *dirty;      // [O.cap] flags, zero at snapshot time
    int32_t *dirty_list; // [>= rejects per chunk]
    int32_t *assign;
    uint8_t *stage;
};

constexpr int RES_THREADS = 1024;

__global__ void __launch_bounds__(RES_THREADS, 1) k_resolve(ResolveArgs a) {
    __shared__ double xs[CCB_MAX_D];
    __shared__ double sumscr[CCB_MAX_D];
    __shared__ double red_d[RES_THREADS / 32];
    __shared__ int red_i[RES_THREADS / 32];
    __shared__ int s_win;
    __shared__ int s_flag;
    const Num nm = a.nm;
    const int D = nm.D, DP = nm.DP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    Ctl *ctl = a.ctl;
    const int n_rej = ctl->n_rej;
    int q = ctl->res_done;
    int reason = RES_DONE;

    for (; q < n_rej; ++q) {
        const int64_t r = a.rej_list[q];
        if (tid < D) xs[tid] = a.X[r * a.ld + tid];
        __syncthreads();
        const int mo = ctl->n_outlier; // uniform: written only by thread 0 before a barrier
        const int nd = ctl->n_dirty;
        // ---- exact distances to every MC modified since the snapshot and every MC created since
        double bd = 0.0;
        int bj = -1;
        const int nlist = nd + (mo - a.mo_snap);
        for (int t = tid; t < nlist; t += RES_THREADS) {
            const int j = t < nd ? a.dirty_list[t] : a.mo_snap + (t - nd);
            const double2 *c = a.O.cw + (size_t)j * DP;
            double acc = 0.0;
            for (int d = 0; d < D; ++d) {
                const double2 cv = c[d];
                double tt = dsub(xs[d], cv.x);
                tt = dmul(tt, tt);
                tt = nm.div_mode ? ddiv(tt, cv.y) : dmul(tt, cv.y);
                acc = dadd(acc, tt);
            }
            if (!(acc != acc) && better(acc, j, bd, bj)) { // NaN = tombstone
                bd = acc;
                bj = j;
            }
        }
        warp_argmin(bd, bj);
        if (lane == 0) {
            red_d[warp] = bd;
            red_i[warp] = bj;
        }
        __syncthreads();
        if (warp == 0) {
            bd = red_d[lane];
            bj = red_i[lane];
            warp_argmin(bd, bj);
            // ---- nearest unmodified MC of the snapshot
            int flag = 0;
            if (lane == 0) {
                const double *td = a.tk_dist + (size_t)(q - a.q_snap) * a.topk;
                const int32_t *ti = a.tk_idx + (size_t)(q - a.q_snap) * a.topk;
                int s = 0;
                for (; s < a.topk; ++s) {
                    const int j = ti[s];
                    if (j < 0) break; // list exhausted: no further snapshot MC exists
                    if (!a.dirty[j]) {
                        if (better(td[s], j, bd, bj)) {
                            bd = td[s];
                            bj = j;
                        }
                        break;
                    }
                }
                if (s == a.topk) flag = 1; // every listed candidate is stale: the best clean MC is unknown
                s_win = bj;
                s_flag = flag;
            }
        }
        __syncthreads();
        if (s_flag) {
            reason = RES_CUT;
            break;
        }
        const int win = s_win;
        bool absorbed = false;
        if (win >= 0) {
            // ---- radius test on the tentative MC (warp 0), commit on success
            if (warp == 0) {
                LaneMc m, o;
                double x[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int d = lane + 32 * h;
                    m.cf1[h] = d < D ? a.O.cf1[(size_t)win * D + d] : 0.0;
                    m.cf2[h] = d < D ? a.O.cf2[(size_t)win * D + d] : 0.0;
                    m.cen[h] = 0.0;
                    x[h] = d < D ? xs[d] : 0.0;
                }
                double wn;
                uint64_t nmask;
                const bool ok = tentative_absorb(m, a.O.w[win], x, nm, sumscr, o, wn, nmask);
                if (ok) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int d = lane + 32 * h;
                        if (d < D) {
                            a.O.cf1[(size_t)win * D + d] = o.cf1[h];
                            a.O.cf2[(size_t)win * D + d] = o.cf2[h];
                            a.O.cen[(size_t)win * D + d] = o.cen[h];
                            double2 cv;
                            cv.x = o.cen[h];
                            cv.y = ((nmask >> d) & 1ull) ? nm.wsel : 1.0;
                            a.O.cw[(size_t)win * DP + d] = cv;
                        }
                    }
                    if (lane == 0) {
                        a.O.w[win] = wn;
                        a.O.mask[win] = nmask;
                        if (win < a.mo_snap && !a.dirty[win]) {
                            a.dirty[win] = 1;
                            a.dirty_list[ctl->n_dirty] = win;
                            ctl->n_dirty = ctl->n_dirty + 1;
                        }
                        if (wn > ctl->max_w_outlier) ctl->max_w_outlier = wn;
                        a.assign[r] = a.O.uid[win];
                        // upgrade test (hddstream.py:413-418)
                        const int pd = nm.cnt_gt1 ? popc64(nmask) : 0;
                        int up = (wn >= nm.beta_mu) && ((int64_t)pd <= nm.pi);
                        if (up && ctl->n_pcore >= a.P.cap) up = 2;
                        s_flag = up;
                        if (a.stage) a.stage[r] = up == 1 ? 2 : 1;
                    }
                }
                if (lane == 0) s_win = ok ? win : -2;
            }
            __syncthreads();
            absorbed = s_win >= 0;
            if (absorbed && s_flag == 2) { // pcore list full: undo is impossible, so the host must never let this happen
                reason = RES_PCAP;
                ++q;
                break;
            }
            if (absorbed && s_flag == 1) {
                // ---- move to the tail of the pcore list with a fresh pcore id; tombstone the outlier slot
                const int pj = ctl->n_pcore;
                for (int d = tid; d < D; d += RES_THREADS) {
                    a.P.cf1[(size_t)pj * D + d] = a.O.cf1[(size_t)win * D + d];
                    a.P.cf2[(size_t)pj * D + d] = a.O.cf2[(size_t)win * D + d];
                    a.P.cen[(size_t)pj * D + d] = a.O.cen[(size_t)win * D + d];
                }
                __syncthreads();
                if (tid == 0) {
                    a.P.w[pj] = a.O.w[win];
                    a.P.mask[pj] = a.O.mask[win];
                    a.P.uid[pj] = a.O.uid[win];
                    a.P.id[pj] = ctl->pcore_last_id;
                    ctl->pcore_last_id += 1;
                    ctl->n_pcore = pj + 1;
                    a.O.w[win] = -1.0; // tombstone marker (a live weight is never negative)
                    a.O.cw[(size_t)win * DP].x = __longlong_as_double(0x7ff8000000000000LL);
                    ctl->n_outlier_alive -= 1;
                    ctl->upgrades += 1;
                }
                reason = RES_UPGRADE;
                ++q;
                break;
            }
        }
        if (!absorbed) {
            // ---- new outlier MC at the tail (hddstream.py:434-462); variance is exactly 0 -> all dims preferred
            if (mo >= a.O.cap) {
                reason = RES_OCAP;
                break;
            }
            const uint64_t full = D >= 64 ? ~0ull : ((1ull << D) - 1ull);
            for (int d = tid; d < DP; d += RES_THREADS) {
                double2 cv;
                if (d < D) {
                    const double xv = xs[d];
                    const double c1 = dadd(0.0, xv);
                    a.O.cf1[(size_t)mo * D + d] = c1;
                    a.O.cf2[(size_t)mo * D + d] = dadd(0.0, dmul(xv, xv));
                    a.O.cen[(size_t)mo * D + d] = ddiv(c1, 1.0);
                    cv.x = ddiv(c1, 1.0);
                    cv.y = (0.0 <= nm.delta2) ? nm.wsel : 1.0;
                } else {
                    cv.x = 0.0;
                    cv.y = 1.0;
                }
                a.O.cw[(size_t)mo * DP + d] = cv;
            }
            if (tid == 0) {
                a.O.w[mo] = 1.0;
                a.O.mask[mo] = (0.0 <= nm.delta2) ? full : 0ull;
                a.O.id[mo] = ctl->outlier_last_id;
                a.O.uid[mo] = (int32_t)ctl->outlier_last_id;
                a.assign[r] = (int32_t)ctl->outlier_last_id;
                if (a.stage) a.stage[r] = 3;
                ctl->outlier_last_id += 1;
                ctl->n_outlier = mo + 1;
                ctl->n_outlier_alive += 1;
                ctl->created += 1;
                if (1.0 > ctl->max_w_outlier) ctl->max_w_outlier = 1.0;
            }
        }
        __syncthreads();
    }
    __syncthreads();
    if (tid == 0) {
        ctl->res_done = q;
        ctl->res_reason = reason;
    }
}

// max weight over live outlier MCs (feeds the reject budget of kernel 2a)
__global__ void k_max_w(const double *w, Ctl *ctl) {
    __shared__ double red[32];
    const int n = ctl->n_outlier;
    double m = -1.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, w[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -1.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) ctl->max_w_outlier = m;
    }
}

// ---------------------------------------------------------------------------------------------------
// Kernel 3.  Plan: one CTA decides, for the decayed weights, which pcore MCs are downgraded and which
// outlier MCs are deleted, reproducing Python's mutate-while-iterating behaviour: an element is examined
// iff the element before it (in the current list) was not just removed (hddstream.py:528-537, 545-549).
// removed[j] = cond[j] && !removed[j-1]  -- a two-state machine, composed across threads.
// Output: new position of every old pcore / outlier slot (-1 = gone; downgraded pcore MCs get a position in
// the NEW outlier list encoded as -2 - newpos).
struct MaintArgs {
    Store P, O, P2, O2; // old and new (ping-pong) stores
    Ctl *ctl;
    int32_t *p_new, *p_fin, *o_new; // [P.cap] phase-1 codes, [P.cap] final codes, [O.cap]
    double f, beta_mu, omicron;
    int64_t pi;
    int32_t D, DP, cnt_gt1;
    double wsel;
};

constexpr int MAINT_THREADS = 1024;

__global__ void __launch_bounds__(MAINT_THREADS, 1) k_maint_plan(MaintArgs a) {
    __shared__ int s_cnt[MAINT_THREADS];
    __shared__ unsigned char s_out0[MAINT_THREADS], s_out1[MAINT_THREADS], s_in[MAINT_THREADS];
    __shared__ int s_nmoved, s_np_new;
    const int tid = threadIdx.x;
    Ctl *ctl = a.ctl;
    const int np = ctl->n_pcore, no = ctl->n_outlier;
    // ---- pcore pass (short list): one thread walks it
    if (tid == 0) {
        int prev_removed = 0, keep = 0, moved = 0;
        for (int j = 0; j < np; ++j) {
            const double w = dmul(a.P.w[j], a.f);
            const int pd = a.cnt_gt1 ? popc64(a.P.mask[j]) : 0;
            const int cond = (w < a.beta_mu) || ((int64_t)pd > a.pi);
            const int rem = cond && !prev_removed;
            a.p_new[j] = rem ? (-2 - moved) : keep;
            a.p_fin[j] = keep; // overwritten below for downgraded MCs
            moved += rem;
            keep += !rem;
            prev_removed = rem;
        }
        s_nmoved = moved;
        s_np_new = keep;
    }
    __syncthreads();
    const int nmoved = s_nmoved;
    // ---- outlier pass over the sequence: live old outliers in order, then the downgraded pcores in order
    const int total = no + nmoved;
    const int per = (total + MAINT_THREADS - 1) / MAINT_THREADS;
    const int lo = min(total, tid * per), hi = min(total, lo + per);
    auto elem = [&](int e, bool &live, bool &cond) {
        if (e < no) {
            const double w0 = a.O.w[e];
            live = w0 >= 0.0; // tombstones carry -1
            cond = live && (dmul(w0, a.f) <= a.omicron);
        } else {
            // e - no-th downgraded pcore: find it (np is small)
            int k = e - no, jj = -1;
            for (int j = 0; j < np; ++j)
                if (a.p_new[j] == -2 - k) {
                    jj = j;
                    break;
                }
            live = true;
            cond = dmul(a.P.w[jj], a.f) <= a.omicron;
        }
    };
    {
        int st0 = 0, st1 = 1; // removed-state of the previous live element, for both possible inputs
        for (int e = lo; e < hi; ++e) {
            bool live, cond;
            elem(e, live, cond);
            if (!live) continue;
            st0 = cond && !st0;
            st1 = cond && !st1;
        }
        s_out0[tid] = (unsigned char)st0;
        s_out1[tid] = (unsigned char)st1;
    }
    __syncthreads();
    if (tid == 0) {
        int st = 0;
        for (int t = 0; t < MAINT_THREADS; ++t) {
            s_in[t] = (unsigned char)st;
            st = st ? s_out1[t] : s_out0[t];
        }
    }
    __syncthreads();
    int kept = 0;
    {
        int st = s_in[tid];
        for (int e = lo; e < hi; ++e) {
            bool live, cond;
            elem(e, live, cond);
            if (!live) continue;
            st = cond && !st;
            kept += !st;
        }
    }
    s_cnt[tid] = kept;
    __syncthreads();
    if (tid == 0) {
        int run = 0;
        for (int t = 0; t < MAINT_THREADS; ++t) {
            const int c = s_cnt[t];
            s_cnt[t] = run;
            run += c;
        }
        ctl->new_n_outlier = run;
        ctl->new_n_pcore = s_np_new;
        ctl->downgraded += nmoved;
    }
    __syncthreads();
    {
        int st = s_in[tid], posn = s_cnt[tid], ndel = 0;
        for (int e = lo; e < hi; ++e) {
            bool live, cond;
            elem(e, live, cond);
            int np_ = -1;
            if (live) {
                st = cond && !st;
                if (!st) np_ = posn++;
                else ++ndel;
            }
            if (e < no) {
                a.o_new[e] = np_;
            } else {
                const int k = e - no;
                for (int j = 0; j < np; ++j)
                    if (a.p_new[j] == -2 - k) {
                        // encode: moved and kept at outlier position np_  ->  -2 - np_ ; moved and deleted -> -1
                        a.p_fin[j] = np_ >= 0 ? (-2 - np_) : -1;
                        break;
                    }
            }
        }
        if (ndel) atomicAdd(reinterpret_cast<unsigned long long *>(&ctl->deleted), (unsigned long long)ndel);
    }
}

// Gather: writes the new (compacted, decayed) stores.  CF1, CF2, W are scaled by f; centroid and preference
// vector are NOT touched (hddstream.py:268-286, asserted by the reference's unittest_hddstream.py:90-121).
__global__ void k_maint_gather(MaintArgs a) {
    const Ctl *ctl = a.ctl;
    const int np = ctl->n_pcore, no = ctl->n_outlier;
    const int D = a.D, DP = a.DP;
    const int64_t total = (int64_t)(np + no) * DP;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int e = (int)(i / DP), d = (int)(i % DP);
        const bool from_p = e < np;
        const int j = from_p ? e : e - np;
        const Store &S = from_p ? a.P : a.O;
        const int code = from_p ? a.p_fin[j] : a.o_new[j];
        if (code == -1) continue;
        const bool to_p = from_p && code >= 0;
        const int nj = to_p ? code : (from_p ? (-2 - code) : code);
        const Store &T = to_p ? a.P2 : a.O2;
        if (d < D) {
            T.cf1[(size_t)nj * D + d] = dmul(S.cf1[(size_t)j * D + d], a.f);
            T.cf2[(size_t)nj * D + d] = dmul(S.cf2[(size_t)j * D + d], a.f);
            T.cen[(size_t)nj * D + d] = S.cen[(size_t)j * D + d];
        }
        if (!to_p) {
            double2 cv;
            cv.x = d < D ? S.cen[(size_t)j * D + d] : 0.0;
            cv.y = (d < D && ((S.mask[j] >> d) & 1ull)) ? a.wsel : 1.0;
            T.cw[(size_t)nj * DP + d] = cv;
        }
        if (d == 0) {
            T.w[nj] = dmul(S.w[j], a.f);
            T.mask[nj] = S.mask[j];
            T.uid[nj] = S.uid[j];
            T.id[nj] = (from_p && !to_p) ? (int64_t)S.uid[j] : S.id[j]; // downgrade: id <- [prev_outlier_id]
        }
    }
}

__global__ void k_maint_finish(Ctl *ctl) {
    ctl->n_pcore = ctl->new_n_pcore;
    ctl->n_outlier = ctl->new_n_outlier;
    ctl->n_outlier_alive = ctl->new_n_outlier;
}

// Rebuilds the packed (centroid, weight) rows of a store (after an import).
__global__ void k_repack_store(Store S, int n, int D, int DP, double wsel) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)n * DP) return;
    const int j = (int)(i / DP), d = (int)(i % DP);
    double2 cv;
    cv.x = d < D ? S.cen[(size_t)j * D + d] : 0.0;
    cv.y = (d < D && ((S.mask[j] >> d) & 1ull)) ? wsel : 1.0;
    S.cw[i] = cv;
}

// ---- scaler (scaling/scaler.py:11-53 -> sklearn MinMaxScaler) ---------------------------------------
// transform: X * scale_ + min_ as TWO roundings (sklearn: `X *= self.scale_; X += self.min_`), in place on the device
// copy of a timepoint, so the host never makes the scaled pass over the data.  HBM-bound: 16 B per element.
__global__ void k_scale_rows(double *__restrict__ X, int64_t N, int64_t ld, int D, const double *__restrict__ scale,
                             const double *__restrict__ shift) {
    const int64_t total = N * D;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / D;
        const int d = (int)(e - r * D);
        double *p = X + r * ld + d;
        *p = dadd(dmul(*p, scale[d]), shift[d]);
    }
}

// fit: column-wise minimum / maximum ignoring NaN (np.nanmin / np.nanmax, what MinMaxScaler.partial_fit uses).
// Doubles are compared through the order-preserving map to unsigned integers, so that one atomicMin / atomicMax per
// column and CTA settles the result exactly.  okey[D] / okey[D + d]: running min / max keys (initialised by the caller).
__device__ __forceinline__ unsigned long long ordered_key(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__global__ void k_colminmax(const double *__restrict__ X, int64_t N, int64_t ld, int D, unsigned long long *okey) {
    __shared__ unsigned long long s_min[CCB_MAX_D], s_max[CCB_MAX_D];
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        s_min[d] = ~0ull;
        s_max[d] = 0ull;
    }
    __syncthreads();
    // a stride that is a multiple of D keeps every thread on one column
    const int64_t total = N * D;
    const int64_t stride = ((int64_t)gridDim.x * blockDim.x + D - 1) / D * D;
    const int64_t e0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e0 < total) {
        const int d = (int)(e0 % D);
        unsigned long long lo = ~0ull, hi = 0ull;
        for (int64_t e = e0; e < total; e += stride) {
            const double v = X[(e / D) * ld + d];
            if (v == v) { // not NaN
                const unsigned long long k = ordered_key(v);
                lo = k < lo ? k : lo;
                hi = k > hi ? k : hi;
            }
        }
        if (lo <= hi) {
            atomicMin(&s_min[d], lo);
            atomicMax(&s_max[d], hi);
        }
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        if (s_min[d] <= s_max[d]) {
            atomicMin(&okey[d], s_min[d]);
            atomicMax(&okey[D + d], s_max[d]);
        }
    }
}
__global__ void k_colminmax_finish(const unsigned long long *okey, int D, double *mn, double *mx) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    auto back = [](unsigned long long k) {
        const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        return __longlong_as_double((long long)b);
    };
    const bool any = okey[d] <= okey[D + d];
    mn[d] = any ? back(okey[d]) : __longlong_as_double(0x7ff8000000000000LL); // all-NaN column: NaN, like np.nanmin
    mx[d] = any ? back(okey[D + d]) : __longlong_as_double(0x7ff8000000000000LL);
}

} // namespace ccb
