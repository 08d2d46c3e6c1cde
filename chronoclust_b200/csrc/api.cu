// api.cu -- host side of the C ABI declared in include/chronoclust_b200.h: device-resident
// structure-of-arrays microcluster stores, chunk orchestration of the ordered kernels, the offline
// pipeline, import/export.  One handle = one device + one stream.  No CPU fallback anywhere: every
// numeric result is produced by the kernels in nearest.cuh / online.cuh / offline.cuh.
#include "../../include/chronoclust_b200.h"
#ifdef CCB_DEBUG
#include "debug.h"
#endif

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "nearest.cuh"
#include "offline.cuh"
#include "online.cuh"
#include "engine.cuh"

using namespace ccb;

namespace {

thread_local std::string g_err;

const int kDPs[] = {4, 8, 12, 16, 24, 32, 40, 48, 64};
int round_dp(int D) {
    for (int v : kDPs)
        if (D <= v) return v;
    return -1;
}
#define CCB_DISPATCH_DP(dp, ...)                    \
    switch (dp) {                                   \
    case 4: { constexpr int kDP = 4; __VA_ARGS__ } break;  \
    case 8: { constexpr int kDP = 8; __VA_ARGS__ } break;  \
    case 12: { constexpr int kDP = 12; __VA_ARGS__ } break; \
    case 16: { constexpr int kDP = 16; __VA_ARGS__ } break; \
    case 24: { constexpr int kDP = 24; __VA_ARGS__ } break; \
    case 32: { constexpr int kDP = 32; __VA_ARGS__ } break; \
    case 40: { constexpr int kDP = 40; __VA_ARGS__ } break; \
    case 48: { constexpr int kDP = 48; __VA_ARGS__ } break; \
    case 64: { constexpr int kDP = 64; __VA_ARGS__ } break; \
    default: break;                                 \
    }

bool is_pow2(double k) {
    if (!(k > 0.0) || std::isinf(k)) return false;
    int e;
    return std::frexp(k, &e) == 0.5;
}

constexpr int MAX_SLABS = 148;

} // namespace

struct ccb_handle {
    ccb_params prm{};
    int D = 0, DP = 0, div_mode = 0, cnt_gt1 = 0;
    double wsel = 1.0;
    cudaStream_t stream = nullptr;
    Store P[2]{}, O[2]{};
    int pcur = 0, ocur = 0;
    Ctl *d_ctl = nullptr, *h_ctl = nullptr;
    double mu = 0, omicron = 0;
    int64_t pi = 0;
    bool have_params = false;
    // ingest buffers
    double *d_X = nullptr;
    size_t x_cap = 0;
    int32_t *d_assign = nullptr;
    uint8_t *d_stage = nullptr;
    size_t n_cap = 0;
    int32_t *d_pnew = nullptr, *d_pfin = nullptr, *d_onew = nullptr; // kernel 3 plan codes
    // block-speculative engine (engine.cuh)
    int bs_bmax = 32768, bs_bmin = 1024, bs_iters = 3;
    BsCtl *d_bc = nullptr, *h_bc = nullptr;
    BsWs ws{};
    std::vector<void *> ws_allocs;
    void *ws_tiles[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int32_t *ws_first = nullptr;
    int ws_first_cap = 0;
    double *d_bs_tk_dist_slab = nullptr;
    int32_t *d_bs_tk_idx_slab = nullptr;
    BsCtl bc_base{}; // counters folded in by ccb_reset
    bool bs_use_graph = true;   // CUDA graph with device-driven WHILE nodes (stream launches when per-kernel timing is on)
    // Two cached graphs: the double-buffered stores flip at every ccb_begin_timepoint with decay, so consecutive timepoints
    // alternate between two sets of (baked-in) pointers -- with one slot the graph was re-captured and re-instantiated at
    // every timepoint (~1 ms each).
    // programmatic dependent launch between the engine's kernels: measured neutral on B200 (profiles/r2zc_pdl_experiment.md:
    // the gaps between the kernels of a round are not launch latency), so it is off; the kernels keep CCB_PDL() (a no-op then)
    bool pdl = false;
    cudaGraph_t bs_graph = nullptr, bs_graph_alt = nullptr;
    cudaGraphExec_t bs_exec = nullptr, bs_exec_alt = nullptr;
    Eng bs_graph_eng_alt{};
    cudaStream_t cap1 = nullptr, cap2 = nullptr, cap3 = nullptr; // capture streams: block loop, round loop, side branch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t copy_stream = nullptr; // host -> device segments of ccb_ingest, ahead of the engine
    cudaStream_t res_stream = nullptr;  // device -> host copies of the per-cell results of a finished segment, behind the engine
    unsigned char *h_res = nullptr;     // page-locked bounce buffer of those results ([n] int32 assignments, then [n] stages)
    size_t res_cap = 0;                 // (rows)
    double *d_scale = nullptr;          // [2][CCB_MAX_D] scale_ / min_ of ccb_ingest_scaled
    cudaEvent_t ev_seg[8] = {nullptr};
    Eng bs_graph_eng{};         // the pointers / capacities the graph was captured with
    EngIo *d_io = nullptr, *h_io = nullptr;
    // offline results (host copies)
    int64_t off_M = 0;
    std::vector<int64_t> cl_off, cl_members;
    std::vector<double> cl_w, cl_cf1, cl_cf2, cl_cen, cl_pref;
    std::vector<int32_t> cl_label;
    // offline workspace: grow-only device buffers, kept between calls (the white-box export reads them lazily)
    struct OffBuf {
        void *p = nullptr;
        size_t bytes = 0;
    };
    enum { OB_CORE, OB_CLS, OB_NBR, OB_WNBR, OB_CNT, OB_BORDER, OB_NBORDER, OB_QUEUE, OB_LABEL, OB_ORDER, OB_CLOFF, OB_NCL,
           OB_SUBMASK, OB_OMASK, OB_D1, OB_D2, OB_DC, OB_DW, OB_DEC, OB_COUNT };
    OffBuf ob[OB_COUNT];
    int off_csr_min_m = 0;
    ccb_stats st{};
    ccb_stats st_base{}; // device-side counters folded in by ccb_reset
    std::string err;
    void *dnrm2 = nullptr;
    // optional per-category GPU timing (CUDA events on the handle's stream)
    bool timing = false;
    struct Ev { cudaEvent_t a, b; int cat; };
    std::vector<Ev> ev_pending;
    std::vector<cudaEvent_t> ev_pool;
    double cat_ms[CCB_NCAT] = {0};
    int64_t cat_n[CCB_NCAT] = {0};
};

namespace {

int fail(ccb_handle *h, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    g_err = buf;
    return code;
}
#define CK(h, call)                                                                                        \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(h, e_ == cudaErrorMemoryAllocation ? CCB_ENOMEM : CCB_ECUDA, "%s:%d %s: %s", __FILE__, \
                        __LINE__, #call, cudaGetErrorString(e_));                                          \
    } while (0)
#define CKL(h) CK(h, cudaGetLastError())

// RAII bracket: records CUDA events around the launches of one category when timing is enabled
struct Timed {
    ccb_handle *h;
    ccb_handle::Ev e{};
    bool on;
    Timed(ccb_handle *h_, int cat) : h(h_), on(h_ && h_->timing) {
        if (!on) return;
        auto get = [&]() {
            cudaEvent_t x;
            if (!h->ev_pool.empty()) {
                x = h->ev_pool.back();
                h->ev_pool.pop_back();
            } else {
                cudaEventCreate(&x);
            }
            return x;
        };
        e.a = get();
        e.b = get();
        e.cat = cat;
        cudaEventRecord(e.a, h->stream);
    }
    void stop() {
        if (!on) return;
        cudaEventRecord(e.b, h->stream);
        h->ev_pending.push_back(e);
        on = false;
    }
    ~Timed() { stop(); }
};
void drain_timing(ccb_handle *h) { // call after a stream synchronize
    for (auto &e : h->ev_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
            h->cat_ms[e.cat] += ms;
            h->cat_n[e.cat] += 1;
        }
        h->ev_pool.push_back(e.a);
        h->ev_pool.push_back(e.b);
    }
    h->ev_pending.clear();
}

int alloc_store(ccb_handle *h, Store &S, int cap) {
    const int D = h->D, DP = h->DP;
    S = Store{};
    S.cap = cap;
    const size_t nd = (size_t)cap * D + 2; // + slack for 16-byte rounded bulk copies
    CK(h, cudaMalloc(&S.cf1, nd * 8));
    CK(h, cudaMalloc(&S.cf2, nd * 8));
    CK(h, cudaMalloc(&S.cen, nd * 8));
    CK(h, cudaMalloc(&S.w, (size_t)cap * 8));
    CK(h, cudaMalloc(&S.mask, (size_t)cap * 8));
    CK(h, cudaMalloc(&S.id, (size_t)cap * 8));
    CK(h, cudaMalloc(&S.uid, (size_t)cap * 4));
    CK(h, cudaMalloc(&S.cw, (size_t)cap * DP * sizeof(double2)));
    return CCB_OK;
}
void free_store(Store &S) {
    cudaFree(S.cf1);
    cudaFree(S.cf2);
    cudaFree(S.cen);
    cudaFree(S.w);
    cudaFree(S.mask);
    cudaFree(S.id);
    cudaFree(S.uid);
    cudaFree(S.cw);
    S = Store{};
}
// grows both ping-pong buffers of a list to >= want entries, preserving the first n of the current one
int grow_store(ccb_handle *h, Store (&S)[2], int cur, int n, int64_t want) {
    if (want <= S[cur].cap) return CCB_OK;
    int64_t cap = S[cur].cap;
    while (cap < want) cap *= 2;
    if (cap > (int64_t)1 << 30) return fail(h, CCB_ELIMIT, "microcluster list would exceed 2^30 entries");
    const int D = h->D, DP = h->DP;
    Store nw;
    int rc = alloc_store(h, nw, (int)cap);
    if (rc) return rc;
    Store &o = S[cur];
    cudaStream_t s = h->stream;
    CK(h, cudaMemcpyAsync(nw.cf1, o.cf1, (size_t)n * D * 8, cudaMemcpyDeviceToDevice, s));
    CK(h, cudaMemcpyAsync(nw.cf2, o.cf2, (size_t)n * D * 8, cudaMemcpyDeviceToDevice, s));
    CK(h, cudaMemcpyAsync(nw.cen, o.cen, (size_t)n * D * 8, cudaMemcpyDeviceToDevice, s));
    CK(h, cudaMemcpyAsync(nw.w, o.w, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    CK(h, cudaMemcpyAsync(nw.mask, o.mask, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    CK(h, cudaMemcpyAsync(nw.id, o.id, (size_t)n * 8, cudaMemcpyDeviceToDevice, s));
    CK(h, cudaMemcpyAsync(nw.uid, o.uid, (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
    CK(h, cudaMemcpyAsync(nw.cw, o.cw, (size_t)n * DP * sizeof(double2), cudaMemcpyDeviceToDevice, s));
    CK(h, cudaStreamSynchronize(s));
    free_store(S[cur]);
    free_store(S[1 - cur]);
    S[cur] = nw;
    rc = alloc_store(h, S[1 - cur], (int)cap);
    return rc;
}

int realloc_aux_for_outlier_cap(ccb_handle *h) {
    const int cap = h->O[h->ocur].cap;
    cudaFree(h->d_onew);
    CK(h, cudaMalloc(&h->d_onew, (size_t)cap * 4));
    return CCB_OK;
}
int realloc_aux_for_pcore_cap(ccb_handle *h) {
    const int cap = h->P[h->pcur].cap;
    cudaFree(h->d_pnew);
    cudaFree(h->d_pfin);
    CK(h, cudaMalloc(&h->d_pnew, (size_t)cap * 4));
    CK(h, cudaMalloc(&h->d_pfin, (size_t)cap * 4));
    return CCB_OK;
}

int sync_ctl(ccb_handle *h) { // device control block -> pinned host mirror
    CK(h, cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->timing) drain_timing(h);
    return CCB_OK;
}
int push_ctl(ccb_handle *h) {
    CK(h, cudaMemcpyAsync(h->d_ctl, h->h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, h->stream));
    return CCB_OK;
}

Num make_num(const ccb_handle *h) {
    Num nm{};
    nm.delta2 = h->prm.delta2;
    nm.k = h->prm.k;
    nm.wsel = h->wsel;
    nm.eps2 = h->prm.eps2;
    nm.beta_mu = h->prm.beta * h->mu; // beta * mu, hddstream.py:413, 524
    nm.pi = h->pi;
    nm.D = h->D;
    nm.DP = h->DP;
    nm.div_mode = h->div_mode;
    nm.cnt_gt1 = h->cnt_gt1;
    // the gate compares count(pref' != 1) with pi; for k == 1 that count is 0 and the gate is vacuous,
    // as it is whenever pi >= D (hddstream.py:315-321)
    nm.pi_active = (h->prm.k != 1.0) && (h->pi < (int64_t)h->D);
    return nm;
}

// Static split of kernel 1: (row groups) x (slabs of the MC axis).  Few cells: enough slabs to fill the GPU twice.
// Many cells: the slab count that balances the waves -- gx CTAs in ceil(gx / resident) waves waste the tail of the last
// one (13 % for the dense 1e6 x 4096 benchmark at one slab); a few slabs make the waves finer.
template <int kDP, int K>
void nearest_static_split(int div_mode, int64_t nrows_max, int M, int max_slabs, int &gx, int &slab_mcs, int &nslab) {
    using Cfg = NearestCfg<kDP, K>;
    gx = (int)((nrows_max + Cfg::CELLS - 1) / Cfg::CELLS);
    int want = gx > 0 ? (2 * 148 + gx - 1) / gx : 1;
    if (want > max_slabs) want = max_slabs;
    const int tiles = (M + Cfg::TM - 1) / Cfg::TM;
    if (want > tiles) want = tiles;
    if (want < 1) want = 1;
    if (gx > 148) {
        int occ = 1;
        if (div_mode) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_nearest<kDP, K, true>, NEAREST_THREADS, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_nearest<kDP, K, false>, NEAREST_THREADS, 0);
        const double resident = 148.0 * std::max(occ, 1);
        auto loss = [&](int sl) {
            const double w = (double)gx * sl / resident;
            return w <= 1.0 ? 0.0 : std::ceil(w) / w - 1.0;
        };
        int best = want;
        for (int sl = want + 1; sl <= std::min(std::min(max_slabs, tiles), want + 7) && loss(best) > 0.03; ++sl)
            if (loss(sl) < loss(best)) best = sl;
        want = best;
    }
    slab_mcs = ((tiles + want - 1) / want) * Cfg::TM;
    nslab = (M + slab_mcs - 1) / slab_mcs;
}

// kernel 1 launch over an explicit cw array
template <int K>
int launch_nearest(ccb_handle *h, cudaStream_t s, int DP, int div_mode, const double *X, const int32_t *rows,
                   const int32_t *nrows_dev, int64_t row_off, int64_t nrows_max, int64_t ld, int D, const double2 *cw,
                   int M, double *slab_dist, int32_t *slab_idx, double *out_dist, int32_t *out_idx, int max_slabs,
                   int *nslab_out, const int32_t *range_dev = nullptr, const int32_t *M_dev = nullptr) {
    int launched = 0;
    CCB_DISPATCH_DP(DP, {
        int gx, slab_mcs, nslab;
        nearest_static_split<kDP, K>(div_mode, nrows_max, M, max_slabs, gx, slab_mcs, nslab);
        dim3 grid(gx, nslab);
        double *od = nslab == 1 ? out_dist : slab_dist;
        int32_t *oi = nslab == 1 ? out_idx : slab_idx;
        if (div_mode)
            k_nearest<kDP, K, true><<<grid, NEAREST_THREADS, 0, s>>>(X, rows, nrows_dev, row_off, nrows_max, ld, D, cw,
                                                                     M, slab_mcs, od, oi, range_dev, M_dev, 0, nullptr);
        else
            k_nearest<kDP, K, false><<<grid, NEAREST_THREADS, 0, s>>>(X, rows, nrows_dev, row_off, nrows_max, ld, D, cw,
                                                                      M, slab_mcs, od, oi, range_dev, M_dev, 0, nullptr);
        launched = 1;
        if (nslab > 1) {
            k_topk_merge<K><<<(unsigned)((nrows_max + 255) / 256), 256, 0, s>>>(slab_dist, slab_idx, nrows_dev, row_off,
                                                                                nrows_max, nslab, out_dist, out_idx, range_dev);
            launched = 2;
        }
        if (nslab_out) *nslab_out = nslab;
    })
    if (!launched) return fail(h, CCB_ELIMIT, "unsupported padded dimensionality %d", DP);
    if (h) h->st.kernel_launches += launched;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, CCB_ECUDA, "k_nearest launch: %s", cudaGetErrorString(e));
    return CCB_OK;
}

// kernel 1 with the device-side work split (block-speculative engine): rows[range_dev[0] .. range_dev[1]) of the
// row list against the first min(M_bound, *M_dev) entries of cw; slab lists at [list position][max_slabs][K]
constexpr int NEAREST_DYN_GRID = 148 * 4;
template <int K>
int launch_nearest_dyn(ccb_handle *h, cudaStream_t s, int DP, int div_mode, const double *X, const int32_t *rows, int64_t ld,
                       int D, const double2 *cw, int M_bound, double *slab_dist, int32_t *slab_idx, double *out_dist,
                       int32_t *out_idx, int max_slabs, int rows_max, const int32_t *range_dev, const int32_t *M_dev,
                       const XRef *xref) {
    int launched = 0;
    CCB_DISPATCH_DP(DP, {
        using Cfg = NearestCfg<kDP, K>;
        if (div_mode)
            k_nearest<kDP, K, true><<<NEAREST_DYN_GRID, NEAREST_THREADS, 0, s>>>(X, rows, nullptr, 0, 0, ld, D, cw, M_bound, 0,
                                                                                 slab_dist, slab_idx, range_dev, M_dev, max_slabs, xref);
        else
            k_nearest<kDP, K, false><<<NEAREST_DYN_GRID, NEAREST_THREADS, 0, s>>>(X, rows, nullptr, 0, 0, ld, D, cw, M_bound, 0,
                                                                                  slab_dist, slab_idx, range_dev, M_dev, max_slabs, xref);
        k_topk_merge_dyn<K><<<std::min(NEAREST_DYN_GRID, (rows_max + 3) / 4), 128, 0, s>>>(
            slab_dist, slab_idx, range_dev, M_dev, M_bound, NEAREST_DYN_GRID, max_slabs, Cfg::TM, Cfg::CELLS,
            Cfg::CAN_SPLIT ? Cfg::TMS : 0, out_dist, out_idx);
        launched = 2;
    })
    if (!launched) return fail(h, CCB_ELIMIT, "unsupported padded dimensionality %d", DP);
    if (h) h->st.kernel_launches += launched;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, CCB_ECUDA, "k_nearest launch: %s", cudaGetErrorString(e));
    return CCB_OK;
}

int ensure_point_buffers(ccb_handle *h, int64_t N) {
    if ((size_t)N > h->n_cap) {
        cudaFree(h->d_assign);
        cudaFree(h->d_stage);
        h->n_cap = 0;
        CK(h, cudaMalloc(&h->d_assign, (size_t)N * 4));
        CK(h, cudaMalloc(&h->d_stage, (size_t)N));
        h->n_cap = (size_t)N;
    }
    return CCB_OK;
}

// ---- block-speculative engine: workspace and block enqueue ------------------------------------------
constexpr int BS_MAX_SLABS = 256; // (k_topk_merge_dyn: a lane walks up to 8 slab lists)

template <typename T>
int ws_alloc(ccb_handle *h, T *&p, size_t n) {
    void *q = nullptr;
    CK(h, cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    h->ws_allocs.push_back(q);
    p = (T *)q;
    return CCB_OK;
}

int ensure_bs_ws(ccb_handle *h) {
    int rc;
    const int B = h->bs_bmax, D = h->D;
    if (!h->d_bc) {
        CCB_DISPATCH_DP(h->DP, {
            CK(h, cudaFuncSetAttribute(k_bs_chain_p<kDP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)ChainPCfg<kDP>::SMEM));
        })
        CK(h, cudaMalloc(&h->d_bc, sizeof(BsCtl)));
        CK(h, cudaMemset(h->d_bc, 0, sizeof(BsCtl)));
        CK(h, cudaMallocHost(&h->h_bc, sizeof(BsCtl)));
        memset(h->h_bc, 0, sizeof(BsCtl));
        CK(h, cudaMalloc(&h->d_io, sizeof(EngIo)));
        CK(h, cudaMallocHost(&h->h_io, sizeof(EngIo)));
        BsWs &w = h->ws;
        w.bmax = B;
#define WSA(field, n) if ((rc = ws_alloc(h, w.field, (n)))) return rc
        WSA(pcand, B); WSA(ospec, B); WSA(tkpos, B); WSA(dec, B); WSA(eff, B); WSA(newrank, B); WSA(pend, B); WSA(vpos, B); WSA(pbest, B);
        WSA(pflag, B); WSA(prej, B); WSA(upf, B);
        w.dp = h->DP;
        w.lsp = 2 * h->DP + 2;
        WSA(ver, (size_t)B * w.lsp); WSA(vcen, (size_t)B * D + 2);
        WSA(vr2, B); WSA(vmask, B);
        WSA(nrows, BS_RMAX); WSA(ncell, BS_RMAX);
        WSA(tk_dist, (size_t)BS_RMAX * BS_TOPK); WSA(tk_idx, (size_t)BS_RMAX * BS_TOPK);
        WSA(hkey, BS_RMAX + 1); WSA(hoff, BS_RMAX + 2); WSA(omem, BS_RMAX + 1); WSA(hrank, BS_RMAX + 2);
        WSA(hfirst, BS_RMAX + 1); WSA(okeys, BS_RMAX);
#undef WSA
        if ((rc = ws_alloc(h, h->d_bs_tk_dist_slab, (size_t)BS_RMAX * BS_MAX_SLABS * BS_TOPK))) return rc;
        if ((rc = ws_alloc(h, h->d_bs_tk_idx_slab, (size_t)BS_RMAX * BS_MAX_SLABS * BS_TOPK))) return rc;
    }
    // per-(tile, pcore key) tables follow the pcore capacity; the modified-flags follow the outlier capacity
    const int stride = h->P[h->pcur].cap;
    if (stride != h->ws.mp_stride) {
        for (void *&q : h->ws_tiles) {
            cudaFree(q);
            q = nullptr;
        }
        const size_t n = ((size_t)B / 32 + 2) * stride;
        const size_t npl = (size_t)B + 4 * (size_t)stride + 8; // every key's segment is padded to a multiple of 4
        const size_t ds = (size_t)h->ws.lsp;
        CK(h, cudaMalloc(&h->ws_tiles[0], n * 4));
        CK(h, cudaMalloc(&h->ws_tiles[1], n * 4));
        CK(h, cudaMalloc(&h->ws_tiles[2], ((size_t)stride + 2) * 4));
        CK(h, cudaMalloc(&h->ws_tiles[3], ((size_t)stride + 2) * 4));
        CK(h, cudaMalloc(&h->ws_tiles[4], npl * 4));
        CK(h, cudaMalloc(&h->ws_tiles[5], npl * ds * 8));
        CK(h, cudaMalloc(&h->ws_tiles[6], npl * ds * 8));
        h->ws.tilecnt = (int32_t *)h->ws_tiles[0];
        h->ws.tbase = (int32_t *)h->ws_tiles[1];
        h->ws.poff = (int32_t *)h->ws_tiles[2];
        h->ws.pcnt = (int32_t *)h->ws_tiles[3];
        h->ws.plist = (int32_t *)h->ws_tiles[4];
        h->ws.xg = (double *)h->ws_tiles[5];
        h->ws.verp = (double *)h->ws_tiles[6];
        h->ws.mp_stride = stride;
        if (h->ws.dbg) {
            cudaFree(h->ws.dbg);
            h->ws.dbg = nullptr;
            CK(h, cudaMalloc(&h->ws.dbg, (size_t)stride * 8 * sizeof(int64_t)));
            CK(h, cudaMemset(h->ws.dbg, 0, (size_t)stride * 8 * sizeof(int64_t)));
        }
    }
    const int ocap = h->O[h->ocur].cap;
    if (ocap != h->ws_first_cap) {
        cudaFree(h->ws_first);
        h->ws_first = nullptr;
        CK(h, cudaMalloc(&h->ws_first, (size_t)ocap * 4));
        CK(h, cudaMemsetAsync(h->ws_first, 0x7f, (size_t)ocap * 4, h->stream)); // 0x7f7f7f7f: "not modified"
        h->ws_first_cap = ocap;
        h->ws.firstmember = h->ws_first;
    }
    return CCB_OK;
}

int sync_bc(ccb_handle *h) { // both control blocks -> pinned host mirrors
    CK(h, cudaMemcpyAsync(h->h_ctl, h->d_ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaMemcpyAsync(h->h_bc, h->d_bc, sizeof(BsCtl), cudaMemcpyDeviceToHost, h->stream));
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->timing) drain_timing(h);
    return CCB_OK;
}

// One block = prologue + refinement rounds + commit.  Every kernel reads its work description from the device-side
// control block, so a converged (or finished) block turns the remaining launches into no-ops.  The three pieces are
// launched either on the handle's stream (bs_iters rounds enqueued per block; used when per-kernel timing is on) or
// captured once into a CUDA graph whose block loop and round loop are WHILE conditional nodes driven from the
// device (k_bs_begin / k_bs_decide / k_bs_commit call cudaGraphSetConditional): no idle launches, no host round trip.
// Kernel launch with (pdl) or without the programmatic-stream-serialization attribute (see CCB_PDL in common.cuh)
template <typename... KA, typename... A>
cudaError_t launch_k(bool pdl, void (*k)(KA...), dim3 g, dim3 b, size_t smem, cudaStream_t s, A &&...a) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = g;
    cfg.blockDim = b;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k, KA(std::forward<A>(a))...);
}

int launch_prologue(ccb_handle *h, const Eng &e, cudaStream_t s) {
    const int g_cells = (h->bs_bmax + BS_THREADS - 1) / BS_THREADS;
    Timed tm(h, CCB_CAT_SPEC);
    const bool pdl = h->pdl;
    k_bs_begin<<<1, 1, 0, s>>>(e);
    CCB_DISPATCH_DP(h->DP, { launch_k(pdl, k_bs_spec<kDP>, g_cells * BS_SPLIT, BS_THREADS, 0, s, e); })
    CKL(h);
    return CCB_OK;
}

// side != nullptr (graph capture): kernel 1 + the speculated outlier decisions run on a side branch next to the
// candidate lists + pcore replay (they touch disjoint data; see the dependency notes in DESIGN.md) and join before
// k_bs_olist.
int launch_round(ccb_handle *h, const Eng &e, cudaStream_t s, int mp_grid, int mo_bound, cudaStream_t side = nullptr) {
    const int B = h->bs_bmax;
    const int g_cells = (B + BS_THREADS - 1) / BS_THREADS;
    const int g_tiles = (B / 32 + 1 + 3) / 4;
    int rc;
    const bool pdl = h->pdl;
    cudaStream_t sa = side ? side : s;
    if (side) {
        CK(h, cudaEventRecord(h->ev_fork, s));
        CK(h, cudaStreamWaitEvent(side, h->ev_fork, 0));
    }
    {
        Timed tm(h, CCB_CAT_NEAREST);
        if ((rc = launch_nearest_dyn<BS_TOPK>(nullptr, sa, h->DP, h->div_mode, e.X, e.ws.nrows, e.ld, h->D, e.O.cw, mo_bound,
                                              h->d_bs_tk_dist_slab, h->d_bs_tk_idx_slab, e.ws.tk_dist, e.ws.tk_idx,
                                              BS_MAX_SLABS, BS_RMAX, &e.bc->tk_lo, &e.bc->Mo0,
                                              reinterpret_cast<const XRef *>(e.io))))
            return fail(h, rc, "%s", ccb_last_error(nullptr));
    }
    {
        Timed tm(h, CCB_CAT_SPEC);
        launch_k(pdl, k_bs_spec_o, BS_RMAX / BS_THREADS, BS_THREADS, 0, sa, e);
    }
    {
        Timed tm(h, CCB_CAT_LISTS);
        launch_k(pdl, k_bs_tilecnt, g_tiles, BS_THREADS, 0, s, e);
        launch_k(pdl, k_bs_pscan, std::max(mp_grid, 1), BS_CTA1, 0, s, e);
        launch_k(pdl, k_bs_pscatter, B / 32 + 1, BS_THREADS, 0, s, e);
    }
    {
        Timed tm(h, CCB_CAT_PCORE);
        CCB_DISPATCH_DP(h->DP, { launch_k(pdl, k_bs_chain_p<kDP>, std::max(mp_grid, 1), BS_CHAINP_THREADS, ChainPCfg<kDP>::SMEM, s, e); })
    }
    if (side) { // join: k_bs_olist needs the speculated outlier decisions, everything below the pcore replay
        CK(h, cudaEventRecord(h->ev_join, side));
        CK(h, cudaStreamWaitEvent(s, h->ev_join, 0));
        // second fork: pcore side of derive + verify next to the outlier-side lists / replay / derive (disjoint cells)
        CK(h, cudaEventRecord(h->ev_fork, s));
        CK(h, cudaStreamWaitEvent(side, h->ev_fork, 0));
    }
    {
        Timed tm(h, CCB_CAT_DERIVE);
        launch_k(pdl, k_bs_derive_p, g_cells, BS_THREADS, 0, sa, e);
    }
    {
        Timed tm(h, CCB_CAT_RESOLVE);
        CCB_DISPATCH_DP(h->DP, { launch_k(pdl, k_bs_verify_p<kDP>, B / 32 + 1, BS_VP_THREADS, 0, sa, e); })
    }
    {
        Timed tm(h, CCB_CAT_OLIST);
        launch_k(pdl, k_bs_olist, BS_OL_CTAS, BS_OL_THREADS, 0, s, e);
    }
    {
        Timed tm(h, CCB_CAT_CHAIN_O);
        CCB_DISPATCH_DP(h->DP, { launch_k(pdl, k_bs_chain_o<kDP>, 148 * 4, BS_THREADS, 0, s, e); }) // CTAs loop over the keys
    }
    {
        Timed tm(h, CCB_CAT_DERIVE);
        launch_k(pdl, k_bs_derive_o, BS_RMAX / BS_THREADS, BS_THREADS, 0, s, e);
    }
    if (side) {
        CK(h, cudaEventRecord(h->ev_join, side));
        CK(h, cudaStreamWaitEvent(s, h->ev_join, 0));
    }
    {
        Timed tm(h, CCB_CAT_RESOLVE);
        CCB_DISPATCH_DP(h->DP, { launch_k(pdl, k_bs_verify_o<kDP>, 148 * 4, BS_THREADS, 0, s, e); })
    }
    {
        Timed tm(h, CCB_CAT_DECIDE);
        launch_k(pdl, k_bs_decide, 1, BS_CTA1, 0, s, e);
    }
    CKL(h);
    return CCB_OK;
}
constexpr int BS_LAUNCHES_PROLOGUE = 2, BS_LAUNCHES_ROUND = 14, BS_LAUNCHES_COMMIT = 1;

int launch_commit(ccb_handle *h, const Eng &e, cudaStream_t s, int mp_grid, cudaStream_t side = nullptr) {
    (void)side;
    const int g_cells = (h->bs_bmax + BS_THREADS - 1) / BS_THREADS;
    const int rows_ctas = (mp_grid + BS_RMAX + 3) / 4;
    Timed tm(h, CCB_CAT_COMMIT);
    launch_k(h->pdl, k_bs_commit, rows_ctas + g_cells, BS_THREADS, 0, s, e, rows_ctas); // rows | cells | (last CTA) finish
    CKL(h);
    return CCB_OK;
}

Eng make_eng(ccb_handle *h) {
    Eng e{};
    e.P = h->P[h->pcur];
    e.O = h->O[h->ocur];
    e.ctl = h->d_ctl;
    e.bc = h->d_bc;
    e.ws = h->ws;
    return e;
}

// ---- CUDA graph of the whole ordered loop: WHILE(block) { prologue; WHILE(round) { round }; commit } ----------
void drop_graph(ccb_handle *h) { // (the current slot)
    if (h->bs_exec) cudaGraphExecDestroy(h->bs_exec);
    if (h->bs_graph) cudaGraphDestroy(h->bs_graph);
    h->bs_exec = nullptr;
    h->bs_graph = nullptr;
}

int build_graph(ccb_handle *h, const Eng &base) {
    drop_graph(h);
    if (!h->cap1) {
        // the main line of a round (lists, replay, outlier-side lists ...) is the critical path: its kernels are captured
        // from high-priority streams, the side branches (kernel 1, pcore-side derive / verify) from a low-priority one, so
        // that a small critical kernel is not queued behind the thousand CTAs of a side kernel running next to it
        int prio_lo = 0, prio_hi = 0;
        CK(h, cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(h, cudaStreamCreateWithPriority(&h->cap1, cudaStreamNonBlocking, prio_hi));
        CK(h, cudaStreamCreateWithPriority(&h->cap2, cudaStreamNonBlocking, prio_hi));
        CK(h, cudaStreamCreateWithPriority(&h->cap3, cudaStreamNonBlocking, prio_lo));
        CK(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        CK(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
    }
    Eng e = base;
    e.io = h->d_io;
    CK(h, cudaGraphCreate(&h->bs_graph, 0));
    cudaGraphConditionalHandle ho, hi;
    CK(h, cudaGraphConditionalHandleCreate(&ho, h->bs_graph, 1, cudaGraphCondAssignDefault));
    CK(h, cudaGraphConditionalHandleCreate(&hi, h->bs_graph, 0, cudaGraphCondAssignDefault));
    e.h_outer = ho;
    e.h_inner = hi;
    cudaGraphNodeParams po = {};
    po.type = cudaGraphNodeTypeConditional;
    po.conditional.handle = ho;
    po.conditional.type = cudaGraphCondTypeWhile;
    po.conditional.size = 1;
    cudaGraphNode_t outer;
    CK(h, cudaGraphAddNode(&outer, h->bs_graph, nullptr, 0, &po));
    cudaGraph_t body_o = po.conditional.phGraph_out[0];
    const int mp_grid = e.P.cap, mo_bound = e.O.cap;
    int rc;
    CK(h, cudaStreamBeginCaptureToGraph(h->cap1, body_o, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    if ((rc = launch_prologue(h, e, h->cap1))) return rc;
    {
        cudaStreamCaptureStatus st;
        cudaGraph_t cg = nullptr;
        const cudaGraphNode_t *deps = nullptr;
        size_t ndeps = 0;
        CK(h, cudaStreamGetCaptureInfo(h->cap1, &st, nullptr, &cg, &deps, &ndeps));
        cudaGraphNodeParams pi = {};
        pi.type = cudaGraphNodeTypeConditional;
        pi.conditional.handle = hi;
        pi.conditional.type = cudaGraphCondTypeWhile;
        pi.conditional.size = 1;
        cudaGraphNode_t inner;
        CK(h, cudaGraphAddNode(&inner, cg, deps, ndeps, &pi));
        CK(h, cudaStreamUpdateCaptureDependencies(h->cap1, &inner, 1, cudaStreamSetCaptureDependencies));
        cudaGraph_t body_i = pi.conditional.phGraph_out[0];
        CK(h, cudaStreamBeginCaptureToGraph(h->cap2, body_i, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
        if ((rc = launch_round(h, e, h->cap2, mp_grid, mo_bound, h->cap3))) return rc;
        CK(h, cudaStreamEndCapture(h->cap2, nullptr));
    }
    if ((rc = launch_commit(h, e, h->cap1, mp_grid, h->cap3))) return rc;
    CK(h, cudaStreamEndCapture(h->cap1, nullptr));
    CK(h, cudaGraphInstantiate(&h->bs_exec, h->bs_graph, 0));
    h->bs_graph_eng = base;
    return CCB_OK;
}

int ensure_bsv_capacity(ccb_handle *h, int blocks_ahead) {
    int rc;
    const Ctl &c = *h->h_ctl;
    const int64_t need_o = (int64_t)c.n_outlier + (int64_t)(blocks_ahead + 1) * BS_RMAX + 1;
    if (need_o > h->O[h->ocur].cap) {
        if ((rc = grow_store(h, h->O, h->ocur, c.n_outlier, need_o))) return rc;
        if ((rc = realloc_aux_for_outlier_cap(h))) return rc;
    }
    if (c.n_pcore + blocks_ahead + 1 > h->P[h->pcur].cap) {
        if ((rc = grow_store(h, h->P, h->pcur, c.n_pcore, c.n_pcore + blocks_ahead + 1))) return rc;
        if ((rc = realloc_aux_for_pcore_cap(h))) return rc;
    }
    return ensure_bs_ws(h);
}

// the ordered loop over cells [0, N) of a device-resident X, block-speculative engine
int ingest_core_bsv(ccb_handle *h, const double *dX, int64_t N, int64_t ld, int32_t *d_assign, uint8_t *d_stage,
                    const std::function<int()> *after_first_launch = nullptr) {
    if (!h->have_params) return fail(h, CCB_ESTATE, "ccb_begin_timepoint must precede ccb_ingest");
    if (N >= ((int64_t)1 << 31)) return fail(h, CCB_ELIMIT, "more than 2^31 - 1 cells in one call");
    cudaStream_t s = h->stream;
    int rc;
    if ((rc = ensure_bs_ws(h))) return rc;
    const bool graph = h->bs_use_graph && !h->timing;
    const int bmax = h->bs_bmax, bmin = std::min(h->bs_bmin, h->bs_bmax);
    const int itmax = graph ? (h->prm.bsv_iters > 0 ? h->bs_iters : 8) : h->bs_iters;
    k_bs_init<<<1, 1, 0, s>>>(h->d_bc, N, itmax, bmin, bmax);
    CKL(h);
    h->st.kernel_launches++;
    if ((rc = sync_bc(h))) return rc;
    EngIo io{};
    io.X = dX;
    io.ld = ld;
    io.assign = d_assign;
    io.stage = d_stage;
    // SAFE / CONTESTED split of the speculation -- a heuristic that never affects results, only how much of the replay
    // takes the exact in-chain radius test and how often verification sends a block into another round.  SAFE = the
    // snapshot says "absorbed" (radius^2 of the tentative MC <= eps^2) and the cell is not far out (snapshot distance
    // <= 4 eps^2).  Established MCs sit right at their radius limit, so a margin below eps^2 marks almost every cell
    // CONTESTED (profiles/r1p_safe_rule_experiment.md: 10x slower); the distance bound singles out the ~1 % background
    // cells, which are exactly the ones whose radius test is open.  Speculative rejection of far cells did not pay
    // either (profiles/r1t_speculative_reject_experiment.md), nor did calling every cell SAFE whose tentative MC keeps a
    // margin below eps^2 whatever its distance (profiles/r2h_slack_rule_experiment.md): every CONTESTED cell takes the
    // exact in-chain test.  The distance bound is flat between 2 and 8 eps^2 (profiles/r2l_theta_experiment.md).
    io.theta = 4.0 * h->prm.eps2;
    io.r2safe = h->prm.eps2;
    io.r2rej = HUGE_VAL;
    io.nm = make_num(h);
    if (graph) {
        *h->h_io = io;
        CK(h, cudaMemcpyAsync(h->d_io, h->h_io, sizeof(EngIo), cudaMemcpyHostToDevice, s));
    }
    int guard = 0;
    while (!h->h_bc->done) {
        const BsCtl &b = *h->h_bc;
        const int64_t left = N - b.pos;
        // blocks to enqueue before the next look at the control block (stream launches); the graph runs until the
        // input is consumed or a store is full, so give it room for a good stretch of blocks
        int G = (int)std::min<int64_t>(16, (left + std::max(b.next_B, 1) - 1) / std::max(b.next_B, 1) + 1);
        if (G < 1) G = 1;
        if ((rc = ensure_bsv_capacity(h, graph ? 32 : G))) return rc;
        if (b.need_grow) {
            h->h_bc->need_grow = 0;
            CK(h, cudaMemcpyAsync(&h->d_bc->need_grow, &h->h_bc->need_grow, sizeof(int32_t), cudaMemcpyHostToDevice, s));
        }
        Eng e = make_eng(h);
        const int64_t pos0 = b.pos;
        const int64_t blocks0 = b.blocks, iters0 = b.iters;
        if (graph) {
            if (!h->bs_exec || memcmp(&h->bs_graph_eng, &e, sizeof(Eng)) != 0) {
                // not the current slot: the other one? (swap); otherwise rebuild the older slot
                std::swap(h->bs_graph, h->bs_graph_alt);
                std::swap(h->bs_exec, h->bs_exec_alt);
                std::swap(h->bs_graph_eng, h->bs_graph_eng_alt);
                if (!h->bs_exec || memcmp(&h->bs_graph_eng, &e, sizeof(Eng)) != 0) {
                    rc = build_graph(h, e);
                    if (rc && h->pdl) { // a driver that cannot take programmatic edges here: the same graph without them
                        cudaGraph_t junk = nullptr;
                        cudaStreamEndCapture(h->cap2, &junk);
                        cudaStreamEndCapture(h->cap1, &junk);
                        cudaGetLastError();
                        drop_graph(h);
                        h->pdl = false;
                        rc = build_graph(h, e);
                    }
                    if (rc) return rc;
                }
            }
            CK(h, cudaGraphLaunch(h->bs_exec, s));
        } else {
            const Ctl &c = *h->h_ctl;
            e.X = io.X, e.ld = io.ld, e.assign = io.assign, e.stage = io.stage, e.theta = io.theta, e.r2safe = io.r2safe, e.r2rej = io.r2rej, e.nm = io.nm;
            for (int g = 0; g < G; ++g) {
                const int mp_grid = c.n_pcore + g + 1;
                const int mo_bound = (int)std::min<int64_t>(c.n_outlier + (int64_t)(g + 1) * BS_RMAX, h->O[h->ocur].cap);
                if ((rc = launch_prologue(h, e, s))) return rc;
                for (int it = 0; it < h->bs_iters; ++it)
                    if ((rc = launch_round(h, e, s, mp_grid, mo_bound))) return rc;
                if ((rc = launch_commit(h, e, s, mp_grid))) return rc;
                h->st.kernel_launches += BS_LAUNCHES_PROLOGUE + h->bs_iters * BS_LAUNCHES_ROUND + BS_LAUNCHES_COMMIT;
            }
        }
        if (after_first_launch) { // the engine is running: the caller queues the next input segment behind it
            const std::function<int()> *f = after_first_launch;
            after_first_launch = nullptr;
            if ((rc = (*f)())) return rc;
        }
        if ((rc = sync_bc(h))) return rc;
        if (graph) // kernels the graph executed: per block prologue + commit, per round the round's kernels
            h->st.kernel_launches += (h->h_bc->blocks - blocks0) * (BS_LAUNCHES_PROLOGUE + BS_LAUNCHES_COMMIT) +
                                     (h->h_bc->iters - iters0) * BS_LAUNCHES_ROUND + BS_LAUNCHES_PROLOGUE;
        if (h->h_bc->pos == pos0 && !h->h_bc->need_grow && !h->h_bc->done && ++guard > 4)
            return fail(h, CCB_ESTATE, "internal: block-speculative engine made no progress at row %lld", (long long)pos0);
        if (h->h_bc->pos != pos0) guard = 0;
    }
    if (after_first_launch && (rc = (*after_first_launch)())) return rc;
    h->st.points += N;
    return CCB_OK;
}

double builtin_nrm2(const double *x, int n) {
    // restatement of OpenBLAS' x86-64 dnrm2: squares, sum and square root in x87 extended precision
    long double s = 0.0L;
    for (int i = 0; i < n; ++i) s += (long double)x[i] * (long double)x[i];
    return (double)sqrtl(s);
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    ~DevBuf() { cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

// grow-only buffer of the handle's offline workspace (contents are NOT preserved across a growth)
int off_buf(ccb_handle *h, int which, size_t bytes) {
    ccb_handle::OffBuf &b = h->ob[which];
    bytes = std::max<size_t>(bytes, 16);
    if (bytes <= b.bytes) return CCB_OK;
    CK(h, cudaStreamSynchronize(h->stream));
    cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
    size_t cap = 256;
    while (cap < bytes) cap *= 2;
    CK(h, cudaMalloc(&b.p, cap));
    b.bytes = cap;
    return CCB_OK;
}

int launch_off_neighbours(ccb_handle *h, cudaStream_t s, const double *cen, int M, int D, int r0, int r1, double E2,
                          uint32_t *nbr, int32_t *cnt, int32_t *border, int border_cap, int32_t *n_border) {
    const int DP = round_dp(D);
    int ok = 0;
    if (r1 > r0 && cudaMemsetAsync(cnt, 0, (size_t)(r1 - r0) * sizeof(int32_t), s) != cudaSuccess)
        return fail(h, CCB_ECUDA, "k_off_neighbours: clearing the row counts failed");
    CCB_DISPATCH_DP(DP, {
        // rows x column slabs: at least ~4 waves of CTAs on 148 SMs so that the tail does not dominate
        const int gx = (r1 - r0 + OffCfg<kDP>::ROWS - 1) / OffCfg<kDP>::ROWS;
        const int ntiles = (M + OffCfg<kDP>::TM - 1) / OffCfg<kDP>::TM;
        int gy = gx > 0 ? std::max(1, std::min(ntiles, (148 * 2 * 4 + gx - 1) / gx)) : 1;
        if (gx > 0) { // ... and a slab count whose last wave is full (391 x 4 CTAs on 296 slots were 5.28 waves)
            int occ = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_off_neighbours<kDP>, OFFN_THREADS, 0);
            const double resident = 148.0 * std::max(occ, 1);
            auto loss = [&](int sl) {
                const double w = (double)gx * sl / resident;
                return w <= 1.0 ? 0.0 : std::ceil(w) / w - 1.0;
            };
            int best = gy;
            for (int sl = gy + 1; sl <= std::min(ntiles, gy + 7) && loss(best) > 0.03; ++sl)
                if (loss(sl) < loss(best)) best = sl;
            gy = best;
        }
        if (gx > 0)
            k_off_neighbours<kDP><<<dim3(gx, gy), OFFN_THREADS, 0, s>>>(cen, M, D, r0, r1, E2, nbr, cnt, border, border_cap,
                                                                          n_border);
        ok = 1;
    })
    if (!ok) return fail(h, CCB_ELIMIT, "unsupported dimensionality %d", D);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, CCB_ECUDA, "k_off_neighbours launch: %s", cudaGetErrorString(e));
    return CCB_OK;
}

// The stateless entry points take their scratch from the device's stream-ordered pool; raise its release threshold once
// so that it keeps what it has instead of handing the memory back to the driver at every synchronisation.
void tune_pool(int device) {
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t keep = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done[device] = true;
}

// kernel 4e: ordered cluster growth.  Small M (the online hot path: tens of pcore MCs): the single-launch bit-row
// kernel.  Large M (config C4): isolated MCs in parallel, CSR lists for the rest, clusters merged by seed rank
// (offline.cuh).  Both produce identical label / order / cl_off / n_cl.  cls [M] and queue [2M + 2] are scratch.
constexpr int OFF_CSR_MIN_M = 2048; // default switch-over; ccb_params.off_csr_min_m / the csr_min_m argument move it

// The ordered growth over a CSR of the weighted neighbourhoods and the merge with the isolated microclusters' clusters by
// seed rank (offline.cuh).  i32: scratch of 7 M + 16 int32; cls [M] zeroed; queue [2 M + 2].
void offc_grow_and_merge(cudaStream_t s, int M, const int64_t *off, const int32_t *col, const uint8_t *core, const uint8_t *iso,
                         const uint64_t *submask, int cnt_gt1, int64_t pi, uint8_t *cls, int32_t *queue, int32_t *i32,
                         int32_t *label, int32_t *order, int32_t *cl_off, int32_t *n_cl, int64_t total_nnz) {
    const size_t m = (size_t)M;
    int32_t *order_s = i32, *cl_off_s = order_s + m, *seed_of = cl_off_s + m + 1, *n_cl_s = seed_of + m + 1,
            *seedflag = n_cl_s + 1, *size_by_node = seedflag + m, *rank = size_by_node + m, *size_by_cluster = rank + m + 1;
    const int tgrid = (M + 255) / 256;
    // short lists (a sparse graph): one warp over the compacted non-isolated microclusters, no block barriers; long lists:
    // the 1024-thread walk.  total_nnz < 0 = unknown.  (cand / its flags / ranks live in buffers the later kernels rewrite.)
    if (total_nnz >= 0 && total_nnz <= (int64_t)M * 16) {
        int32_t *candflag = seedflag, *candrank = rank, *cand = size_by_node;
        k_offc_candflag<<<tgrid, 256, 0, s>>>(M, iso, candflag);
        k_offc_scan<int32_t><<<1, OFFG_THREADS, 0, s>>>(candflag, M, candrank);
        k_offc_candscatter<<<tgrid, 256, 0, s>>>(M, candflag, candrank, cand);
        // (a non-isolated microcluster owns >= 1 CSR entry, so total_nnz bounds the number of candidates, too)
        const size_t smem = (size_t)total_nnz * 22 + 64;
        if (smem <= 200 * 1024) { // the whole non-isolated graph fits one CTA's shared memory: walk it there
            cudaFuncSetAttribute(k_offc_grow_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            k_offc_grow_smem<<<1, 1024, smem, s>>>(cand, candrank, M, off, col, core, submask, cnt_gt1, pi, order_s, cl_off_s,
                                                   seed_of, n_cl_s, (int)total_nnz + 1, (int)total_nnz + 1);
        } else {
            k_offc_grow_warp<<<1, 32, 0, s>>>(cand, candrank + M, off, col, core, submask, cnt_gt1, pi, cls, queue, order_s,
                                              cl_off_s, seed_of, n_cl_s);
        }
    } else {
        k_offc_grow<<<1, OFFG_THREADS, 0, s>>>(M, off, col, core, iso, submask, cnt_gt1, pi, cls, queue, order_s, cl_off_s,
                                               seed_of, n_cl_s);
    }
    k_offc_seeds<<<tgrid, 256, 0, s>>>(M, iso, core, submask, cnt_gt1, pi, seed_of, cl_off_s, n_cl_s, seedflag, size_by_node,
                                       label);
    k_offc_seeds_serial<<<tgrid, 256, 0, s>>>(seed_of, cl_off_s, n_cl_s, seedflag, size_by_node);
    k_offc_scan<int32_t><<<1, OFFG_THREADS, 0, s>>>(seedflag, M, rank);
    cudaMemsetAsync(size_by_cluster, 0, m * 4, s);
    k_offc_sizes<<<tgrid, 256, 0, s>>>(M, seedflag, rank, size_by_node, size_by_cluster, n_cl);
    k_offc_scan<int32_t><<<1, OFFG_THREADS, 0, s>>>(size_by_cluster, M, cl_off);
    k_offc_scatter_iso<<<tgrid, 256, 0, s>>>(M, iso, seedflag, rank, size_by_node, cl_off, order, label);
    k_offc_scatter_serial<<<148, 256, 0, s>>>(seed_of, cl_off_s, n_cl_s, order_s, rank, cl_off, order, label);
}

int launch_off_clusters(ccb_handle *h, cudaStream_t s, int M, const uint32_t *wnbr, const uint8_t *core,
                        const uint64_t *submask, int cnt_gt1, int64_t pi, uint8_t *cls, int32_t *queue, int32_t *label,
                        int32_t *order, int32_t *cl_off, int32_t *n_cl, int *launches, int csr_min_m) {
    const int words = (M + 31) / 32;
    *launches = 0;
    cudaMemsetAsync(cls, 0, (size_t)std::max(M, 1), s);
    auto old_path = [&]() {
        k_off_clusters<<<1, OFFC_THREADS, 0, s>>>(M, wnbr, core, submask, cnt_gt1, pi, cls, queue, label, order, cl_off, n_cl);
        *launches += 1;
        cudaError_t e = cudaGetLastError();
        return e == cudaSuccess ? CCB_OK : fail(h, CCB_ECUDA, "k_off_clusters: %s", cudaGetErrorString(e));
    };
    if (M < (csr_min_m > 0 ? csr_min_m : OFF_CSR_MIN_M) || M == 0) return old_path();
    uint8_t *iso = nullptr;
    int32_t *i32 = nullptr, *col = nullptr;
    int64_t *off = nullptr;
    // int32 scratch: nnz | order_s | cl_off_s (M + 1) | seed_of (M + 1) | n_cl_s (1) | seedflag | size_by_node |
    //                rank (M + 1) | size_by_cluster
    const size_t m = (size_t)M;
    const size_t n32 = 8 * m + 16;
    cudaError_t e;
    if ((e = cudaMallocAsync(&iso, m, s)) != cudaSuccess || (e = cudaMallocAsync(&i32, n32 * 4, s)) != cudaSuccess ||
        (e = cudaMallocAsync(&off, (m + 1) * 8, s)) != cudaSuccess)
        return fail(h, CCB_ENOMEM, "offline scratch: %s", cudaGetErrorString(e));
    int32_t *nnz = i32; // the remaining 7 m + 16 entries are the scratch of offc_grow_and_merge
    auto release = [&]() {
        cudaFreeAsync(iso, s);
        cudaFreeAsync(i32, s);
        cudaFreeAsync(off, s);
        if (col) cudaFreeAsync(col, s);
    };
    const int wgrid = (M + 3) / 4;
    k_offc_rowinfo<<<wgrid, 128, 0, s>>>(wnbr, 0, M, words, iso, nnz);
    k_offc_scan<int64_t><<<1, OFFG_THREADS, 0, s>>>(nnz, M, off);
    *launches += 2;
    int64_t total = 0;
    if ((e = cudaMemcpyAsync(&total, off + M, 8, cudaMemcpyDeviceToHost, s)) != cudaSuccess ||
        (e = cudaStreamSynchronize(s)) != cudaSuccess) {
        release();
        return fail(h, CCB_ECUDA, "offline CSR sizes: %s", cudaGetErrorString(e));
    }
    if (total > (int64_t)INT32_MAX - 1) { // a CSR that large buys nothing over the bit rows
        release();
        return old_path();
    }
    if ((e = cudaMallocAsync(&col, (size_t)std::max<int64_t>(total, 1) * 4, s)) != cudaSuccess) {
        release();
        return fail(h, CCB_ENOMEM, "offline CSR (%lld entries): %s", (long long)total, cudaGetErrorString(e));
    }
    k_offc_fill<<<wgrid, 128, 0, s>>>(wnbr, 0, M, words, iso, off, col);
    *launches += 1;
    offc_grow_and_merge(s, M, off, col, core, iso, submask, cnt_gt1, pi, cls, queue, i32 + m, label, order, cl_off, n_cl, total);
    *launches += 8;
    e = cudaGetLastError();
    release();
    return e == cudaSuccess ? CCB_OK : fail(h, CCB_ECUDA, "offline cluster growth: %s", cudaGetErrorString(e));
}

// Probes of the pcore replay pattern (mode 3 of ccb_fp64_peak): one warp runs the CLEAN-stage schedule of k_bs_chain_p over
// a ring of records in shared memory -- per cell one dependent DADD, the store of the previous version, the load of an
// addend one batch ahead -- under different conditions; sink[8 + v] = cycles per cell of variant v:
//   0 DADD chain only (addends in registers)      1 + addends from LDS      2 + STS of every version
//   3 as 2, while two other warps spin on mbarrier try_wait (the producer / store threads of the kernel)
//   4 as 2 with 64-bit accesses replaced by lane-strided 8-byte accesses of a 26-element record (the real layout)
template <int VAR>
__device__ __forceinline__ double chain_probe_run(double *sm, int lane, long long &cyc) {
    constexpr int LSP = 26, NB = 64, GS = 8, NG = NB / GS, REP = 32;
    double v = 1.0 + lane * 1e-9;
    uint32_t xa = smem_u32(sm) + (lane < LSP ? lane : 0) * 8;
    asm volatile("" : "+r"(xa));
    const bool st_ok = lane < LSP;
    const long long t0 = clock64();
    for (int rep = 0; rep < REP; ++rep) {
        double R[2][GS];
#pragma unroll
        for (int q = 0; q < GS; ++q) R[0][q] = VAR >= 1 ? lds_f64(xa + (q * LSP) * 8) : 1e-12 * q;
#pragma unroll
        for (int k = 0; k < NG; ++k) {
#pragma unroll
            for (int q = 0; q < GS; ++q) {
                const double nv = ccb::dadd(v, R[k & 1][q]);
                if (VAR >= 2 && k + q > 0 && st_ok) sts_f64(xa + ((k * GS + q - 1) * LSP) * 8, v);
                if (k + 1 < NG) R[(k + 1) & 1][q] = VAR >= 1 ? lds_f64(xa + (((k + 1) * GS + q) * LSP) * 8) : 1e-12 * q;
                v = nv;
            }
        }
        if (VAR >= 2 && st_ok) sts_f64(xa + ((NB - 1) * LSP) * 8, v);
    }
    cyc = clock64() - t0;
    return v;
}
__global__ void k_chain_probe(double *sink) {
    __shared__ __align__(16) double sm[64 * 26 + 32];
    __shared__ uint64_t bar;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 64 * 26 + 32; i += blockDim.x) sm[i] = 1e-9 * (i % 97);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
        stop = 0;
    }
    __syncthreads();
    if (warp == 0) {
        long long c0, c1, c2, c3;
        double a = chain_probe_run<0>(sm, lane, c0);
        a += chain_probe_run<1>(sm, lane, c1);
        a += chain_probe_run<2>(sm, lane, c2);
        if (lane == 0) stop = 1; // phase 1 over: spinners start
        __syncwarp();
        for (volatile int w = 0; w < 2000; ++w) {}
        a += chain_probe_run<2>(sm, lane, c3);
        if (lane == 0) {
            const double cells = 32.0 * 64.0;
            sink[8] = c0 / cells;
            sink[9] = c1 / cells;
            sink[10] = c2 / cells;
            sink[11] = c3 / cells;
            sink[7] = a;
            stop = 2;
            mbar_arrive(&bar);
        }
        return;
    }
    // warps 1, 2: idle until phase 1 is over, then lane 0 spins on the mbarrier like the producer / store thread do
    if (lane == 0) {
        while (stop == 0) {}
        mbar_wait(&bar, 0);
    }
}

} // namespace

// =====================================================================================================
extern "C" {

const char *ccb_last_error(const ccb_handle *h) { return h ? h->err.c_str() : g_err.c_str(); }
void *ccb_stream(ccb_handle *h) { return h ? (void *)h->stream : nullptr; }

int ccb_create(const ccb_params *p, ccb_handle **out) {
    if (!p || !out) return fail(nullptr, CCB_EINVAL, "null argument");
    *out = nullptr;
    if (p->D < 1 || p->D > CCB_MAX_D) return fail(nullptr, CCB_ELIMIT, "D=%d outside 1..%d", p->D, CCB_MAX_D);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, CCB_ECUDA, "no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (p->device < 0 || p->device >= ndev) return fail(nullptr, CCB_EINVAL, "device %d of %d", p->device, ndev);
    ccb_handle *h = new (std::nothrow) ccb_handle();
    if (!h) return fail(nullptr, CCB_ENOMEM, "host allocation failed");
    h->prm = *p;
    h->D = p->D;
    h->DP = round_dp(p->D);
    if (p->chunk > 0) h->bs_bmax = std::max(32, (p->chunk + 31) / 32 * 32); // block length cap of the BSV engine
    if (p->bsv_bmin > 0) h->bs_bmin = p->bsv_bmin;
    if (p->bsv_iters > 0) h->bs_iters = std::min(p->bsv_iters, 16);
    h->bs_use_graph = p->bsv_stream == 0;
    h->off_csr_min_m = p->off_csr_min_m;
    h->bs_bmax = std::min(h->bs_bmax, 1 << 20);
    const bool p2 = is_pow2(p->k);
    h->div_mode = p2 ? 0 : 1;
    h->wsel = p2 ? 1.0 / p->k : p->k; // exact reciprocal of a power of two, else the divisor itself
    h->cnt_gt1 = p->k > 1.0;
#define CKC(call)                                                                            \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            int rc_ = fail(nullptr, CCB_ECUDA, "%s: %s", #call, cudaGetErrorString(e_));     \
            ccb_destroy(h);                                                                  \
            return rc_;                                                                      \
        }                                                                                    \
    } while (0)
    CKC(cudaSetDevice(p->device));
    tune_pool(p->device);
    CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    CKC(cudaMalloc(&h->d_ctl, sizeof(Ctl)));
    CKC(cudaMemset(h->d_ctl, 0, sizeof(Ctl)));
    CKC(cudaMallocHost(&h->h_ctl, sizeof(Ctl)));
    memset(h->h_ctl, 0, sizeof(Ctl));
    int rc = 0;
    for (int i = 0; i < 2 && !rc; ++i) rc = alloc_store(h, h->P[i], 256);
    for (int i = 0; i < 2 && !rc; ++i) rc = alloc_store(h, h->O[i], 8192);
    if (!rc) rc = realloc_aux_for_outlier_cap(h);
    if (!rc) rc = realloc_aux_for_pcore_cap(h);
    if (rc) {
        g_err = h->err;
        ccb_destroy(h);
        return rc;
    }
#undef CKC
    *out = h;
    return CCB_OK;
}

void ccb_destroy(ccb_handle *h) {
    if (!h) return;
    cudaSetDevice(h->prm.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int i = 0; i < 2; ++i) {
        free_store(h->P[i]);
        free_store(h->O[i]);
    }
    cudaFree(h->d_ctl);
    if (h->h_ctl) cudaFreeHost(h->h_ctl);
    cudaFree(h->d_X);
    cudaFree(h->d_assign);
    cudaFree(h->d_stage);
    cudaFree(h->d_pnew);
    cudaFree(h->d_pfin);
    cudaFree(h->d_onew);
    drop_graph(h);
    std::swap(h->bs_graph, h->bs_graph_alt);
    std::swap(h->bs_exec, h->bs_exec_alt);
    drop_graph(h);
    if (h->cap1) cudaStreamDestroy(h->cap1);
    if (h->cap2) cudaStreamDestroy(h->cap2);
    if (h->cap3) cudaStreamDestroy(h->cap3);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->res_stream) cudaStreamDestroy(h->res_stream);
    if (h->h_res) cudaFreeHost(h->h_res);
    cudaFree(h->d_scale);
    for (auto &ev : h->ev_seg)
        if (ev) cudaEventDestroy(ev);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    cudaFree(h->d_io);
    if (h->h_io) cudaFreeHost(h->h_io);
    cudaFree(h->d_bc);
    if (h->h_bc) cudaFreeHost(h->h_bc);
    for (void *q : h->ws_allocs) cudaFree(q);
    for (void *q : h->ws_tiles) cudaFree(q);
    cudaFree(h->ws.dbg);
    cudaFree(h->ws_first);
    for (auto &b : h->ob) cudaFree(b.p);
    for (auto &e : h->ev_pending) {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    for (auto &e : h->ev_pool) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int ccb_get_stats(const ccb_handle *h, ccb_stats *out) {
    if (!h || !out) return fail(nullptr, CCB_EINVAL, "null argument");
    // the device-side counters are as of the last control-block read-back (every chunk / ccb_counts)
    *out = h->st;
    const Ctl &c = *h->h_ctl;
    out->upgrades = h->st_base.upgrades + c.upgrades;
    out->created = h->st_base.created + c.created;
    out->downgraded = h->st_base.downgraded + c.downgraded;
    out->deleted = h->st_base.deleted + c.deleted;
    if (h->h_bc) {
        const BsCtl &b = *h->h_bc, &bb = h->bc_base;
        out->bsv_blocks = bb.blocks + b.blocks;
        out->bsv_rounds = bb.iters + b.iters;
        out->bsv_mismatches = bb.mismatches + b.mismatches;
        out->bsv_cuts_unknown = bb.cuts_unknown + b.cuts_unknown;
        out->bsv_cuts_rounds = bb.cuts_iter + b.cuts_iter;
        out->bsv_cuts_capacity = bb.cuts_cap + b.cuts_cap;
        out->bsv_late_topk = bb.tk_late + b.tk_late;
        out->bsv_outlier_stage_cells = bb.rejects + b.rejects;
        out->bsv_replayed_cells = bb.replayed + b.replayed;
        out->bsv_light_rounds = bb.rounds_light + b.rounds_light;
        out->bsv_serial_cells = bb.serial_cells + b.serial_cells;
        out->nearest_pairs += bb.pairs + b.pairs;
    }
    out->bsv_pdl = (h->bs_use_graph && h->pdl && (h->bs_exec || h->bs_exec_alt)) ? 1 : 0;
    return CCB_OK;
}

#ifdef CCB_DEBUG
// Diagnostics build only (python chronoclust_b200/build.py --debug -> libchronoclust_b200_debug.so, declared in
// csrc/debug.h): neither symbol exists in the product library.
int ccb_debug_set(ccb_handle *h, int32_t mode) {
    if (!h) return fail(h, CCB_EINVAL, "null argument");
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaMemcpyToSymbol(g_bs_dbg_mode, &mode, sizeof(int)));
    return CCB_OK;
}

int64_t ccb_debug_trace(ccb_handle *h, int64_t *out, int64_t max_records) {
    if (!h || !out) return -1;
    cudaStreamSynchronize(h->stream);
    int n = 0;
    cudaMemcpyFromSymbol(&n, g_trace_n, sizeof(int));
    const int64_t m = std::min<int64_t>(std::min<int64_t>(n, CCB_TRACE_MAX), max_records);
    if (m > 0) cudaMemcpyFromSymbol(out, g_trace, (size_t)m * CCB_TRACE_WORDS * sizeof(long long));
    const int zero = 0;
    cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(int));
    return n;
}

int ccb_debug_counters(ccb_handle *h, int64_t *out16, int32_t reset) {
    if (!h || !out16) return -1;
    cudaStreamSynchronize(h->stream);
    cudaMemcpyFromSymbol(out16, g_dbg_cnt, 16 * sizeof(long long));
    if (reset) {
        const long long z[16] = {0};
        cudaMemcpyToSymbol(g_dbg_cnt, z, sizeof(z));
    }
    return 0;
}

int ccb_debug_chain(ccb_handle *h, int64_t *out, int32_t max_keys) {
    if (!h || !out) return fail(h, CCB_EINVAL, "null argument");
    if (!h->ws.dbg) { // first call switches the counters on
        if (h->ws.mp_stride <= 0) return fail(h, CCB_ESTATE, "no block-speculative workspace yet");
        CK(h, cudaStreamSynchronize(h->stream));
        CK(h, cudaMalloc(&h->ws.dbg, (size_t)h->ws.mp_stride * 8 * sizeof(int64_t)));
        CK(h, cudaMemset(h->ws.dbg, 0, (size_t)h->ws.mp_stride * 8 * sizeof(int64_t)));
        memset(out, 0, (size_t)max_keys * 8 * sizeof(int64_t));
        return CCB_OK;
    }
    CK(h, cudaStreamSynchronize(h->stream));
    const int n = std::min<int>(max_keys, h->ws.mp_stride);
    CK(h, cudaMemcpy(out, h->ws.dbg, (size_t)n * 8 * sizeof(int64_t), cudaMemcpyDeviceToHost));
    return CCB_OK;
}
#endif

int ccb_reset(ccb_handle *h) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    CK(h, cudaSetDevice(h->prm.device));
    int rc = sync_ctl(h);
    if (rc) return rc;
    {
        const Ctl &c = *h->h_ctl;
        h->st_base.upgrades += c.upgrades;
        h->st_base.created += c.created;
        h->st_base.downgraded += c.downgraded;
        h->st_base.deleted += c.deleted;
    }
    memset(h->h_ctl, 0, sizeof(Ctl));
    rc = push_ctl(h);
    if (rc) return rc;
    if (h->d_bc) {
        CK(h, cudaMemcpyAsync(h->h_bc, h->d_bc, sizeof(BsCtl), cudaMemcpyDeviceToHost, h->stream));
        CK(h, cudaStreamSynchronize(h->stream));
        const BsCtl &b = *h->h_bc;
        BsCtl &bb = h->bc_base;
        bb.blocks += b.blocks, bb.iters += b.iters, bb.mismatches += b.mismatches, bb.cuts_unknown += b.cuts_unknown;
        bb.cuts_iter += b.cuts_iter, bb.cuts_cap += b.cuts_cap, bb.tk_late += b.tk_late, bb.rejects += b.rejects;
        bb.replayed += b.replayed, bb.pairs += b.pairs, bb.rounds_light += b.rounds_light, bb.serial_cells += b.serial_cells;
        const int32_t keep_B = b.next_B;
        memset(h->h_bc, 0, sizeof(BsCtl));
        h->h_bc->next_B = keep_B;
        CK(h, cudaMemcpyAsync(h->d_bc, h->h_bc, sizeof(BsCtl), cudaMemcpyHostToDevice, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    h->have_params = false;
    h->off_M = 0;
    h->cl_off.assign(1, 0);
    h->cl_members.clear();
    h->cl_w.clear();
    return CCB_OK;
}

int ccb_enable_timing(ccb_handle *h, int32_t on) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    h->timing = on != 0;
    return CCB_OK;
}

int ccb_get_timing(ccb_handle *h, double ms[CCB_NCAT], int64_t launches[CCB_NCAT], int32_t reset) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    CK(h, cudaSetDevice(h->prm.device));
    CK(h, cudaStreamSynchronize(h->stream));
    drain_timing(h);
    for (int i = 0; i < CCB_NCAT; ++i) {
        if (ms) ms[i] = h->cat_ms[i];
        if (launches) launches[i] = h->cat_n[i];
        if (reset) {
            h->cat_ms[i] = 0;
            h->cat_n[i] = 0;
        }
    }
    return CCB_OK;
}

int ccb_set_dnrm2(ccb_handle *h, void *fn) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    h->dnrm2 = fn;
    return CCB_OK;
}

int ccb_begin_timepoint(ccb_handle *h, double mu, double omicron, int64_t pi, int32_t decay, double decay_factor) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    CK(h, cudaSetDevice(h->prm.device));
    h->mu = mu;
    h->omicron = omicron;
    h->pi = pi;
    h->have_params = true;
    if (!decay) return CCB_OK;
    cudaStream_t s = h->stream;
    MaintArgs ma{};
    ma.P = h->P[h->pcur];
    ma.O = h->O[h->ocur];
    ma.P2 = h->P[1 - h->pcur];
    ma.O2 = h->O[1 - h->ocur];
    ma.ctl = h->d_ctl;
    ma.p_new = h->d_pnew;
    ma.p_fin = h->d_pfin;
    ma.o_new = h->d_onew;
    ma.f = decay_factor;
    ma.beta_mu = h->prm.beta * mu;
    ma.omicron = omicron;
    ma.pi = pi;
    ma.D = h->D;
    ma.DP = h->DP;
    ma.cnt_gt1 = h->cnt_gt1;
    ma.wsel = h->wsel;
    int rc;
    if ((rc = sync_ctl(h))) return rc;
    // every downgraded pcore MC lands in the outlier list: make room first
    const int64_t need_o = (int64_t)h->h_ctl->n_outlier + h->h_ctl->n_pcore;
    if (need_o > h->O[h->ocur].cap) {
        if ((rc = grow_store(h, h->O, h->ocur, h->h_ctl->n_outlier, need_o))) return rc;
        if ((rc = realloc_aux_for_outlier_cap(h))) return rc;
        ma.O = h->O[h->ocur];
        ma.O2 = h->O[1 - h->ocur];
        ma.o_new = h->d_onew;
    }
    Timed tm(h, CCB_CAT_MAINT);
    k_maint_plan<<<1, MAINT_THREADS, 0, s>>>(ma);
    CKL(h);
    const int64_t total = (int64_t)(h->h_ctl->n_pcore + h->h_ctl->n_outlier) * h->DP;
    const int gx = (int)std::min<int64_t>(std::max<int64_t>((total + 255) / 256, 1), 148 * 8);
    k_maint_gather<<<gx, 256, 0, s>>>(ma);
    CKL(h);
    k_maint_finish<<<1, 1, 0, s>>>(h->d_ctl);
    CKL(h);
    h->st.kernel_launches += 3;
    h->pcur = 1 - h->pcur;
    h->ocur = 1 - h->ocur;
    tm.stop();
    return sync_ctl(h);
}

int ccb_ingest_device(ccb_handle *h, const double *X_dev, int64_t N, int64_t ld, int32_t *assign_uid_dev,
                      uint8_t *stage_dev) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    if (N < 0 || ld < h->D || (!X_dev && N > 0) || (!assign_uid_dev && N > 0)) return fail(h, CCB_EINVAL, "bad ingest arguments");
    CK(h, cudaSetDevice(h->prm.device));
    return ingest_core_bsv(h, X_dev, N, ld, assign_uid_dev, stage_dev);
}

static int ingest_host(ccb_handle *h, const double *X, int64_t N, int64_t ld, int32_t *assign_uid, uint8_t *stage,
                       const double *scale, const double *shift);

int ccb_ingest(ccb_handle *h, const double *X, int64_t N, int64_t ld, int32_t *assign_uid, uint8_t *stage) {
    return ingest_host(h, X, N, ld, assign_uid, stage, nullptr, nullptr);
}

int ccb_ingest_scaled(ccb_handle *h, const double *X_raw, int64_t N, int64_t ld, const double *scale, const double *min_,
                      int32_t *assign_uid, uint8_t *stage) {
    if (!scale || !min_) return fail(h, CCB_EINVAL, "null scaler vectors");
    return ingest_host(h, X_raw, N, ld, assign_uid, stage, scale, min_);
}

// host buffers in, per-cell results out; scale != nullptr: the rows are min-max scaled on the device first
// (X * scale + shift, two roundings: sklearn MinMaxScaler.transform), segment by segment behind their copies
static int ingest_host(ccb_handle *h, const double *X, int64_t N, int64_t ld, int32_t *assign_uid, uint8_t *stage,
                       const double *scale, const double *shift) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    if (N < 0 || ld < h->D || (!X && N > 0) || (!assign_uid && N > 0)) return fail(h, CCB_EINVAL, "bad ingest arguments");
    if (N == 0) return CCB_OK;
    CK(h, cudaSetDevice(h->prm.device));
    const size_t need = (size_t)N * ld;
    if (need > h->x_cap) {
        cudaFree(h->d_X);
        h->x_cap = 0;
        CK(h, cudaMalloc(&h->d_X, (need + 2) * 8));
        h->x_cap = need;
    }
    int rc = ensure_point_buffers(h, N);
    if (rc) return rc;
    if (scale) {
        if (!h->d_scale) CK(h, cudaMalloc(&h->d_scale, 2 * CCB_MAX_D * sizeof(double)));
        CK(h, cudaMemcpyAsync(h->d_scale, scale, (size_t)h->D * 8, cudaMemcpyHostToDevice, h->stream));
        CK(h, cudaMemcpyAsync(h->d_scale + CCB_MAX_D, shift, (size_t)h->D * 8, cudaMemcpyHostToDevice, h->stream));
    }
    auto scale_rows = [&](int64_t r0, int64_t n) { // on the handle's stream, behind the wait for the rows' copy
        if (!scale || n <= 0) return;
        const int64_t total = n * h->D;
        k_scale_rows<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16), 256, 0, h->stream>>>(
            h->d_X + r0 * ld, n, ld, h->D, h->d_scale, h->d_scale + CCB_MAX_D);
        h->st.kernel_launches++;
    };
    // segments of the input: one block's worth of rows first, then doubling, the last one takes the rest -- the engine
    // starts after a 3 MB copy instead of a 24 MB one and the copies (5x faster than the engine) stay ahead of it
    // The per-cell results travel back the same way, one segment behind the engine: through a page-locked bounce buffer
    // (the caller's arrays are ordinary pageable memory as a rule -- numpy -- and a device -> pageable copy runs at a
    // fraction of the link rate), copied out by this thread while the engine works on the next segment.  A short last
    // segment keeps the part that cannot overlap (its results) small.
    int64_t seg_end[8];
    int nseg = 1;
    seg_end[0] = N;
    if (!h->timing && N >= 262144) {
        int64_t len = h->bs_bmax, at = 0;
        nseg = 0;
        while (nseg < 6 && at + len + len < N) {
            at += len;
            seg_end[nseg++] = at;
            len *= 2;
        }
        const int64_t tail = 2 * (int64_t)h->bs_bmax;
        if (N - at > 4 * tail) seg_end[nseg++] = N - tail;
        seg_end[nseg++] = N;
    }
    if (nseg == 1) {
        {
            Timed tm(h, CCB_CAT_COPY);
            CK(h, cudaMemcpyAsync(h->d_X, X, need * 8, cudaMemcpyHostToDevice, h->stream));
        }
        scale_rows(0, N);
        rc = ingest_core_bsv(h, h->d_X, N, ld, h->d_assign, h->d_stage);
        if (rc) return rc;
    } else {
        // The ordered engine consumes cells front to back, so the input travels in segments on a copy stream and the
        // engine starts as soon as the first one has landed: segment k + 1 is queued right after the engine has been
        // launched on segment k (with pageable host memory the staging copy then overlaps the engine, too).
        if (!h->copy_stream) {
            CK(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
            for (auto &ev : h->ev_seg) CK(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        }
        CK(h, cudaEventRecord(h->ev_seg[0], h->stream)); // earlier work on the handle's stream (previous readers of d_X)
        CK(h, cudaStreamWaitEvent(h->copy_stream, h->ev_seg[0], 0));
        auto seg_rows = [&](int k, int64_t &r0, int64_t &n) {
            r0 = k ? seg_end[k - 1] : 0;
            n = seg_end[k] - r0;
        };
        auto copy_seg = [&](int k) -> int {
            int64_t r0, n;
            seg_rows(k, r0, n);
            if (n > 0)
                CK(h, cudaMemcpyAsync(h->d_X + r0 * ld, X + r0 * ld, (size_t)n * ld * 8, cudaMemcpyHostToDevice, h->copy_stream));
            CK(h, cudaEventRecord(h->ev_seg[k], h->copy_stream));
            return CCB_OK;
        };
        if (!h->res_stream) CK(h, cudaStreamCreateWithFlags(&h->res_stream, cudaStreamNonBlocking));
        if ((size_t)N > h->res_cap) {
            if (h->h_res) cudaFreeHost(h->h_res);
            h->h_res = nullptr;
            h->res_cap = 0;
            CK(h, cudaHostAlloc((void **)&h->h_res, (size_t)N * 5, cudaHostAllocDefault));
            h->res_cap = (size_t)N;
        }
        int32_t *ba = reinterpret_cast<int32_t *>(h->h_res);
        uint8_t *bs = h->h_res + (size_t)N * 4;
        // results of segment k (its engine run is complete: ingest_core_bsv returns after a synchronisation) -> caller
        auto results_seg = [&](int k) -> int {
            int64_t r0, n;
            seg_rows(k, r0, n);
            if (n <= 0) return CCB_OK;
            CK(h, cudaMemcpyAsync(ba + r0, h->d_assign + r0, (size_t)n * 4, cudaMemcpyDeviceToHost, h->res_stream));
            if (stage) CK(h, cudaMemcpyAsync(bs + r0, h->d_stage + r0, (size_t)n, cudaMemcpyDeviceToHost, h->res_stream));
            CK(h, cudaStreamSynchronize(h->res_stream));
            memcpy(assign_uid + r0, ba + r0, (size_t)n * 4);
            if (stage) memcpy(stage + r0, bs + r0, (size_t)n);
            return CCB_OK;
        };
        if ((rc = copy_seg(0))) return rc;
        for (int k = 0; k < nseg; ++k) {
            int64_t r0, n;
            seg_rows(k, r0, n);
            CK(h, cudaStreamWaitEvent(h->stream, h->ev_seg[k], 0));
            scale_rows(r0, n);
            // once the engine is running on segment k: queue the input of segment k + 1, fetch the results of segment k - 1
            const std::function<int()> next = [&]() -> int {
                int r = k + 1 < nseg ? copy_seg(k + 1) : CCB_OK;
                if (!r && k > 0) r = results_seg(k - 1);
                return r;
            };
            if ((rc = ingest_core_bsv(h, h->d_X + r0 * ld, n, ld, h->d_assign + r0, h->d_stage + r0, &next))) return rc;
        }
        if ((rc = results_seg(nseg - 1))) return rc;
        return CCB_OK;
    }
    {
        Timed tm(h, CCB_CAT_COPY);
        CK(h, cudaMemcpyAsync(assign_uid, h->d_assign, (size_t)N * 4, cudaMemcpyDeviceToHost, h->stream));
        if (stage) CK(h, cudaMemcpyAsync(stage, h->d_stage, (size_t)N, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(h, cudaStreamSynchronize(h->stream));
    if (h->timing) drain_timing(h);
    return CCB_OK;
}

int ccb_counts(ccb_handle *h, int64_t out[4]) {
    if (!h || !out) return fail(nullptr, CCB_EINVAL, "null argument");
    CK(h, cudaSetDevice(h->prm.device));
    int rc = sync_ctl(h);
    if (rc) return rc;
    out[0] = h->h_ctl->n_pcore;
    out[1] = h->h_ctl->n_outlier_alive;
    out[2] = h->h_ctl->pcore_last_id;
    out[3] = h->h_ctl->outlier_last_id;
    return CCB_OK;
}

int ccb_export_list(ccb_handle *h, int32_t which, int64_t *ids, int64_t *uids, double *w, double *cf1, double *cf2,
                    double *cen, double *pref) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    CK(h, cudaSetDevice(h->prm.device));
    int rc = sync_ctl(h);
    if (rc) return rc;
    const Store &S = which ? h->O[h->ocur] : h->P[h->pcur];
    const int n = which ? h->h_ctl->n_outlier : h->h_ctl->n_pcore;
    const int D = h->D;
    if (n == 0) return CCB_OK;
    std::vector<double> hw(n), h1((size_t)n * D), h2((size_t)n * D), hc((size_t)n * D);
    std::vector<uint64_t> hm(n);
    std::vector<int64_t> hid(n);
    std::vector<int32_t> hu(n);
    CK(h, cudaMemcpy(hw.data(), S.w, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(h1.data(), S.cf1, (size_t)n * D * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(h2.data(), S.cf2, (size_t)n * D * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(hc.data(), S.cen, (size_t)n * D * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(hm.data(), S.mask, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(hid.data(), S.id, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(hu.data(), S.uid, (size_t)n * 4, cudaMemcpyDeviceToHost));
    int o = 0;
    for (int j = 0; j < n; ++j) {
        if (which && hw[j] < 0.0) continue; // tombstone left by an upgrade
        if (ids) ids[o] = hid[j];
        if (uids) uids[o] = hu[j];
        if (w) w[o] = hw[j];
        for (int d = 0; d < D; ++d) {
            if (cf1) cf1[(size_t)o * D + d] = h1[(size_t)j * D + d];
            if (cf2) cf2[(size_t)o * D + d] = h2[(size_t)j * D + d];
            if (cen) cen[(size_t)o * D + d] = hc[(size_t)j * D + d];
            if (pref) pref[(size_t)o * D + d] = ((hm[j] >> d) & 1ull) ? h->prm.k : 1.0;
        }
        ++o;
    }
    return CCB_OK;
}

int ccb_import_list(ccb_handle *h, int32_t which, int64_t n, const int64_t *ids, const int64_t *uids, const double *w,
                    const double *cf1, const double *cf2, const double *cen, const double *pref) {
    if (!h || n < 0) return fail(nullptr, CCB_EINVAL, "bad argument");
    CK(h, cudaSetDevice(h->prm.device));
    int rc = sync_ctl(h);
    if (rc) return rc;
    const int D = h->D;
    if (which) {
        if ((rc = grow_store(h, h->O, h->ocur, 0, std::max<int64_t>(n, 1)))) return rc;
        if ((rc = realloc_aux_for_outlier_cap(h))) return rc;
    } else {
        if ((rc = grow_store(h, h->P, h->pcur, 0, std::max<int64_t>(n, 1)))) return rc;
        if ((rc = realloc_aux_for_pcore_cap(h))) return rc;
    }
    Store &S = which ? h->O[h->ocur] : h->P[h->pcur];
    std::vector<uint64_t> hm(n);
    std::vector<int32_t> hu(n);
    for (int64_t j = 0; j < n; ++j) {
        uint64_t m = 0;
        for (int d = 0; d < D; ++d)
            if (pref[(size_t)j * D + d] == h->prm.k && h->prm.k != 1.0) m |= 1ull << d;
        hm[j] = m;
        if (uids[j] < 0 || uids[j] >= ((int64_t)1 << 31)) return fail(h, CCB_ELIMIT, "uid outside int32");
        hu[j] = (int32_t)uids[j];
    }
    if (n) {
        CK(h, cudaMemcpy(S.w, w, (size_t)n * 8, cudaMemcpyHostToDevice));
        CK(h, cudaMemcpy(S.cf1, cf1, (size_t)n * D * 8, cudaMemcpyHostToDevice));
        CK(h, cudaMemcpy(S.cf2, cf2, (size_t)n * D * 8, cudaMemcpyHostToDevice));
        CK(h, cudaMemcpy(S.cen, cen, (size_t)n * D * 8, cudaMemcpyHostToDevice));
        CK(h, cudaMemcpy(S.mask, hm.data(), (size_t)n * 8, cudaMemcpyHostToDevice));
        CK(h, cudaMemcpy(S.id, ids, (size_t)n * 8, cudaMemcpyHostToDevice));
        CK(h, cudaMemcpy(S.uid, hu.data(), (size_t)n * 4, cudaMemcpyHostToDevice));
        k_repack_store<<<(unsigned)(((size_t)n * h->DP + 255) / 256), 256, 0, h->stream>>>(S, (int)n, D, h->DP, h->wsel);
        CKL(h);
        h->st.kernel_launches++;
    }
    if (which) {
        h->h_ctl->n_outlier = (int32_t)n;
        h->h_ctl->n_outlier_alive = (int32_t)n;
    } else {
        h->h_ctl->n_pcore = (int32_t)n;
    }
    if ((rc = push_ctl(h))) return rc;
    CK(h, cudaStreamSynchronize(h->stream));
    return CCB_OK;
}

int ccb_set_counters(ccb_handle *h, int64_t pcore_last_id, int64_t outlier_last_id) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    CK(h, cudaSetDevice(h->prm.device));
    int rc = sync_ctl(h);
    if (rc) return rc;
    h->h_ctl->pcore_last_id = pcore_last_id;
    h->h_ctl->outlier_last_id = outlier_last_id;
    if ((rc = push_ctl(h))) return rc;
    CK(h, cudaStreamSynchronize(h->stream));
    return CCB_OK;
}

// ---- offline ------------------------------------------------------------------------------------------
int ccb_offline(ccb_handle *h, int64_t *n_clusters) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    if (!h->have_params) return fail(h, CCB_ESTATE, "ccb_begin_timepoint must precede ccb_offline");
    CK(h, cudaSetDevice(h->prm.device));
    int rc = sync_ctl(h);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const int M = h->h_ctl->n_pcore, D = h->D;
    const Store &P = h->P[h->pcur];
    const int words = (M + 31) / 32;
    h->off_M = M;
    h->cl_off.assign(1, 0);
    h->cl_members.clear();
    h->cl_w.clear();
    h->cl_cf1.clear();
    h->cl_cf2.clear();
    h->cl_cen.clear();
    h->cl_pref.clear();
    h->cl_label.assign(M, -1);
    if (n_clusters) *n_clusters = 0;
    if (M == 0) return CCB_OK;

    const double E2 = h->prm.upsilon_eps2;
    Timed tm_off(h, CCB_CAT_OFFLINE);
    uint8_t *core, *cls;
    uint32_t *nbr, *wnbr;
    int32_t *cnt, *border, *nborder, *queue, *label, *order, *cloff, *ncl;
    uint64_t *submask;
    int border_cap = (int)std::max<size_t>(h->ob[ccb_handle::OB_BORDER].bytes / 8, (size_t)1 << 16);
#define OBUF(var, which, n)                                                      \
    if ((rc = off_buf(h, ccb_handle::which, (size_t)(n) * sizeof(*var)))) return rc; \
    var = (decltype(var))h->ob[ccb_handle::which].p
    OBUF(core, OB_CORE, M);
    OBUF(cls, OB_CLS, M);
    OBUF(nbr, OB_NBR, (size_t)M * words);
    OBUF(wnbr, OB_WNBR, (size_t)M * words);
    OBUF(cnt, OB_CNT, M);
    OBUF(border, OB_BORDER, 2 * (size_t)border_cap);
    OBUF(nborder, OB_NBORDER, 1);
    OBUF(queue, OB_QUEUE, 2 * (size_t)M + 2);
    OBUF(label, OB_LABEL, M);
    OBUF(order, OB_ORDER, M);
    OBUF(cloff, OB_CLOFF, (size_t)M + 2);
    OBUF(ncl, OB_NCL, 1);
    OBUF(submask, OB_SUBMASK, M);
    CK(h, cudaMemsetAsync(submask, 0, (size_t)M * 8, s));
    CK(h, cudaMemsetAsync(cls, 0, (size_t)M, s));

    k_off_core<<<(M + 127) / 128, 128, 0, s>>>(P.cf1, P.cf2, P.w, P.mask, M, D, h->prm.k, h->wsel, h->div_mode, h->cnt_gt1,
                                               h->prm.eps2, h->mu, h->pi, core);
    CKL(h);
    h->st.kernel_launches++;
    int32_t nb = 0;
    for (;;) { // the borderline list grows to whatever the data needs (e.g. many coincident centroids with upsilon = 0)
        CK(h, cudaMemsetAsync(nborder, 0, 4, s));
        if ((rc = launch_off_neighbours(h, s, P.cen, M, D, 0, M, E2, nbr, cnt, border, border_cap, nborder))) return rc;
        h->st.kernel_launches++;
        CK(h, cudaMemcpyAsync(&nb, nborder, 4, cudaMemcpyDeviceToHost, s));
        CK(h, cudaStreamSynchronize(s));
        if (nb <= border_cap) break;
        border_cap = nb;
        OBUF(border, OB_BORDER, 2 * (size_t)border_cap);
    }
    if (nb > 0) {
        // settle the pairs inside the guard band with the BLAS dnrm2 the reference itself calls
        std::vector<int32_t> pairs(2 * (size_t)nb);
        std::vector<double> hc((size_t)M * D), x(D);
        std::vector<uint8_t> dec(nb);
        CK(h, cudaMemcpy(pairs.data(), border, pairs.size() * 4, cudaMemcpyDeviceToHost));
        CK(h, cudaMemcpy(hc.data(), P.cen, hc.size() * 8, cudaMemcpyDeviceToHost));
        typedef double (*nrm2_fn)(int *, double *, int *);
        for (int i = 0; i < nb; ++i) {
            const int p_ = pairs[2 * i], q_ = pairs[2 * i + 1];
            for (int d = 0; d < D; ++d) x[d] = hc[(size_t)q_ * D + d] - hc[(size_t)p_ * D + d];
            double r;
            if (h->dnrm2) {
                int n = D, inc = 1;
                r = ((nrm2_fn)h->dnrm2)(&n, x.data(), &inc);
            } else {
                r = builtin_nrm2(x.data(), D);
            }
            dec[i] = r <= h->prm.upsilon_eps;
        }
        uint8_t *ddec;
        OBUF(ddec, OB_DEC, nb);
        CK(h, cudaMemcpyAsync(ddec, dec.data(), nb, cudaMemcpyHostToDevice, s));
        k_off_patch<<<(nb + 127) / 128, 128, 0, s>>>(nbr, cnt, border, ddec, nb, 0, words);
        CKL(h);
        CK(h, cudaStreamSynchronize(s)); // dec (host) is read by the copy above
        h->st.borderline_pairs += nb;
        h->st.kernel_launches++;
    }
    {
        const int64_t t = (int64_t)M * D;
        k_off_subspace<<<(unsigned)((t + 127) / 128), 128, 0, s>>>(P.cen, M, D, 0, M, nbr, cnt, h->prm.delta, submask);
        CKL(h);
        k_off_weighted<<<(unsigned)((M + 3) / 4), OFFW_THREADS, 0, s>>>(P.cen, M, D, 0, M, nbr, submask, h->prm.k, E2, wnbr);
        CKL(h);
        int nl = 0, rc2;
        if ((rc2 = launch_off_clusters(h, s, M, wnbr, core, submask, h->cnt_gt1, h->pi, cls, queue, label, order, cloff, ncl, &nl,
                                       h->off_csr_min_m)))
            return rc2;
        h->st.kernel_launches += 2 + nl;
    }
    int32_t nc_raw = 0;
    CK(h, cudaMemcpyAsync(&nc_raw, ncl, 4, cudaMemcpyDeviceToHost, s));
    CK(h, cudaStreamSynchronize(s));
    std::vector<int32_t> hl(M), ho(M), hoff((size_t)nc_raw + 1);
    std::vector<double> kw(nc_raw), k1((size_t)nc_raw * D), k2((size_t)nc_raw * D), kc((size_t)nc_raw * D);
    std::vector<uint64_t> km(nc_raw);
    if (nc_raw > 0) {
        double *d1, *d2, *dc, *dw;
        uint64_t *omask;
        OBUF(d1, OB_D1, (size_t)nc_raw * D);
        OBUF(d2, OB_D2, (size_t)nc_raw * D);
        OBUF(dc, OB_DC, (size_t)nc_raw * D);
        OBUF(dw, OB_DW, nc_raw);
        OBUF(omask, OB_OMASK, nc_raw);
        k_off_cluster_cf<<<nc_raw, 64, 0, s>>>(P.cf1, P.cf2, P.w, D, order, cloff, h->prm.delta2, d1, d2, dc, omask, dw);
        CKL(h);
        h->st.kernel_launches++;
        CK(h, cudaMemcpyAsync(kw.data(), dw, (size_t)nc_raw * 8, cudaMemcpyDeviceToHost, s));
        CK(h, cudaMemcpyAsync(k1.data(), d1, k1.size() * 8, cudaMemcpyDeviceToHost, s));
        CK(h, cudaMemcpyAsync(k2.data(), d2, k2.size() * 8, cudaMemcpyDeviceToHost, s));
        CK(h, cudaMemcpyAsync(kc.data(), dc, kc.size() * 8, cudaMemcpyDeviceToHost, s));
        CK(h, cudaMemcpyAsync(km.data(), omask, (size_t)nc_raw * 8, cudaMemcpyDeviceToHost, s));
    }
#undef OBUF
    tm_off.stop();
    CK(h, cudaMemcpyAsync(hl.data(), label, (size_t)M * 4, cudaMemcpyDeviceToHost, s));
    CK(h, cudaMemcpyAsync(ho.data(), order, (size_t)M * 4, cudaMemcpyDeviceToHost, s));
    CK(h, cudaMemcpyAsync(hoff.data(), cloff, hoff.size() * 4, cudaMemcpyDeviceToHost, s));
    std::vector<int64_t> ids(M);
    CK(h, cudaMemcpyAsync(ids.data(), P.id, (size_t)M * 8, cudaMemcpyDeviceToHost, s));
    CK(h, cudaStreamSynchronize(s));
    if (h->timing) drain_timing(h);
    // keep clusters whose weight is > 0 (predecon.py:83); relabel
    std::vector<int32_t> remap(nc_raw, -1);
    for (int c = 0; c < nc_raw; ++c) {
        if (!(kw[c] > 0.0)) continue;
        remap[c] = (int32_t)h->cl_w.size();
        h->cl_w.push_back(kw[c]);
        for (int i = hoff[c]; i < hoff[c + 1]; ++i) h->cl_members.push_back(ids[ho[i]]);
        h->cl_off.push_back((int64_t)h->cl_members.size());
        for (int d = 0; d < D; ++d) {
            h->cl_cf1.push_back(k1[(size_t)c * D + d]);
            h->cl_cf2.push_back(k2[(size_t)c * D + d]);
            h->cl_cen.push_back(kc[(size_t)c * D + d]);
            h->cl_pref.push_back(((km[c] >> d) & 1ull) ? h->prm.k : 1.0);
        }
    }
    for (int j = 0; j < M; ++j) h->cl_label[j] = hl[j] >= 0 ? remap[hl[j]] : -1;
    if (n_clusters) *n_clusters = (int64_t)h->cl_w.size();
    return CCB_OK;
}

int ccb_cluster_sizes(ccb_handle *h, int64_t out[3]) {
    if (!h || !out) return fail(nullptr, CCB_EINVAL, "null argument");
    out[0] = (int64_t)h->cl_w.size();
    out[1] = (int64_t)h->cl_members.size();
    out[2] = h->off_M;
    return CCB_OK;
}

int ccb_export_clusters(ccb_handle *h, int64_t *off, int64_t *members, double *w, double *cf1, double *cf2, double *cen,
                        double *pref, int32_t *label) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    auto cp = [](auto *dst, const auto &v) {
        if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(v[0]));
    };
    cp(off, h->cl_off);
    cp(members, h->cl_members);
    cp(w, h->cl_w);
    cp(cf1, h->cl_cf1);
    cp(cf2, h->cl_cf2);
    cp(cen, h->cl_cen);
    cp(pref, h->cl_pref);
    cp(label, h->cl_label);
    return CCB_OK;
}

int ccb_export_offline(ccb_handle *h, uint8_t *core, uint8_t *nbr, uint8_t *wnbr, double *subw) {
    if (!h) return fail(nullptr, CCB_EINVAL, "null handle");
    const int64_t M = h->off_M;
    if (M == 0) return CCB_OK;
    CK(h, cudaSetDevice(h->prm.device));
    // read back from the offline workspace of the last ccb_offline (kept on the device until the next one)
    const int D = h->D, words = (int)((M + 31) / 32);
    std::vector<uint8_t> hcore(M);
    std::vector<uint32_t> hn((size_t)M * words), hw((size_t)M * words);
    std::vector<uint64_t> hs(M);
    CK(h, cudaStreamSynchronize(h->stream));
    CK(h, cudaMemcpy(hcore.data(), h->ob[ccb_handle::OB_CORE].p, (size_t)M, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(hn.data(), h->ob[ccb_handle::OB_NBR].p, hn.size() * 4, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(hw.data(), h->ob[ccb_handle::OB_WNBR].p, hw.size() * 4, cudaMemcpyDeviceToHost));
    CK(h, cudaMemcpy(hs.data(), h->ob[ccb_handle::OB_SUBMASK].p, (size_t)M * 8, cudaMemcpyDeviceToHost));
    for (int64_t p = 0; p < M; ++p) {
        if (core) core[p] = hcore[p];
        for (int64_t q = 0; q < M; ++q) {
            if (nbr) nbr[p * M + q] = (hn[p * words + (q >> 5)] >> (q & 31)) & 1u;
            if (wnbr) wnbr[p * M + q] = (hw[p * words + (q >> 5)] >> (q & 31)) & 1u;
        }
        if (subw)
            for (int d = 0; d < D; ++d) subw[p * D + d] = ((hs[p] >> d) & 1ull) ? h->prm.k : 1.0;
    }
    return CCB_OK;
}

// ---- FP64 pipe microbenchmark (roofline denominators; MEASURED_PEAKS.json has no FP64 figure) ----------
__global__ void k_fp64_peak(int mode, int iters, double *sink) {
    double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3;
    double a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
    const double m = 1.0000000001, c = 1e-12;
    if (mode == 0) { // fused multiply-add stream: 2 flop per instruction
        for (int i = 0; i < iters; ++i) {
            a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
            a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
        }
    } else { // separate multiply and add, never contracted: 1 flop per instruction (what parity allows)
        for (int i = 0; i < iters; ++i) {
            a0 = __dadd_rn(__dmul_rn(a0, m), c); a1 = __dadd_rn(__dmul_rn(a1, m), c);
            a2 = __dadd_rn(__dmul_rn(a2, m), c); a3 = __dadd_rn(__dmul_rn(a3, m), c);
            a4 = __dadd_rn(__dmul_rn(a4, m), c); a5 = __dadd_rn(__dmul_rn(a5, m), c);
            a6 = __dadd_rn(__dmul_rn(a6, m), c); a7 = __dadd_rn(__dmul_rn(a7, m), c);
        }
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) sink[0] = r;
}

// single-warp latency probes (cycles per dependent step), results in sink[1..]
__global__ void k_fp64_latency(double *sink) {
    __shared__ double sm[64 * 26];
    const int lane = threadIdx.x;
    for (int i = lane; i < 64 * 26; i += 32) sm[i] = 1e-9 * i;
    __syncwarp();
    double a = lane * 1e-9 + 1.0;
    const double c = 1e-12, m = 1.0000000001;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) a = __dadd_rn(a, c);
    long long t1 = clock64();
    const double r_add = (double)(t1 - t0) / 1024.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) a = __dmul_rn(a, m);
    t1 = clock64();
    const double r_mul = (double)(t1 - t0) / 1024.0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) a = __fma_rn(a, m, c);
    t1 = clock64();
    const double r_fma = (double)(t1 - t0) / 1024.0;
    // the replay pattern: 8 loads, 8 chained adds, 8 stores in place
    t0 = clock64();
    for (int g = 0; g < 128; ++g) {
        double av[8];
        const int base = (g & 7) * 8 * 26 + (lane < 26 ? lane : 0);
#pragma unroll
        for (int q = 0; q < 8; ++q) av[q] = sm[base + q * 26];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            a = __dadd_rn(a, av[q]);
            if (lane < 26) sm[base + q * 26] = a;
        }
    }
    t1 = clock64();
    const double r_rep = (double)(t1 - t0) / 1024.0;
    float f = (float)a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < 1024; ++i) f = __fadd_rn(f, 1e-6f);
    t1 = clock64();
    const double r_fadd = (double)(t1 - t0) / 1024.0;
    if (lane == 0) {
        sink[1] = r_add;
        sink[2] = r_mul;
        sink[3] = r_fma;
        sink[4] = r_rep;
        sink[5] = r_fadd;
        sink[6] = a + f;
    }
}

int ccb_fp64_peak(int32_t device, void *stream, int32_t mode, int32_t iters, int32_t blocks, double *sink,
                  double *flops_out) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (mode == 3) { // probes of the replay pattern: sink[8..11] = cycles per cell (see k_chain_probe)
        k_chain_probe<<<1, 96, 0, (cudaStream_t)stream>>>(sink);
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "k_chain_probe: %s", cudaGetErrorString(e));
        if (flops_out) *flops_out = 0.0;
        return CCB_OK;
    }
    if (mode == 2) { // latency probes: sink[1..5] = cycles per dependent DADD, DMUL, DFMA, replay step, FADD
        k_fp64_latency<<<1, 32, 0, (cudaStream_t)stream>>>(sink);
        e = cudaGetLastError();
        if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "k_fp64_latency: %s", cudaGetErrorString(e));
        if (flops_out) *flops_out = 0.0;
        return CCB_OK;
    }
    k_fp64_peak<<<blocks, 256, 0, (cudaStream_t)stream>>>(mode, iters, sink);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "k_fp64_peak: %s", cudaGetErrorString(e));
    if (flops_out) *flops_out = (double)blocks * 256.0 * (double)iters * 8.0 * 2.0; // 8 chains x (mul + add)
    return CCB_OK;
}

// ---- stateless device-pointer entry points ------------------------------------------------------------
int ccb_nearest(int32_t device, void *stream, const double *X, int64_t N, int64_t ld, int32_t D, const double *cen,
                const uint64_t *prefmask, int64_t M, double k, int32_t *slot, double *dist) {
    if (D < 1 || D > CCB_MAX_D || N < 0 || M < 0 || ld < D) return fail(nullptr, CCB_EINVAL, "bad ccb_nearest arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0) return CCB_OK;
    if (M == 0) {
        cudaMemsetAsync(slot, 0xff, (size_t)N * 4, s);
        return CCB_OK;
    }
    const int DP = round_dp(D);
    const bool p2 = is_pow2(k);
    const double wsel = p2 ? 1.0 / k : k;
    double2 *cw = nullptr;
    double *sd = nullptr;
    int32_t *si = nullptr;
    // scratch: packed rows + per-slab candidates, sized by the split the launch will use
    tune_pool(device);
    int max_slabs = 1;
    {
        int ok = 0;
        CCB_DISPATCH_DP(DP, {
            int gx, slab_mcs;
            nearest_static_split<kDP, 1>(p2 ? 0 : 1, N, (int)M, MAX_SLABS, gx, slab_mcs, max_slabs);
            ok = 1;
        })
        if (!ok) return fail(nullptr, CCB_ELIMIT, "unsupported dimensionality %d", D);
    }
    if ((e = cudaMallocAsync(&cw, (size_t)M * DP * sizeof(double2), s)) != cudaSuccess ||
        (e = cudaMallocAsync(&sd, (size_t)N * max_slabs * 8, s)) != cudaSuccess ||
        (e = cudaMallocAsync(&si, (size_t)N * max_slabs * 4, s)) != cudaSuccess)
        return fail(nullptr, CCB_ENOMEM, "scratch allocation: %s", cudaGetErrorString(e));
    k_pack_cw<<<(unsigned)(((size_t)M * DP + 255) / 256), 256, 0, s>>>(cen, prefmask, M, D, DP, wsel, cw);
    int rc = launch_nearest<1>(nullptr, s, DP, p2 ? 0 : 1, X, nullptr, nullptr, 0, N, ld, D, cw, (int)M, sd, si, dist, slot,
                               max_slabs, nullptr);
    cudaFreeAsync(cw, s);
    cudaFreeAsync(sd, s);
    cudaFreeAsync(si, s);
    return rc;
}

int ccb_assoc_nearest(int32_t device, void *stream, const double *cur_cen, const uint64_t *cur_prefmask, int64_t Q,
                      const double *prev_cen, int64_t P, int32_t D, double k, int32_t *best, double *dist) {
    return ccb_assoc_nearest2(device, stream, cur_cen, cur_prefmask, Q, prev_cen, P, D, k, best, dist, nullptr);
}

int ccb_assoc_nearest2(int32_t device, void *stream, const double *cur_cen, const uint64_t *cur_prefmask, int64_t Q,
                       const double *prev_cen, int64_t P, int32_t D, double k, int32_t *best, double *dist, double *dist2) {
    if (D < 1 || D > CCB_MAX_D || Q < 0 || P < 0 || Q >= ((int64_t)1 << 31) || P >= ((int64_t)1 << 31))
        return fail(nullptr, CCB_EINVAL, "bad ccb_assoc_nearest arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaStream_t s = (cudaStream_t)stream;
    if (Q == 0) return CCB_OK;
    if (P == 0) {
        cudaMemsetAsync(best, 0xff, (size_t)Q * 4, s);
        return CCB_OK;
    }
    const int DP = round_dp(D);
    const bool p2 = is_pow2(k);
    int ok = 0;
    CCB_DISPATCH_DP(DP, {
        const int gx = (int)((Q + ASSOC_THREADS - 1) / ASSOC_THREADS);
        if (p2) k_assoc<kDP, false><<<gx, ASSOC_THREADS, 0, s>>>(cur_cen, cur_prefmask, (int)Q, prev_cen, (int)P, D, k, 1.0 / k, best, dist, dist2);
        else k_assoc<kDP, true><<<gx, ASSOC_THREADS, 0, s>>>(cur_cen, cur_prefmask, (int)Q, prev_cen, (int)P, D, k, k, best, dist, dist2);
        ok = 1;
    })
    if (!ok) return fail(nullptr, CCB_ELIMIT, "unsupported dimensionality %d", D);
    e = cudaGetLastError();
    return e == cudaSuccess ? CCB_OK : fail(nullptr, CCB_ECUDA, "k_assoc launch: %s", cudaGetErrorString(e));
}

int ccb_colminmax(int32_t device, void *stream, const double *X_dev, int64_t N, int64_t ld, int32_t D, double *min_dev,
                  double *max_dev) {
    if (D < 1 || D > CCB_MAX_D || N < 0 || ld < D) return fail(nullptr, CCB_EINVAL, "bad ccb_colminmax arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaStream_t s = (cudaStream_t)stream;
    tune_pool(device);
    unsigned long long *okey = nullptr;
    if ((e = cudaMallocAsync(&okey, 2 * (size_t)D * 8, s)) != cudaSuccess)
        return fail(nullptr, CCB_ENOMEM, "scratch allocation: %s", cudaGetErrorString(e));
    cudaMemsetAsync(okey, 0xff, (size_t)D * 8, s);   // running minima: largest key
    cudaMemsetAsync(okey + D, 0x00, (size_t)D * 8, s); // running maxima: smallest key
    const int64_t total = N * D;
    if (total > 0)
        k_colminmax<<<(unsigned)std::min<int64_t>((total + 255) / 256, 148 * 8), 256, 0, s>>>(X_dev, N, ld, D, okey);
    k_colminmax_finish<<<1, 64, 0, s>>>(okey, D, min_dev, max_dev);
    e = cudaGetLastError();
    cudaFreeAsync(okey, s);
    return e == cudaSuccess ? CCB_OK : fail(nullptr, CCB_ECUDA, "k_colminmax: %s", cudaGetErrorString(e));
}

int ccb_off_neighbours(int32_t device, void *stream, const double *cen, int64_t M, int32_t D, int64_t r0, int64_t r1,
                       double E2, uint32_t *nbr, int32_t *cnt, int32_t *border, int32_t border_cap, int32_t *n_border) {
    if (D < 1 || D > CCB_MAX_D || M < 0 || r0 < 0 || r1 < r0 || r1 > M) return fail(nullptr, CCB_EINVAL, "bad arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    return launch_off_neighbours(nullptr, (cudaStream_t)stream, cen, (int)M, D, (int)r0, (int)r1, E2, nbr, cnt, border,
                                 border_cap, n_border);
}

int ccb_off_patch(int32_t device, void *stream, uint32_t *nbr, int32_t *cnt, const int32_t *pairs, const uint8_t *decision,
                  int32_t n, int64_t r0, int64_t M) {
    if (n < 0 || M < 0) return fail(nullptr, CCB_EINVAL, "bad arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (n == 0) return CCB_OK;
    k_off_patch<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(nbr, cnt, pairs, decision, n, (int)r0, (int)((M + 31) / 32));
    e = cudaGetLastError();
    return e == cudaSuccess ? CCB_OK : fail(nullptr, CCB_ECUDA, "k_off_patch: %s", cudaGetErrorString(e));
}

int ccb_off_subspace(int32_t device, void *stream, const double *cen, int64_t M, int32_t D, int64_t r0, int64_t r1,
                     const uint32_t *nbr, const int32_t *cnt, double delta, uint64_t *submask) {
    if (D < 1 || D > CCB_MAX_D || M < 0 || r0 < 0 || r1 < r0 || r1 > M) return fail(nullptr, CCB_EINVAL, "bad arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaStream_t s = (cudaStream_t)stream;
    if (r1 == r0) return CCB_OK;
    cudaMemsetAsync(submask, 0, (size_t)(r1 - r0) * 8, s);
    const int64_t t = (r1 - r0) * D;
    k_off_subspace<<<(unsigned)((t + 127) / 128), 128, 0, s>>>(cen, (int)M, D, (int)r0, (int)r1, nbr, cnt, delta, submask);
    e = cudaGetLastError();
    return e == cudaSuccess ? CCB_OK : fail(nullptr, CCB_ECUDA, "k_off_subspace: %s", cudaGetErrorString(e));
}

int ccb_off_weighted(int32_t device, void *stream, const double *cen, int64_t M, int32_t D, int64_t r0, int64_t r1,
                     const uint32_t *nbr, const uint64_t *submask_all, double k, double E2, uint32_t *wnbr) {
    if (D < 1 || D > CCB_MAX_D || M < 0 || r0 < 0 || r1 < r0 || r1 > M) return fail(nullptr, CCB_EINVAL, "bad arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (r1 == r0) return CCB_OK;
    k_off_weighted<<<(unsigned)((r1 - r0 + 3) / 4), OFFW_THREADS, 0, (cudaStream_t)stream>>>(cen, (int)M, D, (int)r0, (int)r1, nbr,
                                                                                             submask_all, k, E2, wnbr);
    e = cudaGetLastError();
    return e == cudaSuccess ? CCB_OK : fail(nullptr, CCB_ECUDA, "k_off_weighted: %s", cudaGetErrorString(e));
}

int ccb_off_clusters(int32_t device, void *stream, int64_t M, const uint32_t *wnbr, const uint8_t *core,
                     const uint64_t *submask_all, double k, int64_t pi, int32_t csr_min_m, int32_t *label, int32_t *order,
                     int32_t *cl_off, int32_t *n_cl) {
    if (M < 0) return fail(nullptr, CCB_EINVAL, "bad arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaStream_t s = (cudaStream_t)stream;
    tune_pool(device);
    uint8_t *cls = nullptr;
    int32_t *queue = nullptr;
    if ((e = cudaMallocAsync(&cls, (size_t)std::max<int64_t>(M, 1), s)) != cudaSuccess ||
        (e = cudaMallocAsync(&queue, (2 * (size_t)M + 2) * 4, s)) != cudaSuccess)
        return fail(nullptr, CCB_ENOMEM, "scratch allocation: %s", cudaGetErrorString(e));
    int nl = 0;
    const int rc = launch_off_clusters(nullptr, s, (int)M, wnbr, core, submask_all, k > 1.0, pi, cls, queue, label, order, cl_off,
                                       n_cl, &nl, csr_min_m);
    cudaFreeAsync(cls, s);
    cudaFreeAsync(queue, s);
    return rc;
}

int ccb_offc_rowinfo(int32_t device, void *stream, const uint32_t *wnbr_rows, int64_t M, int64_t r0, int64_t r1, uint8_t *iso,
                     int32_t *nnz) {
    if (M < 0 || r0 < 0 || r1 < r0 || r1 > M) return fail(nullptr, CCB_EINVAL, "bad arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (r1 > r0)
        k_offc_rowinfo<<<(unsigned)((r1 - r0 + 3) / 4), 128, 0, (cudaStream_t)stream>>>(wnbr_rows, (int)r0, (int)r1,
                                                                                       (int)((M + 31) / 32), iso, nnz);
    e = cudaGetLastError();
    return e == cudaSuccess ? CCB_OK : fail(nullptr, CCB_ECUDA, "k_offc_rowinfo: %s", cudaGetErrorString(e));
}

int ccb_offc_fill(int32_t device, void *stream, const uint32_t *wnbr_rows, int64_t M, int64_t r0, int64_t r1,
                  const uint8_t *iso_all, const int64_t *off_all, int32_t *col) {
    if (M < 0 || r0 < 0 || r1 < r0 || r1 > M) return fail(nullptr, CCB_EINVAL, "bad arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    if (r1 > r0)
        k_offc_fill<<<(unsigned)((r1 - r0 + 3) / 4), 128, 0, (cudaStream_t)stream>>>(wnbr_rows, (int)r0, (int)r1,
                                                                                    (int)((M + 31) / 32), iso_all, off_all, col);
    e = cudaGetLastError();
    return e == cudaSuccess ? CCB_OK : fail(nullptr, CCB_ECUDA, "k_offc_fill: %s", cudaGetErrorString(e));
}

int ccb_off_clusters_csr(int32_t device, void *stream, int64_t M, const int64_t *off, const int32_t *col, const uint8_t *iso,
                         const uint8_t *core, const uint64_t *submask_all, double k, int64_t pi, int32_t *label, int32_t *order,
                         int32_t *cl_off, int32_t *n_cl) {
    if (M < 0 || M >= ((int64_t)1 << 31)) return fail(nullptr, CCB_EINVAL, "bad arguments");
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CCB_ECUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
    cudaStream_t s = (cudaStream_t)stream;
    tune_pool(device);
    if (M == 0) {
        cudaMemsetAsync(n_cl, 0, 4, s);
        cudaMemsetAsync(cl_off, 0, 4, s);
        return CCB_OK;
    }
    uint8_t *cls = nullptr;
    int32_t *queue = nullptr, *i32 = nullptr;
    const size_t m = (size_t)M;
    if ((e = cudaMallocAsync(&cls, m, s)) != cudaSuccess || (e = cudaMallocAsync(&queue, (2 * m + 2) * 4, s)) != cudaSuccess ||
        (e = cudaMallocAsync(&i32, (7 * m + 16) * 4, s)) != cudaSuccess)
        return fail(nullptr, CCB_ENOMEM, "scratch allocation: %s", cudaGetErrorString(e));
    cudaMemsetAsync(cls, 0, m, s);
    int64_t total = -1; // list lengths decide between the one-warp and the 1024-thread walk
    if (cudaMemcpyAsync(&total, off + M, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
        total = -1;
    offc_grow_and_merge(s, (int)M, off, col, core, iso, submask_all, k > 1.0, pi, cls, queue, i32, label, order, cl_off, n_cl,
                        total);
    e = cudaGetLastError();
    cudaFreeAsync(cls, s);
    cudaFreeAsync(queue, s);
    cudaFreeAsync(i32, s);
    return e == cudaSuccess ? CCB_OK : fail(nullptr, CCB_ECUDA, "offline cluster growth (CSR): %s", cudaGetErrorString(e));
}

} // extern "C"
