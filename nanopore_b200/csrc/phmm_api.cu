// phmm_api.cu -- host side of libphmm_sm100.so: the C ABI of include/phmm.h.
//
// Replaces the per-read process fan-out of the reference
// (nanopore/analyses/utils.py:557-609): guide cigars are turned into anchor
// runs and DP regions on the host (integer work, O(#cigar ops)), everything
// else runs in the kernels of phmm_kernels.cuh.  There is no CPU fallback and
// nothing here links or calls oracle/.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/phmm.h"
#include "phmm_fb2.cuh"
#include "phmm_decode_w.cuh"

using namespace phmm;

namespace {

// Both base arrays carry BASE_PAD bytes of N (code 4) either side: the unguarded cell loops of k_fb2 read the base before
// a region's first and after its last without a range test (the transitions those bases would emit start or end outside
// the matrix, i.e. at -inf, so the value does not matter; the address must be valid).
constexpr size_t BASE_PAD = 16;

thread_local std::string g_create_error;
thread_local int64_t g_held_bytes = 0;      // device bytes held by DevBufs of this thread's contexts

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    ~DevBuf() { release(); }
    void release() { if (p) { cudaFree(p); g_held_bytes -= (int64_t)cap; p = nullptr; cap = 0; } }
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { (void)cudaGetLastError(); e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) { cap = want; g_held_bytes += (int64_t)want; } else p = nullptr;
        return e;
    }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

struct BatchState {
    bool prepared = false, ran = false, expect = false;
    int64_t n_reads = 0;
    phmm_params params;
    DevParams dp;
    std::vector<Region> regions;
    std::vector<Run> runs;
    std::vector<RegionGeom> geom;
    std::vector<int32_t> order;
    std::vector<int64_t> read_first_region;   // n_reads + 1
    std::vector<int64_t> read_lx, read_ly;
    int nw = 1;
    int fb_slots = 0, dec_slots = 0;
    int32_t nd_stride = 0;                       // ints per decode slot in the per-diagonal arrays
    int64_t ring_cells = 0; int32_t dcap = 0, bw = 0;
    int32_t max_lx = 0, max_ly = 0, max_nd = 0, max_pairs = 0;
    int64_t total_pair_cap = 0, total_mrun_cap = 0;
    int pair_factor = 8;
    // windowed kernel (k_fb2)
    bool fast = false;
    std::vector<int64_t> tb_off;              // n_regions + 1
    std::vector<int64_t> rec_off;             // n_regions + 1: first diagonal record of each region
    int64_t ring_doubles = 0; int32_t wcap = 0, wg = 0, tcap = 0;
    int32_t ccap = 1;                            // posterior candidates per slot (k_fb2)
    int32_t cell_doubles = 1, total_extra = 5;   // ring doubles per cell / extra on total-probability diagonals
    size_t fb2_smem = 0;
    phmm_batch_stats stats;
};

}  // namespace

struct phmm_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;
    DevModel model;
    int64_t mem_budget = 0;
    bool force_legacy = false, decode_full_sweep = false, decode_block = false;
    int opt_warps = 0, opt_wcap = 0, opt_dbg = 0, opt_ccap = 0;
    double opt_est_eps = 0.02;             // tests: run the first-generation kernel (k_fwdbwd) instead of k_fb2
    DevBuf d_ref; int64_t ref_len = -1;
    DevBuf d_reads, d_regions, d_runs, d_geom, d_order, d_counter;
    DevBuf d_fring, d_dtab, d_bring, d_dots;
    DevBuf d_tboff, d_tbp, d_ring, d_wide, d_fsave, d_totals, d_recs, d_recoff, d_cand;
    DevBuf d_px, d_py, d_pw, d_npairs;
    DevBuf d_expT, d_expE, d_expLL;
    DevBuf d_sumx, d_sumy, d_dstart, d_dfill, d_sidx, d_wre, d_pred, d_colmap, d_sring, d_lring;
    DevBuf d_by, d_blo, d_bhi, d_fblist;         // k_decode_w: sorted y, envelope per diagonal, regions left to k_decode
    DevBuf d_mrx, d_mry, d_mrn, d_nmruns, d_score;
    DevBuf d_cx, d_cy, d_cn, d_coff;
    DevBuf d_baseexp, d_readmask; int64_t baseexp_len = -1; int32_t baseexp_tables = 0;   // tables of 5 x reference length sums of posterior mass per read base (A C G T other)
    BatchState b;
};

namespace {

int fail(phmm_ctx *ctx, int code, const std::string &msg) {
    ctx->err = msg;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            (void)cudaGetLastError();                                                              \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? PHMM_E_NOMEM : PHMM_E_CUDA,         \
                        std::string(#call) + ": " + cudaGetErrorString(e_));                       \
        }                                                                                          \
    } while (0)

const double NEG_INF = -INFINITY;

// stateMachine5 in log space
void build_model(DevModel &m, const double *trans, const double *emis) {
    memset(&m, 0, sizeof(m));
    if (!trans || !emis) {
        for (int i = 0; i < 25; i++) m.tr[i] = NEG_INF;
        m.tr[S_M * 5 + S_M] = -0.030064059121770816;
        m.tr[S_SX * 5 + S_M] = m.tr[S_SY * 5 + S_M] = -1.272871422049609;
        m.tr[S_LX * 5 + S_M] = m.tr[S_LY * 5 + S_M] = -5.673280173170473;
        m.tr[S_M * 5 + S_SX] = m.tr[S_M * 5 + S_SY] = -4.34381910900448;
        m.tr[S_SX * 5 + S_SX] = m.tr[S_SY * 5 + S_SY] = -0.3388262689231553;
        m.tr[S_SX * 5 + S_SY] = m.tr[S_SY * 5 + S_SX] = -4.910694825551255;
        m.tr[S_M * 5 + S_LX] = m.tr[S_M * 5 + S_LY] = -6.30810595366929;
        m.tr[S_LX * 5 + S_LX] = m.tr[S_LY * 5 + S_LY] = -0.003442492794189331;
        for (int x = 0; x < 4; x++) {
            for (int y = 0; y < 4; y++) {
                double v = -4.5691014376830479;                       // transversion
                if (x == y) v = -2.1149196655034745;
                else if ((x ^ y) == 2) v = -3.9833860032220842;       // A<->G, C<->T
                m.eM[x * 5 + y] = v;
            }
            m.eX[x] = m.eY[x] = -1.6094379124341003;
        }
    } else {
        for (int i = 0; i < 25; i++) m.tr[i] = log(trans[i]);
        for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++) m.eM[x * 5 + y] = log(emis[x * 4 + y]);
        // gap emissions: marginal of the 4x4 tables of both gap states of a kind, normalised
        double gx[4] = {0, 0, 0, 0}, gy[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) { gx[i] += emis[S_SX * 16 + i * 4 + j]; gx[i] += emis[S_LX * 16 + i * 4 + j]; }
        for (int i = 0; i < 4; i++)
            for (int j = 0; j < 4; j++) { gy[j] += emis[S_SY * 16 + i * 4 + j]; gy[j] += emis[S_LY * 16 + i * 4 + j]; }
        double tx = 0.0, ty = 0.0;
        for (int i = 0; i < 4; i++) { tx += gx[i]; ty += gy[i]; }
        for (int i = 0; i < 4; i++) { m.eX[i] = log(gx[i] / tx); m.eY[i] = log(gy[i] / ty); }
    }
    for (int i = 0; i < 5; i++) m.eM[4 * 5 + i] = m.eM[i * 5 + 4] = -2.772588722;   // N: log(1/16)
    m.eX[4] = m.eY[4] = -1.386294361;                                               // N: log(1/4)
    for (int s = 0; s < 5; s++) m.endp[s] = m.tr[s * 5 + S_M];
    m.rendp[S_M] = m.tr[S_M * 5 + S_LX];
    m.rendp[S_SX] = m.tr[S_M * 5 + S_LX];
    m.rendp[S_SY] = m.tr[S_M * 5 + S_LY];
    m.rendp[S_LX] = m.tr[S_LX * 5 + S_LX];
    m.rendp[S_LY] = m.tr[S_LY * 5 + S_LY];
    m.has_switch = (m.tr[S_SX * 5 + S_SY] != NEG_INF || m.tr[S_SY * 5 + S_SX] != NEG_INF) ? 1 : 0;
}

int check_params(phmm_ctx *ctx, const phmm_params *p) {
    if (!p) return fail(ctx, PHMM_E_ARG, "params is NULL");
    if (p->band < 0 || (p->band & 1)) return fail(ctx, PHMM_E_ARG, "band (diagonalExpansion) must be even and >= 0");
    if (p->anchor_trim < 0) return fail(ctx, PHMM_E_ARG, "anchor_trim must be >= 0");
    if (p->split_side < 1) return fail(ctx, PHMM_E_ARG, "split_side must be >= 1");
    if (p->tb_diags < 1 || p->tb_diags + 2 > 256) return fail(ctx, PHMM_E_ARG, "tb_diags must be in [1, 254]");
    if (p->min_diags < 2 || p->tb_diags + 1 >= p->min_diags)
        return fail(ctx, PHMM_E_ARG, "need min_diags >= 2 and tb_diags + 1 < min_diags");
    if (!(p->threshold >= 0.0 && p->threshold <= 1.0)) return fail(ctx, PHMM_E_ARG, "threshold must be in [0,1]");
    return PHMM_OK;
}

// Splits one anchor-free block if its area exceeds side^2 (getSplitPoints).
struct SplitState { int64_t x1 = 0, y1 = 0; std::vector<int64_t> quads; };

void split_block(SplitState &s, int64_t x2, int64_t y2, int64_t x3, int64_t y3, int64_t side) {
    const int64_t bx = x3 - x2, by = y3 - y2;
    if (bx * by > side * side) {
        const int64_t hx = std::min(bx / 2, side), hy = std::min(by / 2, side);
        s.quads.insert(s.quads.end(), {s.x1, s.y1, x2 + hx, y2 + hy});
        s.x1 = x3 - hx; s.y1 = y3 - hy;
    }
}

// Host planning of one read: guide cigar -> anchor runs -> regions.
int plan_read(phmm_ctx *ctx, int64_t read, const uint32_t *ops, int64_t nops, int64_t lX, int64_t lY,
              int64_t ref_abs, int64_t read_abs, const phmm_params &p, std::vector<Region> &regions, std::vector<Run> &runs) {
    struct ARun { int64_t x, y, n; };
    std::vector<ARun> ar;
    int64_t x = 0, y = 0;
    for (int64_t i = 0; i < nops; i++) {
        const int64_t len = ops[i] >> 2; const int code = ops[i] & 3;
        if (code == 0) {
            if (len > 2 * (int64_t)p.anchor_trim) ar.push_back({x + p.anchor_trim, y + p.anchor_trim, len - 2 * p.anchor_trim});
            x += len; y += len;
        } else if (code == 1) y += len;
        else if (code == 2) x += len;
        else return fail(ctx, PHMM_E_ARG, "cigar op code 3 is not M/I/D (read " + std::to_string(read) + ")");
    }
    // region origins, run offsets and posterior coordinates are int32 (Region::x1/y1, phmm_posteriors::ref_pos)
    if (lX > 0x7ffffff0 || lY > 0x7ffffff0)
        return fail(ctx, PHMM_E_ARG, "read " + std::to_string(read) + ": windows longer than 2^31 - 16 bases are not supported");
    if (x != lX || y != lY)
        return fail(ctx, PHMM_E_ARG, "guide cigar of read " + std::to_string(read) + " spans " + std::to_string(x) + "x" +
                                         std::to_string(y) + " but the sequences are " + std::to_string(lX) + "x" + std::to_string(lY));
    // strictly increasing anchors are implied by a single cigar; merge touching runs
    SplitState sp;
    int64_t x2 = 0, y2 = 0;
    for (const ARun &r : ar) {
        split_block(sp, x2, y2, r.x, r.y, p.split_side);
        x2 = r.x + r.n; y2 = r.y + r.n;
    }
    split_block(sp, x2, y2, lX, lY, p.split_side);
    sp.quads.insert(sp.quads.end(), {sp.x1, sp.y1, lX, lY});
    const size_t nreg = sp.quads.size() / 4;
    size_t j = 0;
    for (size_t i = 0; i < nreg; i++) {
        Region g;
        memset(&g, 0, sizeof(g));
        const int64_t x1 = sp.quads[4 * i], y1 = sp.quads[4 * i + 1], xe = sp.quads[4 * i + 2], ye = sp.quads[4 * i + 3];
        if (xe - x1 > 0x3fffffff || ye - y1 > 0x3fffffff) return fail(ctx, PHMM_E_ARG, "region too large");
        g.xoff = ref_abs + x1; g.yoff = read_abs + y1;
        g.lx = (int32_t)(xe - x1); g.ly = (int32_t)(ye - y1);
        g.read = (int32_t)read; g.x1 = (int32_t)x1; g.y1 = (int32_t)y1;
        g.ragged_left = i > 0; g.ragged_right = i + 1 < nreg;
        g.run0 = (int32_t)runs.size();
        while (j < ar.size() && ar[j].x + ar[j].y < xe + ye) {
            const ARun &r = ar[j];
            if (r.x < x1 || r.y < y1 || r.x + r.n > xe || r.y + r.n > ye)
                return fail(ctx, PHMM_E_ARG, "anchor run crosses a region boundary (read " + std::to_string(read) + ")");
            runs.push_back({(int32_t)(r.x - x1), (int32_t)(r.y - y1), (int32_t)r.n});
            j++;
        }
        g.nrun = (int32_t)runs.size() - g.run0;
        regions.push_back(g);
    }
    if (j != ar.size()) return fail(ctx, PHMM_E_ARG, "unassigned anchors (read " + std::to_string(read) + ")");
    return PHMM_OK;
}

template <int NW>
int launch_fwdbwd(phmm_ctx *ctx, const FbArgs &a, int slots, bool expect) {
    const bool sw = ctx->model.has_switch != 0;
    if (expect) {
        if (sw) k_fwdbwd<NW, true, true><<<slots, NW * 32, 0, ctx->stream>>>(a);
        else k_fwdbwd<NW, false, true><<<slots, NW * 32, 0, ctx->stream>>>(a);
    } else {
        if (sw) k_fwdbwd<NW, true, false><<<slots, NW * 32, 0, ctx->stream>>>(a);
        else k_fwdbwd<NW, false, false><<<slots, NW * 32, 0, ctx->stream>>>(a);
    }
    CK(cudaGetLastError());
    return PHMM_OK;
}

template <int NW>
int occupancy_fwdbwd(bool sw, bool expect) {
    int n = 0;
    if (expect) {
        if (sw) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_fwdbwd<NW, true, true>, NW * 32, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_fwdbwd<NW, false, true>, NW * 32, 0);
    } else {
        if (sw) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_fwdbwd<NW, true, false>, NW * 32, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_fwdbwd<NW, false, false>, NW * 32, 0);
    }
    return n > 0 ? n : 1;
}

template <int NW, bool SW, bool EX>
int fb2_occupancy_t(size_t smem) {
    int n = 0;
    if (cudaFuncSetAttribute(k_fb2<NW, SW, EX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_fb2<NW, SW, EX>, NW * 32, smem) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return n;
}

template <int NW, bool SW, bool EX>
int fb2_launch_t(const Fb2Args &a, int slots, size_t smem, cudaStream_t st) {
    k_fb2<NW, SW, EX><<<slots, NW * 32, smem, st>>>(a);
    return 0;
}

#define FB2_DISPATCH(FN, ...)                                                                        \
    do {                                                                                             \
        if (nw == 2) {                                                                               \
            if (ex) return sw ? FN<2, true, true>(__VA_ARGS__) : FN<2, false, true>(__VA_ARGS__);    \
            return sw ? FN<2, true, false>(__VA_ARGS__) : FN<2, false, false>(__VA_ARGS__);          \
        }                                                                                            \
        if (nw == 4) {                                                                               \
            if (ex) return sw ? FN<PHMM_NW4, true, true>(__VA_ARGS__) : FN<PHMM_NW4, false, true>(__VA_ARGS__);    \
            return sw ? FN<PHMM_NW4, true, false>(__VA_ARGS__) : FN<PHMM_NW4, false, false>(__VA_ARGS__);          \
        }                                                                                            \
        if (ex) return sw ? FN<8, true, true>(__VA_ARGS__) : FN<8, false, true>(__VA_ARGS__);        \
        return sw ? FN<8, true, false>(__VA_ARGS__) : FN<8, false, false>(__VA_ARGS__);              \
    } while (0)

int fb2_occupancy(int nw, bool sw, bool ex, size_t smem) { FB2_DISPATCH(fb2_occupancy_t, smem); }
int fb2_launch(int nw, bool sw, bool ex, const Fb2Args &a, int slots, size_t smem, cudaStream_t st) { FB2_DISPATCH(fb2_launch_t, a, slots, smem, st); }

int32_t pow2_at_least(int32_t v) { int32_t p = 1; while (p < v) p <<= 1; return p; }

int64_t budget(phmm_ctx *ctx) {
    if (ctx->mem_budget > 0) return ctx->mem_budget;
    size_t fr = 0, tot = 0;
    cudaMemGetInfo(&fr, &tot);
    // what is free now plus what this library already holds and will reuse
    return (int64_t)(((double)fr + (double)g_held_bytes) * 0.8);
}

void release_scratch(phmm_ctx *ctx) {
    DevBuf *all[] = {&ctx->d_fring, &ctx->d_dtab, &ctx->d_bring, &ctx->d_dots, &ctx->d_ring, &ctx->d_wide, &ctx->d_fsave, &ctx->d_totals,
                     &ctx->d_recs, &ctx->d_recoff, &ctx->d_cand, &ctx->d_px, &ctx->d_py, &ctx->d_pw, &ctx->d_npairs, &ctx->d_expT, &ctx->d_expE,
                     &ctx->d_expLL, &ctx->d_sumx, &ctx->d_sumy, &ctx->d_dstart, &ctx->d_dfill, &ctx->d_sidx, &ctx->d_wre, &ctx->d_pred,
                     &ctx->d_colmap, &ctx->d_sring, &ctx->d_lring, &ctx->d_by, &ctx->d_blo, &ctx->d_bhi, &ctx->d_fblist, &ctx->d_mrx, &ctx->d_mry, &ctx->d_mrn, &ctx->d_nmruns, &ctx->d_score,
                     &ctx->d_cx, &ctx->d_cy, &ctx->d_cn, &ctx->d_coff};
    for (DevBuf *d : all) d->release();
}

// Sizes scratch and (re)allocates it.  pair capacities scale with pair_factor.
int plan_memory(phmm_ctx *ctx) {
    BatchState &b = ctx->b;
    const int64_t nreg = (int64_t)b.regions.size();
    int64_t cells = 0, diags = 0;
    b.ring_cells = 2; b.dcap = 4; b.bw = 1; b.max_lx = 1; b.max_ly = 1; b.max_nd = 1; b.max_pairs = 1;
    b.total_pair_cap = 0; b.total_mrun_cap = 0;
    for (int64_t i = 0; i < nreg; i++) {
        Region &r = b.regions[i];
        const RegionGeom &g = b.geom[i];
        cells += g.cells; diags += g.diagonals;
        b.ring_cells = std::max<int64_t>(b.ring_cells, g.max_live_cells + g.max_width + 2);
        b.dcap = std::max<int32_t>(b.dcap, g.max_live_diags + 4);
        b.bw = std::max<int32_t>(b.bw, g.max_width);
        b.max_lx = std::max(b.max_lx, r.lx); b.max_ly = std::max(b.max_ly, r.ly);
        b.max_nd = std::max(b.max_nd, r.lx + r.ly);
        const int64_t mn = std::min(r.lx, r.ly);
        int64_t pc = (int64_t)b.pair_factor * mn + 1024;
        pc = std::min<int64_t>(pc, g.cells);                 // never more pairs than cells
        if (pc > 0x7fffffff) return fail(ctx, PHMM_E_ARG, "region too large");
        r.pair_cap = (int32_t)pc; r.pair_off = b.total_pair_cap; b.total_pair_cap += pc;
        r.mrun_cap = (int32_t)(mn + 1); r.mrun_off = b.total_mrun_cap; b.total_mrun_cap += mn + 1;
        b.max_pairs = std::max<int32_t>(b.max_pairs, r.pair_cap);
    }
    if (b.ring_cells > 0x7ffffff0) return fail(ctx, PHMM_E_ARG, "forward window too large for one region");
    b.stats.cells = cells; b.stats.diagonals = diags; b.stats.n_regions = nreg; b.stats.n_reads = b.n_reads;
    // width class -> warps per region
    const double avgw = diags > 0 ? (double)cells / (double)diags : 1.0;
    const bool sw = ctx->model.has_switch != 0;
    int64_t max_live_doubles = 0;
    for (int64_t i = 0; i < nreg; i++) max_live_doubles = std::max(max_live_doubles, b.geom[i].max_live_doubles);
    // the windowed kernel needs a total-probability schedule that looks one traceback point ahead
    // E-step layout: 10 doubles per live cell, 11 where a total is evaluated (all forward and backward values)
    b.cell_doubles = b.expect ? 10 : 1; b.total_extra = b.expect ? 1 : 5;
    const int64_t live_need = b.expect ? (b.ring_cells + 2) * 11 : max_live_doubles;
    b.fast = !ctx->force_legacy && b.params.min_diags >= 2 * (b.params.tb_diags + 1) + 2 &&
             live_need + 4 * 11 * (int64_t)b.bw + 16 < (b.expect ? 0x08000000 : 0x7ffffff0);   // E-step: up to 1 GiB of ring per region
    int occ = 1;
    int64_t slot_bytes = 0;
    if (b.fast) {
        // measured on B200 (profiles/r01_tune_warps.txt): 4 compute warps x 4 resident regions per SM beat 8 x 2 even at
        // mean width 158 (fewer warps idle at the per-diagonal barrier); 8 only pays for very wide bands
        b.nw = ctx->opt_warps ? ctx->opt_warps : (avgw <= 48.0 ? 2 : (avgw <= 320.0 ? 4 : 8));
        b.wg = pow2_at_least(b.bw);
        b.wcap = ctx->opt_wcap ? ctx->opt_wcap : std::max<int32_t>(64, std::min<int32_t>(512, b.wg));
        // bump allocation with wrap-around wastes at most one diagonal's worth at the end of the ring
        b.ring_doubles = live_need + 4 * 11 * (int64_t)b.bw + 16;
        b.dcap += 4;
        b.tcap = b.dcap / TOTAL_EVERY + 4;
        // one 16-byte record per diagonal of every region (k_records)
        b.rec_off.assign(nreg + 1, 0);
        for (int64_t i = 0; i < nreg; i++) b.rec_off[i + 1] = b.rec_off[i] + (int64_t)b.regions[i].lx + b.regions[i].ly + 1;
        b.fb2_smem = (size_t)2 * CS * b.wcap * 8;      // the diagonal buffers; tables and record FIFOs are static
        occ = fb2_occupancy(b.nw, sw, b.expect, b.fb2_smem);
        if (occ < 1) return fail(ctx, PHMM_E_CUDA, "k_fb2 does not fit on this device");
        slot_bytes = b.ring_doubles * 8 + (int64_t)4 * NS * b.wg * 8 +
                     (int64_t)2 * CS * b.wcap * 8 + ((int64_t)b.tcap + b.wg) * 8 + (b.expect ? 8 : (ctx->opt_ccap > 0 ? (int64_t)ctx->opt_ccap : 2 * (int64_t)b.max_pairs + 1024) * 8);
    } else {
        b.nw = avgw <= 40.0 ? 1 : (avgw <= 96.0 ? 2 : 4);
        occ = b.nw == 1 ? occupancy_fwdbwd<1>(sw, b.expect) : b.nw == 2 ? occupancy_fwdbwd<2>(sw, b.expect) : occupancy_fwdbwd<4>(sw, b.expect);
        slot_bytes = b.ring_cells * NS * 8 + (int64_t)b.dcap * sizeof(DiagRec) + (int64_t)3 * b.bw * NS * 8 + (int64_t)2 * b.bw * 8;
    }
    int64_t want = (int64_t)ctx->sm_count * occ;
    // fixed allocations
    const int64_t fixed = b.total_pair_cap * 12 + b.total_mrun_cap * 12 + nreg * (sizeof(Region) + sizeof(RegionGeom) + 64) +
                          (b.fast ? b.rec_off[nreg] * (int64_t)sizeof(DiagRec) + nreg * 8 : 0);
    // decode slots serve both kernels: k_decode_w (one warp per region, DW_BLOCKS per SM) and k_decode for the regions it
    // leaves; they share sumx, sumy, dstart, dfill / nxt, sidx / bx, wre / bwr and pred
    b.nd_stride = (b.max_nd + 4 + 3) & ~3;
    const int64_t dec_slot_bytes = (int64_t)(b.max_lx + 1) * 4 + (int64_t)(b.max_ly + 1) * 4 + (int64_t)b.nd_stride * 16 +
                                   (int64_t)(b.max_pairs + 1) * 20 + (int64_t)2 * (b.max_lx + 2) * 8 + (int64_t)4 * b.bw * 12;
    int64_t avail = budget(ctx) - fixed;
    if (avail < slot_bytes + dec_slot_bytes) return fail(ctx, PHMM_E_NOMEM, "memory budget too small for one region of this batch");
    // measured (profiles/r01b_phase_breakdown.txt, tune15): 2-warp decode blocks, 16 per SM, beat 4 warps x 8
    int64_t dec_want = (int64_t)ctx->sm_count * std::max(DW_BLOCKS, b.nw >= 2 ? 16 : 8);
    dec_want = std::min<int64_t>(dec_want, nreg);
    dec_want = std::max<int64_t>(1, std::min<int64_t>(dec_want, (avail / 4) / dec_slot_bytes));
    avail -= dec_want * dec_slot_bytes;
    want = std::min<int64_t>(want, nreg);
    want = std::max<int64_t>(1, std::min<int64_t>(want, avail / slot_bytes));
    b.fb_slots = (int)want; b.dec_slots = (int)dec_want;
    b.stats.slot_bytes = slot_bytes; b.stats.n_slots = want;

    // Scratch of an earlier batch that is larger than this plan needs stays allocated (DevBuf only grows), so the sum can
    // exceed the budget when consecutive batches differ a lot (mixed read lengths): on the first failed allocation every
    // scratch buffer is released and the plan is allocated afresh.
    auto alloc_all = [&]() -> cudaError_t {
#define TRY(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { (void)cudaGetLastError(); return e_; } } while (0)
        if (b.fast) {
            ctx->d_fring.release(); ctx->d_bring.release(); ctx->d_dots.release(); ctx->d_dtab.release();
            TRY(ctx->d_recs.ensure((size_t)b.rec_off[nreg] * sizeof(DiagRec) + 64));
            TRY(ctx->d_recoff.ensure((size_t)(nreg + 1) * 8));
            TRY(ctx->d_ring.ensure((size_t)want * b.ring_doubles * 8));
            TRY(ctx->d_wide.ensure((size_t)want * 4 * NS * b.wg * 8));
            TRY(ctx->d_fsave.ensure((size_t)want * 2 * CS * b.wcap * 8));
            TRY(ctx->d_totals.ensure((size_t)want * ((size_t)b.tcap + b.wg) * 8));
            b.ccap = b.expect ? 1 : (ctx->opt_ccap > 0 ? ctx->opt_ccap : 2 * b.max_pairs + 1024);
            TRY(ctx->d_cand.ensure((size_t)want * (size_t)b.ccap * 8));
        } else {
            ctx->d_ring.release(); ctx->d_wide.release(); ctx->d_recs.release();
            TRY(ctx->d_dtab.ensure((size_t)want * b.dcap * sizeof(DiagRec)));
            TRY(ctx->d_fring.ensure((size_t)want * b.ring_cells * NS * 8));
            TRY(ctx->d_bring.ensure((size_t)want * 3 * b.bw * NS * 8));
            TRY(ctx->d_dots.ensure((size_t)want * 2 * b.bw * 8));
        }
        TRY(ctx->d_px.ensure((size_t)b.total_pair_cap * 4 + 16));
        TRY(ctx->d_py.ensure((size_t)b.total_pair_cap * 4 + 16));
        TRY(ctx->d_pw.ensure((size_t)b.total_pair_cap * 4 + 16));
        TRY(ctx->d_npairs.ensure((size_t)nreg * 4 + 16));
        if (b.expect) {
            TRY(ctx->d_expT.ensure((size_t)nreg * 25 * 8));
            TRY(ctx->d_expE.ensure((size_t)nreg * 80 * 8));
            TRY(ctx->d_expLL.ensure((size_t)nreg * 8));
        } else {
            TRY(ctx->d_sumx.ensure((size_t)dec_want * (b.max_lx + 1) * 4));
            TRY(ctx->d_sumy.ensure((size_t)dec_want * (b.max_ly + 1) * 4));
            TRY(ctx->d_dstart.ensure((size_t)dec_want * b.nd_stride * 4));
            TRY(ctx->d_dfill.ensure((size_t)dec_want * b.nd_stride * 4));
            TRY(ctx->d_blo.ensure((size_t)dec_want * b.nd_stride * 4));
            TRY(ctx->d_bhi.ensure((size_t)dec_want * b.nd_stride * 4));
            TRY(ctx->d_by.ensure((size_t)dec_want * (b.max_pairs + 1) * 4));
            TRY(ctx->d_fblist.ensure((size_t)nreg * 4 + 16));
            TRY(ctx->d_sidx.ensure((size_t)dec_want * (b.max_pairs + 1) * 4));
            TRY(ctx->d_wre.ensure((size_t)dec_want * (b.max_pairs + 1) * 8));
            TRY(ctx->d_pred.ensure((size_t)dec_want * (b.max_pairs + 1) * 4));
            TRY(ctx->d_colmap.ensure((size_t)dec_want * 2 * (b.max_lx + 2) * 8));
            TRY(ctx->d_sring.ensure((size_t)dec_want * 4 * b.bw * 8));
            TRY(ctx->d_lring.ensure((size_t)dec_want * 4 * b.bw * 4));
            TRY(ctx->d_mrx.ensure((size_t)b.total_mrun_cap * 4 + 16));
            TRY(ctx->d_mry.ensure((size_t)b.total_mrun_cap * 4 + 16));
            TRY(ctx->d_mrn.ensure((size_t)b.total_mrun_cap * 4 + 16));
            TRY(ctx->d_nmruns.ensure((size_t)nreg * 4 + 16));
            TRY(ctx->d_score.ensure((size_t)nreg * 8 + 16));
        }
        return cudaSuccess;
#undef TRY
    };
    if (alloc_all() != cudaSuccess) {
        release_scratch(ctx);
        b.ccap = 1;
        cudaError_t e2 = alloc_all();
        if (e2 != cudaSuccess)
            return fail(ctx, e2 == cudaErrorMemoryAllocation ? PHMM_E_NOMEM : PHMM_E_CUDA, std::string("scratch allocation: ") + cudaGetErrorString(e2));
    }
    // regions carry the plan (pair_off, caps): upload
    CK(cudaMemcpyAsync(ctx->d_regions.p, b.regions.data(), nreg * sizeof(Region), cudaMemcpyHostToDevice, ctx->stream));
    if (b.fast) {
        // diagonal records of every region (data independent; once per prepared batch)
        CK(cudaMemcpyAsync(ctx->d_recoff.p, b.rec_off.data(), (size_t)(nreg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        k_records<<<(unsigned)((nreg + 63) / 64), 64, 0, ctx->stream>>>(
            ctx->d_regions.as<Region>(), ctx->d_runs.as<Run>(), (int)nreg, b.dp, ctx->d_tboff.as<int64_t>(), ctx->d_tbp.as<int32_t>(),
            reinterpret_cast<const int32_t *>(ctx->d_geom.as<char>() + offsetof(RegionGeom, tracebacks)), (int)(sizeof(RegionGeom) / 4),
            b.ring_doubles, b.wcap, b.cell_doubles, b.total_extra, ctx->d_recoff.as<int64_t>(), ctx->d_recs.as<DiagRec>());
        CK(cudaGetLastError());
    }
    return PHMM_OK;
}

int do_prepare(phmm_ctx *ctx, int64_t n_reads, const uint8_t *read_bases, const int64_t *read_off,
               const int64_t *ref_start, const int64_t *ref_end, const uint32_t *in_ops, const int64_t *in_off,
               const phmm_params *params, bool expect) {
    BatchState &b = ctx->b;
    b.prepared = false; b.ran = false; b.expect = expect;
    int rc = check_params(ctx, params);
    if (rc) return rc;
    if (ctx->ref_len < 0) return fail(ctx, PHMM_E_STATE, "phmm_set_reference has not been called");
    if (n_reads < 0) return fail(ctx, PHMM_E_ARG, "n_reads < 0");
    if (n_reads > 0 && (!read_off || !ref_start || !ref_end || !in_off)) return fail(ctx, PHMM_E_ARG, "NULL input array");
    CK(cudaSetDevice(ctx->device));
    b.params = *params;
    b.n_reads = n_reads;
    b.pair_factor = 8;
    memset(&b.stats, 0, sizeof(b.stats));
    b.regions.clear(); b.runs.clear(); b.geom.clear(); b.order.clear();
    b.read_first_region.assign(n_reads + 1, 0);
    b.read_lx.assign(n_reads, 0); b.read_ly.assign(n_reads, 0);
    for (int64_t i = 0; i < n_reads; i++) {
        const int64_t lY = read_off[i + 1] - read_off[i];
        const int64_t lX = ref_end[i] - ref_start[i];
        if (lY < 0 || lX < 0 || ref_start[i] < 0 || ref_end[i] > ctx->ref_len)
            return fail(ctx, PHMM_E_ARG, "read " + std::to_string(i) + ": coordinates outside the reference or negative length");
        if (lX > 0x3fffffff || lY > 0x3fffffff)      // region origins and posterior positions are 32-bit
            return fail(ctx, PHMM_E_ARG, "read " + std::to_string(i) + ": window or read longer than 2^30 - 1 bases");
        b.read_first_region[i] = (int64_t)b.regions.size();
        b.read_lx[i] = lX; b.read_ly[i] = lY;
        rc = plan_read(ctx, i, in_ops + in_off[i], in_off[i + 1] - in_off[i], lX, lY, ref_start[i], read_off[i], *params, b.regions, b.runs);
        if (rc) return rc;
    }
    b.read_first_region[n_reads] = (int64_t)b.regions.size();
    const int64_t nreg = (int64_t)b.regions.size();
    if (nreg > 0x7ffffff0) return fail(ctx, PHMM_E_ARG, "too many regions in one batch");
    b.dp.expansion = params->band; b.dp.min_diags = params->min_diags; b.dp.tb_diags = params->tb_diags; b.dp.pad = 0;
    b.dp.threshold = params->threshold;
    b.dp.lp_skip = params->threshold > 0.0 ? log(params->threshold) - 1e-3 : -INFINITY;
    b.dp.gap_gamma = params->gap_gamma; b.dp.match_gamma = params->match_gamma;
    if (nreg == 0) { b.prepared = true; return PHMM_OK; }

    const int64_t total_read = n_reads ? read_off[n_reads] : 0;
    CK(ctx->d_reads.ensure((size_t)total_read + 2 * BASE_PAD));
    CK(ctx->d_regions.ensure((size_t)nreg * sizeof(Region)));
    CK(ctx->d_runs.ensure((size_t)(b.runs.size() + 1) * sizeof(Run)));
    CK(ctx->d_geom.ensure((size_t)nreg * sizeof(RegionGeom)));
    CK(ctx->d_order.ensure((size_t)nreg * 4));
    CK(ctx->d_counter.ensure(64));
    // upper bound of traceback points per region: consecutive points are >= min_diags - tb_diags - 1 apart
    b.tb_off.assign(nreg + 1, 0);
    {
        const int64_t gap = std::max<int64_t>(1, (int64_t)params->min_diags - params->tb_diags - 1);
        for (int64_t i = 0; i < nreg; i++) b.tb_off[i + 1] = b.tb_off[i] + ((int64_t)b.regions[i].lx + b.regions[i].ly) / gap + 2;
    }
    CK(ctx->d_tboff.ensure((size_t)(nreg + 1) * 8));
    CK(ctx->d_tbp.ensure((size_t)b.tb_off[nreg] * 4 + 16));
    CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_tboff.p, b.tb_off.data(), (size_t)(nreg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_reads.p, 4, (size_t)total_read + 2 * BASE_PAD, ctx->stream));
    if (total_read) CK(cudaMemcpyAsync(ctx->d_reads.as<uint8_t>() + BASE_PAD, read_bases, (size_t)total_read, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_regions.p, b.regions.data(), nreg * sizeof(Region), cudaMemcpyHostToDevice, ctx->stream));
    if (!b.runs.empty()) CK(cudaMemcpyAsync(ctx->d_runs.p, b.runs.data(), b.runs.size() * sizeof(Run), cudaMemcpyHostToDevice, ctx->stream));
    k_geometry<<<(unsigned)((nreg + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_regions.as<Region>(), ctx->d_runs.as<Run>(), (int)nreg, b.dp,
                                                                         ctx->d_geom.as<RegionGeom>(), ctx->d_tboff.as<int64_t>(),
                                                                         ctx->d_tbp.as<int32_t>());
    CK(cudaGetLastError());
    CK(cudaEventRecord(ctx->ev[1], ctx->stream));
    b.stats.h2d_bytes = total_read + nreg * (int64_t)sizeof(Region) * 2 + (int64_t)(b.runs.size() * sizeof(Run)) + nreg * 4;
    b.stats.d2h_bytes = nreg * (int64_t)sizeof(RegionGeom);
    b.geom.resize(nreg);
    CK(cudaMemcpyAsync(b.geom.data(), ctx->d_geom.p, nreg * sizeof(RegionGeom), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]);
    b.stats.ms_geometry = ms;
    // longest regions first
    b.order.resize(nreg);
    std::iota(b.order.begin(), b.order.end(), 0);
    std::stable_sort(b.order.begin(), b.order.end(), [&](int32_t x, int32_t y) { return b.geom[x].cells > b.geom[y].cells; });
    CK(cudaMemcpyAsync(ctx->d_order.p, b.order.data(), nreg * 4, cudaMemcpyHostToDevice, ctx->stream));
    rc = plan_memory(ctx);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    b.prepared = true;
    return PHMM_OK;
}

int do_run(phmm_ctx *ctx) {
    BatchState &b = ctx->b;
    if (!b.prepared) return fail(ctx, PHMM_E_STATE, "no batch prepared");
    const int64_t nreg = (int64_t)b.regions.size();
    b.stats.launches = b.fast ? 2 : 1;   // geometry (+ diagonal records)
    b.stats.run_launches = 0;
    if (nreg == 0) { b.ran = true; return PHMM_OK; }
    CK(cudaSetDevice(ctx->device));
    FbArgs fa;
    memset(&fa, 0, sizeof(fa));
    fa.ref = ctx->d_ref.as<uint8_t>() + BASE_PAD; fa.reads = ctx->d_reads.as<uint8_t>() + BASE_PAD;
    fa.regions = ctx->d_regions.as<Region>(); fa.runs = ctx->d_runs.as<Run>(); fa.order = ctx->d_order.as<int32_t>();
    fa.n_regions = (int32_t)nreg; fa.counter = ctx->d_counter.as<int32_t>();
    fa.m = ctx->model; fa.p = b.dp;
    fa.fring = ctx->d_fring.as<double>(); fa.ring_cells = b.ring_cells;
    fa.dtab = ctx->d_dtab.as<DiagRec>(); fa.dcap = b.dcap;
    fa.bring = ctx->d_bring.as<double>(); fa.bw = b.bw; fa.dots = ctx->d_dots.as<double>();
    fa.px = ctx->d_px.as<int32_t>(); fa.py = ctx->d_py.as<int32_t>(); fa.pw = ctx->d_pw.as<int32_t>();
    fa.npairs = ctx->d_npairs.as<int32_t>();
    fa.expT = ctx->d_expT.as<unsigned long long>(); fa.expE = ctx->d_expE.as<unsigned long long>(); fa.expLL = ctx->d_expLL.as<double>();
    CK(cudaMemsetAsync(ctx->d_counter.p, 0, 64, ctx->stream));
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    int rc = PHMM_OK;
    if (b.fast) {
        Fb2Args f2;
        memset(&f2, 0, sizeof(f2));
        f2.ref = fa.ref; f2.reads = fa.reads; f2.regions = fa.regions; f2.runs = fa.runs; f2.order = fa.order;
        f2.n_regions = fa.n_regions; f2.counter = fa.counter; f2.m = ctx->model; f2.p = b.dp;
        f2.tb_off = ctx->d_tboff.as<int64_t>(); f2.tbp = ctx->d_tbp.as<int32_t>();
        f2.ntb = reinterpret_cast<const int32_t *>(ctx->d_geom.as<char>() + offsetof(RegionGeom, tracebacks));
        f2.ntb_stride = (int32_t)(sizeof(RegionGeom) / 4);
        f2.ring = ctx->d_ring.as<double>(); f2.ring_doubles = b.ring_doubles;
        f2.recs = ctx->d_recs.as<DiagRec>(); f2.rec_off = ctx->d_recoff.as<int64_t>();
        f2.wide = ctx->d_wide.as<double>(); f2.wg = b.wg;
        f2.fsave = ctx->d_fsave.as<double>(); f2.totals = ctx->d_totals.as<double>(); f2.tcap = b.tcap;
        f2.wcap = b.wcap;
        f2.cand = ctx->d_cand.as<long long>(); f2.ccap = b.ccap; f2.est_eps = ctx->opt_est_eps;
        f2.dbg = ctx->opt_dbg;
        f2.px = fa.px; f2.py = fa.py; f2.pw = fa.pw; f2.npairs = fa.npairs;
        f2.expT = fa.expT; f2.expE = fa.expE; f2.expLL = fa.expLL;
        fb2_launch(b.nw, ctx->model.has_switch != 0, b.expect, f2, b.fb_slots, b.fb2_smem, ctx->stream);
        CK(cudaGetLastError());
    } else {
        rc = b.nw == 1 ? launch_fwdbwd<1>(ctx, fa, b.fb_slots, b.expect)
           : b.nw == 2 ? launch_fwdbwd<2>(ctx, fa, b.fb_slots, b.expect)
                       : launch_fwdbwd<4>(ctx, fa, b.fb_slots, b.expect);
    }
    if (rc) return rc;
    b.stats.launches++; b.stats.run_launches++;
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    if (!b.expect) {
        DecArgs da;
        memset(&da, 0, sizeof(da));
        da.regions = fa.regions; da.runs = fa.runs; da.order = fa.order; da.n_regions = fa.n_regions;
        da.counter = ctx->d_counter.as<int32_t>() + 8;
        da.p = b.dp;
        da.px = fa.px; da.py = fa.py; da.pw = fa.pw; da.npairs = fa.npairs;
        da.sumx = ctx->d_sumx.as<int32_t>(); da.sumy = ctx->d_sumy.as<int32_t>(); da.max_lx = b.max_lx; da.max_ly = b.max_ly;
        da.dstart = ctx->d_dstart.as<int32_t>(); da.dfill = ctx->d_dfill.as<int32_t>(); da.max_nd = b.max_nd;
        da.sidx = ctx->d_sidx.as<int32_t>(); da.wre = ctx->d_wre.as<int64_t>(); da.pred = ctx->d_pred.as<int32_t>(); da.max_pairs = b.max_pairs;
        da.colmap = ctx->d_colmap.as<int64_t>();
        da.sring = ctx->d_sring.as<int64_t>(); da.lring = ctx->d_lring.as<int32_t>(); da.bw = b.bw;
        da.mrx = ctx->d_mrx.as<int32_t>(); da.mry = ctx->d_mry.as<int32_t>(); da.mrn = ctx->d_mrn.as<int32_t>();
        da.nmruns = ctx->d_nmruns.as<int32_t>(); da.score = ctx->d_score.as<int64_t>();
        da.regular = reinterpret_cast<const int32_t *>(ctx->d_geom.as<char>() + offsetof(RegionGeom, regular));
        da.regular_stride = ctx->decode_full_sweep ? 0 : (int32_t)(sizeof(RegionGeom) / 4);
        if (ctx->decode_full_sweep) da.regular = ctx->d_counter.as<int32_t>() + 4;      // a zero: no region takes the shortcut
        const bool block_only = ctx->decode_full_sweep || ctx->decode_block;
        if (!block_only) {
            // the envelope kernel first; what it cannot take (irregular band, very wide envelope) is left in d_fblist
            DecWArgs dw;
            memset(&dw, 0, sizeof(dw));
            dw.regions = da.regions; dw.order = da.order; dw.n_regions = da.n_regions;
            dw.counter = ctx->d_counter.as<int32_t>() + 9;
            dw.p = b.dp;
            dw.px = da.px; dw.py = da.py; dw.pw = da.pw; dw.npairs = da.npairs;
            dw.regular = da.regular; dw.regular_stride = da.regular_stride;
            dw.sumx = da.sumx; dw.sumy = da.sumy; dw.max_lx = b.max_lx; dw.max_ly = b.max_ly;
            dw.dstart = da.dstart; dw.nxt = da.dfill; dw.blo = ctx->d_blo.as<int32_t>(); dw.bhi = ctx->d_bhi.as<int32_t>();
            dw.nd_stride = b.nd_stride;
            dw.bx = da.sidx; dw.by = ctx->d_by.as<int32_t>(); dw.bwr = da.wre; dw.pred = da.pred; dw.max_pairs = b.max_pairs;
            dw.fb_list = ctx->d_fblist.as<int32_t>(); dw.fb_count = ctx->d_counter.as<int32_t>() + 10;
            dw.mrx = da.mrx; dw.mry = da.mry; dw.mrn = da.mrn; dw.nmruns = da.nmruns; dw.score = da.score;
            k_decode_w<<<b.dec_slots, 32, 0, ctx->stream>>>(dw);
            CK(cudaGetLastError());
            b.stats.launches++; b.stats.run_launches++;
            da.order = dw.fb_list; da.n_dev = dw.fb_count;
        }
        const int old_slots = std::min<int>(b.dec_slots, ctx->sm_count * (b.nw >= 2 ? 16 : 8));
        if (b.nw == 1) k_decode<1><<<old_slots, 32, 0, ctx->stream>>>(da);
        else k_decode<2><<<old_slots, 64, 0, ctx->stream>>>(da);
        CK(cudaGetLastError());
        b.stats.launches++; b.stats.run_launches++;
    }
    CK(cudaEventRecord(ctx->ev[4], ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    float m1 = 0.f, m2 = 0.f;
    cudaEventElapsedTime(&m1, ctx->ev[2], ctx->ev[3]);
    cudaEventElapsedTime(&m2, ctx->ev[3], ctx->ev[4]);
    b.stats.ms_fwdbwd = m1; b.stats.ms_decode = m2; b.stats.ms_total = m1 + m2;
    b.ran = true;
    return PHMM_OK;
}

// Re-runs with larger pair buffers while any region overflowed.  Returns the per-region pair counts.
int run_until_fits(phmm_ctx *ctx, std::vector<int32_t> &npairs) {
    BatchState &b = ctx->b;
    const int64_t nreg = (int64_t)b.regions.size();
    for (int attempt = 0; attempt < 6; attempt++) {
        if (!b.ran) { int rc = do_run(ctx); if (rc) return rc; }
        npairs.resize(nreg);
        if (nreg) CK(cudaMemcpy(npairs.data(), ctx->d_npairs.p, nreg * 4, cudaMemcpyDeviceToHost));
        bool over = false;
        for (int64_t i = 0; i < nreg; i++) if (npairs[i] > b.regions[i].pair_cap) { over = true; break; }
        if (!over) return PHMM_OK;
        b.pair_factor *= 4;
        int rc = plan_memory(ctx);
        if (rc) return rc;
        CK(cudaStreamSynchronize(ctx->stream));
        b.ran = false;
    }
    return fail(ctx, PHMM_E_NOMEM, "posterior pair buffers overflowed repeatedly");
}

void *xmalloc(size_t n) { return malloc(n ? n : 1); }

}  // namespace

extern "C" {

int phmm_version(void) { return PHMM_VERSION; }

void phmm_default_params(phmm_params *p) {
    if (!p) return;
    p->band = 10; p->anchor_trim = 14; p->split_side = 3000; p->min_diags = 1000; p->tb_diags = 40;
    p->threshold = 0.01; p->gap_gamma = 0.5; p->match_gamma = 0.0;
}

const char *phmm_create_error(void) { return g_create_error.c_str(); }

phmm_ctx *phmm_create(int device, const double *trans, const double *emis, int model_type) {
    g_create_error.clear();
    if ((trans == nullptr) != (emis == nullptr)) { g_create_error = "trans and emis must both be given or both be NULL"; return nullptr; }
    if (model_type != 0 && model_type != 1) { g_create_error = "model_type must be 0 (fiveState) or 1 (fiveStateAsymmetric)"; return nullptr; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        (void)cudaGetLastError();
        g_create_error = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                         " (libphmm_sm100 has no CPU fallback)";
        return nullptr;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device ordinal out of range"; return nullptr; }
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        g_create_error = "cudaSetDevice / cudaGetDeviceProperties failed";
        return nullptr;
    }
    if (prop.major != 10) {
        g_create_error = std::string("device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor) +
                         "; this library is built for sm_100a only";
        return nullptr;
    }
    phmm_ctx *ctx = new phmm_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        g_create_error = "cudaStreamCreate failed"; delete ctx; return nullptr;
    }
    ctx->stream = ctx->own_stream;
    for (auto &ev : ctx->ev) cudaEventCreate(&ev);
    build_model(ctx->model, trans, emis);
    return ctx;
}

void phmm_destroy(phmm_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto &ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char *phmm_last_error(phmm_ctx *ctx) { return ctx ? ctx->err.c_str() : "NULL ctx"; }

int phmm_set_stream(phmm_ctx *ctx, void *cuda_stream) {
    if (!ctx) return PHMM_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return PHMM_OK;
}

int phmm_set_model(phmm_ctx *ctx, const double *trans, const double *emis, int model_type) {
    if (!ctx) return PHMM_E_ARG;
    if ((trans == nullptr) != (emis == nullptr)) return fail(ctx, PHMM_E_ARG, "trans and emis must both be given or both be NULL");
    if (model_type != 0 && model_type != 1) return fail(ctx, PHMM_E_ARG, "model_type must be 0 or 1");
    build_model(ctx->model, trans, emis);
    ctx->b.ran = false;                     // the plan of a prepared batch (regions, geometry, records) does not depend on the model
    return PHMM_OK;
}

int phmm_set_reference(phmm_ctx *ctx, const uint8_t *bases, int64_t n) {
    if (!ctx) return PHMM_E_ARG;
    if (n < 0 || (n > 0 && !bases)) return fail(ctx, PHMM_E_ARG, "bad reference");
    CK(cudaSetDevice(ctx->device));
    CK(ctx->d_ref.ensure((size_t)n + 2 * BASE_PAD));
    CK(cudaMemset(ctx->d_ref.p, 4, (size_t)n + 2 * BASE_PAD));
    if (n) CK(cudaMemcpy(ctx->d_ref.as<uint8_t>() + BASE_PAD, bases, (size_t)n, cudaMemcpyHostToDevice));
    ctx->ref_len = n;
    ctx->baseexp_len = -1; ctx->baseexp_tables = 0;
    ctx->b.prepared = false; ctx->b.ran = false;
    return PHMM_OK;
}

int phmm_set_option(phmm_ctx *ctx, const char *name, int64_t value) {
    if (!ctx || !name) return PHMM_E_ARG;
    const std::string n(name);
    if (n == "legacy_kernel") ctx->force_legacy = value != 0;
    else if (n == "decode_full_sweep") ctx->decode_full_sweep = value != 0;
    else if (n == "decode_block") ctx->decode_block = value != 0;
    else if (n == "warps") {
        if (value != 0 && value != 2 && value != 4 && value != 8) return fail(ctx, PHMM_E_ARG, "warps must be 0, 2, 4 or 8");
        ctx->opt_warps = (int)value;
    } else if (n == "candidate_cap") {                 // tests: posterior candidates per window before the full re-read (0 = automatic)
        if (value < 0 || value > 0x7fffffff) return fail(ctx, PHMM_E_ARG, "candidate_cap must be >= 0");
        ctx->opt_ccap = (int)value;
    } else if (n == "candidate_eps_ppm") {             // tests: tolerance of the candidate shortcut in 1e-6 log units (default 20000)
        if (value < 0) return fail(ctx, PHMM_E_ARG, "candidate_eps_ppm must be >= 0");
        ctx->opt_est_eps = (double)value * 1e-6;
    } else if (n == "timing_experiment") {
#ifdef PHMM_TUNE
        ctx->opt_dbg = (int)value;
#else
        if (value != 0) return fail(ctx, PHMM_E_ARG, "timing_experiment exists only in a library built with -DPHMM_TUNE (scripts/tune.py)");
#endif
    } else if (n == "smem_columns") {
        if (value != 0 && (value < 64 || value > 1024 || (value & (value - 1)))) return fail(ctx, PHMM_E_ARG, "smem_columns must be 0 or a power of two in [64, 1024]");
        ctx->opt_wcap = (int)value;
    } else return fail(ctx, PHMM_E_ARG, "unknown option " + n);
    ctx->b.prepared = false; ctx->b.ran = false;
    return PHMM_OK;
}

int phmm_set_memory_budget(phmm_ctx *ctx, int64_t bytes) {
    if (!ctx) return PHMM_E_ARG;
    ctx->mem_budget = bytes;
    return PHMM_OK;
}

int phmm_batch_prepare(phmm_ctx *ctx, int64_t n_reads, const uint8_t *read_bases, const int64_t *read_off,
                       const int64_t *ref_start, const int64_t *ref_end, const uint32_t *in_cigar_ops,
                       const int64_t *in_cigar_off, const phmm_params *params) {
    if (!ctx) return PHMM_E_ARG;
    try {
        return do_prepare(ctx, n_reads, read_bases, read_off, ref_start, ref_end, in_cigar_ops, in_cigar_off, params, false);
    } catch (const std::exception &ex) { return fail(ctx, PHMM_E_NOMEM, ex.what()); }
}

int phmm_batch_run(phmm_ctx *ctx) {
    if (!ctx) return PHMM_E_ARG;
    try { return do_run(ctx); } catch (const std::exception &ex) { return fail(ctx, PHMM_E_NOMEM, ex.what()); }
}

int phmm_batch_get_stats(phmm_ctx *ctx, phmm_batch_stats *out) {
    if (!ctx || !out) return PHMM_E_ARG;
    *out = ctx->b.stats;
    return PHMM_OK;
}

void phmm_free(void *p) { free(p); }

void phmm_free_posteriors(phmm_posteriors *post) {
    if (!post) return;
    free(post->off); free(post->ref_pos); free(post->read_pos); free(post->prob_1e7);
    memset(post, 0, sizeof(*post));
}

int phmm_batch_fetch(phmm_ctx *ctx, uint32_t **out_cigar_ops, int64_t **out_cigar_off, phmm_posteriors *post) {
    if (!ctx) return PHMM_E_ARG;
    if (!out_cigar_ops || !out_cigar_off) return fail(ctx, PHMM_E_ARG, "NULL output pointer");
    BatchState &b = ctx->b;
    if (!b.prepared || b.expect) return fail(ctx, PHMM_E_STATE, "no realignment batch prepared");
    try {
        CK(cudaSetDevice(ctx->device));
        const int64_t nreg = (int64_t)b.regions.size();
        std::vector<int32_t> npairs;
        int rc = run_until_fits(ctx, npairs);
        if (rc) return rc;
        // match runs: counts -> offsets -> compact -> D2H
        std::vector<int32_t> nm(nreg);
        std::vector<int64_t> moff(nreg + 1, 0);
        std::vector<int32_t> hx, hy, hn;
        if (nreg) {
            CK(cudaMemcpy(nm.data(), ctx->d_nmruns.p, nreg * 4, cudaMemcpyDeviceToHost));
            for (int64_t i = 0; i < nreg; i++) {
                if (nm[i] > b.regions[i].mrun_cap) return fail(ctx, PHMM_E_CUDA, "match-run buffer overflow (internal)");
                moff[i + 1] = moff[i] + nm[i];
            }
            const int64_t tot = moff[nreg];
            CK(ctx->d_coff.ensure((size_t)(nreg + 1) * 8));
            CK(ctx->d_cx.ensure((size_t)tot * 4 + 16)); CK(ctx->d_cy.ensure((size_t)tot * 4 + 16)); CK(ctx->d_cn.ensure((size_t)tot * 4 + 16));
            CK(cudaMemcpyAsync(ctx->d_coff.p, moff.data(), (nreg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
            k_compact<<<(unsigned)nreg, 128, 0, ctx->stream>>>(ctx->d_regions.as<Region>(), ctx->d_nmruns.as<int32_t>(), ctx->d_coff.as<int64_t>(),
                                                                 (int)nreg, ctx->d_mrx.as<int32_t>(), ctx->d_mry.as<int32_t>(), ctx->d_mrn.as<int32_t>(),
                                                                 ctx->d_cx.as<int32_t>(), ctx->d_cy.as<int32_t>(), ctx->d_cn.as<int32_t>());
            CK(cudaGetLastError());
            b.stats.launches++;
            hx.resize(tot); hy.resize(tot); hn.resize(tot);
            if (tot) {
                CK(cudaMemcpyAsync(hx.data(), ctx->d_cx.p, tot * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(hy.data(), ctx->d_cy.p, tot * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(hn.data(), ctx->d_cn.p, tot * 4, cudaMemcpyDeviceToHost, ctx->stream));
            }
            CK(cudaStreamSynchronize(ctx->stream));
        }
        // CIGAR assembly: reference-only gap (D) before read-only gap (I) between matched runs.  Reads are
        // independent: a counting pass and a writing pass, both spread over the host threads.
        std::vector<int64_t> off(b.n_reads + 1, 0);
        // emits the ops of read r through `put(op)`; `last` carries the previous op for merging equal codes
        auto assemble = [&](int64_t r, auto &&put) {
            uint32_t cur = 0; bool have = false;
            auto push = [&](int code, int64_t len) {
                while (len > 0) {
                    if (have && (int)(cur & 3) == code && (int64_t)(cur >> 2) + len <= 0x3fffffff) {
                        cur = (uint32_t)((((int64_t)(cur >> 2) + len) << 2) | code);
                        return;
                    }
                    if (have) put(cur);
                    const int64_t l = std::min<int64_t>(len, 0x3fffffff);
                    cur = (uint32_t)((l << 2) | code); have = true;
                    len -= l;
                }
            };
            int64_t pxx = -1, pyy = -1;
            for (int64_t g = b.read_first_region[r]; g < b.read_first_region[r + 1]; g++) {
                const Region &reg = b.regions[g];
                for (int64_t k = moff[g + 1] - 1; k >= moff[g]; k--) {       // stored in reverse
                    const int64_t x = reg.x1 + hx[k], y = reg.y1 + hy[k], n = hn[k];
                    push(2, x - pxx - 1);
                    push(1, y - pyy - 1);
                    push(0, n);
                    pxx = x + n - 1; pyy = y + n - 1;
                }
            }
            push(2, b.read_lx[r] - pxx - 1);
            push(1, b.read_ly[r] - pyy - 1);
            if (have) put(cur);
        };
        const int nthr = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>(16, std::thread::hardware_concurrency()), b.n_reads / 256));
        auto parallel_reads = [&](auto &&fn) {
            if (nthr <= 1) { fn(0, b.n_reads); return; }
            std::vector<std::thread> th;
            const int64_t chunk = (b.n_reads + nthr - 1) / nthr;
            for (int t = 0; t < nthr; t++) {
                const int64_t r0 = t * chunk, r1 = std::min<int64_t>(b.n_reads, r0 + chunk);
                if (r0 < r1) th.emplace_back([&fn, r0, r1]() { fn(r0, r1); });
            }
            for (auto &t : th) t.join();
        };
        parallel_reads([&](int64_t r0, int64_t r1) {
            for (int64_t r = r0; r < r1; r++) { int64_t c = 0; assemble(r, [&](uint32_t) { c++; }); off[r + 1] = c; }
        });
        for (int64_t r = 0; r < b.n_reads; r++) off[r + 1] += off[r];
        const int64_t total_ops = off[b.n_reads];
        *out_cigar_ops = (uint32_t *)xmalloc((size_t)total_ops * 4);
        *out_cigar_off = (int64_t *)xmalloc(off.size() * 8);
        if (!*out_cigar_ops || !*out_cigar_off) return fail(ctx, PHMM_E_NOMEM, "host allocation failed");
        {
            uint32_t *const dst = *out_cigar_ops;
            parallel_reads([&](int64_t r0, int64_t r1) {
                for (int64_t r = r0; r < r1; r++) { uint32_t *w = dst + off[r]; assemble(r, [&](uint32_t v) { *w++ = v; }); }
            });
        }
        memcpy(*out_cigar_off, off.data(), off.size() * 8);
        int64_t tp = 0;
        for (int64_t i = 0; i < nreg; i++) tp += npairs[i];
        b.stats.pairs = tp;
        b.stats.d2h_bytes = nreg * (int64_t)sizeof(RegionGeom) + nreg * 8 + moff[nreg] * 12 + (post ? tp * 12 : 0);
        if (post) {
            memset(post, 0, sizeof(*post));
            post->n = tp;
            post->off = (int64_t *)xmalloc((b.n_reads + 1) * 8);
            post->ref_pos = (int32_t *)xmalloc(tp * 4); post->read_pos = (int32_t *)xmalloc(tp * 4); post->prob_1e7 = (int32_t *)xmalloc(tp * 4);
            if (!post->off || !post->ref_pos || !post->read_pos || !post->prob_1e7) return fail(ctx, PHMM_E_NOMEM, "host allocation failed");
            // compact pairs on the device with the same gather kernel
            std::vector<int64_t> poff(nreg + 1, 0);
            for (int64_t i = 0; i < nreg; i++) poff[i + 1] = poff[i] + npairs[i];
            std::vector<int32_t> qx(tp), qy(tp), qw(tp);
            if (tp) {
                // k_compact reads mrun_off/mrun_cap: give it a region table whose mrun fields alias the pair fields
                std::vector<Region> alias(b.regions);
                for (auto &g : alias) { g.mrun_off = g.pair_off; g.mrun_cap = g.pair_cap; }
                DevBuf d_alias;
                CK(d_alias.ensure(alias.size() * sizeof(Region)));
                CK(ctx->d_coff.ensure((size_t)(nreg + 1) * 8));
                CK(ctx->d_cx.ensure((size_t)tp * 4 + 16)); CK(ctx->d_cy.ensure((size_t)tp * 4 + 16)); CK(ctx->d_cn.ensure((size_t)tp * 4 + 16));
                CK(cudaMemcpyAsync(d_alias.p, alias.data(), alias.size() * sizeof(Region), cudaMemcpyHostToDevice, ctx->stream));
                CK(cudaMemcpyAsync(ctx->d_coff.p, poff.data(), (nreg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
                k_compact<<<(unsigned)nreg, 128, 0, ctx->stream>>>(d_alias.as<Region>(), ctx->d_npairs.as<int32_t>(), ctx->d_coff.as<int64_t>(), (int)nreg,
                                                                     ctx->d_px.as<int32_t>(), ctx->d_py.as<int32_t>(), ctx->d_pw.as<int32_t>(),
                                                                     ctx->d_cx.as<int32_t>(), ctx->d_cy.as<int32_t>(), ctx->d_cn.as<int32_t>());
                CK(cudaGetLastError());
                b.stats.launches++;
                CK(cudaMemcpyAsync(qx.data(), ctx->d_cx.p, tp * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(qy.data(), ctx->d_cy.p, tp * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaMemcpyAsync(qw.data(), ctx->d_cn.p, tp * 4, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaStreamSynchronize(ctx->stream));
            }
            // per read: its regions' pairs in canonical (ref_pos, read_pos) order; reads are independent -> host threads
            post->off[0] = 0;
            for (int64_t r = 0; r < b.n_reads; r++)
                post->off[r + 1] = post->off[r] + (poff[b.read_first_region[r + 1]] - poff[b.read_first_region[r]]);
            parallel_reads([&](int64_t r0, int64_t r1) {
                std::vector<std::pair<uint64_t, int32_t>> tmp;
                for (int64_t r = r0; r < r1; r++) {
                    tmp.clear();
                    for (int64_t g = b.read_first_region[r]; g < b.read_first_region[r + 1]; g++) {
                        const Region &reg = b.regions[g];
                        for (int64_t k = poff[g]; k < poff[g + 1]; k++)
                            tmp.emplace_back(((uint64_t)(uint32_t)(reg.x1 + qx[k]) << 32) | (uint32_t)(reg.y1 + qy[k]), qw[k]);
                    }
                    std::sort(tmp.begin(), tmp.end());                 // a cell occurs once: keys are distinct
                    int64_t o = post->off[r];
                    for (const auto &e : tmp) {
                        post->ref_pos[o] = (int32_t)(e.first >> 32); post->read_pos[o] = (int32_t)(uint32_t)e.first; post->prob_1e7[o] = e.second;
                        o++;
                    }
                }
            });
        }
        return PHMM_OK;
    } catch (const std::exception &ex) { return fail(ctx, PHMM_E_NOMEM, ex.what()); }
}

int phmm_base_expectations_reset(phmm_ctx *ctx, int32_t n_tables) {
    if (!ctx) return PHMM_E_ARG;
    if (ctx->ref_len < 0) return fail(ctx, PHMM_E_STATE, "no reference set");
    if (n_tables < 1 || n_tables > 4096) return fail(ctx, PHMM_E_ARG, "n_tables must be in 1..4096");
    CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)n_tables * (size_t)ctx->ref_len * 5 * 8 + 64;
    CK(ctx->d_baseexp.ensure(bytes));
    CK(cudaMemsetAsync(ctx->d_baseexp.p, 0, bytes, ctx->stream));
    ctx->baseexp_len = ctx->ref_len; ctx->baseexp_tables = n_tables;
    return PHMM_OK;
}

int phmm_batch_add_base_expectations(phmm_ctx *ctx, const uint8_t *read_mask, int32_t table) {
    if (!ctx) return PHMM_E_ARG;
    if (table < 0 || table >= ctx->baseexp_tables) return fail(ctx, PHMM_E_ARG, "no such table (phmm_base_expectations_reset sizes them)");
    BatchState &b = ctx->b;
    if (!b.prepared || b.expect) return fail(ctx, PHMM_E_STATE, "no realignment batch prepared");
    if (ctx->baseexp_len != ctx->ref_len) return fail(ctx, PHMM_E_STATE, "phmm_base_expectations_reset has not been called for this reference");
    try {
        CK(cudaSetDevice(ctx->device));
        const int64_t nreg = (int64_t)b.regions.size();
        if (nreg == 0) return PHMM_OK;
        std::vector<int32_t> npairs;
        int rc = run_until_fits(ctx, npairs);
        if (rc) return rc;
        const uint8_t *d_mask = nullptr;
        if (read_mask) {
            CK(ctx->d_readmask.ensure((size_t)b.n_reads + 16));
            CK(cudaMemcpyAsync(ctx->d_readmask.p, read_mask, (size_t)b.n_reads, cudaMemcpyHostToDevice, ctx->stream));
            d_mask = ctx->d_readmask.as<uint8_t>();
        }
        const unsigned grid = (unsigned)std::min<int64_t>(nreg, (int64_t)ctx->sm_count * 16);
        k_base_expect<<<grid, 256, 0, ctx->stream>>>(ctx->d_regions.as<Region>(), ctx->d_npairs.as<int32_t>(), (int)nreg,
                                                     ctx->d_reads.as<uint8_t>() + BASE_PAD, d_mask, ctx->d_px.as<int32_t>(),
                                                     ctx->d_py.as<int32_t>(), ctx->d_pw.as<int32_t>(),
                                                     ctx->d_baseexp.as<unsigned long long>() + (size_t)table * (size_t)ctx->ref_len * 5);
        CK(cudaGetLastError());
        b.stats.launches++;
        CK(cudaStreamSynchronize(ctx->stream));      // read_mask is the caller's again
        return PHMM_OK;
    } catch (const std::exception &ex) { return fail(ctx, PHMM_E_NOMEM, ex.what()); }
}

int phmm_base_expectations_fetch(phmm_ctx *ctx, int32_t table, int64_t *out, int64_t n) {
    if (!ctx || !out) return PHMM_E_ARG;
    if (table < 0 || table >= ctx->baseexp_tables) return fail(ctx, PHMM_E_ARG, "no such table");
    if (ctx->baseexp_len < 0 || ctx->baseexp_len != ctx->ref_len) return fail(ctx, PHMM_E_STATE, "no base expectations accumulated");
    if (n != 5 * ctx->ref_len) return fail(ctx, PHMM_E_ARG, "out must hold 5 values per reference base");
    CK(cudaSetDevice(ctx->device));
    if (n) CK(cudaMemcpyAsync(out, ctx->d_baseexp.as<unsigned long long>() + (size_t)table * (size_t)n, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return PHMM_OK;
}

int phmm_realign_batch(phmm_ctx *ctx, int64_t n_reads, const uint8_t *read_bases, const int64_t *read_off,
                       const int64_t *ref_start, const int64_t *ref_end, const uint32_t *in_cigar_ops,
                       const int64_t *in_cigar_off, const phmm_params *params, uint32_t **out_cigar_ops,
                       int64_t **out_cigar_off, phmm_posteriors *post) {
    int rc = phmm_batch_prepare(ctx, n_reads, read_bases, read_off, ref_start, ref_end, in_cigar_ops, in_cigar_off, params);
    if (rc) return rc;
    rc = phmm_batch_run(ctx);
    if (rc) return rc;
    return phmm_batch_fetch(ctx, out_cigar_ops, out_cigar_off, post);
}

// E-step of one batch as exact integers: out_hi[k] + out_lo[k] / 2^32 for the 105 expectations (the kernels
// accumulate in 2^-32 fixed point), out_hi[105] + out_lo[105] / 2^20 for the summed log-likelihood (per-region
// values rounded to 2^-20).  Integer sums are independent of the order of regions, reads, calls and ranks.
// Sums the per-region statistics of the E-step that just ran.
static int expectations_reduce(phmm_ctx *ctx, int64_t out_hi[106], int64_t out_lo[106]) {
    BatchState &b = ctx->b;
    const int64_t nreg = (int64_t)b.regions.size();
    std::vector<unsigned long long> T(nreg * 25), E(nreg * 80);
    std::vector<double> LL(nreg);
    if (nreg) {
        CK(cudaMemcpy(T.data(), ctx->d_expT.p, nreg * 25 * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(E.data(), ctx->d_expE.p, nreg * 80 * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(LL.data(), ctx->d_expLL.p, nreg * 8, cudaMemcpyDeviceToHost));
    }
    b.stats.d2h_bytes += nreg * (25 + 80 + 1) * 8;
    __int128 acc[106];
    for (int k = 0; k < 106; k++) acc[k] = 0;
    for (int64_t g = 0; g < nreg; g++) {
        // a sub-problem no path of the model can explain (zero transitions in a sparse or random-start HMM) has no
        // finite likelihood and its expectations are meaningless: report it instead of summing garbage
        if (!std::isfinite(LL[g]))
            return fail(ctx, PHMM_E_ARG, "read " + std::to_string(b.regions[g].read) + " has zero probability under the model (region " +
                                             std::to_string(g) + ": log-likelihood is not finite)");
        for (int k = 0; k < 25; k++) acc[k] += (__int128)(int64_t)T[g * 25 + k];
        for (int k = 0; k < 80; k++) acc[25 + k] += (__int128)(int64_t)E[g * 80 + k];
        acc[105] += (__int128)llrint(LL[g] * 1048576.0);
    }
    for (int k = 0; k < 106; k++) {
        const int bits = k < 105 ? 32 : 20;
        out_hi[k] = (int64_t)(acc[k] >> bits);                                  // arithmetic shift: floor
        out_lo[k] = (int64_t)(acc[k] - ((__int128)out_hi[k] << bits));         // in [0, 2^bits)
    }
    return PHMM_OK;
}

// E-step of one batch as exact integers: out_hi[k] + out_lo[k] / 2^32 for the 105 expectations (the kernels
// accumulate in 2^-32 fixed point), out_hi[105] + out_lo[105] / 2^20 for the summed log-likelihood (per-region
// values rounded to 2^-20).  Integer sums are independent of the order of regions, reads, calls and ranks.
static int expectations_fixed(phmm_ctx *ctx, int64_t n_reads, const uint8_t *read_bases, const int64_t *read_off,
                              const int64_t *ref_start, const int64_t *ref_end, const uint32_t *in_cigar_ops,
                              const int64_t *in_cigar_off, const phmm_params *params, int64_t out_hi[106], int64_t out_lo[106]) {
    int rc = do_prepare(ctx, n_reads, read_bases, read_off, ref_start, ref_end, in_cigar_ops, in_cigar_off, params, true);
    if (rc) return rc;
    rc = do_run(ctx);
    if (rc) return rc;
    return expectations_reduce(ctx, out_hi, out_lo);
}

int phmm_expectations_prepare(phmm_ctx *ctx, int64_t n_reads, const uint8_t *read_bases, const int64_t *read_off,
                              const int64_t *ref_start, const int64_t *ref_end, const uint32_t *in_cigar_ops,
                              const int64_t *in_cigar_off, const phmm_params *params) {
    if (!ctx) return PHMM_E_ARG;
    try {
        return do_prepare(ctx, n_reads, read_bases, read_off, ref_start, ref_end, in_cigar_ops, in_cigar_off, params, true);
    } catch (const std::exception &ex) { return fail(ctx, PHMM_E_NOMEM, ex.what()); }
}

int phmm_expectations_run_fixed(phmm_ctx *ctx, int64_t out_hi[106], int64_t out_lo[106]) {
    if (!ctx) return PHMM_E_ARG;
    if (!out_hi || !out_lo) return fail(ctx, PHMM_E_ARG, "out_hi / out_lo is NULL");
    if (!ctx->b.prepared || !ctx->b.expect) return fail(ctx, PHMM_E_STATE, "no E-step batch prepared (phmm_expectations_prepare)");
    try {
        int rc = do_run(ctx);
        if (rc) return rc;
        return expectations_reduce(ctx, out_hi, out_lo);
    } catch (const std::exception &ex) { return fail(ctx, PHMM_E_NOMEM, ex.what()); }
}

int phmm_expectations_batch_fixed(phmm_ctx *ctx, int64_t n_reads, const uint8_t *read_bases, const int64_t *read_off,
                                  const int64_t *ref_start, const int64_t *ref_end, const uint32_t *in_cigar_ops,
                                  const int64_t *in_cigar_off, const phmm_params *params, int64_t out_hi[106],
                                  int64_t out_lo[106]) {
    if (!ctx) return PHMM_E_ARG;
    if (!out_hi || !out_lo) return fail(ctx, PHMM_E_ARG, "out_hi / out_lo is NULL");
    try {
        return expectations_fixed(ctx, n_reads, read_bases, read_off, ref_start, ref_end, in_cigar_ops, in_cigar_off, params, out_hi, out_lo);
    } catch (const std::exception &ex) { return fail(ctx, PHMM_E_NOMEM, ex.what()); }
}

int phmm_expectations_batch(phmm_ctx *ctx, int64_t n_reads, const uint8_t *read_bases, const int64_t *read_off,
                            const int64_t *ref_start, const int64_t *ref_end, const uint32_t *in_cigar_ops,
                            const int64_t *in_cigar_off, const phmm_params *params, double out_stats[106]) {
    if (!ctx) return PHMM_E_ARG;
    if (!out_stats) return fail(ctx, PHMM_E_ARG, "out_stats is NULL");
    try {
        int64_t hi[106], lo[106];
        int rc = expectations_fixed(ctx, n_reads, read_bases, read_off, ref_start, ref_end, in_cigar_ops, in_cigar_off, params, hi, lo);
        if (rc) return rc;
        for (int k = 0; k < 105; k++) out_stats[k] = (double)hi[k] + (double)lo[k] / 4294967296.0;
        out_stats[105] = (double)hi[105] + (double)lo[105] / 1048576.0;
        return PHMM_OK;
    } catch (const std::exception &ex) { return fail(ctx, PHMM_E_NOMEM, ex.what()); }
}

}  // extern "C"
