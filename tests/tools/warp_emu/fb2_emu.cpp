// fb2_emu.cpp -- runs the UNCHANGED sources of k_geometry, k_records and k_fb2 (nanopore_b200/csrc/phmm_kernels.cuh,
// phmm_fb2.cuh) on the host, one DP region at a time: the threads of the block are fibers that switch at every
// __syncthreads (warp_emu.h).  The planning a launch needs (ring size, records, traceback points) mirrors
// plan_memory() / do_run() of phmm_api.cu for a single region.  tests/test_fb2_emulated.py compares the posterior pairs
// with the checker's, bit for bit.  Test infrastructure only.  Build: g++ -O1 -ffp-contract=off.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "cuda_runtime.h"
#include "warp_emu.h"
#include "../../../nanopore_b200/csrc/phmm_fb2.cuh"

using namespace phmm;

static int pow2_at_least(int v) { int p = 1; while (p < v) p <<= 1; return p; }

// model60: tr[25], eM[25], eX[5], eY[5] in log space (oracle.Model().dump()).  region: x1, y1, x2, y2, ragged_left,
// ragged_right in window / read coordinates; runs: n_runs x (x, y, n) region-local anchor runs.  Returns the number of
// posterior pairs (written to px, py, pw in region-local sequence coordinates, unsorted), -1 if cap is too small.
//
// EXPECT: the Baum-Welch E-step instead (k_fb2<.., true>): expT[25] / expE[80] receive the region's 2^-32 fixed-point counts,
// expLL its summed log-likelihood; no pairs.
template <int NW, bool EXPECT>
static int run_region(const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY, const int64_t *region, int n_runs, const int32_t *runs_xyn,
                      const double *model60, int band, int min_diags, int tb_diags, double threshold, int wcap_opt, unsigned seed,
                      int32_t *px, int32_t *py, int32_t *pw, int cap, int64_t *cells,
                      unsigned long long *expT = nullptr, unsigned long long *expE = nullptr, double *expLL = nullptr) {
    constexpr int PAD = 16;
    std::vector<uint8_t> refp(lX + 2 * PAD, 4), readp(lY + 2 * PAD, 4);
    memcpy(refp.data() + PAD, X, lX);
    memcpy(readp.data() + PAD, Y, lY);
    Region reg;
    memset(&reg, 0, sizeof(reg));
    reg.xoff = region[0]; reg.yoff = region[1];
    reg.lx = (int32_t)(region[2] - region[0]); reg.ly = (int32_t)(region[3] - region[1]);
    reg.ragged_left = (int32_t)region[4]; reg.ragged_right = (int32_t)region[5];
    reg.run0 = 0; reg.nrun = n_runs; reg.pair_off = 0; reg.pair_cap = cap;
    std::vector<Run> runs(n_runs + 1);
    for (int i = 0; i < n_runs; i++) runs[i] = Run{runs_xyn[3 * i], runs_xyn[3 * i + 1], runs_xyn[3 * i + 2]};
    DevParams dp;
    memset(&dp, 0, sizeof(dp));
    dp.expansion = band; dp.min_diags = min_diags; dp.tb_diags = tb_diags;
    dp.threshold = threshold; dp.lp_skip = threshold > 0.0 ? log(threshold) - 1e-3 : -INFINITY;
    DevModel m;
    memset(&m, 0, sizeof(m));
    memcpy(m.tr, model60, 25 * 8); memcpy(m.eM, model60 + 25, 25 * 8); memcpy(m.eX, model60 + 50, 5 * 8); memcpy(m.eY, model60 + 55, 5 * 8);
    for (int s = 0; s < 5; s++) m.endp[s] = m.tr[s * 5 + S_M];                 // as build_model() of phmm_api.cu
    m.rendp[S_M] = m.tr[S_M * 5 + S_LX]; m.rendp[S_SX] = m.tr[S_M * 5 + S_LX]; m.rendp[S_SY] = m.tr[S_M * 5 + S_LY];
    m.rendp[S_LX] = m.tr[S_LX * 5 + S_LX]; m.rendp[S_LY] = m.tr[S_LY * 5 + S_LY];
    const bool sw = m.tr[S_SX * 5 + S_SY] != -INFINITY || m.tr[S_SY * 5 + S_SX] != -INFINITY;
    m.has_switch = sw;
    // geometry
    const int nd = reg.lx + reg.ly;
    const int64_t gap = std::max<int64_t>(1, (int64_t)min_diags - tb_diags - 1);
    std::vector<int64_t> tb_off = {0, nd / gap + 2};
    std::vector<int32_t> tbp(tb_off[1] + 4, 0);
    RegionGeom geom;
    memset(&geom, 0, sizeof(geom));
    warp_emu::run_block(1, 0, [&]() { k_geometry(&reg, runs.data(), 1, dp, &geom, tb_off.data(), tbp.data()); });
    *cells = geom.cells;
    if (nd == 0) return 0;
    // plan (plan_memory of phmm_api.cu for one region; realignment layout: 1 double per cell, 5 more on total diagonals;
    // E-step layout: 10 doubles per live cell, 11 where a total is evaluated)
    const int bw = geom.max_width, wg = pow2_at_least(bw);
    const int wcap = wcap_opt ? wcap_opt : std::max(64, std::min(512, wg));
    const int64_t ring_cells = std::max<int64_t>(2, geom.max_live_cells + geom.max_width + 2);
    const int64_t live_need = EXPECT ? (ring_cells + 2) * 11 : geom.max_live_doubles;
    const int64_t ring_doubles = live_need + 4 * 11 * (int64_t)bw + 16;
    const int dcap = std::max(4, geom.max_live_diags + 4) + 4, tcap = dcap / TOTAL_EVERY + 4;
    std::vector<int64_t> rec_off = {0, (int64_t)nd + 1};
    std::vector<DiagRec> recs(nd + 8);
    const int32_t *ntb = &geom.tracebacks;
    warp_emu::run_block(1, 0, [&]() {
        k_records(&reg, runs.data(), 1, dp, tb_off.data(), tbp.data(), ntb, (int)(sizeof(RegionGeom) / 4), ring_doubles, wcap,
                  EXPECT ? 10 : 1, EXPECT ? 1 : 5, rec_off.data(), recs.data());
    });
    const int ccap = EXPECT ? 1 : 2 * cap + 1024;
    std::vector<double> ring(ring_doubles + 16), wide((size_t)4 * NS * wg + 16), fsave((size_t)2 * CS * wcap + 16), totals((size_t)tcap + wg + 16);
    std::vector<long long> cand(ccap + 16);
    std::vector<unsigned char> smem((size_t)2 * CS * wcap * 8 + 64);
    int32_t order = 0, counter = 0, npairs = 0;
    Fb2Args a;
    memset(&a, 0, sizeof(a));
    a.ref = refp.data() + PAD; a.reads = readp.data() + PAD;
    a.regions = &reg; a.runs = runs.data(); a.order = &order; a.n_regions = 1; a.counter = &counter;
    a.m = m; a.p = dp;
    a.tb_off = tb_off.data(); a.tbp = tbp.data(); a.ntb = ntb; a.ntb_stride = (int32_t)(sizeof(RegionGeom) / 4);
    a.ring = ring.data(); a.ring_doubles = ring_doubles;
    a.recs = recs.data(); a.rec_off = rec_off.data();
    a.wide = wide.data(); a.wg = wg; a.fsave = fsave.data(); a.totals = totals.data(); a.tcap = tcap;
    a.cand = cand.data(); a.ccap = ccap; a.est_eps = 0.02; a.wcap = wcap;
    a.px = px; a.py = py; a.pw = pw; a.npairs = &npairs;
    a.expT = expT; a.expE = expE; a.expLL = expLL;
    warp_emu::dyn_smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem.data() + 15) & ~(uintptr_t)15);
    if (sw) warp_emu::run_block(NW * 32, seed, [&]() { k_fb2<NW, true, EXPECT>(a); });
    else warp_emu::run_block(NW * 32, seed, [&]() { k_fb2<NW, false, EXPECT>(a); });
    return npairs > cap ? -1 : npairs;
}

extern "C" int emu_fb2_region(const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY, const int64_t *region, int n_runs,
                              const int32_t *runs_xyn, const double *model60, int band, int min_diags, int tb_diags, double threshold,
                              int warps, int wcap, unsigned seed, int32_t *px, int32_t *py, int32_t *pw, int cap, int64_t *cells) {
    if (warps == 2) return run_region<2, false>(X, lX, Y, lY, region, n_runs, runs_xyn, model60, band, min_diags, tb_diags, threshold, wcap, seed, px, py, pw, cap, cells);
    if (warps == 8) return run_region<8, false>(X, lX, Y, lY, region, n_runs, runs_xyn, model60, band, min_diags, tb_diags, threshold, wcap, seed, px, py, pw, cap, cells);
    return run_region<4, false>(X, lX, Y, lY, region, n_runs, runs_xyn, model60, band, min_diags, tb_diags, threshold, wcap, seed, px, py, pw, cap, cells);
}

extern "C" int emu_fb2_region_expect(const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY, const int64_t *region, int n_runs,
                                     const int32_t *runs_xyn, const double *model60, int band, int min_diags, int tb_diags, int warps,
                                     int wcap, unsigned seed, unsigned long long *expT, unsigned long long *expE, double *expLL, int64_t *cells) {
    int32_t dummy[4];
    if (nullptr == expT || nullptr == expE || nullptr == expLL) return -2;
    for (int k = 0; k < 25; k++) expT[k] = 0;
    for (int k = 0; k < 80; k++) expE[k] = 0;
    *expLL = 0.0;
    if (warps == 2) return run_region<2, true>(X, lX, Y, lY, region, n_runs, runs_xyn, model60, band, min_diags, tb_diags, 0.01, wcap, seed, dummy, dummy, dummy, 1, cells, expT, expE, expLL);
    return run_region<4, true>(X, lX, Y, lY, region, n_runs, runs_xyn, model60, band, min_diags, tb_diags, 0.01, wcap, seed, dummy, dummy, dummy, 1, cells, expT, expE, expLL);
}

// The first-generation kernel k_fwdbwd<NW> (whole forward window in HBM; option legacy_kernel and the fall-back for very
// long E-step windows): one region, planning as plan_memory() does for it.
template <int NW>
static int run_region_legacy(const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY, const int64_t *region, int n_runs, const int32_t *runs_xyn,
                             const double *model60, int band, int min_diags, int tb_diags, double threshold, unsigned seed,
                             int32_t *px, int32_t *py, int32_t *pw, int cap, int64_t *cells) {
    constexpr int PAD = 16;
    std::vector<uint8_t> refp(lX + 2 * PAD, 4), readp(lY + 2 * PAD, 4);
    memcpy(refp.data() + PAD, X, lX);
    memcpy(readp.data() + PAD, Y, lY);
    Region reg;
    memset(&reg, 0, sizeof(reg));
    reg.xoff = region[0]; reg.yoff = region[1];
    reg.lx = (int32_t)(region[2] - region[0]); reg.ly = (int32_t)(region[3] - region[1]);
    reg.ragged_left = (int32_t)region[4]; reg.ragged_right = (int32_t)region[5];
    reg.nrun = n_runs; reg.pair_cap = cap;
    std::vector<Run> runs(n_runs + 1);
    for (int i = 0; i < n_runs; i++) runs[i] = Run{runs_xyn[3 * i], runs_xyn[3 * i + 1], runs_xyn[3 * i + 2]};
    DevParams dp;
    memset(&dp, 0, sizeof(dp));
    dp.expansion = band; dp.min_diags = min_diags; dp.tb_diags = tb_diags;
    dp.threshold = threshold; dp.lp_skip = threshold > 0.0 ? log(threshold) - 1e-3 : -INFINITY;
    DevModel m;
    memset(&m, 0, sizeof(m));
    memcpy(m.tr, model60, 25 * 8); memcpy(m.eM, model60 + 25, 25 * 8); memcpy(m.eX, model60 + 50, 5 * 8); memcpy(m.eY, model60 + 55, 5 * 8);
    for (int s = 0; s < 5; s++) m.endp[s] = m.tr[s * 5 + S_M];
    m.rendp[S_M] = m.tr[S_M * 5 + S_LX]; m.rendp[S_SX] = m.tr[S_M * 5 + S_LX]; m.rendp[S_SY] = m.tr[S_M * 5 + S_LY];
    m.rendp[S_LX] = m.tr[S_LX * 5 + S_LX]; m.rendp[S_LY] = m.tr[S_LY * 5 + S_LY];
    const bool sw = m.tr[S_SX * 5 + S_SY] != -INFINITY || m.tr[S_SY * 5 + S_SX] != -INFINITY;
    m.has_switch = sw;
    const int nd = reg.lx + reg.ly;
    const int64_t gap = std::max<int64_t>(1, (int64_t)min_diags - tb_diags - 1);
    std::vector<int64_t> tb_off = {0, nd / gap + 2};
    std::vector<int32_t> tbp(tb_off[1] + 4, 0);
    RegionGeom geom;
    memset(&geom, 0, sizeof(geom));
    warp_emu::run_block(1, 0, [&]() { k_geometry(&reg, runs.data(), 1, dp, &geom, tb_off.data(), tbp.data()); });
    *cells = geom.cells;
    if (nd == 0) return 0;
    const int64_t ring_cells = std::max<int64_t>(2, geom.max_live_cells + geom.max_width + 2);
    const int dcap = std::max(4, geom.max_live_diags + 4), bw = std::max(1, geom.max_width);
    std::vector<double> fring((size_t)ring_cells * NS + 16), bring((size_t)3 * bw * NS + 16), dots((size_t)2 * bw + 16);
    std::vector<DiagRec> dtab(dcap + 4);
    int32_t order = 0, counter = 0, npairs = 0;
    FbArgs a;
    memset(&a, 0, sizeof(a));
    a.ref = refp.data() + PAD; a.reads = readp.data() + PAD;
    a.regions = &reg; a.runs = runs.data(); a.order = &order; a.n_regions = 1; a.counter = &counter;
    a.m = m; a.p = dp;
    a.fring = fring.data(); a.ring_cells = ring_cells; a.dtab = dtab.data(); a.dcap = dcap;
    a.bring = bring.data(); a.bw = bw; a.dots = dots.data();
    a.px = px; a.py = py; a.pw = pw; a.npairs = &npairs;
    if (sw) warp_emu::run_block(NW * 32, seed, [&]() { k_fwdbwd<NW, true, false>(a); });
    else warp_emu::run_block(NW * 32, seed, [&]() { k_fwdbwd<NW, false, false>(a); });
    return npairs > cap ? -1 : npairs;
}

extern "C" int emu_fwdbwd_region(const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY, const int64_t *region, int n_runs,
                                 const int32_t *runs_xyn, const double *model60, int band, int min_diags, int tb_diags, double threshold,
                                 int warps, int wcap_unused, unsigned seed, int32_t *px, int32_t *py, int32_t *pw, int cap, int64_t *cells) {
    (void)wcap_unused;
    if (warps == 1) return run_region_legacy<1>(X, lX, Y, lY, region, n_runs, runs_xyn, model60, band, min_diags, tb_diags, threshold, seed, px, py, pw, cap, cells);
    if (warps == 2) return run_region_legacy<2>(X, lX, Y, lY, region, n_runs, runs_xyn, model60, band, min_diags, tb_diags, threshold, seed, px, py, pw, cap, cells);
    return run_region_legacy<4>(X, lX, Y, lY, region, n_runs, runs_xyn, model60, band, min_diags, tb_diags, threshold, seed, px, py, pw, cap, cells);
}
