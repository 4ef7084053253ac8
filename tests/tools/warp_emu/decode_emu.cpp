// decode_emu.cpp -- runs the UNCHANGED source of k_decode<2> (nanopore_b200/csrc/phmm_kernels.cuh: the block-per-region
// MEA chain decode on the band of the forward sweep, the kernel that takes the regions k_decode_w leaves) on the host,
// one region per call, 64 fibers (warp_emu.h); k_geometry supplies the band's width and its `regular` flag as on the
// device.  tests/test_decode_w_emulated.py compares the chain with the checker's.  Test infrastructure only.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "cuda_runtime.h"
#include "warp_emu.h"
#include "../../../nanopore_b200/csrc/phmm_kernels.cuh"

using namespace phmm;

// pairs in region-local sequence coordinates; runs: n_runs x (x, y, n) region-local anchor runs of the band.
// full_sweep != 0: no skipping of pairless stretches (option decode_full_sweep).  Returns the number of match runs
// (reverse order in mrx / mry / mrn).
extern "C" int emu_decode_block(int lx, int ly, int n_runs, const int32_t *runs_xyn, int band, int np, const int32_t *px, const int32_t *py,
                                const int32_t *pw, double gap_gamma, double match_gamma, int full_sweep, unsigned seed, int32_t *mrx,
                                int32_t *mry, int32_t *mrn, int mrun_cap, int64_t *score, int32_t *regular_out) {
    Region reg;
    memset(&reg, 0, sizeof(reg));
    reg.lx = lx; reg.ly = ly; reg.run0 = 0; reg.nrun = n_runs; reg.pair_off = 0; reg.pair_cap = np; reg.mrun_off = 0; reg.mrun_cap = mrun_cap;
    std::vector<Run> runs(n_runs + 1);
    for (int i = 0; i < n_runs; i++) runs[i] = Run{runs_xyn[3 * i], runs_xyn[3 * i + 1], runs_xyn[3 * i + 2]};
    DevParams dp;
    memset(&dp, 0, sizeof(dp));
    dp.expansion = band; dp.min_diags = 1000; dp.tb_diags = 40; dp.threshold = 0.01; dp.gap_gamma = gap_gamma; dp.match_gamma = match_gamma;
    const int nd = lx + ly;
    std::vector<int64_t> tb_off = {0, nd / 959 + 2};
    std::vector<int32_t> tbp(tb_off[1] + 4, 0);
    RegionGeom geom;
    memset(&geom, 0, sizeof(geom));
    warp_emu::run_block(1, 0, [&]() { k_geometry(&reg, runs.data(), 1, dp, &geom, tb_off.data(), tbp.data()); });
    if (regular_out) *regular_out = geom.regular;
    const int bw = std::max(1, geom.max_width);
    std::vector<int32_t> sumx(lx + 1, 0x5a5a5a5a), sumy(ly + 1, 0x5a5a5a5a), dstart(nd + 4, 0x5a5a5a5a), dfill(nd + 4, 0x5a5a5a5a), sidx(np + 1), pred(np + 1),
        lring((size_t)4 * bw);
    std::vector<int64_t> wre(np + 1), colmap((size_t)2 * (lx + 2), 0x5a5a5a5a5a5a5a5aLL), sring((size_t)4 * bw);
    int32_t order = 0, counter = 0, npairs = np, nmruns = -7, zero = 0;
    DecArgs a;
    memset(&a, 0, sizeof(a));
    a.regions = &reg; a.runs = runs.data(); a.order = &order; a.n_regions = 1; a.counter = &counter; a.p = dp;
    a.px = px; a.py = py; a.pw = pw; a.npairs = &npairs;
    a.sumx = sumx.data(); a.sumy = sumy.data(); a.max_lx = lx; a.max_ly = ly;
    a.dstart = dstart.data(); a.dfill = dfill.data(); a.max_nd = nd;
    a.sidx = sidx.data(); a.wre = wre.data(); a.pred = pred.data(); a.max_pairs = np;
    a.colmap = colmap.data(); a.sring = sring.data(); a.lring = lring.data(); a.bw = bw;
    a.regular = full_sweep ? &zero : &geom.regular; a.regular_stride = 0;
    a.mrx = mrx; a.mry = mry; a.mrn = mrn; a.nmruns = &nmruns; a.score = score;
    warp_emu::run_block(64, seed, [&]() { k_decode<2>(a); });
    return nmruns;
}
