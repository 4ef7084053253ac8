// decode_w_emu.cpp -- runs the UNCHANGED source of k_decode_w (nanopore_b200/csrc/phmm_decode_w.cuh) on the host: the
// 32 lanes of its one-warp block are 32 fibers that switch at every warp synchronisation point (see cuda_runtime.h in
// this directory).  tests/test_decode_w_emulated.py feeds it the checker's posterior pairs and compares the chain it
// returns with the checker's.  Test infrastructure only.
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#include "cuda_runtime.h"
#include "warp_emu.h"
#include "../../../nanopore_b200/csrc/phmm_decode_w.cuh"

// One region: pairs (px, py, pw) in region-local sequence coordinates.  Returns the number of match runs written
// (reverse order) or -1 when the kernel left the region to k_decode; *score receives the chain score.
extern "C" int emu_decode_w(int lx, int ly, int np, const int32_t *px, const int32_t *py, const int32_t *pw, double gap_gamma,
                            double match_gamma, int regular, unsigned seed, int32_t *mrx, int32_t *mry, int32_t *mrn, int mrun_cap,
                            int64_t *score, int32_t *envelope_cells) {
    using namespace phmm;
    Region reg;
    memset(&reg, 0, sizeof(reg));
    reg.lx = lx; reg.ly = ly; reg.pair_off = 0; reg.pair_cap = np; reg.mrun_off = 0; reg.mrun_cap = mrun_cap;
    const int nd = lx + ly, stride = (nd + 4 + 3) & ~3;
    std::vector<int32_t> sumx(lx + 1), sumy(ly + 1), dstart(stride), nxt(stride), blo(stride), bhi(stride), bx(np + 1), by(np + 1), pred(np + 1);
    std::vector<int64_t> bwr(np + 1);
    // stale scratch of an earlier region must not matter
    for (auto *v : {&sumx, &sumy, &dstart, &nxt, &blo, &bhi, &bx, &by, &pred}) for (auto &e : *v) e = 0x5a5a5a5a;
    int32_t order = 0, counter = 0, npairs = np, fb_list[1] = {-1}, fb_count = 0, nmruns = -7;
    DecWArgs a;
    memset(&a, 0, sizeof(a));
    a.regions = &reg; a.order = &order; a.n_regions = 1; a.counter = &counter;
    a.p.gap_gamma = gap_gamma; a.p.match_gamma = match_gamma;
    a.px = px; a.py = py; a.pw = pw; a.npairs = &npairs;
    a.regular = &regular; a.regular_stride = 0;
    a.sumx = sumx.data(); a.sumy = sumy.data(); a.max_lx = lx; a.max_ly = ly;
    a.dstart = dstart.data(); a.nxt = nxt.data(); a.blo = blo.data(); a.bhi = bhi.data(); a.nd_stride = stride;
    a.bx = bx.data(); a.by = by.data(); a.bwr = bwr.data(); a.pred = pred.data(); a.max_pairs = np;
    a.fb_list = fb_list; a.fb_count = &fb_count;
    a.mrx = mrx; a.mry = mry; a.mrn = mrn; a.nmruns = &nmruns; a.score = score;
    warp_emu::run_block(32, seed, [&]() { phmm::k_decode_w(a); });
    if (envelope_cells) {
        long c = 0;
        for (int d = 0; d <= nd; d++) c += bhi[d] - blo[d] + 1;
        *envelope_cells = (int32_t)std::min<long>(c, 0x7fffffff);
    }
    if (fb_count) return -1;
    return nmruns;
}
