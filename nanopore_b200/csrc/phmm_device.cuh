// phmm_device.cuh -- device-side types, log-space arithmetic and the band
// iterator shared by the kernels of libphmm_sm100.so (sm_100a only).
//
// Algorithm provenance: SURVEY.md Appendix A (cPecan pair-HMM of
// benedictpaten/cactus, absent from the reference tree); call-site contract
// from reference nanopore/analyses/utils.py:587.  Independent of oracle/.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace phmm {

constexpr int NS = 5;                 // match, shortGapX, shortGapY, longGapX, longGapY (utils.py:617)
constexpr int S_M = 0, S_SX = 1, S_SY = 2, S_LX = 3, S_LY = 4;
constexpr int PROB_1 = 10000000;      // posterior quantum 1e-7

struct Run {                          // run of consecutive diagonal anchors, region-local sequence coords
    int32_t x, y, n;
};

struct Region {                       // one banded DP sub-problem (after splitting)
    int64_t xoff;                     // index of region origin in the device reference array
    int64_t yoff;                     // index of region origin in the device read array
    int64_t pair_off;                 // first slot of this region in the pair arrays
    int64_t mrun_off;                 // first slot in the match-run output
    int32_t lx, ly;
    int32_t run0, nrun;
    int32_t read, x1, y1;
    int32_t pair_cap, mrun_cap;
    int32_t ragged_left, ragged_right;
    int32_t pad;
};

struct RegionGeom {                   // written by the geometry kernel
    int64_t cells;
    int64_t max_live_cells;           // peak cells of forward window between tracebacks (incl. wrap slack)
    int32_t max_width;
    int32_t max_live_diags;
    int32_t tracebacks;
    int32_t diagonals;
    int64_t max_live_doubles;         // peak of the windowed kernel's ring (1 or 6 doubles per cell, see phmm_fb2.cuh)
    int32_t regular;                  // 1: both band edges move right by 0 or 1 cell per diagonal, so every band cell has a
                                      //    predecessor and a successor in the band (reachable from (0,0), reaches (lx,ly))
    int32_t pad;
};

struct DevModel {                     // log-space stateMachine5
    double tr[25];                    // [from*5+to]
    double eM[25];                    // [x*5+y], N row/col = log(1/16)
    double eX[5], eY[5];
    double endp[5], rendp[5];
    int32_t has_switch;
    int32_t pad;
};

struct DevParams {
    int32_t expansion, min_diags, tb_diags, pad;
    double threshold;
    double lp_skip;                   // log(threshold) - margin: below this exp() is not evaluated
    double gap_gamma, match_gamma;
};

struct DiagRec {                      // one live diagonal of the forward window
    int32_t off;                      // ring offset (cells in k_fwdbwd, doubles in k_fb2)
    int32_t xlo;                      // first x of the diagonal
    int32_t w;                        // cells
    int32_t pad;                      // k_fb2: 1 = the total probability is evaluated on this diagonal
};

// Diagonals on which a traceback window evaluates the total probability: every
// 10th, counted down from the window's first posterior diagonal.
constexpr int TOTAL_EVERY = 10;

#define PHMM_NEG_INF (__longlong_as_double(0xfff0000000000000LL))

// The dynamic shared-memory array of a kernel.  tests/tools/warp_emu/ (the host emulation the CPU tests run kernel
// sources under) defines it as a pointer to a host buffer before including the kernel headers.
#ifndef PHMM_DYN_SHARED
#define PHMM_DYN_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// ---------------------------------------------------------------------------
// logAdd: sonLib's piecewise-cubic log(exp(x)+exp(y)), Horner
// steps fused.  Branch-free; bit-exact with the scalar definition:
//   x<y ? (x==-inf || y-x>=7.5 ? y : lookup(y-x)+x)
//       : (y==-inf || x-y>=7.5 ? x : lookup(x-y)+y)
// ---------------------------------------------------------------------------
__device__ __forceinline__ double lookup_cubic(double t) {
    double c3, c2, c1, c0;
    if (t <= 1.00) {
        c3 = -0.009350833524763; c2 = 0.130659527668286; c1 = 0.498799810682272; c0 = 0.693203116424741;
    } else if (t <= 2.50) {
        c3 = -0.014532321752540; c2 = 0.139942324101744; c1 = 0.495635523139337; c0 = 0.692140569840976;
    } else if (t <= 4.50) {
        c3 = -0.004605031767994; c2 = 0.063427417320019; c1 = 0.695956496475118; c0 = 0.514272634594009;
    } else {
        c3 = -0.000458661602210; c2 = 0.009695946122598; c1 = 0.930734667215156; c0 = 0.168037164329057;
    }
    return fma(fma(fma(c3, t, c2), t, c1), t, c0);
}

__device__ __forceinline__ double logadd(double x, double y) {
    const double d = x - y;                    // NaN only when both are -inf
    const bool lt = x < y;
    const double mn = lt ? x : y;
    const double mx = lt ? y : x;
    const double ad = fabs(d);                 // larger - smaller, exactly
    const double r = lookup_cubic(ad) + mn;
    return (ad < 7.5) ? r : mx;                // ad is +inf / NaN when the smaller operand is -inf
}

// exp() from IEEE primitives only; same operation sequence as the checker's.
__device__ __forceinline__ double exp_det(double x) {
    if (!(x > -700.0)) return 0.0;
    if (x > 700.0) return __longlong_as_double(0x7ff0000000000000LL);
    const double SHIFT = 6755399441055744.0;
    double t = fma(x, 1.4426950408889634, SHIFT);
    double kd = t - SHIFT;
    int k = __double2loint(t);
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// ---------------------------------------------------------------------------
// Band iterator (band_construct).  Between consecutive anchors
// p and n the band is the rectangle [px-e/2, nx+e/2] x [py-e/2, ny+e/2]
// clipped to the matrix; diagonal d of that interval covers
// x in [max(xL, d-yL), min(xU, d-yU)].  All threads of a block step an
// identical copy (warp-uniform integer work).
// ---------------------------------------------------------------------------
struct BandIter {
    const Run *runs;
    int nrun, ri, k;
    int rx, ry, rn;       // current run
    int lx, ly, e2;
    int nx, ny, nxay;
    int xL, yL, xU, yU;

    __device__ __forceinline__ void load_run() {
        if (ri < nrun) { Run r = runs[ri]; rx = r.x; ry = r.y; rn = r.n; }
    }
    __device__ __forceinline__ void advance() {
        const int px = nx, py = ny;
        if (ri < nrun) {
            nx = rx + k + 1; ny = ry + k + 1;
            if (++k == rn) { ri++; k = 0; load_run(); }
        } else { nx = lx; ny = ly; }
        nxay = nx + ny;
        xL = min(max(px - e2, 0), lx);
        yL = min(max(ny + e2, 0), ly);
        xU = min(max(nx + e2, 0), lx);
        yU = min(max(py - e2, 0), ly);
    }
    __device__ __forceinline__ void init(const Run *runs_, int nrun_, int lx_, int ly_, int expansion) {
        runs = runs_; nrun = nrun_; ri = 0; k = 0; lx = lx_; ly = ly_; e2 = expansion >> 1;
        nx = 0; ny = 0; rx = ry = rn = 0;
        load_run();
        advance();                       // diagonal 0 is the single cell (0,0); interval 0 starts at d = 1
    }
    // diagonal d >= 1, must be called with d increasing by one
    __device__ __forceinline__ void diag(int d, int &xlo, int &w) {
        xlo = max(xL, d - yL);
        const int xhi = min(xU, d - yU);
        w = xhi - xlo + 1;
        if (d == nxay) advance();
    }
};

}  // namespace phmm
