// phmm_fb2.cuh -- k_fb2: the windowed forward / backward / posterior kernel of the realignment path.
//
// One thread block per DP region (work queue, longest first).  Per anti-diagonal the block's threads
// each own one cell (x = xlo + tid, + NT, ...) and the two previous diagonals live in shared memory,
// updated in place: cell (d, x) sits in column (x - (d >> 1)) & mask of the buffer of parity d & 1, which
// is exactly the column of its `middle` predecessor (d-2, x-1), read by the same thread just before.
// Diagonals wider than the shared-memory buffer fall back to a global buffer of the same shape.
//
// HBM holds, per live diagonal, only what a later phase needs (the "ring", recycled per traceback window):
//     F_M  (forward match state)        every diagonal        -> posterior, step-over term of the total
//     B_M  (backward match state)       every diagonal        -> posterior, step-over term of the total
//     F_sX F_sY F_lX F_lY + cell dots   every 10th diagonal   -> total probability
// i.e. 2 doubles per cell, 7 on the diagonals where the window evaluates the total probability.
//
// A traceback window runs in four block-wide phases:
//   1 backward sweep from the traceback diagonal, storing B_M and the per-cell dot products,
//   2 total probabilities: one thread per total diagonal folds the dots left to right (the scalar
//     order; logAdd is not associative), independent diagonals in parallel,
//   3 posterior match probabilities >= threshold -> pairs,
//   4 forward state restored, sweep continues.
//
// Replaces the body of `cactus_realign` (reference nanopore/analyses/utils.py:587).  fp64 throughout,
// every operation in the association order of the scalar definition: bit-exact against the CPU checker.
#pragma once
#include "phmm_kernels.cuh"

namespace phmm {

struct Fb2Args {
    const uint8_t *ref;
    const uint8_t *reads;
    const Region *regions;
    const Run *runs;
    const int32_t *order;
    int32_t n_regions;
    int32_t *counter;
    DevModel m;
    DevParams p;
    const int64_t *tb_off;       // traceback points of region r: tbp[tb_off[r] .. tb_off[r+1])
    const int32_t *tbp;
    const int32_t *ntb;          // RegionGeom::tracebacks, strided
    int32_t ntb_stride;          // in int32 units
    // per-slot scratch (slot = blockIdx.x)
    double *ring;   int64_t ring_doubles;
    DiagRec *dtab;  int32_t dcap;
    double *wide;   int32_t wg;  // 4 x 5 x wg doubles: F even/odd, B even/odd for diagonals wider than wcap
    double *fsave;               // 2 x 5 x wcap doubles: forward state across a traceback window
    double *totals; int32_t tcap;
    int32_t wcap;                // shared-memory columns per state (power of two)
    // outputs
    int32_t *px, *py, *pw;
    int32_t *npairs;
};

// log(exp(x)+exp(y)), sonLib's piecewise cubic.  Same value, bit for bit, as
// phmm::logadd; written for the instruction mix that is fastest on sm_100a (measured with
// scripts/probes/logadd_probe.cu): threshold tests on the fp64 pipe, coefficients from a 128-byte
// shared-memory table (one row per segment), ordering and range tests on the integer pipe.
__device__ __forceinline__ double logadd_t(double x, double y, const char *ctab) {
    const double d = x - y;
    const int dh = __double2hiint(d);
    const bool lt = dh < 0;                      // x < y  (d is NaN only when both are -inf: either order is right)
    const double mn = lt ? x : y, mx = lt ? y : x;
    const double t = fabs(d);
    int off = 0;
    if (t > 1.0) off = 32;
    if (t > 2.5) off = 64;
    if (t > 4.5) off = 96;
    const double2 c32 = *reinterpret_cast<const double2 *>(ctab + off);
    const double2 c10 = *reinterpret_cast<const double2 *>(ctab + off + 16);
    const double r = fma(fma(fma(c32.x, t, c32.y), t, c10.x), t, c10.y) + mn;
    return ((unsigned)dh * 2u < 0x401E0000u * 2u) ? r : mx;      // |d| < 7.5; false for inf / NaN
}

struct DBuf {                   // where one diagonal's 5 x w values live
    double *p;                  // state s at p + s * cap
    int cap, mask;
};

template <bool SWITCH>
__device__ __forceinline__ void fwd_cell2(const DevModel &m, const EmisTables &t, const char *ctab,
                                          const DBuf &b1, int xlo1, int w1, int h1,
                                          const DBuf &b2, int xlo2, int w2, int h2,
                                          int x, int cX, int cY, double out[NS]) {
    const bool okl = (unsigned)(x - 1 - xlo1) < (unsigned)w1;
    const bool oku = (unsigned)(x - xlo1) < (unsigned)w1;
    const bool okm = (unsigned)(x - 1 - xlo2) < (unsigned)w2;
    const int cl = (x - 1 - h1) & b1.mask;
    const int cu = (x - h1) & b1.mask;
    const int cm = (x - 1 - h2) & b2.mask;
    const double eXc = t.eX[cX], eYc = t.eY[cY], eMc = t.eM[cX * 5 + cY];
    const double *F1 = b1.p, *F2 = b2.p;
    const int c1 = b1.cap, c2 = b2.cap;
    {
        const double Ml = ldv(F1, cl, okl), sXl = ldv(F1 + c1, cl, okl), lXl = ldv(F1 + 3 * c1, cl, okl);
        double a = Ml + (eXc + m.tr[S_M * 5 + S_SX]);
        a = logadd_t(a, sXl + (eXc + m.tr[S_SX * 5 + S_SX]), ctab);
        if (SWITCH) { const double sYl = ldv(F1 + 2 * c1, cl, okl); a = logadd_t(a, sYl + (eXc + m.tr[S_SY * 5 + S_SX]), ctab); }
        out[S_SX] = a;
        double b = Ml + (eXc + m.tr[S_M * 5 + S_LX]);
        b = logadd_t(b, lXl + (eXc + m.tr[S_LX * 5 + S_LX]), ctab);
        out[S_LX] = b;
    }
    {
        double a = ldv(F2, cm, okm) + (eMc + m.tr[S_M * 5 + S_M]);
        a = logadd_t(a, ldv(F2 + c2, cm, okm) + (eMc + m.tr[S_SX * 5 + S_M]), ctab);
        a = logadd_t(a, ldv(F2 + 2 * c2, cm, okm) + (eMc + m.tr[S_SY * 5 + S_M]), ctab);
        a = logadd_t(a, ldv(F2 + 3 * c2, cm, okm) + (eMc + m.tr[S_LX * 5 + S_M]), ctab);
        a = logadd_t(a, ldv(F2 + 4 * c2, cm, okm) + (eMc + m.tr[S_LY * 5 + S_M]), ctab);
        out[S_M] = a;
    }
    {
        const double Mu = ldv(F1, cu, oku), sYu = ldv(F1 + 2 * c1, cu, oku), lYu = ldv(F1 + 4 * c1, cu, oku);
        double a = Mu + (eYc + m.tr[S_M * 5 + S_SY]);
        a = logadd_t(a, sYu + (eYc + m.tr[S_SY * 5 + S_SY]), ctab);
        if (SWITCH) { const double sXu = ldv(F1 + c1, cu, oku); a = logadd_t(a, sXu + (eYc + m.tr[S_SX * 5 + S_SY]), ctab); }
        out[S_SY] = a;
        double b = Mu + (eYc + m.tr[S_M * 5 + S_LY]);
        b = logadd_t(b, lYu + (eYc + m.tr[S_LY * 5 + S_LY]), ctab);
        out[S_LY] = b;
    }
}

template <bool SWITCH>
__device__ __forceinline__ void bwd_cell2(const DevModel &m, const EmisTables &t, const char *ctab,
                                          const DBuf &b1, int xlo1, int w1, int h1,
                                          const DBuf &b2, int xlo2, int w2, int h2,
                                          int x, int cXn, int cYn, double out[NS]) {
    const bool oku = (unsigned)(x - xlo1) < (unsigned)w1;          // successor (x, y+1)
    const bool okl = (unsigned)(x + 1 - xlo1) < (unsigned)w1;      // successor (x+1, y)
    const bool okm = (unsigned)(x + 1 - xlo2) < (unsigned)w2;      // successor (x+1, y+1)
    const int cu = (x - h1) & b1.mask;
    const int cl = (x + 1 - h1) & b1.mask;
    const int cm = (x + 1 - h2) & b2.mask;
    const double eXn = t.eX[cXn], eYn = t.eY[cYn], eMn = t.eM[cXn * 5 + cYn];
    const double *B1 = b1.p, *B2 = b2.p;
    const int c1 = b1.cap;
    const double Bm = ldv(B2, cm, okm);
    const double BsY = ldv(B1 + 2 * c1, cu, oku), BlY = ldv(B1 + 4 * c1, cu, oku);
    const double BsX = ldv(B1 + c1, cl, okl), BlX = ldv(B1 + 3 * c1, cl, okl);
    {
        double a = Bm + (eMn + m.tr[S_M * 5 + S_M]);
        a = logadd_t(a, BsY + (eYn + m.tr[S_M * 5 + S_SY]), ctab);
        a = logadd_t(a, BlY + (eYn + m.tr[S_M * 5 + S_LY]), ctab);
        a = logadd_t(a, BsX + (eXn + m.tr[S_M * 5 + S_SX]), ctab);
        a = logadd_t(a, BlX + (eXn + m.tr[S_M * 5 + S_LX]), ctab);
        out[S_M] = a;
    }
    {
        double a = Bm + (eMn + m.tr[S_SX * 5 + S_M]);
        if (SWITCH) a = logadd_t(a, BsY + (eYn + m.tr[S_SX * 5 + S_SY]), ctab);
        a = logadd_t(a, BsX + (eXn + m.tr[S_SX * 5 + S_SX]), ctab);
        out[S_SX] = a;
    }
    {
        double a = Bm + (eMn + m.tr[S_SY * 5 + S_M]);
        a = logadd_t(a, BsY + (eYn + m.tr[S_SY * 5 + S_SY]), ctab);
        if (SWITCH) a = logadd_t(a, BsX + (eXn + m.tr[S_SY * 5 + S_SX]), ctab);
        out[S_SY] = a;
    }
    {
        double a = Bm + (eMn + m.tr[S_LX * 5 + S_M]);
        a = logadd_t(a, BlX + (eXn + m.tr[S_LX * 5 + S_LX]), ctab);
        out[S_LX] = a;
    }
    {
        double a = Bm + (eMn + m.tr[S_LY * 5 + S_M]);
        a = logadd_t(a, BlY + (eYn + m.tr[S_LY * 5 + S_LY]), ctab);
        out[S_LY] = a;
    }
}

// left-to-right logAdd fold of n values produced by f(i), starting from -inf (dpDiagonal_dotProduct order)
template <typename F>
__device__ __forceinline__ double fold_seq(int n, const char *ctab, F f) {
    double t = PHMM_NEG_INF;
    for (int i = 0; i < n; i++) t = logadd_t(t, f(i), ctab);
    return t;
}

template <int NW, bool SWITCH>
__global__ void __launch_bounds__(NW * 32) k_fb2(const __grid_constant__ Fb2Args a) {
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const int wcap = a.wcap;
    double *const sbuf = reinterpret_cast<double *>(smem_raw);                  // [2][5][wcap]
    double *const sct = sbuf + 2 * NS * wcap;                                    // 4 rows x (c3 c2 c1 c0)
    EmisTables &tab = *reinterpret_cast<EmisTables *>(sct + 16);
    __shared__ int s_region;
    __shared__ int s_npairs;

    if (tid == 0) {
        sct[0] = -0.009350833524763; sct[1] = 0.130659527668286; sct[2] = 0.498799810682272; sct[3] = 0.693203116424741;
        sct[4] = -0.014532321752540; sct[5] = 0.139942324101744; sct[6] = 0.495635523139337; sct[7] = 0.692140569840976;
        sct[8] = -0.004605031767994; sct[9] = 0.063427417320019; sct[10] = 0.695956496475118; sct[11] = 0.514272634594009;
        sct[12] = -0.000458661602210; sct[13] = 0.009695946122598; sct[14] = 0.930734667215156; sct[15] = 0.168037164329057;
    }
    for (int i = tid; i < 25; i += NT) tab.eM[i] = a.m.eM[i];
    if (tid < 5) { tab.eX[tid] = a.m.eX[tid]; tab.eY[tid] = a.m.eY[tid]; }
    const char *const ctab = reinterpret_cast<const char *>(sct);

    const int slot = blockIdx.x;
    double *const ring = a.ring + (int64_t)slot * a.ring_doubles;
    DiagRec *const dt = a.dtab + (int64_t)slot * a.dcap;
    double *const wide = a.wide + (int64_t)slot * 4 * NS * a.wg;
    double *const fsave = a.fsave + (int64_t)slot * 2 * NS * wcap;
    double *const totals = a.totals + (int64_t)slot * a.tcap;
    const int e = a.p.expansion;
    const int tbd = a.p.tb_diags + 1;

    // storage of a diagonal of width w and parity par: shared memory when it fits, else the slot's wide buffer
    auto fbuf = [&](int w, int par) -> DBuf {
        DBuf b;
        if (w <= wcap) { b.p = sbuf + par * NS * wcap; b.cap = wcap; b.mask = wcap - 1; }
        else { b.p = wide + (int64_t)par * NS * a.wg; b.cap = a.wg; b.mask = a.wg - 1; }
        return b;
    };
    auto bbuf = [&](int w, int par) -> DBuf {
        DBuf b;
        if (w <= wcap) { b.p = sbuf + par * NS * wcap; b.cap = wcap; b.mask = wcap - 1; }
        else { b.p = wide + (int64_t)(2 + par) * NS * a.wg; b.cap = a.wg; b.mask = a.wg - 1; }
        return b;
    };

    for (;;) {
        __syncthreads();
        if (tid == 0) s_region = atomicAdd(a.counter, 1);
        __syncthreads();
        const int qi = s_region;
        if (qi >= a.n_regions) break;
        const int ridx = a.order[qi];
        const Region reg = a.regions[ridx];
        const uint8_t *X = a.ref + reg.xoff;
        const uint8_t *Y = a.reads + reg.yoff;
        const int lx = reg.lx, ly = reg.ly, nd = lx + ly;
        if (tid == 0) s_npairs = 0;
        if (nd > 0) {
            const int32_t *tb = a.tbp + a.tb_off[ridx];
            const int ntb = a.ntb[(int64_t)ridx * a.ntb_stride];
            int tk = 0;                                               // index of the upcoming traceback point
            int P = tb[0];
            int TF = P - (P == nd ? 0 : tbd);
            int Pn = ntb > 1 ? tb[1] : nd;
            int TFn = Pn - (Pn == nd ? 0 : tbd);
            BandIter it;
            it.init(a.runs + reg.run0, reg.nrun, lx, ly, e);
            // diagonal 0: the single cell (0,0), column 0 of the even buffer
            if (tid < NS) {
                double v;
                if (reg.ragged_left) v = (tid == S_LX || tid == S_LY) ? 0.0 : PHMM_NEG_INF;
                else v = (tid == S_M) ? 0.0 : PHMM_NEG_INF;
                sbuf[tid * wcap] = v;
            }
            int xlo1 = 0, w1 = 1;                                     // diagonal d-1
            int xlo2 = 0, w2 = 0;                                     // diagonal d-2 (w2 = 0: absent)
            int roff = 0, rsz = 0;                                    // ring entry of diagonal d-1
            int traced_to = 0;
            __syncthreads();
            for (int d = 1; d <= nd; d++) {
                int xlo, w;
                it.diag(d, xlo, w);
                const int tf = d <= TF ? TF : TFn;
                const bool tot = (tf - d) % TOTAL_EVERY == 0;
                const int es = w * (tot ? 7 : 2);
                int off = roff + rsz;
                if ((int64_t)off + es > a.ring_doubles) off = 0;
                if (tid == 0) { DiagRec rc; rc.off = off; rc.xlo = xlo; rc.w = w; rc.pad = tot ? 1 : 0; dt[d % a.dcap] = rc; }
                {
                    const DBuf b1 = fbuf(w1, (d - 1) & 1);
                    const DBuf b2 = fbuf(w2 > 0 ? w2 : 1, d & 1);
                    const DBuf b0 = fbuf(w, d & 1);
                    const int h0 = d >> 1, h1 = (d - 1) >> 1, h2 = (d - 2) >> 1;
                    double *const rg = ring + off;
                    for (int i = tid; i < w; i += NT) {
                        const int x = xlo + i, y = d - x;
                        const int cX = x >= 1 ? X[x - 1] : 4;
                        const int cY = y >= 1 ? Y[y - 1] : 4;
                        double o[NS];
                        fwd_cell2<SWITCH>(a.m, tab, ctab, b1, xlo1, w1, h1, b2, xlo2, w2, h2, x, cX, cY, o);
                        const int c0 = (x - h0) & b0.mask;
#pragma unroll
                        for (int s = 0; s < NS; s++) b0.p[s * b0.cap + c0] = o[s];
                        rg[i] = o[S_M];
                        if (tot) {
#pragma unroll
                            for (int s = 1; s < NS; s++) rg[(s + 1) * w + i] = o[s];
                        }
                    }
                }
                __syncthreads();
                if (d == P) {
                    // ------------------------- traceback window (traced_to, d] -------------------------
                    const bool at_end = d == nd;
                    const int traced_from = TF;
                    const double *endv = (at_end && !reg.ragged_right) ? a.m.endp : a.m.rendp;
                    if (!at_end) {
                        for (int i = tid; i < 2 * NS * wcap; i += NT) fsave[i] = sbuf[i];
                        __syncthreads();
                    }
                    // phase 1: backward sweep
                    {
                        int bxlo1 = 0, bw1 = 0, bxlo2 = 0, bw2 = 0;       // diagonals dd+1, dd+2
                        DiagRec rc = dt[d % a.dcap];
                        for (int dd = d; dd > traced_to; dd--) {
                            DiagRec nxt = rc;
                            if (dd - 1 > traced_to) nxt = dt[(dd - 1) % a.dcap];      // prefetch for the next step
                            const DBuf b0 = bbuf(rc.w, dd & 1);
                            const DBuf b1 = bbuf(bw1 > 0 ? bw1 : 1, (dd + 1) & 1);
                            const DBuf b2 = bbuf(bw2 > 0 ? bw2 : 1, dd & 1);
                            const int h0 = dd >> 1, h1 = (dd + 1) >> 1, h2 = (dd + 2) >> 1;
                            double *const rg = ring + rc.off;
                            const bool dots = rc.pad != 0 && dd <= traced_from;
                            for (int i = tid; i < rc.w; i += NT) {
                                const int x = rc.xlo + i, y = dd - x;
                                double o[NS];
                                if (dd < d) {
                                    const int cXn = x < lx ? X[x] : 4;
                                    const int cYn = y < ly ? Y[y] : 4;
                                    bwd_cell2<SWITCH>(a.m, tab, ctab, b1, bxlo1, bw1, h1, b2, bxlo2, bw2, h2, x, cXn, cYn, o);
                                } else {
#pragma unroll
                                    for (int s = 0; s < NS; s++) o[s] = endv[s];
                                }
                                const int c0 = (x - h0) & b0.mask;
#pragma unroll
                                for (int s = 0; s < NS; s++) b0.p[s * b0.cap + c0] = o[s];
                                rg[rc.w + i] = o[S_M];
                                if (dots) {
                                    double t = rg[i] + o[S_M];
#pragma unroll
                                    for (int s = 1; s < NS; s++) t = logadd_t(t, rg[(s + 1) * rc.w + i] + o[s], ctab);
                                    rg[6 * rc.w + i] = t;
                                }
                            }
                            bxlo2 = bxlo1; bw2 = bw1;
                            bxlo1 = rc.xlo; bw1 = rc.w;
                            rc = nxt;
                            __syncthreads();
                        }
                    }
                    // phase 2: total probabilities, one thread per total diagonal
                    const int nk = traced_from > traced_to ? (traced_from - traced_to - 1) / TOTAL_EVERY + 1 : 0;
                    for (int k = tid; k < nk; k += NT) {
                        const int dd = traced_from - TOTAL_EVERY * k;
                        const DiagRec rc = dt[dd % a.dcap];
                        const double *cd = ring + rc.off + 6 * rc.w;
                        double total = fold_seq(rc.w, ctab, [&](int i) { return cd[i]; });
                        if (dd < d) {
                            const DiagRec r1 = dt[(dd + 1) % a.dcap];
                            const double *f1 = ring + r1.off, *b1 = f1 + r1.w;
                            const double t1 = fold_seq(r1.w, ctab, [&](int i) { return f1[i] + b1[i]; });
                            total = logadd_t(total, t1, ctab);
                        }
                        totals[k] = total;
                    }
                    __syncthreads();
                    // phase 3: posterior match probabilities, one warp per diagonal
                    for (int dd = traced_from - (tid >> 5); dd > traced_to; dd -= NW) {
                        const DiagRec rc = dt[dd % a.dcap];
                        const double total = totals[(traced_from - dd) / TOTAL_EVERY];
                        const double *fm = ring + rc.off, *bm = fm + rc.w;
                        for (int i = tid & 31; i < rc.w; i += 32) {
                            const int x = rc.xlo + i, y = dd - x;
                            if (x > 0 && y > 0) {
                                const double lp = (fm[i] + bm[i]) - total;
                                if (lp >= a.p.lp_skip) {
                                    double pr = exp_det(lp);
                                    if (pr >= a.p.threshold) {
                                        if (pr > 1.0) pr = 1.0;
                                        const int wq = (int)floor(pr * (double)PROB_1);
                                        const int slotp = atomicAdd(&s_npairs, 1);
                                        if (slotp < reg.pair_cap) {
                                            a.px[reg.pair_off + slotp] = x - 1;
                                            a.py[reg.pair_off + slotp] = y - 1;
                                            a.pw[reg.pair_off + slotp] = wq;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    __syncthreads();
                    // phase 4: forward state back, next window
                    if (!at_end) {
                        for (int i = tid; i < 2 * NS * wcap; i += NT) sbuf[i] = fsave[i];
                        __syncthreads();
                    }
                    traced_to = traced_from;
                    tk++;
                    P = Pn; TF = TFn;
                    Pn = tk + 1 < ntb ? tb[tk + 1] : nd;
                    TFn = Pn - (Pn == nd ? 0 : tbd);
                }
                xlo2 = xlo1; w2 = w1;
                xlo1 = xlo; w1 = w;
                roff = off; rsz = es;
            }
        }
        __syncthreads();
        if (tid == 0) a.npairs[ridx] = s_npairs;
    }
}

}  // namespace phmm
