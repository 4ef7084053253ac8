// phmm_fb2.cuh -- k_fb2: the windowed forward / backward / posterior kernel of the realignment path.
//
// One thread block per DP region (work queue, longest first): NW compute warps + 1 producer warp.
//
// Compute warps.  Per anti-diagonal each thread owns one cell (x = xlo + tid, + NC, ...).  The two
// previous diagonals live in shared memory as [column][5 states] and are updated in place: cell (d, x)
// sits in column (x - (d >> 1)) & mask of the buffer of parity d & 1, which is exactly the column of its
// `middle` predecessor (d-2, x-1), read by the same thread just before it is overwritten.  Diagonals
// wider than the shared-memory buffer fall back to a global buffer of the same shape.
//
// Producer warp.  Walks the band geometry (anchor runs -> first x and width of every diagonal), allocates
// the diagonal's space in the HBM ring, decides whether the window will evaluate the total probability on
// it, and publishes that record a few diagonals ahead through a shared-memory FIFO; on the way back it
// prefetches the records of the live window from HBM into the same kind of FIFO.  The compute warps read
// one 16-byte record per diagonal instead of re-deriving the geometry 256 times.
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

constexpr int FB2_RQ = 16;          // record FIFO entries (power of two)
constexpr int FB2_PRE = 6;          // how many diagonals the producer runs ahead (< FB2_RQ - 2)
constexpr int REC_TOT = 1;          // DiagRec::pad bits
constexpr int REC_WIDE = 2;

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
    double *wide;   int32_t wg;  // 4 x wg x 5 doubles: F even/odd, B even/odd for diagonals wider than wcap
    double *fsave;               // 2 x wcap x 5 doubles: forward state across a traceback window
    double *totals; int32_t tcap;
    int32_t wcap;                // shared-memory columns (power of two)
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

__device__ __forceinline__ double ldc(const double *col, int s, bool ok) {
    return ok ? col[s] : PHMM_NEG_INF;
}

// Forward cell from its three predecessor columns (5 states each): lower = (x-1,y), upper = (x,y-1) on
// diagonal d-1, middle = (x-1,y-1) on d-2.  Transition order of SURVEY.md A.4.
template <bool SWITCH>
__device__ __forceinline__ void fwd_cell2(const DevModel &m, const EmisTables &t, const char *ctab,
                                          const double *pl, bool okl, const double *pu, bool oku,
                                          const double *pm, bool okm, int cX, int cY, double out[NS]) {
    const double eXc = t.eX[cX], eYc = t.eY[cY], eMc = t.eM[cX * 5 + cY];
    {
        const double Ml = ldc(pl, S_M, okl), sXl = ldc(pl, S_SX, okl), lXl = ldc(pl, S_LX, okl);
        double a = Ml + (eXc + m.tr[S_M * 5 + S_SX]);
        a = logadd_t(a, sXl + (eXc + m.tr[S_SX * 5 + S_SX]), ctab);
        if (SWITCH) { const double sYl = ldc(pl, S_SY, okl); a = logadd_t(a, sYl + (eXc + m.tr[S_SY * 5 + S_SX]), ctab); }
        out[S_SX] = a;
        double b = Ml + (eXc + m.tr[S_M * 5 + S_LX]);
        b = logadd_t(b, lXl + (eXc + m.tr[S_LX * 5 + S_LX]), ctab);
        out[S_LX] = b;
    }
    {
        double a = ldc(pm, S_M, okm) + (eMc + m.tr[S_M * 5 + S_M]);
        a = logadd_t(a, ldc(pm, S_SX, okm) + (eMc + m.tr[S_SX * 5 + S_M]), ctab);
        a = logadd_t(a, ldc(pm, S_SY, okm) + (eMc + m.tr[S_SY * 5 + S_M]), ctab);
        a = logadd_t(a, ldc(pm, S_LX, okm) + (eMc + m.tr[S_LX * 5 + S_M]), ctab);
        a = logadd_t(a, ldc(pm, S_LY, okm) + (eMc + m.tr[S_LY * 5 + S_M]), ctab);
        out[S_M] = a;
    }
    {
        const double Mu = ldc(pu, S_M, oku), sYu = ldc(pu, S_SY, oku), lYu = ldc(pu, S_LY, oku);
        double a = Mu + (eYc + m.tr[S_M * 5 + S_SY]);
        a = logadd_t(a, sYu + (eYc + m.tr[S_SY * 5 + S_SY]), ctab);
        if (SWITCH) { const double sXu = ldc(pu, S_SX, oku); a = logadd_t(a, sXu + (eYc + m.tr[S_SX * 5 + S_SY]), ctab); }
        out[S_SY] = a;
        double b = Mu + (eYc + m.tr[S_M * 5 + S_LY]);
        b = logadd_t(b, lYu + (eYc + m.tr[S_LY * 5 + S_LY]), ctab);
        out[S_LY] = b;
    }
}

// Backward cell from its three successor columns: pu = (x, y+1), pl = (x+1, y) on diagonal d+1,
// pm = (x+1, y+1) on d+2.  cXn = X[x], cYn = Y[y]: the symbols those steps consume.
template <bool SWITCH>
__device__ __forceinline__ void bwd_cell2(const DevModel &m, const EmisTables &t, const char *ctab,
                                          const double *pl, bool okl, const double *pu, bool oku,
                                          const double *pm, bool okm, int cXn, int cYn, double out[NS]) {
    const double eXn = t.eX[cXn], eYn = t.eY[cYn], eMn = t.eM[cXn * 5 + cYn];
    const double Bm = ldc(pm, S_M, okm);
    const double BsY = ldc(pu, S_SY, oku), BlY = ldc(pu, S_LY, oku);
    const double BsX = ldc(pl, S_SX, okl), BlX = ldc(pl, S_LX, okl);
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

// Producer lane state (shared memory, touched by one thread): band walk, ring allocator, schedule of
// total-probability diagonals (needs the next two traceback points).
struct ProdState {
    BandIter it;
    int d;                      // last diagonal generated
    int roff, rsz;              // ring entry of that diagonal
    int tk, P, TF, Pn, TFn;
};

__device__ __noinline__ void fb2_produce_init(ProdState *ps, const Run *runs, int nrun, int lx, int ly, int expansion,
                                              const int32_t *tb, int ntb, int tbd) {
    const int nd = lx + ly;
    BandIter it;
    it.init(runs, nrun, lx, ly, expansion);
    ps->it = it;
    ps->d = 0; ps->roff = 0; ps->rsz = 0; ps->tk = 0;
    const int P = tb[0];
    ps->P = P;
    ps->TF = P - (P == nd ? 0 : tbd);
    const int Pn = ntb > 1 ? tb[1] : nd;
    ps->Pn = Pn;
    ps->TFn = Pn - (Pn == nd ? 0 : tbd);
}

// generates the records of the next `count` diagonals (stops at nd)
__device__ __noinline__ void fb2_produce(ProdState *ps, int count, int nd, const int32_t *tb, int ntb, int tbd,
                                         int64_t ring_doubles, int wcap, DiagRec *srec, DiagRec *dt, int dcap) {
    BandIter it = ps->it;
    int d = ps->d, roff = ps->roff, rsz = ps->rsz, tk = ps->tk, P = ps->P, TF = ps->TF, Pn = ps->Pn, TFn = ps->TFn;
    for (int k = 0; k < count && d < nd; k++) {
        d++;
        int xlo, w;
        it.diag(d, xlo, w);
        const int tf = d <= TF ? TF : TFn;
        const bool tot = (tf - d) % TOTAL_EVERY == 0;
        const int es = w * (tot ? 7 : 2);
        int off = roff + rsz;
        if ((int64_t)off + es > ring_doubles) off = 0;
        roff = off; rsz = es;
        if (d == P) {
            tk++;
            P = Pn; TF = TFn;
            Pn = tk + 1 < ntb ? tb[tk + 1] : nd;
            TFn = Pn - (Pn == nd ? 0 : tbd);
        }
        DiagRec rc; rc.off = off; rc.xlo = xlo; rc.w = w; rc.pad = (tot ? REC_TOT : 0) | (w > wcap ? REC_WIDE : 0);
        srec[d & (FB2_RQ - 1)] = rc;
        dt[d % dcap] = rc;
    }
    ps->it = it;
    ps->d = d; ps->roff = roff; ps->rsz = rsz; ps->tk = tk; ps->P = P; ps->TF = TF; ps->Pn = Pn; ps->TFn = TFn;
}

constexpr int fb2_min_blocks(int nw) { return nw == 8 ? 2 : (nw == 4 ? 4 : 6); }

template <int NW, bool SWITCH>
__global__ void __launch_bounds__((NW + 1) * 32, fb2_min_blocks(NW)) k_fb2(const __grid_constant__ Fb2Args a) {
    constexpr int NC = NW * 32;            // compute threads
    constexpr int NTA = NC + 32;           // + producer warp
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    const bool producer = tid >= NC;
    const int wcap = a.wcap, cmask = wcap - 1;
    double *const sbuf = reinterpret_cast<double *>(smem_raw);                  // [2][wcap][5]
    double *const sct = sbuf + 2 * NS * wcap;                                    // 4 rows x (c3 c2 c1 c0)
    EmisTables &tab = *reinterpret_cast<EmisTables *>(sct + 16);
    DiagRec *const srec = reinterpret_cast<DiagRec *>(sct + 16 + 36);            // [FB2_RQ]
    __shared__ int s_region;
    __shared__ int s_npairs;
    __shared__ ProdState s_prod;

    if (tid == 0) {
        sct[0] = -0.009350833524763; sct[1] = 0.130659527668286; sct[2] = 0.498799810682272; sct[3] = 0.693203116424741;
        sct[4] = -0.014532321752540; sct[5] = 0.139942324101744; sct[6] = 0.495635523139337; sct[7] = 0.692140569840976;
        sct[8] = -0.004605031767994; sct[9] = 0.063427417320019; sct[10] = 0.695956496475118; sct[11] = 0.514272634594009;
        sct[12] = -0.000458661602210; sct[13] = 0.009695946122598; sct[14] = 0.930734667215156; sct[15] = 0.168037164329057;
    }
    for (int i = tid; i < 25; i += NTA) tab.eM[i] = a.m.eM[i];
    if (tid < 5) { tab.eX[tid] = a.m.eX[tid]; tab.eY[tid] = a.m.eY[tid]; }
    const char *const ctab = reinterpret_cast<const char *>(sct);

    const int slot = blockIdx.x;
    double *const ring = a.ring + (int64_t)slot * a.ring_doubles;
    DiagRec *const dt = a.dtab + (int64_t)slot * a.dcap;
    double *const wide = a.wide + (int64_t)slot * 4 * NS * a.wg;
    double *const fsave = a.fsave + (int64_t)slot * 2 * NS * wcap;
    double *const totals = a.totals + (int64_t)slot * a.tcap;
    const int wgmask = a.wg - 1;
    const int tbd = a.p.tb_diags + 1;

    // column of cell x of a diagonal with half-index h, in shared memory or in the wide buffer `wb` (0..3)
    auto col_ptr = [&](bool is_wide, int par, int wb, int x, int h) -> double * {
        if (!is_wide) return sbuf + ((par * wcap) + ((x - h) & cmask)) * NS;
        return wide + ((int64_t)(wb + par) * a.wg + ((x - h) & wgmask)) * NS;
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
            if (tid == NC) {
                fb2_produce_init(&s_prod, a.runs + reg.run0, reg.nrun, lx, ly, a.p.expansion, tb, ntb, tbd);
                fb2_produce(&s_prod, FB2_PRE, nd, tb, ntb, tbd, a.ring_doubles, wcap, srec, dt, a.dcap);
            }
            // diagonal 0: the single cell (0,0), column 0 of the even buffer
            if (tid < NS) {
                double v;
                if (reg.ragged_left) v = (tid == S_LX || tid == S_LY) ? 0.0 : PHMM_NEG_INF;
                else v = (tid == S_M) ? 0.0 : PHMM_NEG_INF;
                sbuf[tid] = v;
            }
            int xlo1 = 0, w1 = 1, f1 = 0;                             // diagonal d-1
            int xlo2 = 0, w2 = 0, f2 = 0;                             // diagonal d-2 (w2 = 0: absent)
            int traced_to = 0;
            int tk = 0;                                               // upcoming traceback point (all threads)
            int P = tb[0];
            __syncthreads();
            for (int d = 1; d <= nd; d++) {
                const DiagRec rc = srec[d & (FB2_RQ - 1)];
                const int xlo = rc.xlo, w = rc.w;
                if (producer) {
                    if (tid == NC) fb2_produce(&s_prod, 1, nd, tb, ntb, tbd, a.ring_doubles, wcap, srec, dt, a.dcap);   // diagonal d + FB2_PRE
                } else {
                    const bool tot = (rc.pad & REC_TOT) != 0;
                    const int h0 = d >> 1, h1 = (d - 1) >> 1;
                    double *const rg = ring + rc.off;
                    const int par = d & 1;
                    if (!((rc.pad | f1 | f2) & REC_WIDE)) {
                        // fast path: the three diagonals are in shared memory
                        const double *const b1 = sbuf + (par ^ 1) * wcap * NS;
                        double *const b0 = sbuf + par * wcap * NS;
                        for (int i = tid; i < w; i += NC) {
                            const int x = xlo + i, y = d - x;
                            const int cX = x >= 1 ? X[x - 1] : 4;
                            const int cY = y >= 1 ? Y[y - 1] : 4;
                            const bool okl = (unsigned)(x - 1 - xlo1) < (unsigned)w1;
                            const bool oku = (unsigned)(x - xlo1) < (unsigned)w1;
                            const bool okm = (unsigned)(x - 1 - xlo2) < (unsigned)w2;
                            double *const p0 = b0 + ((x - h0) & cmask) * NS;              // own column == middle's
                            const double *const pl = b1 + ((x - 1 - h1) & cmask) * NS;
                            const double *const pu = b1 + ((x - h1) & cmask) * NS;
                            double o[NS];
                            fwd_cell2<SWITCH>(a.m, tab, ctab, pl, okl, pu, oku, p0, okm, cX, cY, o);
#pragma unroll
                            for (int s = 0; s < NS; s++) p0[s] = o[s];
                            rg[i] = o[S_M];
                            if (tot) {
#pragma unroll
                                for (int s = 1; s < NS; s++) rg[(s + 1) * w + i] = o[s];
                            }
                        }
                    } else {
                        const bool wd0 = (rc.pad & REC_WIDE) != 0, wd1 = (f1 & REC_WIDE) != 0, wd2 = (f2 & REC_WIDE) != 0;
                        const int h2 = (d - 2) >> 1;
                        for (int i = tid; i < w; i += NC) {
                            const int x = xlo + i, y = d - x;
                            const int cX = x >= 1 ? X[x - 1] : 4;
                            const int cY = y >= 1 ? Y[y - 1] : 4;
                            const bool okl = (unsigned)(x - 1 - xlo1) < (unsigned)w1;
                            const bool oku = (unsigned)(x - xlo1) < (unsigned)w1;
                            const bool okm = (unsigned)(x - 1 - xlo2) < (unsigned)w2;
                            double *const p0 = col_ptr(wd0, par, 0, x, h0);
                            const double *const pl = col_ptr(wd1, par ^ 1, 0, x - 1, h1);
                            const double *const pu = col_ptr(wd1, par ^ 1, 0, x, h1);
                            const double *const pm = col_ptr(wd2, par, 0, x - 1, h2);
                            double o[NS];
                            fwd_cell2<SWITCH>(a.m, tab, ctab, pl, okl, pu, oku, pm, okm, cX, cY, o);
#pragma unroll
                            for (int s = 0; s < NS; s++) p0[s] = o[s];
                            rg[i] = o[S_M];
                            if (tot) {
#pragma unroll
                                for (int s = 1; s < NS; s++) rg[(s + 1) * w + i] = o[s];
                            }
                        }
                    }
                }
                __syncthreads();
                if (d == P) {
                    // ------------------------- traceback window (traced_to, d] -------------------------
                    const bool at_end = d == nd;
                    const int traced_from = d - (at_end ? 0 : tbd);
                    const double *endv = (at_end && !reg.ragged_right) ? a.m.endp : a.m.rendp;
                    DiagRec *const srb = srec + FB2_RQ;                                  // backward FIFO
                    if (!at_end) {
                        for (int i = tid; i < 2 * NS * wcap; i += NTA) fsave[i] = sbuf[i];
                    }
                    if (producer) {
                        for (int k = 0; k < FB2_PRE; k++) {
                            const int dd = d - k;
                            if (dd > traced_to && tid == NC) srb[dd & (FB2_RQ - 1)] = dt[dd % a.dcap];
                        }
                    }
                    __syncthreads();
                    // phase 1: backward sweep
                    {
                        int bxlo1 = 0, bw1 = 0, bf1 = 0, bxlo2 = 0, bw2 = 0, bf2 = 0;      // diagonals dd+1, dd+2
                        for (int dd = d; dd > traced_to; dd--) {
                            const DiagRec rb = srb[dd & (FB2_RQ - 1)];
                            if (producer) {
                                const int dn = dd - FB2_PRE;
                                if (dn > traced_to && tid == NC) srb[dn & (FB2_RQ - 1)] = dt[dn % a.dcap];
                            } else {
                                const int h0 = dd >> 1, h1 = (dd + 1) >> 1;
                                const int par = dd & 1;
                                double *const rg = ring + rb.off;
                                const bool dots = (rb.pad & REC_TOT) != 0 && dd <= traced_from;
                                const bool fastb = !((rb.pad | bf1 | bf2) & REC_WIDE);
                                const bool wd0 = (rb.pad & REC_WIDE) != 0, wd1 = (bf1 & REC_WIDE) != 0, wd2 = (bf2 & REC_WIDE) != 0;
                                const int h2 = (dd + 2) >> 1;
                                for (int i = tid; i < rb.w; i += NC) {
                                    const int x = rb.xlo + i, y = dd - x;
                                    double o[NS];
                                    double *p0;
                                    if (fastb) p0 = sbuf + (par * wcap + ((x - h0) & cmask)) * NS;
                                    else p0 = col_ptr(wd0, par, 2, x, h0);
                                    if (dd < d) {
                                        const int cXn = x < lx ? X[x] : 4;
                                        const int cYn = y < ly ? Y[y] : 4;
                                        const bool oku = (unsigned)(x - bxlo1) < (unsigned)bw1;
                                        const bool okl = (unsigned)(x + 1 - bxlo1) < (unsigned)bw1;
                                        const bool okm = (unsigned)(x + 1 - bxlo2) < (unsigned)bw2;
                                        if (fastb) {
                                            const double *const b1 = sbuf + (par ^ 1) * wcap * NS;
                                            const double *const pu = b1 + ((x - h1) & cmask) * NS;
                                            const double *const pl = b1 + ((x + 1 - h1) & cmask) * NS;
                                            bwd_cell2<SWITCH>(a.m, tab, ctab, pl, okl, pu, oku, p0, okm, cXn, cYn, o);
                                        } else {
                                            const double *const pu = col_ptr(wd1, par ^ 1, 2, x, h1);
                                            const double *const pl = col_ptr(wd1, par ^ 1, 2, x + 1, h1);
                                            const double *const pm = col_ptr(wd2, par, 2, x + 1, h2);
                                            bwd_cell2<SWITCH>(a.m, tab, ctab, pl, okl, pu, oku, pm, okm, cXn, cYn, o);
                                        }
                                    } else {
#pragma unroll
                                        for (int s = 0; s < NS; s++) o[s] = endv[s];
                                    }
#pragma unroll
                                    for (int s = 0; s < NS; s++) p0[s] = o[s];
                                    rg[rb.w + i] = o[S_M];
                                    if (dots) {
                                        double t = rg[i] + o[S_M];
#pragma unroll
                                        for (int s = 1; s < NS; s++) t = logadd_t(t, rg[(s + 1) * rb.w + i] + o[s], ctab);
                                        rg[6 * rb.w + i] = t;
                                    }
                                }
                            }
                            bxlo2 = bxlo1; bw2 = bw1; bf2 = bf1;
                            bxlo1 = rb.xlo; bw1 = rb.w; bf1 = rb.pad;
                            __syncthreads();
                        }
                    }
                    // phase 2: total probabilities, one thread per total diagonal
                    const int nk = traced_from > traced_to ? (traced_from - traced_to - 1) / TOTAL_EVERY + 1 : 0;
                    for (int k = tid; k < nk; k += NTA) {
                        const int dd = traced_from - TOTAL_EVERY * k;
                        const DiagRec r0 = dt[dd % a.dcap];
                        const double *cd = ring + r0.off + 6 * r0.w;
                        double total = fold_seq(r0.w, ctab, [&](int i) { return cd[i]; });
                        if (dd < d) {
                            const DiagRec r1 = dt[(dd + 1) % a.dcap];
                            const double *fm = ring + r1.off, *bm = fm + r1.w;
                            const double t1 = fold_seq(r1.w, ctab, [&](int i) { return fm[i] + bm[i]; });
                            total = logadd_t(total, t1, ctab);
                        }
                        totals[k] = total;
                    }
                    __syncthreads();
                    // phase 3: posterior match probabilities, one warp per diagonal
                    for (int dd = traced_from - (tid >> 5); dd > traced_to; dd -= NW + 1) {
                        const DiagRec r0 = dt[dd % a.dcap];
                        const double total = totals[(traced_from - dd) / TOTAL_EVERY];
                        const double *fm = ring + r0.off, *bm = fm + r0.w;
                        for (int i = tid & 31; i < r0.w; i += 32) {
                            const int x = r0.xlo + i, y = dd - x;
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
                        for (int i = tid; i < 2 * NS * wcap; i += NTA) sbuf[i] = fsave[i];
                        __syncthreads();
                    }
                    traced_to = traced_from;
                    tk++;
                    P = tk < ntb ? tb[tk] : nd + 1;
                }
                xlo2 = xlo1; w2 = w1; f2 = f1;
                xlo1 = xlo; w1 = w; f1 = rc.pad;
            }
        }
        __syncthreads();
        if (tid == 0) a.npairs[ridx] = s_npairs;
    }
}

}  // namespace phmm
