// phmm_fb2.cuh -- k_fb2: the windowed forward / backward / posterior kernel of the realignment path.
//
// One thread block per DP region (work queue, longest first), NW warps, all of them computing cells.
//
// Per anti-diagonal each thread owns the cells of its columns.  The two previous diagonals live in shared memory as
// [column][5 states] and are updated in place: cell (d, x) sits in column (x - (d >> 1)) & mask of the buffer of
// parity d & 1, which is exactly the column of its `middle` predecessor (d-2, x-1), read by the same thread just
// before it is overwritten.  Diagonals wider than the shared-memory buffer fall back to a global buffer of the
// same shape.
//
// Diagonal records.  The band geometry (first x and width of every diagonal), the diagonal's slot in the HBM ring,
// whether the window evaluates the total probability on it and whether it may take the unguarded fast path are
// data independent: k_records (one thread per region, run once per prepared batch) writes one 16-byte record per
// diagonal to HBM, and the first warp of each block streams them 32 at a time into a shared-memory FIFO, two
// batches ahead (the load of one batch is consumed a batch later, so nobody waits for it).  No thread of the hot
// kernel re-derives geometry.
//
// HBM holds, per live diagonal, only what a later phase needs (the "ring", recycled per traceback window):
//     F_M, overwritten by F_M + B_M on the way back   every diagonal       -> posterior, step-over term of the total
//     F_sX F_sY F_lX F_lY + cell dots                 every 10th diagonal  -> total probability
// i.e. 1 double per cell, 6 on the diagonals where the window evaluates the total probability.
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

constexpr int FB2_RQ = 64;          // record FIFO entries (power of two): two batches
constexpr int FB2_BATCH = 32;       // records streamed per refill (one per lane of the first warp)

constexpr int REC_TOT = 1;          // DiagRec::pad bits: the window evaluates the total probability here
constexpr int REC_WIDE = 2;         //   wider than the shared-memory buffer: lives in the global fallback buffer
constexpr int REC_FAST3 = 4;        //   diagonals d-2, d-1, d (+-1 column) fit the shared-memory columns without aliasing
#ifndef PHMM_CS
#define PHMM_CS 5
#endif
constexpr int CS = PHMM_CS;               // doubles per shared-memory column: 5 (five resident regions per SM at 512 columns) or 6 (padded: 16-byte loads)
constexpr int TS = 6;                     // doubles per row of the (emission + transition) tables: 5 used, 16-byte aligned pairs
constexpr int TG_S = 0, TG_SS = 1, TG_L = 2, TG_LL = 3, TG_SW = 4;   // slots of a gap row: M->s, s->s, M->l, l->l, other s->s (switch)
#ifndef PHMM_TM_T
#define PHMM_TM_T 2             // layout of the s -> M table, see tm_index / ld_mat
#endif
constexpr int TM_PAIRS = PHMM_TM_T == 2 ? 37 : 25;       // entries per state of the transposed table (2: bank-distinct pair index)
constexpr int TM_DOUBLES = PHMM_TM_T == 2 ? 5 * TM_PAIRS + 1 : 25 * TS;
constexpr int FB2_TAB = 16 + TM_DOUBLES + 5 * TS + 5 * TS;   // logAdd coefficients + the three (emission + transition) tables


// Timing experiments (scripts/tune.py) switch phases of the kernel off and make its results wrong: they exist only in
// a library built with -DPHMM_TUNE, never in the shipped one.
#ifdef PHMM_TUNE
#define FB2_DBG(bits) ((a.dbg & (bits)) != 0)
#else
#define FB2_DBG(bits) false
#endif

// Rotation of the warp -> cells assignment within a diagonal (fast paths): warp k takes the cells [32 (k ^ r), ..) with
// r = diagonal mod NW.  Hardware warp w of a block runs on scheduler w % 4; without the rotation warp 0 of every resident
// block takes the extra round of each diagonal wider than the block and its scheduler saturates while the others idle
// at the barrier (profiles/r02a_k_fb2_ncu_metrics.txt: issue active 92 % max, 30 % min over the schedulers).
#ifndef PHMM_NO_ROT
#define FB2_ROT(diag) ((NW & (NW - 1)) ? 0 : ((((diag) >> FB2_ROT_SHIFT) & (NW - 1)) << 5))
#else
#define FB2_ROT(diag) 0
#endif
#ifndef FB2_ROT_SHIFT
#define FB2_ROT_SHIFT 0
#endif

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
    const DiagRec *recs;         // records of region r: recs[rec_off[r] + d], d = 1 .. lx + ly
    const int64_t *rec_off;
    double *wide;   int32_t wg;  // 4 x wg x 5 doubles: F even/odd, B even/odd for diagonals wider than wcap
    double *fsave;               // 2 x wcap x CS doubles: forward state across a traceback window
    double *totals; int32_t tcap; // per slot: tcap totals, then wg sums F_M + B_M of the diagonal just above the posterior range
    long long *cand; int32_t ccap; // per slot: posterior candidates (diagonal << 32 | cell) of the current window
    double est_eps;                // candidates: cells within log(threshold) - est_eps of the window's first total (default 0.02)
    int32_t wcap;                // shared-memory columns (power of two)
    // outputs
    int32_t *px, *py, *pw;
    int32_t *npairs;
    unsigned long long *expT;    // EXPECT: per region 25 transition counts (2^-32 fixed point)
    unsigned long long *expE;    // EXPECT: per region 80 emission counts
    double *expLL;               // EXPECT: per region summed total log-probability
    int32_t dbg;                 // PHMM_TUNE builds only: timing experiments (results are wrong when non-zero)
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
#if PHMM_RANGE_INT
    return ((unsigned)dh * 2u < 0x401E0000u * 2u) ? r : mx;      // |d| < 7.5; false for inf / NaN
#else
    return (t < 7.5) ? r : mx;                                   // false for inf / NaN (the smaller operand is -inf)
#endif
}

// The same coefficient rows in constant memory (tuning variants: the shared-memory pipe is the busiest unit of k_fb2 --
// 9.5 wavefronts per cell, half of them these rows -- and the constant cache is a different path).
__constant__ double c_logadd[16] = {
    -0.009350833524763, 0.130659527668286, 0.498799810682272, 0.693203116424741,
    -0.014532321752540, 0.139942324101744, 0.495635523139337, 0.692140569840976,
    -0.004605031767994, 0.063427417320019, 0.695956496475118, 0.514272634594009,
    -0.000458661602210, 0.009695946122598, 0.930734667215156, 0.168037164329057};

__device__ __forceinline__ double logadd_k(double x, double y) {
    const double d = x - y;
    const int dh = __double2hiint(d);
    const bool lt = dh < 0;
    const double mn = lt ? x : y, mx = lt ? y : x;
    const double t = fabs(d);
    int off = 0;
    if (t > 1.0) off = 4;
    if (t > 2.5) off = 8;
    if (t > 4.5) off = 12;
    const double r = fma(fma(fma(c_logadd[off], t, c_logadd[off + 1]), t, c_logadd[off + 2]), t, c_logadd[off + 3]) + mn;
    return (t < 7.5) ? r : mx;
}

#ifndef PHMM_COEF_CONST
#define PHMM_COEF_CONST 0       // 0: rows from shared memory; 1: all from constant memory; 2: the match-state chains only
#endif
#if PHMM_COEF_CONST == 1
#define LA_M(x, y, ctab) logadd_k(x, y)
#define LA_G(x, y, ctab) logadd_k(x, y)
#elif PHMM_COEF_CONST == 2
#define LA_M(x, y, ctab) logadd_k(x, y)
#define LA_G(x, y, ctab) logadd_t(x, y, ctab)
#elif PHMM_COEF_CONST == 3      // the gap-state chains (off the critical path of a cell) from constant memory
#define LA_M(x, y, ctab) logadd_t(x, y, ctab)
#define LA_G(x, y, ctab) logadd_k(x, y)
#else
#define LA_M(x, y, ctab) logadd_t(x, y, ctab)
#define LA_G(x, y, ctab) logadd_t(x, y, ctab)
#endif
#ifndef PHMM_TAB_SUM
#define PHMM_TAB_SUM 0          // 1: shared memory holds the emissions only; (emission + transition) is added per cell, the
#endif                          //    transition coming from the kernel parameters (constant bank operand of the DADD)
                                // 2: that for the two gap tables only (their chains are off a cell's critical path); s -> M stays a table

// Shared-memory tables of (emission + transition) sums, one padded row of CS doubles per symbol (pair); the
// scalar definition adds `from + (eP + tP)`, so the parenthesised sum can be taken once per model.
//   tM[cX*5+cY][s] = eM[cX][cY] + tr[s -> M]                              s = M, sX, sY, lX, lY
//   tX[cX][TG_*]   = eX[cX] + tr[M->sX], tr[sX->sX], tr[M->lX], tr[lX->lX], tr[sY->sX]
//   tY[cY][TG_*]   = eY[cY] + tr[M->sY], tr[sY->sY], tr[M->lY], tr[lY->lY], tr[sX->sY]
// Rows are TS = 6 doubles apart so that the pairs a cell uses together load as one 16-byte access.
struct Tabs {
    const double *tM, *tX, *tY;
    const char *ctab;
};
constexpr int TROW = PHMM_TAB_SUM ? 1 : TS;   // doubles per table row
// PHMM_TM_T: layout of the s -> M table.  0: one 48-byte row per base pair (cX * 5 + cY), read as 16 + 16 + 8 bytes: lanes whose
// pairs differ by 8 rows collide (16 common rows, 8 bank groups of a 16-byte load).  1: transposed, tM[s][cX * 5 + cY], five 8-byte
// loads whose bank is 2 * pair mod 32: pairs 16 apart collide.  2: transposed with the pair index cX * 4 + cY (+ 16 when cY is N),
// so that the 16 A/C/G/T pairs sit in 16 different bank pairs and only N meets a conflict.  (Bank conflicts were 9 % of the
// shared-memory wavefronts of k_fb2 with layout 0, profiles/r02p_k_fb2_ncu_metrics.txt.)  Measured, same box, 2368 bench reads
// (profiles/r02x_tune_table_layout.txt): layout 0 1026.6 ms, 1 1025 ms, 2 1009.4 ms at a launch bound of 6; 999.6 / 990.1 / 965.2 ms
// at a bound of 5 -- layout 2 with 96 registers is 6.3 % faster than what round 2 shipped before.
constexpr int TMROW = (PHMM_TM_T || PHMM_TAB_SUM == 1) ? 1 : TS;   // doubles between the entries of two base pairs in the s -> M table
__device__ __forceinline__ int tm_index(int cX, int cY) {
#if PHMM_TM_T == 2
    return cX * 4 + cY + (cY >> 2) * 16;      // 0..15 A/C/G/T pairs, 16..19 (N, y), 20 + 4 x (x, N), 36 (N, N)
#else
    return (cX * 5 + cY) * TMROW;
#endif
}

// One column of 5 state values.  GUARD: scalar loads through an in-band test (any buffer, any stride);
// otherwise unguarded loads (16-byte ones when the columns are padded, CS = 6) from a CS-strided shared-memory column whose out-of-band
// neighbours are kept at -inf.
struct ColV { double M, sX, sY, lX, lY; };

template <bool GUARD>
__device__ __forceinline__ ColV ld_col(const double *p, bool ok) {
    ColV c;
    if (GUARD) {
        c.M = ok ? p[S_M] : PHMM_NEG_INF; c.sX = ok ? p[S_SX] : PHMM_NEG_INF; c.sY = ok ? p[S_SY] : PHMM_NEG_INF;
        c.lX = ok ? p[S_LX] : PHMM_NEG_INF; c.lY = ok ? p[S_LY] : PHMM_NEG_INF;
    } else if (CS % 2 == 0) {
        const double2 a = *reinterpret_cast<const double2 *>(p);
        const double2 b = *reinterpret_cast<const double2 *>(p + 2);
        c.M = a.x; c.sX = a.y; c.sY = b.x; c.lX = b.y; c.lY = p[4];
    } else {
        c.M = p[0]; c.sX = p[1]; c.sY = p[2]; c.lX = p[3]; c.lY = p[4];
    }
    return c;
}

template <bool VEC>
__device__ __forceinline__ void st_col(double *p, const double o[NS]) {
    if (VEC && CS % 2 == 0) {
        *reinterpret_cast<double2 *>(p) = make_double2(o[0], o[1]);
        *reinterpret_cast<double2 *>(p + 2) = make_double2(o[2], o[3]);
        p[4] = o[4];
    } else {
#pragma unroll
        for (int s = 0; s < NS; s++) p[s] = o[s];
    }
}

// Rows of the (emission + transition) tables, loaded as 16-byte pairs.
struct GapRow { double s, ss, l, ll, sw; };    // M->short, short->short, M->long, long->long, other short->short (switch)
struct MatRow { double m, sx, sy, lx, ly; };   // s -> M for s = M, sX, sY, lX, lY
#if PHMM_TAB_SUM
template <bool SWITCH, bool ISX>
__device__ __forceinline__ GapRow ld_gap(const double *r, const DevModel &m) {
    const double e = r[0];
    constexpr int S = ISX ? S_SX : S_SY, L = ISX ? S_LX : S_LY, O = ISX ? S_SY : S_SX;
    GapRow g;
    g.s = e + m.tr[S_M * 5 + S]; g.ss = e + m.tr[S * 5 + S]; g.l = e + m.tr[S_M * 5 + L]; g.ll = e + m.tr[L * 5 + L];
    g.sw = SWITCH ? e + m.tr[O * 5 + S] : 0.0;
    return g;
}
#else
// Rows of the (emission + transition) tables, loaded as 16-byte pairs.
template <bool SWITCH, bool ISX>
__device__ __forceinline__ GapRow ld_gap(const double *r, const DevModel &) {
    const double2 a = *reinterpret_cast<const double2 *>(r + TG_S);
    const double2 b = *reinterpret_cast<const double2 *>(r + TG_L);
    GapRow g; g.s = a.x; g.ss = a.y; g.l = b.x; g.ll = b.y; g.sw = SWITCH ? r[TG_SW] : 0.0;
    return g;
}
#endif
#if PHMM_TAB_SUM == 1
__device__ __forceinline__ MatRow ld_mat(const double *r, const DevModel &m) {
    const double e = r[0];
    MatRow q;
    q.m = e + m.tr[S_M * 5 + S_M]; q.sx = e + m.tr[S_SX * 5 + S_M]; q.sy = e + m.tr[S_SY * 5 + S_M];
    q.lx = e + m.tr[S_LX * 5 + S_M]; q.ly = e + m.tr[S_LY * 5 + S_M];
    return q;
}
#else
__device__ __forceinline__ MatRow ld_mat(const double *r, const DevModel &) {
#if PHMM_TM_T
    MatRow m; m.m = r[0]; m.sx = r[TM_PAIRS]; m.sy = r[2 * TM_PAIRS]; m.lx = r[3 * TM_PAIRS]; m.ly = r[4 * TM_PAIRS];
    return m;
#else
    const double2 a = *reinterpret_cast<const double2 *>(r);
    const double2 b = *reinterpret_cast<const double2 *>(r + 2);
    MatRow m; m.m = a.x; m.sx = a.y; m.sy = b.x; m.lx = b.y; m.ly = r[4];
    return m;
#endif
}
#endif

// Forward cell from register values.  L = lower (x-1,y), U = upper (x,y-1) on diagonal d-1, C = middle (x-1,y-1)
// on d-2.  Transition order of SURVEY.md A.4.
template <bool SWITCH>
__device__ __forceinline__ void fwd_cell_r(const Tabs &t, const DevModel &md, int cX, int cY, double LM, double LsX, double LsY, double LlX,
                                           const ColV &C, double UM, double UsX, double UsY, double UlY, double out[NS]) {
    const char *ctab = t.ctab;
    const GapRow gx = ld_gap<SWITCH, true>(t.tX + cX * TROW, md), gy = ld_gap<SWITCH, false>(t.tY + cY * TROW, md);
    const MatRow gm = ld_mat(t.tM + tm_index(cX, cY), md);
    {
        double a = LM + gx.s;
        a = LA_G(a, LsX + gx.ss, ctab);
        if (SWITCH) a = LA_G(a, LsY + gx.sw, ctab);
        out[S_SX] = a;
        double b = LM + gx.l;
        b = LA_G(b, LlX + gx.ll, ctab);
        out[S_LX] = b;
    }
    {
        double a = C.M + gm.m;
        a = LA_M(a, C.sX + gm.sx, ctab);
        a = LA_M(a, C.sY + gm.sy, ctab);
        a = LA_M(a, C.lX + gm.lx, ctab);
        a = LA_M(a, C.lY + gm.ly, ctab);
        out[S_M] = a;
    }
    {
        double a = UM + gy.s;
        a = LA_G(a, UsY + gy.ss, ctab);
        if (SWITCH) a = LA_G(a, UsX + gy.sw, ctab);
        out[S_SY] = a;
        double b = UM + gy.l;
        b = LA_G(b, UlY + gy.ll, ctab);
        out[S_LY] = b;
    }
}

// Backward cell from register values: Bm = B_M of (x+1,y+1) on d+2; BsY, BlY of (x,y+1) and BsX, BlX of (x+1,y) on d+1.
// cXn = X[x], cYn = Y[y]: the symbols those steps consume.
template <bool SWITCH>
__device__ __forceinline__ void bwd_cell_r(const Tabs &t, const DevModel &md, int cXn, int cYn, double Bm, double BsX, double BlX, double BsY,
                                           double BlY, double out[NS]) {
    const char *ctab = t.ctab;
    const GapRow gx = ld_gap<SWITCH, true>(t.tX + cXn * TROW, md), gy = ld_gap<SWITCH, false>(t.tY + cYn * TROW, md);
    const MatRow gm = ld_mat(t.tM + tm_index(cXn, cYn), md);
    {
        double a = Bm + gm.m;
        a = LA_M(a, BsY + gy.s, ctab);
        a = LA_M(a, BlY + gy.l, ctab);
        a = LA_M(a, BsX + gx.s, ctab);
        a = LA_M(a, BlX + gx.l, ctab);
        out[S_M] = a;
    }
    {
        double a = Bm + gm.sx;
        if (SWITCH) a = LA_G(a, BsY + gy.sw, ctab);
        a = LA_G(a, BsX + gx.ss, ctab);
        out[S_SX] = a;
    }
    {
        double a = Bm + gm.sy;
        a = LA_G(a, BsY + gy.ss, ctab);
        if (SWITCH) a = LA_G(a, BsX + gx.sw, ctab);
        out[S_SY] = a;
    }
    {
        double a = Bm + gm.lx;
        a = LA_G(a, BlX + gx.ll, ctab);
        out[S_LX] = a;
    }
    {
        double a = Bm + gm.ly;
        a = LA_G(a, BlY + gy.ll, ctab);
        out[S_LY] = a;
    }
}

// Forward cell from its three predecessor columns: lower = (x-1,y), upper = (x,y-1) on diagonal d-1,
// middle = (x-1,y-1) on d-2.
template <bool SWITCH, bool GUARD>
__device__ __forceinline__ void fwd_cell3(const Tabs &t, const DevModel &md, const double *pl, bool okl, const double *pu, bool oku,
                                          const double *pm, bool okm, int cX, int cY, double out[NS]) {
    const ColV L = ld_col<GUARD>(pl, okl), C = ld_col<GUARD>(pm, okm), U = ld_col<GUARD>(pu, oku);
    fwd_cell_r<SWITCH>(t, md, cX, cY, L.M, L.sX, L.sY, L.lX, C, U.M, U.sX, U.sY, U.lY, out);
}

// Backward cell from its three successor columns: pu = (x, y+1), pl = (x+1, y) on diagonal d+1,
// pm = (x+1, y+1) on d+2.
template <bool SWITCH, bool GUARD>
__device__ __forceinline__ void bwd_cell3(const Tabs &t, const DevModel &md, const double *pl, bool okl, const double *pu, bool oku,
                                          const double *pm, bool okm, int cXn, int cYn, double out[NS]) {
    const double Bm = (!GUARD || okm) ? pm[S_M] : PHMM_NEG_INF;
    const double BsY = (!GUARD || oku) ? pu[S_SY] : PHMM_NEG_INF, BlY = (!GUARD || oku) ? pu[S_LY] : PHMM_NEG_INF;
    const double BsX = (!GUARD || okl) ? pl[S_SX] : PHMM_NEG_INF, BlX = (!GUARD || okl) ? pl[S_LX] : PHMM_NEG_INF;
    bwd_cell_r<SWITCH>(t, md, cXn, cYn, Bm, BsX, BlX, BsY, BlY, out);
}

// left-to-right logAdd fold of n values produced by f(i), starting from -inf (dpDiagonal_dotProduct order)
template <typename F>
__device__ __forceinline__ double fold_seq(int n, const char *ctab, F f) {
    double t = PHMM_NEG_INF;
    for (int i = 0; i < n; i++) t = logadd_t(t, f(i), ctab);
    return t;
}

// ---------------------------------------------------------------------------
// k_records: one thread per region walks the band once and writes the diagonal records of the whole region.
// Ring slots are handed out by a bump allocator that wraps; the ring is sized (host) for the largest live window.
// The schedule of total-probability diagonals needs the traceback points (k_geometry): every TOTAL_EVERY-th
// diagonal counted down from the first posterior diagonal of the window the diagonal belongs to.
// ---------------------------------------------------------------------------
__global__ void k_records(const Region *regions, const Run *runs, int n_regions, DevParams p, const int64_t *tb_off,
                          const int32_t *tbp, const int32_t *ntb, int ntb_stride, int64_t ring_doubles, int wcap,
                          int cell_doubles, int total_extra, const int64_t *rec_off, DiagRec *recs) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_regions) return;
    const Region reg = regions[r];
    const int nd = reg.lx + reg.ly;
    if (nd == 0) return;
    const int32_t *tb = tbp + tb_off[r];
    const int ntbr = ntb[(int64_t)r * ntb_stride];
    const int tbd = p.tb_diags + 1;
    DiagRec *out = recs + rec_off[r];
    BandIter it;
    it.init(runs + reg.run0, reg.nrun, reg.lx, reg.ly, p.expansion);
    // diagonal 0 (the single cell (0,0)) gets a slot too: the E-step reads its forward values back
    int roff = 0, rsz = cell_doubles, tk = 0;
    { DiagRec r0; r0.off = 0; r0.xlo = 0; r0.w = 1; r0.pad = 0; out[0] = r0; }
    int P = tb[0];
    int TF = P - (P == nd ? 0 : tbd);
    int Pn = ntbr > 1 ? tb[1] : nd;
    int TFn = Pn - (Pn == nd ? 0 : tbd);
    int c1lo = 0, c1hi = 0;                     // diagonal 0: the single cell (0,0), column 0
    int c2lo = 1, c2hi = 0;                     // no diagonal -1
    int wf1 = 0, wf2 = 0;
    for (int d = 1; d <= nd; d++) {
        int xlo, w;
        it.diag(d, xlo, w);
        const int tf = d <= TF ? TF : TFn;
        const bool tot = (tf - d) % TOTAL_EVERY == 0;
        const int es = w * (cell_doubles + (tot ? total_extra : 0));
        int off = roff + rsz;
        if ((int64_t)off + es > ring_doubles) off = 0;
        roff = off; rsz = es;
        if (d == P) {
            tk++;
            P = Pn; TF = TFn;
            Pn = tk + 1 < ntbr ? tb[tk + 1] : nd;
            TFn = Pn - (Pn == nd ? 0 : tbd);
        }
        // columns (x - (d >> 1)) of this and the two previous diagonals, one spare either side: do they alias?
        const int clo = xlo - (d >> 1), chi = clo + w - 1;
        const int wf = w > wcap ? REC_WIDE : 0;
        int lo = min(clo, c1lo), hi = max(chi, c1hi);
        if (c2lo <= c2hi) { lo = min(lo, c2lo); hi = max(hi, c2hi); }
        // the unguarded loops read, around every cell, columns of the two neighbouring diagonals; only the band of those
        // diagonals plus ONE sentinel column either side is guaranteed to hold a value or -inf (FB2 sentinels), so every
        // such read must stay within that reach -- forward into d (from d-1, d-2) and backward into d-2 (from d-1, d)
        const int dl = (d & 1) ? -1 : 0;
        const bool have2 = c2lo <= c2hi;
        const bool fwd_ok = clo + dl >= c1lo - 1 && chi + dl + 1 <= c1hi + 1 && (!have2 || (clo >= c2lo - 1 && chi <= c2hi + 1));
        const bool bwd_ok = !have2 || (c2lo + dl >= c1lo - 1 && c2hi + dl + 1 <= c1hi + 1 && c2lo >= clo - 1 && c2hi <= chi + 1);
        const bool fast3 = !(wf | wf1 | wf2) && (hi - lo + 3 <= wcap) && fwd_ok && bwd_ok;
        DiagRec rc; rc.off = off; rc.xlo = xlo; rc.w = w; rc.pad = (tot ? REC_TOT : 0) | wf | (fast3 ? REC_FAST3 : 0);
        out[d] = rc;
        c2lo = c1lo; c2hi = c1hi; wf2 = wf1;
        c1lo = clo; c1hi = chi; wf1 = wf;
    }
}

__device__ __forceinline__ DiagRec ld_rec(const DiagRec *p) {
    const int4 v = *reinterpret_cast<const int4 *>(p);
    DiagRec r; r.off = v.x; r.xlo = v.y; r.w = v.z; r.pad = v.w;
    return r;
}

#ifndef PHMM_MB4
#define PHMM_MB4 5              // launch bound of the 4-warp kernel: 96 registers; shared memory holds 5 blocks per SM at 512 columns
#endif                          // anyway.  A bound of 6 (80 registers) measured +1.5 % in profiles/r02g_tune_warps.txt and -2.6 % after the
                                // later changes (profiles/r02x_tune_table_layout.txt); with the transposed table 5 wins by 4.4 %
#ifndef PHMM_NW4
#define PHMM_NW4 4              // warps of the kernel the planner's "4 warps" class launches (tuning builds: 5, 6)
#endif
#ifndef PHMM_MB8
#define PHMM_MB8 2
#endif
#ifndef PHMM_MBE
#define PHMM_MBE 4              // launch bound of the 4-warp E-step kernel (126 registers at 4)
#endif
#ifndef PHMM_MBE2
#define PHMM_MBE2 8             // ... of the 2-warp E-step kernel
#endif
constexpr int fb2_min_blocks(int nw, bool expect) {
    return expect ? (nw == 8 ? 2 : (nw == PHMM_NW4 ? (PHMM_NW4 == 4 ? PHMM_MBE : PHMM_MB4) : PHMM_MBE2)) : (nw == 8 ? PHMM_MB8 : (nw == PHMM_NW4 ? PHMM_MB4 : 8));
}

// Shared-memory diagonal buffers.  Cell (d, x) lives in column (x - (d >> 1)) & (wcap - 1) of the buffer of parity
// d & 1, so a cell overwrites its own `middle` predecessor (d-2, x-1) and its `lower` / `upper` predecessors sit in
// the same and the adjacent column of the other buffer.  INVARIANT kept for both buffers: the column either side of
// the band of the diagonal a buffer holds contains -inf (two sentinel columns, written after the diagonal's cells).
// With it the recurrences read their neighbours without any in-band test as long as (REC_FAST3, decided by k_records)
// the columns of three consecutive diagonals plus one either side do not alias modulo wcap and every neighbour read
// lands inside band +- 1 of the diagonal it reads.  Other diagonals take the guarded path and then clear every
// out-of-band column of their buffer (a superset of the sentinels).
// EXPECT (Baum-Welch E-step, replaces `cactus_realign --outputExpectations`): the ring keeps all five forward AND
// backward values of every cell of the live window (10 doubles per cell, 11 where a total is evaluated), the
// posterior phase is replaced by an expectation phase over the same diagonals -- every cell independent, no barriers
// -- that evaluates the 15 transitions of SURVEY.md A.8 with the window's totals and accumulates 2^-32 fixed-point
// counts (transitions in per-thread registers, emissions in shared memory).
template <int NW, bool SWITCH, bool EXPECT>
__global__ void __launch_bounds__(NW * 32, fb2_min_blocks(NW, EXPECT)) k_fb2(const __grid_constant__ Fb2Args a) {
    constexpr int NC = NW * 32;            // threads; all compute
    constexpr int NTA = NC;
    PHMM_DYN_SHARED(smem_raw);
    const int tid = threadIdx.x;
    const int wcap = a.wcap, cmask = wcap - 1;
    double *const sbuf = reinterpret_cast<double *>(smem_raw);                  // [2][wcap][CS]
    // tables and record FIFOs at link-time constant shared addresses: their offsets fold into the LDS immediates
    __shared__ __align__(16) double s_tab[FB2_TAB];
    __shared__ __align__(16) DiagRec s_rec[2 * FB2_RQ];
    double *const sct = s_tab;                                                   // 4 rows x (c3 c2 c1 c0)
    double *const stM = sct + 16, *const stX = stM + TM_DOUBLES, *const stY = stX + 5 * TS;
    DiagRec *const srec = s_rec;                                                 // [2][FB2_RQ]
    __shared__ int s_region;
    __shared__ int s_npairs;
    __shared__ double s_est;                           // total of the window's first posterior diagonal
    __shared__ int s_ncand, s_estok;
    const double EST_EPS = a.est_eps;                  // how far a window's totals may lie from s_est (verified per window)
    __shared__ EmisTables etab;                        // EXPECT
    __shared__ unsigned long long sT[EXPECT ? 25 : 1];
    __shared__ unsigned long long sE[EXPECT ? 80 : 1];
    constexpr int DOT = EXPECT ? 10 : 5;               // block of the per-cell dot products within a diagonal's ring entry
    if (EXPECT) {
        for (int i = tid; i < 25; i += NC) etab.eM[i] = a.m.eM[i];
        if (tid < 5) { etab.eX[tid] = a.m.eX[tid]; etab.eY[tid] = a.m.eY[tid]; }
    }

    if (tid == 0) {
        sct[0] = -0.009350833524763; sct[1] = 0.130659527668286; sct[2] = 0.498799810682272; sct[3] = 0.693203116424741;
        sct[4] = -0.014532321752540; sct[5] = 0.139942324101744; sct[6] = 0.495635523139337; sct[7] = 0.692140569840976;
        sct[8] = -0.004605031767994; sct[9] = 0.063427417320019; sct[10] = 0.695956496475118; sct[11] = 0.514272634594009;
        sct[12] = -0.000458661602210; sct[13] = 0.009695946122598; sct[14] = 0.930734667215156; sct[15] = 0.168037164329057;
    }
#if PHMM_TAB_SUM == 1
    for (int i = tid; i < 25; i += NTA) stM[PHMM_TM_T == 2 ? tm_index(i / 5, i % 5) : i] = a.m.eM[i];
#else
    for (int i = tid; i < 25 * TS; i += NTA) {
        const int r = i / TS, s = i - r * TS;
#if PHMM_TM_T
        if (s < NS) stM[s * TM_PAIRS + (PHMM_TM_T == 2 ? tm_index(r / 5, r % 5) : r)] = a.m.eM[r] + a.m.tr[s * 5 + S_M];
#else
        stM[i] = s < NS ? a.m.eM[r] + a.m.tr[s * 5 + S_M] : 0.0;
#endif
    }
#endif
#if PHMM_TAB_SUM
    if (tid < 5) { stX[tid] = a.m.eX[tid]; stY[tid] = a.m.eY[tid]; }
#else
    if (tid < 5 * TS) {
        const int c = tid / TS, s = tid - c * TS;
        // slots TG_S, TG_SS, TG_L, TG_LL, TG_SW
        const int fx[5] = {S_M * 5 + S_SX, S_SX * 5 + S_SX, S_M * 5 + S_LX, S_LX * 5 + S_LX, S_SY * 5 + S_SX};
        const int fy[5] = {S_M * 5 + S_SY, S_SY * 5 + S_SY, S_M * 5 + S_LY, S_LY * 5 + S_LY, S_SX * 5 + S_SY};
        stX[tid] = s < NS ? a.m.eX[c] + a.m.tr[fx[s < NS ? s : 0]] : 0.0;
        stY[tid] = s < NS ? a.m.eY[c] + a.m.tr[fy[s < NS ? s : 0]] : 0.0;
    }
#endif
    Tabs tabs;
    tabs.tM = stM; tabs.tX = stX; tabs.tY = stY; tabs.ctab = reinterpret_cast<const char *>(sct);
    const char *const ctab = tabs.ctab;

    const int slot = blockIdx.x;
    double *const ring = a.ring + (int64_t)slot * a.ring_doubles;
    double *const wide = a.wide + (int64_t)slot * 4 * NS * a.wg;
    double *const fsave = a.fsave + (int64_t)slot * 2 * CS * wcap;
    double *const totals = a.totals + (int64_t)slot * (a.tcap + a.wg);
    double *const ovs = totals + a.tcap;
    long long *const cand = a.cand + (int64_t)slot * a.ccap;
    const double lp_lo = a.p.lp_skip - EST_EPS;
    const int wgmask = a.wg - 1;
    const int tbd = a.p.tb_diags + 1;

    // column of cell x of a diagonal with half-index h, in shared memory or in the wide buffer `wb` (0..3)
    auto col_ptr = [&](bool is_wide, int par, int wb, int x, int h) -> double * {
        if (!is_wide) return sbuf + ((par * wcap) + ((x - h) & cmask)) * CS;
        return wide + ((int64_t)(wb + par) * a.wg + ((x - h) & wgmask)) * NS;
    };
    // -inf into the columns [c0, c1] of buffer `par`, cooperatively by `nthr` threads
    auto clear_cols = [&](int par, int c0, int c1, int lane, int nthr) {
        for (int c = c0 + lane; c <= c1; c += nthr) {
            double *q = sbuf + (par * wcap + (c & cmask)) * CS;
#pragma unroll
            for (int s = 0; s < NS; s++) q[s] = PHMM_NEG_INF;
        }
    };
    // every column of buffer `par` outside [clo, clo + w) (band columns, unwrapped; w <= wcap)
    auto rg_of = [&](const DiagRec &r) -> double * { return ring + r.off; };
    auto clear_outside = [&](int par, int clo, int w, int lane, int nthr) { clear_cols(par, clo + w, clo + wcap - 1, lane, nthr); };

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
        double ll = 0.0;                                            // thread 0
        if (EXPECT) {
            for (int i = tid; i < 25; i += NC) sT[i] = 0ull;
            for (int i = tid; i < 80; i += NC) sE[i] = 0ull;
        }
        if (nd > 0) {
            const int32_t *tb = a.tbp + a.tb_off[ridx];
            const int ntb = a.ntb[(int64_t)ridx * a.ntb_stride];
            const DiagRec *const rec = a.recs + a.rec_off[ridx];
            DiagRec pre;                                             // lanes of warp 0: record in flight for the batch after next
            pre.off = pre.xlo = pre.w = pre.pad = 0;
            if (tid < FB2_BATCH) {
                if (1 + tid <= nd) srec[(1 + tid) & (FB2_RQ - 1)] = ld_rec(rec + 1 + tid);
                if (1 + FB2_BATCH + tid <= nd) pre = ld_rec(rec + 1 + FB2_BATCH + tid);
            }
            // both buffers -inf, then diagonal 0: the single cell (0,0), column 0 of the even buffer
            for (int i = tid; i < 2 * CS * wcap; i += NTA) sbuf[i] = PHMM_NEG_INF;
            __syncthreads();
            if (tid < NS) {
                double v;
                if (reg.ragged_left) v = (tid == S_LX || tid == S_LY) ? 0.0 : PHMM_NEG_INF;
                else v = (tid == S_M) ? 0.0 : PHMM_NEG_INF;
                sbuf[tid] = v;
                if (EXPECT) ring[tid] = v;                            // diagonal 0: slot 0 of the ring, width 1
            }
            int xlo1 = 0, w1 = 1, f1 = 0;                             // diagonal d-1
            int xlo2 = 0, w2 = 0, f2 = 0;                             // diagonal d-2 (w2 = 0: absent)
            int traced_to = 0;
            int tk = 0;                                               // upcoming traceback point (all threads)
            int P = tb[0];
            __syncthreads();
            DiagRec rn = srec[1 & (FB2_RQ - 1)];                       // record of the next diagonal, read one barrier ahead
            for (int d = 1; d <= nd; d++) {
                if (((d - 1) & (FB2_BATCH - 1)) == 0 && tid < FB2_BATCH) {
                    // records d .. d+31 are in the FIFO; publish d+32 .. d+63 (loaded a batch ago), fetch d+64 .. d+95
                    if (d + FB2_BATCH + tid <= nd) srec[(d + FB2_BATCH + tid) & (FB2_RQ - 1)] = pre;
                    if (d + 2 * FB2_BATCH + tid <= nd) pre = ld_rec(rec + d + 2 * FB2_BATCH + tid);
                }
                const DiagRec rc = rn;
                const int xlo = rc.xlo, w = rc.w;
                const int h0 = d >> 1, par = d & 1;
                const int clo = xlo - h0;                             // first column of this diagonal (unwrapped)
                const bool fast = (rc.pad & REC_FAST3) != 0;
                {
                    const bool tot = (rc.pad & REC_TOT) != 0;
                    double *const rg = ring + rc.off;
                    if (FB2_DBG(32)) {
                    } else if (fast) {
                        // the three diagonals are in shared memory and every out-of-band neighbour reads -inf
                        double *const b0 = sbuf + par * wcap * CS;
                        const double *const b1 = sbuf + (par ^ 1) * wcap * CS;
                        const int dl = par ? -1 : 0;                  // lower is in column c-1 (odd d) or c (even d); upper one further
                        // which warp takes which 32 cells rotates with the diagonal: the warps that get a second (third ..)
                        // round of a diagonal wider than the block sit on all four schedulers of the SM in turn
                        for (int i = tid ^ FB2_ROT(d); i < w; i += NC) {
                            const int x = xlo + i, y = d - x;
                            // x = 0 / y = 0: the byte before the region (padded arrays); it only meets -inf predecessors
                            const int cX = X[x - 1], cY = Y[y - 1];
                            const int c = clo + i;
                            double *const p0 = b0 + (c & cmask) * CS;                      // own column == middle's
                            const double *const pl = b1 + ((c + dl) & cmask) * CS;
                            const double *const pu = b1 + ((c + dl + 1) & cmask) * CS;
                            double o[NS];
                            fwd_cell3<SWITCH, false>(tabs, a.m, pl, true, pu, true, p0, true, cX, cY, o);
                            st_col<true>(p0, o);
                            if (!FB2_DBG(2)) {
                            rg[i] = o[S_M];
                            if (tot || EXPECT) {
#pragma unroll
                                for (int s = 1; s < NS; s++) rg[s * w + i] = o[s];
                            }
                            }
                        }
                    } else {
                        const bool wd0 = (rc.pad & REC_WIDE) != 0, wd1 = (f1 & REC_WIDE) != 0, wd2 = (f2 & REC_WIDE) != 0;
                        const int h1 = (d - 1) >> 1, h2 = (d - 2) >> 1;
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
                            fwd_cell3<SWITCH, true>(tabs, a.m, pl, okl, pu, oku, pm, okm, cX, cY, o);
                            st_col<false>(p0, o);
                            rg[i] = o[S_M];
                            if (tot || EXPECT) {
#pragma unroll
                                for (int s = 1; s < NS; s++) rg[s * w + i] = o[s];
                            }
                        }
                    }
                }
                {
                    if (fast) {
                        // sentinels: -inf in the column either side of the band (nobody reads them during d)
                        if (tid < 2 * NS) {
                            const int side = tid >= NS ? 1 : 0;
                            sbuf[(par * wcap + ((side ? clo + w : clo - 1) & cmask)) * CS + (tid - side * NS)] = PHMM_NEG_INF;
                        }
                    } else if (!(rc.pad & REC_WIDE)) {
                        // guarded diagonal held in shared memory: restore the invariant for its buffer (the cleared columns
                        // are outside the band, and the only column of this buffer a cell reads is its own)
                        clear_outside(par, clo, w, tid, NC);
                    }
                    if (d < nd) rn = srec[(d + 1) & (FB2_RQ - 1)];      // published at least one barrier ago; its latency hides in this one
                    __syncthreads();
                }
                if (d == P && FB2_DBG(16)) {
                    traced_to = d - (d == nd ? 0 : tbd);
                    tk++;
                    P = tk < ntb ? tb[tk] : nd + 1;
                } else if (d == P) {
                    // ------------------------- traceback window (traced_to, d] -------------------------
                    const bool at_end = d == nd;
                    const int traced_from = d - (at_end ? 0 : tbd);
                    const double *endv = (at_end && !reg.ragged_right) ? a.m.endp : a.m.rendp;
                    DiagRec *const srb = srec + FB2_RQ;                                  // backward FIFO
                    if (!at_end) {
                        for (int i = tid; i < 2 * CS * wcap; i += NTA) fsave[i] = sbuf[i];
                    }
                    if (tid == 0) { s_ncand = 0; s_estok = 1; }
                    double est = 0.0;                                                    // = s_est once the sweep has passed traced_from
                    DiagRec preb;                                                        // backward counterpart of `pre`
                    preb.off = preb.xlo = preb.w = preb.pad = 0;
                    if (tid < FB2_BATCH) {
                        if (d - tid >= 1) srb[(d - tid) & (FB2_RQ - 1)] = ld_rec(rec + d - tid);
                        if (d - FB2_BATCH - tid >= 1) preb = ld_rec(rec + d - FB2_BATCH - tid);
                    }
                    __syncthreads();
                    // the backward sweep reuses the two buffers: all -inf before its first diagonal
                    for (int i = tid; i < 2 * CS * wcap; i += NTA) sbuf[i] = PHMM_NEG_INF;
                    __syncthreads();
                    // phase 1: backward sweep
                    {
                        int bxlo1 = 0, bw1 = 0, bf1 = 0, bxlo2 = 0, bw2 = 0, bf2 = 0;      // diagonals dd+1, dd+2
                        DiagRec rbn = srb[d & (FB2_RQ - 1)];
                        for (int dd = d; dd > traced_to; dd--) {
                            if (((d - dd) & (FB2_BATCH - 1)) == 0 && tid < FB2_BATCH) {
                                if (dd - FB2_BATCH - tid >= 1) srb[(dd - FB2_BATCH - tid) & (FB2_RQ - 1)] = preb;
                                if (dd - 2 * FB2_BATCH - tid >= 1) preb = ld_rec(rec + dd - 2 * FB2_BATCH - tid);
                            }
                            const DiagRec rb = rbn;
                            const int hb0 = dd >> 1, bpar = dd & 1;
                            const int bclo = rb.xlo - hb0;
                            // fast: diagonals dd, dd+1, dd+2 are a FAST3 triple (flag of dd+2) and all exist
                            const bool bfast = dd + 2 <= d && (bf2 & REC_FAST3) != 0;
                            {
                                double *const rg = ring + rb.off;
                                const bool dots = (rb.pad & REC_TOT) != 0 && dd <= traced_from;
                                if (FB2_DBG(256)) {
                                } else if (bfast) {
                                    double *const b0 = sbuf + bpar * wcap * CS;
                                    const double *const b1 = sbuf + (bpar ^ 1) * wcap * CS;
                                    const int du = bpar ? -1 : 0;     // (x, y+1) is in column c-1 (odd dd) or c (even dd); (x+1, y) one further
                                    for (int i = tid ^ FB2_ROT(dd); i < rb.w; i += NC) {
                                        const int x = rb.xlo + i, y = dd - x;
                                        const int cXn = X[x], cYn = Y[y];              // x = lx / y = ly: the byte after the region, meets -inf only
                                        const int c = bclo + i;
                                        const double fM = rg[i];                                   // issued early, used last
                                        double *const p0 = b0 + (c & cmask) * CS;                  // own column == (x+1, y+1)'s
                                        const double *const pu = b1 + ((c + du) & cmask) * CS;
                                        const double *const pl = b1 + ((c + du + 1) & cmask) * CS;
                                        double o[NS];
                                        bwd_cell3<SWITCH, false>(tabs, a.m, pl, true, pu, true, p0, true, cXn, cYn, o);
                                        st_col<true>(p0, o);
                                        const double sM = fM + o[S_M];
                                        if (EXPECT) {
#pragma unroll
                                            for (int s = 0; s < NS; s++) rg[(NS + s) * rb.w + i] = o[s];
                                        } else {
                                            // F_M is not needed again below traced_from; above it the next window sweeps the
                                            // diagonal once more, and only the one next to the range feeds a total
                                            if (dd <= traced_from) rg[i] = sM;
                                            else if (dd == traced_from + 1) ovs[i] = sM;
                                            // posterior candidate: could reach the threshold for any total within EST_EPS of est
                                            if (dd < traced_from && sM - est >= lp_lo) {
                                                const int cs = atomicAdd(&s_ncand, 1);
                                                if (cs < a.ccap) cand[cs] = ((long long)dd << 32) | (unsigned)i;
                                            }
                                        }
                                        if (dots) {
                                            double t = sM;
#pragma unroll
                                            for (int s = 1; s < NS; s++) t = logadd_t(t, rg[s * rb.w + i] + o[s], ctab);
                                            rg[DOT * rb.w + i] = t;
                                        }
                                    }
                                } else {
                                    const bool wd0 = (rb.pad & REC_WIDE) != 0, wd1 = (bf1 & REC_WIDE) != 0, wd2 = (bf2 & REC_WIDE) != 0;
                                    const int h1 = (dd + 1) >> 1, h2 = (dd + 2) >> 1;
                                    for (int i = tid; i < rb.w; i += NC) {
                                        const int x = rb.xlo + i, y = dd - x;
                                        double o[NS];
                                        double *const p0 = col_ptr(wd0, bpar, 2, x, hb0);
                                        if (dd < d) {
                                            const int cXn = x < lx ? X[x] : 4;
                                            const int cYn = y < ly ? Y[y] : 4;
                                            const bool oku = (unsigned)(x - bxlo1) < (unsigned)bw1;
                                            const bool okl = (unsigned)(x + 1 - bxlo1) < (unsigned)bw1;
                                            const bool okm = (unsigned)(x + 1 - bxlo2) < (unsigned)bw2;
                                            const double *const pu = col_ptr(wd1, bpar ^ 1, 2, x, h1);
                                            const double *const pl = col_ptr(wd1, bpar ^ 1, 2, x + 1, h1);
                                            const double *const pm = col_ptr(wd2, bpar, 2, x + 1, h2);
                                            bwd_cell3<SWITCH, true>(tabs, a.m, pl, okl, pu, oku, pm, okm, cXn, cYn, o);
                                        } else {
#pragma unroll
                                            for (int s = 0; s < NS; s++) o[s] = endv[s];
                                        }
                                        st_col<false>(p0, o);
                                        const double sM = rg[i] + o[S_M];
                                        if (EXPECT) {
#pragma unroll
                                            for (int s = 0; s < NS; s++) rg[(NS + s) * rb.w + i] = o[s];
                                        } else {
                                            if (dd <= traced_from) rg[i] = sM;
                                            else if (dd == traced_from + 1) ovs[i] = sM;
                                            if (dd < traced_from && sM - est >= lp_lo) {
                                                const int cs = atomicAdd(&s_ncand, 1);
                                                if (cs < a.ccap) cand[cs] = ((long long)dd << 32) | (unsigned)i;
                                            }
                                        }
                                        if (dots) {
                                            double t = sM;
#pragma unroll
                                            for (int s = 1; s < NS; s++) t = logadd_t(t, rg[s * rb.w + i] + o[s], ctab);
                                            rg[DOT * rb.w + i] = t;
                                        }
                                    }
                                }
                            }
                            {
                                if (bfast) {
                                    if (tid < 2 * NS) {
                                        const int side = tid >= NS ? 1 : 0;
                                        sbuf[(bpar * wcap + ((side ? bclo + rb.w : bclo - 1) & cmask)) * CS + (tid - side * NS)] = PHMM_NEG_INF;
                                    }
                                } else if (!(rb.pad & REC_WIDE)) {
                                    clear_outside(bpar, bclo, rb.w, tid, NC);
                                }
                                if (dd - 1 >= 1) rbn = srb[(dd - 1) & (FB2_RQ - 1)];
                                __syncthreads();
                            }
                            if (!EXPECT && dd == traced_from) {
                                // the window's first total, evaluated as phase 2 will (same folds): the yardstick of the
                                // candidate test for all diagonals below
                                if (tid == 0) {
                                    const double *cd = rg_of(rb) + DOT * rb.w;
                                    double total = fold_seq(rb.w, ctab, [&](int i) { return cd[i]; });
                                    if (dd < d) {
                                        const int w1e = bw1;                         // width of diagonal dd+1 (previous iteration)
                                        total = logadd_t(total, fold_seq(w1e, ctab, [&](int i) { return ovs[i]; }), ctab);
                                    }
                                    s_est = total;
                                }
                                __syncthreads();
                                est = s_est;
                            }
                            bxlo2 = bxlo1; bw2 = bw1; bf2 = bf1;
                            bxlo1 = rb.xlo; bw1 = rb.w; bf1 = rb.pad;
                        }
                    }
                    // phase 2: total probabilities, one thread per total diagonal
                    const int nk = traced_from > traced_to ? (traced_from - traced_to - 1) / TOTAL_EVERY + 1 : 0;
                    for (int k = tid; k < (FB2_DBG(64) ? 0 : nk); k += NTA) {
                        const int dd = traced_from - TOTAL_EVERY * k;
                        const DiagRec r0 = rec[dd];
                        const double *cd = ring + r0.off + DOT * r0.w;
                        double total = fold_seq(r0.w, ctab, [&](int i) { return cd[i]; });
                        if (dd < d) {
                            const DiagRec r1 = rec[dd + 1];
                            const double *sm = dd + 1 > traced_from ? ovs : ring + r1.off;
                            const double *fm1 = ring + r1.off, *bm1 = fm1 + NS * r1.w;                     // EXPECT: F_M and B_M apart
                            const double t1 = EXPECT ? fold_seq(r1.w, ctab, [&](int i) { return fm1[i] + bm1[i]; })
                                                     : fold_seq(r1.w, ctab, [&](int i) { return sm[i]; });
                            total = logadd_t(total, t1, ctab);
                        }
                        totals[k] = total;
                        if (!EXPECT && !(fabs(total - s_est) <= EST_EPS)) s_estok = 0;     // the candidate test was not safe: full scan
                    }
                    __syncthreads();
                    if (EXPECT) {
                        // phase 3E: expectations, one warp per diagonal, cells independent; log-likelihood = the totals of the
                        // posterior diagonals added in descending order (the scalar order)
                        if (tid == 0)
                            for (int dd = traced_from; dd > traced_to; dd--) ll += totals[(traced_from - dd) / TOTAL_EVERY];
                        unsigned long long accT[EXP_NT];                // this thread's transition counts of the window:
#pragma unroll
                        for (int k = 0; k < EXP_NT; k++) accT[k] = 0ull; // registers only while this phase runs
                        for (int dd = traced_from - (tid >> 5); dd > traced_to; dd -= NW) {
                            const DiagRec r0 = rec[dd], r1 = rec[dd - 1];
                            DiagRec r2 = r1;
                            int w2e = 0;
                            if (dd - 2 >= traced_to) { r2 = rec[dd - 2]; w2e = r2.w; }     // older forward diagonals are gone in the scalar schedule
                            const double total = totals[(traced_from - dd) / TOTAL_EVERY];
                            const double *bd = ring + r0.off + NS * r0.w;
                            for (int i = tid & 31; i < r0.w; i += 32) {
                                const int x = r0.xlo + i, y = dd - x;
                                const int cX = x >= 1 ? X[x - 1] : 4;
                                const int cY = y >= 1 ? Y[y - 1] : 4;
                                double B[NS];
#pragma unroll
                                for (int s = 0; s < NS; s++) B[s] = bd[s * r0.w + i];
                                expect_cell<SWITCH>(a.m, etab, ring + r1.off, r1.xlo, r1.w, ring + r2.off, r2.xlo, w2e, x, cX, cY, B, total, accT, sE);
                            }
                        }
#pragma unroll
                        for (int k = 0; k < EXP_NT; k++)
                            if (accT[k]) atomicAdd(&sT[EXP_SLOT_TR[k]], accT[k]);
                    }
                    // phase 3: posterior match probabilities >= threshold.  When every total of the window lies within EST_EPS
                    // of its first one (verified above), only the first diagonal and the candidates collected during the
                    // backward sweep can pass; otherwise (or when the candidate list overflowed) every cell is re-read.
                    auto emit = [&](int dd, int i, int xlo_, double sM, double total) {
                        const int x = xlo_ + i, y = dd - x;
                        if (x > 0 && y > 0) {
                            const double lp = sM - total;
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
                    };
                    const int ncand = s_ncand;
                    if (!EXPECT && s_estok && ncand <= a.ccap && !FB2_DBG(128 | 512) && traced_from > traced_to) {
                        {
                            const DiagRec r0 = rec[traced_from];
                            const double *sm = ring + r0.off;
                            for (int i = tid; i < r0.w; i += NC) emit(traced_from, i, r0.xlo, sm[i], totals[0]);
                        }
                        for (int c = tid; c < ncand; c += NC) {
                            const long long v = cand[c];
                            const int dd = (int)(v >> 32), i = (int)(unsigned)v;
                            const DiagRec r0 = rec[dd];
                            emit(dd, i, r0.xlo, ring[r0.off + i], totals[(traced_from - dd) / TOTAL_EVERY]);
                        }
                    } else
                    for (int dd = EXPECT ? traced_to : traced_from - (tid >> 5); dd > (FB2_DBG(128) ? traced_from : traced_to); dd -= NW) {
                        const DiagRec r0 = rec[dd];
                        const double total = totals[(traced_from - dd) / TOTAL_EVERY];
                        const double *sm = ring + r0.off;
                        for (int i0 = tid & 31; i0 < r0.w; i0 += 128) {
                            double sv[4];
#pragma unroll
                            for (int u = 0; u < 4; u++) sv[u] = i0 + 32 * u < r0.w ? sm[i0 + 32 * u] : PHMM_NEG_INF;   // 4 loads in flight
#pragma unroll
                            for (int u = 0; u < 4; u++)
                                if (i0 + 32 * u < r0.w) emit(dd, i0 + 32 * u, r0.xlo, sv[u], total);
                        }
                    }
                    __syncthreads();
                    // phase 4: forward state back, next window
                    if (!at_end) {
                        for (int i = tid; i < 2 * CS * wcap; i += NTA) sbuf[i] = fsave[i];
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
        if (tid == 0) {
            a.npairs[ridx] = s_npairs;
            if (EXPECT) a.expLL[ridx] = ll;
        }
        if (EXPECT) {
            for (int i = tid; i < 25; i += NC) a.expT[(int64_t)ridx * 25 + i] = sT[i];
            for (int i = tid; i < 80; i += NC) a.expE[(int64_t)ridx * 80 + i] = sE[i];
        }
    }
}

}  // namespace phmm
