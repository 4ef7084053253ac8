/*
 * phmm_oracle.c -- CPU oracle for the pair-HMM realignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in nanopore_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, as the checker or the timed
 * CPU baseline -- never as the product path.
 *
 * PARITY UNPINNED.  The arithmetic this file restates lives in two git
 * submodules of the reference that are EMPTY in the build environment:
 *   benedictpaten/cactus @ db284869e4d6b656484b3241b54f3e840dbab8a9
 *       (bar/impl/pairwiseAligner.c, stateMachine.c, cactus_realign.c)
 *   benedictpaten/sonLib @ 5cbc1583797e567900b53ccd50f0b8e72b973d44
 *       (logAdd, cigar I/O)
 * (reference: .SUBMODULES.json:23-29,58-64; submodules/cactus and
 * submodules/sonLib contain no files).  The reference's own tests hold no
 * golden vector for the DP.  This file therefore restates the *published
 * algorithm* of those modules and anchors on the
 * reference's call sites:
 *   nanopore/analyses/utils.py:587      cactus_realign invocation + flags
 *   nanopore/analyses/utils.py:509-528  EM options
 *   nanopore/analyses/utils.py:597-605  only the returned ops are consumed
 *   nanopore/analyses/marginAlignSnpCaller.py:149  "refPos readPos prob"
 * It is validated independently (tests/test_oracle_*.py) by an unbanded
 * exact log-space DP in numpy and by invariants, not by the upstream binary.
 *
 * Documented deviations from the recalled upstream arithmetic:
 *   (1) the cubic in lookup() is evaluated with fused multiply-add (fma) in
 *       every Horner step (upstream, built by gcc for x86-64 without -mfma,
 *       rounds the product and the sum separately);
 *   (2) exp() of posterior log-probabilities uses po_exp() below (an fma-only
 *       routine whose every operation is an IEEE-754 primitive, so that the
 *       CUDA kernels can reproduce it bit for bit) instead of libm exp;
 *   (3) the "make pairs ordered" step is an exact maximum-expected-accuracy
 *       chain DP over the reweighted pairs (upstream anneals greedily over
 *       the same weights);
 *   (4) Baum-Welch sufficient statistics are accumulated per sub-problem in
 *       64-bit fixed point (units of 2^-32) so that sums are independent of
 *       evaluation order (upstream adds doubles).
 *
 * Each of (1)-(3) has a switch (po_set_upstream_arithmetic, bits PO_UP_*) that
 * replaces it by the recalled upstream behaviour, so that a differential run
 * against a real cactus_realign binary (tests/tools/differential.py) can tell which
 * deviation a mismatch comes from.  The CUDA library implements the default
 * (switches off) arithmetic only.
 *
 * States (utils.py:617): 0 match, 1 shortGapX, 2 shortGapY,
 * 3 longGapX, 4 longGapY.  X = cigar target = reference, Y = read.
 * Symbols: A=0 C=1 G=2 T=3, anything else 4 (N).
 * Cigar ops use SAM codes: 0 = M, 1 = I (read only / gap in X), 2 = D
 * (reference only / gap in Y) (utils.py:173,602).
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define NS 5
#define ST_M 0
#define ST_SX 1
#define ST_SY 2
#define ST_LX 3
#define ST_LY 4
#define PROB_1 10000000LL            /* PAIR_ALIGNMENT_PROB_1 */
#define LOG_ZERO (-INFINITY)
#define EXPECT_SCALE 4294967296.0    /* 2^32 fixed-point unit of the E-step */

/* ------------------------------------------------------------------ */
/* Arithmetic helpers (sonLib logAdd)                    */
/* ------------------------------------------------------------------ */

static int g_exact_logadd = 0; /* tests only: replace the cubic by log1p(exp()) */

void po_set_exact_logadd(int on) { g_exact_logadd = on; }

/* Differential runs only: undo the documented deviations one by one. */
#define PO_UP_UNFUSED_HORNER 1 /* (1) product and sum of every Horner step rounded separately (x86-64 gcc without -mfma) */
#define PO_UP_LIBM_EXP 2       /* (2) libm exp() for posterior and expectation probabilities */
#define PO_UP_GREEDY_ORDER 4   /* (3) pairs made ordered greedily by descending weight instead of by the exact chain DP */
static int g_upstream = 0;

void po_set_upstream_arithmetic(int flags) { g_upstream = flags; }
int po_get_upstream_arithmetic(void) { return g_upstream; }

/* this file is compiled with -ffp-contract=off: a * b + c below is two roundings */
static inline double po_lookup_unfused(double x) {
    if (x <= 1.00)
        return ((-0.009350833524763 * x + 0.130659527668286) * x + 0.498799810682272) * x + 0.693203116424741;
    if (x <= 2.50)
        return ((-0.014532321752540 * x + 0.139942324101744) * x + 0.495635523139337) * x + 0.692140569840976;
    if (x <= 4.50)
        return ((-0.004605031767994 * x + 0.063427417320019) * x + 0.695956496475118) * x + 0.514272634594009;
    return ((-0.000458661602210 * x + 0.009695946122598) * x + 0.930734667215156) * x + 0.168037164329057;
}

static inline double po_lookup(double x) {
    /* piecewise cubic fit of log(1+exp(-x))+x ... i.e. log(exp(x)+1), x in [0,7.5) */
    if (g_upstream & PO_UP_UNFUSED_HORNER) return po_lookup_unfused(x);
    if (x <= 1.00)
        return fma(fma(fma(-0.009350833524763, x, 0.130659527668286), x, 0.498799810682272), x, 0.693203116424741);
    if (x <= 2.50)
        return fma(fma(fma(-0.014532321752540, x, 0.139942324101744), x, 0.495635523139337), x, 0.692140569840976);
    if (x <= 4.50)
        return fma(fma(fma(-0.004605031767994, x, 0.063427417320019), x, 0.695956496475118), x, 0.514272634594009);
    return fma(fma(fma(-0.000458661602210, x, 0.009695946122598), x, 0.930734667215156), x, 0.168037164329057);
}

double po_logadd(double x, double y) {
    if (g_exact_logadd) {
        if (x == LOG_ZERO) return y;
        if (y == LOG_ZERO) return x;
        return x > y ? x + log1p(exp(y - x)) : y + log1p(exp(x - y));
    }
    if (x < y)
        return (x == LOG_ZERO || y - x >= 7.5) ? y : po_lookup(y - x) + x;
    return (y == LOG_ZERO || x - y >= 7.5) ? x : po_lookup(x - y) + y;
}

/* exp() built only from IEEE primitives (fma, add, integer ops): identical
 * instruction-for-instruction in the CUDA kernels.  |rel err| < 2 ulp. */
double po_exp(double x) {
    if (g_upstream & PO_UP_LIBM_EXP) return exp(x);
    if (!(x > -700.0)) return 0.0;         /* also catches NaN, -inf */
    if (x > 700.0) return INFINITY;
    const double SHIFT = 6755399441055744.0; /* 1.5 * 2^52 */
    double t = fma(x, 1.4426950408889634, SHIFT);
    double kd = t - SHIFT;
    union { double d; int64_t i; } u;
    u.d = t;
    int32_t k = (int32_t)(u.i & 0xffffffffLL);
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;          /* 1/13! */
    p = fma(p, r, 2.08767569878681e-09);         /* 1/12! */
    p = fma(p, r, 2.505210838544172e-08);        /* 1/11! */
    p = fma(p, r, 2.755731922398589e-07);        /* 1/10! */
    p = fma(p, r, 2.7557319223985893e-06);       /* 1/9!  */
    p = fma(p, r, 2.48015873015873e-05);         /* 1/8!  */
    p = fma(p, r, 1.984126984126984e-04);        /* 1/7!  */
    p = fma(p, r, 1.388888888888889e-03);        /* 1/6!  */
    p = fma(p, r, 8.333333333333333e-03);        /* 1/5!  */
    p = fma(p, r, 4.1666666666666664e-02);       /* 1/4!  */
    p = fma(p, r, 1.6666666666666666e-01);       /* 1/3!  */
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    u.d = p;
    u.i += ((int64_t)k) << 52;                   /* scale by 2^k (result is normal for |x|<=700) */
    return u.d;
}

/* ------------------------------------------------------------------ */
/* Model (stateMachine5)                      */
/* ------------------------------------------------------------------ */

typedef struct po_model {
    double tr[NS][NS];   /* log transition, [from][to] */
    double eM[5][5];     /* log match emission [x][y], N row/col = log(1/16) */
    double eX[5];        /* log gap-in-Y emission of a reference base (states 1,3) */
    double eY[5];        /* log gap-in-X emission of a read base (states 2,4) */
    double start[NS], rstart[NS], end[NS], rend[NS];
} po_model;

static void model_finish(po_model *m) {
    for (int i = 0; i < 5; i++) {
        m->eM[4][i] = m->eM[i][4] = -2.772588722;  /* log(0.0625) as upstream spells it */
    }
    m->eX[4] = m->eY[4] = -1.386294361;            /* log(0.25) */
    for (int s = 0; s < NS; s++) {
        m->start[s] = (s == ST_M) ? 0.0 : LOG_ZERO;
        m->rstart[s] = (s == ST_LX || s == ST_LY) ? 0.0 : LOG_ZERO;
        m->end[s] = m->tr[s][ST_M];
    }
    m->rend[ST_M] = m->tr[ST_M][ST_LX];
    m->rend[ST_SX] = m->tr[ST_M][ST_LX];
    m->rend[ST_SY] = m->tr[ST_M][ST_LY];
    m->rend[ST_LX] = m->tr[ST_LX][ST_LX];
    m->rend[ST_LY] = m->tr[ST_LY][ST_LY];
}

/* trans: 25 probabilities row-major from*5+to; emis: 80 probabilities
 * state*16 + x*4 + y (blasr_hmm_0.txt:1-2).  NULL,NULL -> stock model. */
po_model *po_model_create(const double *trans, const double *emis) {
    po_model *m = (po_model *)calloc(1, sizeof(po_model));
    if (!trans || !emis) {
        for (int i = 0; i < NS; i++) for (int j = 0; j < NS; j++) m->tr[i][j] = LOG_ZERO;
        m->tr[ST_M][ST_M] = -0.030064059121770816;
        m->tr[ST_SX][ST_M] = m->tr[ST_SY][ST_M] = -1.272871422049609;
        m->tr[ST_LX][ST_M] = m->tr[ST_LY][ST_M] = -5.673280173170473;
        m->tr[ST_M][ST_SX] = m->tr[ST_M][ST_SY] = -4.34381910900448;
        m->tr[ST_SX][ST_SX] = m->tr[ST_SY][ST_SY] = -0.3388262689231553;
        m->tr[ST_SX][ST_SY] = m->tr[ST_SY][ST_SX] = -4.910694825551255;
        m->tr[ST_M][ST_LX] = m->tr[ST_M][ST_LY] = -6.30810595366929;
        m->tr[ST_LX][ST_LX] = m->tr[ST_LY][ST_LY] = -0.003442492794189331;
        for (int x = 0; x < 4; x++) {
            for (int y = 0; y < 4; y++) {
                if (x == y) m->eM[x][y] = -2.1149196655034745;
                else if ((x ^ y) == 2) m->eM[x][y] = -3.9833860032220842; /* A<->G, C<->T */
                else m->eM[x][y] = -4.5691014376830479;
            }
            m->eX[x] = m->eY[x] = -1.6094379124341003;
        }
    } else {
        for (int i = 0; i < NS; i++) for (int j = 0; j < NS; j++) m->tr[i][j] = log(trans[i * NS + j]);
        for (int x = 0; x < 4; x++) for (int y = 0; y < 4; y++) m->eM[x][y] = log(emis[x * 4 + y]);
        /* gap emissions: marginal of the 4x4 table over the other sequence,
         * summed over the two gap states of that kind, then normalised */
        double gx[4] = {0, 0, 0, 0}, gy[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; i++) {
            for (int j = 0; j < 4; j++) {
                gx[i] += emis[ST_SX * 16 + i * 4 + j];
                gx[i] += emis[ST_LX * 16 + i * 4 + j];
            }
        }
        for (int i = 0; i < 4; i++) {
            for (int j = 0; j < 4; j++) {
                gy[j] += emis[ST_SY * 16 + i * 4 + j];
                gy[j] += emis[ST_LY * 16 + i * 4 + j];
            }
        }
        double tx = 0.0, ty = 0.0;
        for (int i = 0; i < 4; i++) { tx += gx[i]; ty += gy[i]; }
        for (int i = 0; i < 4; i++) { m->eX[i] = log(gx[i] / tx); m->eY[i] = log(gy[i] / ty); }
    }
    model_finish(m);
    return m;
}

void po_model_destroy(po_model *m) { free(m); }

/* flat dump used by tests to cross-check the GPU library's model tables:
 * out[0..24] tr, [25..49] eM(5x5), [50..54] eX, [55..59] eY */
void po_model_dump(const po_model *m, double *out) {
    memcpy(out, m->tr, 25 * sizeof(double));
    memcpy(out + 25, m->eM, 25 * sizeof(double));
    memcpy(out + 50, m->eX, 5 * sizeof(double));
    memcpy(out + 55, m->eY, 5 * sizeof(double));
}

/* ------------------------------------------------------------------ */
/* Parameters                                                           */
/* ------------------------------------------------------------------ */

typedef struct po_params {
    int64_t expansion;      /* --diagonalExpansion (even)                 utils.py:587 */
    int64_t trim;           /* constraintDiagonalTrim (upstream default 14)            */
    int64_t split_side;     /* --splitMatrixBiggerThanThis, a side; area = side^2      */
    int64_t min_diags;      /* minDiagsBetweenTraceBack (1000)                         */
    int64_t tb_diags;       /* traceBackDiagonals (40)                                 */
    double threshold;       /* posterior threshold (0.01)                              */
    double gap_gamma;       /* --gapGamma   abstractMapper.py:25                       */
    double match_gamma;     /* --matchGamma abstractMapper.py:25                       */
} po_params;

void po_params_default(po_params *p) {
    p->expansion = 10; p->trim = 14; p->split_side = 3000;
    p->min_diags = 1000; p->tb_diags = 40; p->threshold = 0.01;
    p->gap_gamma = 0.5; p->match_gamma = 0.0;
}

/* ------------------------------------------------------------------ */
/* Band (band_construct)                                 */
/* ------------------------------------------------------------------ */

typedef struct { int64_t xay, xmyL, xmyR; } diag_t;

static inline int64_t dg_x(int64_t xay, int64_t xmy) { return (xay + xmy) / 2; }
static inline int64_t dg_y(int64_t xay, int64_t xmy) { return (xay - xmy) / 2; }
static inline int64_t dg_width(diag_t d) { return (d.xmyR - d.xmyL) / 2 + 1; }

static int64_t avoid_off_by_one(int64_t xay, int64_t xmy) { return ((xay + xmy) % 2 == 0) ? xmy : xmy + 1; }
static void bound_p(int64_t *xmy, int64_t i, int64_t j, int64_t k) { if (i < j) *xmy += 2 * (j - i) * k; }
static int64_t bound_coord(int64_t z, int64_t lz) { return z < 0 ? 0 : (z > lz ? lz : z); }

static diag_t band_diag(int64_t xay, int64_t xL, int64_t yL, int64_t xU, int64_t yU) {
    int64_t xmyL = avoid_off_by_one(xay, xL - yL);
    int64_t xmyR = avoid_off_by_one(xay, xU - yU);
    bound_p(&xmyL, dg_x(xay, xmyL), xL, 1);
    bound_p(&xmyL, yL, dg_y(xay, xmyL), 1);
    bound_p(&xmyR, xU, dg_x(xay, xmyR), -1);
    bound_p(&xmyR, dg_y(xay, xmyR), yU, -1);
    diag_t d = { xay, xmyL, xmyR };
    return d;
}

/* anchors: (x,y) sequence coordinates, strictly increasing in both */
static diag_t *band_build(const int64_t *ax, const int64_t *ay, int64_t na, int64_t lX, int64_t lY, int64_t e) {
    diag_t *dg = (diag_t *)malloc(sizeof(diag_t) * (size_t)(lX + lY + 1));
    int64_t ai = 0, xay = 0, pxay = 0, pxmy = 0, nxay = 0, nxmy = 0;
    int64_t xL = 0, yL = 0, xU = 0, yU = 0;
    while (xay <= lX + lY) {
        dg[xay] = band_diag(xay, xL, yL, xU, yU);
        if (nxay == xay++) {
            pxay = nxay; pxmy = nxmy;
            int64_t x = lX, y = lY;
            if (ai < na) { x = ax[ai] + 1; y = ay[ai] + 1; ai++; }
            nxay = x + y; nxmy = x - y;
            xL = bound_coord(dg_x(pxay, pxmy - e), lX);
            yL = bound_coord(dg_y(nxay, nxmy - e), lY);
            xU = bound_coord(dg_x(nxay, nxmy + e), lX);
            yU = bound_coord(dg_y(pxay, pxmy + e), lY);
        }
    }
    return dg;
}

/* ------------------------------------------------------------------ */
/* DP diagonals                                                         */
/* ------------------------------------------------------------------ */

typedef struct { diag_t d; double *c; } dpd_t;   /* c: width * NS, cell-major */

static dpd_t *dpd_new(diag_t d, const double *init /* NS values or NULL -> LOG_ZERO */) {
    dpd_t *q = (dpd_t *)malloc(sizeof(dpd_t));
    int64_t w = dg_width(d);
    q->d = d;
    q->c = (double *)malloc(sizeof(double) * (size_t)(w * NS));
    for (int64_t i = 0; i < w; i++) for (int s = 0; s < NS; s++) q->c[i * NS + s] = init ? init[s] : LOG_ZERO;
    return q;
}
static void dpd_free(dpd_t *q) { if (q) { free(q->c); free(q); } }
static inline double *dpd_cell(dpd_t *q, int64_t xmy) {
    if (!q || xmy < q->d.xmyL || xmy > q->d.xmyR) return NULL;
    return q->c + ((xmy - q->d.xmyL) / 2) * NS;
}

typedef void (*trans_fn)(double *from, double *to, int f, int t, double eP, double tP, int cX, int cY, void *extra);

/* cell_calculate: enumeration order of SURVEY.md A.4 */
static inline void cell_calc(const po_model *m, double *cur, double *lower, double *middle, double *upper,
                             int cX, int cY, trans_fn fn, void *extra) {
    if (lower) {
        double eP = m->eX[cX];
        fn(lower, cur, ST_M, ST_SX, eP, m->tr[ST_M][ST_SX], cX, cY, extra);
        fn(lower, cur, ST_SX, ST_SX, eP, m->tr[ST_SX][ST_SX], cX, cY, extra);
        fn(lower, cur, ST_SY, ST_SX, eP, m->tr[ST_SY][ST_SX], cX, cY, extra);
        fn(lower, cur, ST_M, ST_LX, eP, m->tr[ST_M][ST_LX], cX, cY, extra);
        fn(lower, cur, ST_LX, ST_LX, eP, m->tr[ST_LX][ST_LX], cX, cY, extra);
    }
    if (middle) {
        double eP = m->eM[cX][cY];
        fn(middle, cur, ST_M, ST_M, eP, m->tr[ST_M][ST_M], cX, cY, extra);
        fn(middle, cur, ST_SX, ST_M, eP, m->tr[ST_SX][ST_M], cX, cY, extra);
        fn(middle, cur, ST_SY, ST_M, eP, m->tr[ST_SY][ST_M], cX, cY, extra);
        fn(middle, cur, ST_LX, ST_M, eP, m->tr[ST_LX][ST_M], cX, cY, extra);
        fn(middle, cur, ST_LY, ST_M, eP, m->tr[ST_LY][ST_M], cX, cY, extra);
    }
    if (upper) {
        double eP = m->eY[cY];
        fn(upper, cur, ST_M, ST_SY, eP, m->tr[ST_M][ST_SY], cX, cY, extra);
        fn(upper, cur, ST_SY, ST_SY, eP, m->tr[ST_SY][ST_SY], cX, cY, extra);
        fn(upper, cur, ST_SX, ST_SY, eP, m->tr[ST_SX][ST_SY], cX, cY, extra);
        fn(upper, cur, ST_M, ST_LY, eP, m->tr[ST_M][ST_LY], cX, cY, extra);
        fn(upper, cur, ST_LY, ST_LY, eP, m->tr[ST_LY][ST_LY], cX, cY, extra);
    }
}

static void tr_forward(double *from, double *to, int f, int t, double eP, double tP, int cX, int cY, void *extra) {
    (void)cX; (void)cY; (void)extra;
    to[t] = po_logadd(to[t], from[f] + (eP + tP));
}
static void tr_backward(double *from, double *to, int f, int t, double eP, double tP, int cX, int cY, void *extra) {
    (void)cX; (void)cY; (void)extra;
    from[f] = po_logadd(from[f], to[t] + (eP + tP));
}

typedef struct {
    double total;
    int64_t T[25];     /* fixed point, units 2^-32 */
    int64_t E[80];
} expect_acc;

static void tr_expect(double *from, double *to, int f, int t, double eP, double tP, int cX, int cY, void *extra) {
    expect_acc *a = (expect_acc *)extra;
    double p = po_exp(from[f] + to[t] + (eP + tP) - a->total);
    int64_t q = (int64_t)floor(p * EXPECT_SCALE);
    a->T[f * NS + t] += q;
    if (cX < 4 && cY < 4) a->E[t * 16 + cX * 4 + cY] += q;
}

/* run cell_calc over one diagonal: cur on xay, m1 on xay-1, m2 on xay-2 */
static void diag_calc(const po_model *m, dpd_t *cur, dpd_t *m1, dpd_t *m2, const uint8_t *X, const uint8_t *Y,
                      trans_fn fn, void *extra) {
    diag_t d = cur->d;
    for (int64_t xmy = d.xmyL; xmy <= d.xmyR; xmy += 2) {
        int64_t ix = dg_x(d.xay, xmy) - 1, iy = dg_y(d.xay, xmy) - 1;
        int cX = ix >= 0 ? X[ix] : 4, cY = iy >= 0 ? Y[iy] : 4;
        cell_calc(m, dpd_cell(cur, xmy), dpd_cell(m1, xmy - 1), dpd_cell(m2, xmy), dpd_cell(m1, xmy + 1), cX, cY, fn, extra);
    }
}

static double cell_dot(const double *a, const double *b) {
    double t = a[0] + b[0];
    for (int s = 1; s < NS; s++) t = po_logadd(t, a[s] + b[s]);
    return t;
}
static double diag_dot(dpd_t *a, dpd_t *b) {
    double t = LOG_ZERO;
    for (int64_t xmy = a->d.xmyL; xmy <= a->d.xmyR; xmy += 2) t = po_logadd(t, cell_dot(dpd_cell(a, xmy), dpd_cell(b, xmy)));
    return t;
}

/* diagonalCalculationTotalProbability (A.6): paths through diagonal xay plus
 * paths that step over it with a match from xay-1 to xay+1 */
static double diag_total(const po_model *m, int64_t xay, dpd_t **F, dpd_t **B, int64_t nd, const uint8_t *X, const uint8_t *Y) {
    double total = diag_dot(F[xay], B[xay]);
    dpd_t *fm1 = xay >= 1 ? F[xay - 1] : NULL;
    dpd_t *bp1 = xay + 1 <= nd ? B[xay + 1] : NULL;
    if (fm1 && bp1) {
        dpd_t *md = dpd_new(bp1->d, NULL);
        diag_calc(m, md, NULL, fm1, X, Y, tr_forward, NULL);
        total = po_logadd(total, diag_dot(md, bp1));
        dpd_free(md);
    }
    return total;
}

/* growable pair list: (x, y, w) sequence coordinates, w in 1e-7 units */
typedef struct { int64_t n, cap; int64_t *x, *y, *w; } pairs_t;
static void pairs_push(pairs_t *p, int64_t x, int64_t y, int64_t w) {
    if (p->n == p->cap) {
        p->cap = p->cap ? p->cap * 2 : 1024;
        p->x = (int64_t *)realloc(p->x, sizeof(int64_t) * (size_t)p->cap);
        p->y = (int64_t *)realloc(p->y, sizeof(int64_t) * (size_t)p->cap);
        p->w = (int64_t *)realloc(p->w, sizeof(int64_t) * (size_t)p->cap);
    }
    p->x[p->n] = x; p->y[p->n] = y; p->w[p->n] = w; p->n++;
}
static void pairs_free(pairs_t *p) { free(p->x); free(p->y); free(p->w); memset(p, 0, sizeof(*p)); }

/* diagonalCalculationPosteriorMatchProbs (A.6) */
static void diag_posteriors(int64_t xay, dpd_t *f, dpd_t *b, double total, double threshold, pairs_t *out, int64_t offx, int64_t offy) {
    for (int64_t xmy = f->d.xmyL; xmy <= f->d.xmyR; xmy += 2) {
        int64_t x = dg_x(xay, xmy), y = dg_y(xay, xmy);
        if (x > 0 && y > 0) {
            double *cf = dpd_cell(f, xmy), *cb = dpd_cell(b, xmy);
            double p = po_exp((cf[ST_M] + cb[ST_M]) - total);
            if (p >= threshold) {
                if (p > 1.0) p = 1.0;
                pairs_push(out, x - 1 + offx, y - 1 + offy, (int64_t)floor(p * (double)PROB_1));
            }
        }
    }
}

typedef struct {
    int64_t cells;          /* DP cells in the band (sum of diagonal widths) */
    int64_t diagonals;
    int64_t tracebacks;
    int64_t max_live_cells; /* peak forward cells held between tracebacks */
} po_stats;

/* getPosteriorProbsWithBanding (A.6).  mode 0: posterior pairs into out;
 * mode 1: expectations into acc (loglik accumulated into *loglik). */
static void posteriors_banded(const po_model *m, const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY,
                              const int64_t *ax, const int64_t *ay, int64_t na, const po_params *p,
                              int ragged_left, int ragged_right, int mode, pairs_t *out, int64_t offx, int64_t offy,
                              expect_acc *acc, double *loglik, po_stats *st) {
    int64_t nd = lX + lY;
    if (nd == 0) return;
    diag_t *band = band_build(ax, ay, na, lX, lY, p->expansion);
    dpd_t **F = (dpd_t **)calloc((size_t)(nd + 2), sizeof(dpd_t *));
    dpd_t **B = (dpd_t **)calloc((size_t)(nd + 2), sizeof(dpd_t *));
    F[0] = dpd_new(band[0], ragged_left ? m->rstart : m->start);
    int64_t traced_to = 0, live = dg_width(band[0]);
    if (st) { st->cells += dg_width(band[0]); st->diagonals += nd + 1; }
    for (int64_t xay = 1; xay <= nd; xay++) {
        diag_t d = band[xay];
        F[xay] = dpd_new(d, NULL);
        diag_calc(m, F[xay], F[xay - 1], xay >= 2 ? F[xay - 2] : NULL, X, Y, tr_forward, NULL);
        live += dg_width(d);
        if (st) { st->cells += dg_width(d); if (live > st->max_live_cells) st->max_live_cells = live; }
        int at_end = (xay == nd);
        int tb_point = xay >= traced_to + p->min_diags && dg_width(d) <= p->expansion * 2 + 1;
        if (!(at_end || tb_point)) continue;
        if (st) st->tracebacks++;
        B[xay] = dpd_new(d, (at_end && !ragged_right) ? m->end : m->rend);
        if (xay > traced_to + 1) B[xay - 1] = dpd_new(band[xay - 1], NULL);
        int64_t traced_from = xay - (at_end ? 0 : p->tb_diags + 1);
        double total = LOG_ZERO;
        int64_t ncalc = 0;
        for (int64_t d2 = xay; d2 > traced_to; d2--) {
            if (d2 > traced_to + 2) B[d2 - 2] = dpd_new(band[d2 - 2], NULL);
            if (d2 > traced_to + 1)
                diag_calc(m, B[d2], B[d2 - 1], d2 >= 2 ? B[d2 - 2] : NULL, X, Y, tr_backward, NULL);
            if (d2 <= traced_from) {
                if (ncalc++ % 10 == 0) total = diag_total(m, d2, F, B, nd, X, Y);
                if (mode == 0) {
                    diag_posteriors(d2, F[d2], B[d2], total, p->threshold, out, offx, offy);
                } else {
                    acc->total = total;
                    *loglik += total;
                    diag_calc(m, B[d2], F[d2 - 1], d2 >= 2 ? F[d2 - 2] : NULL, X, Y, tr_expect, acc);
                }
                if (d2 < traced_from || at_end) { live -= dg_width(F[d2]->d); dpd_free(F[d2]); F[d2] = NULL; }
            }
            if (d2 + 1 <= nd) { dpd_free(B[d2 + 1]); B[d2 + 1] = NULL; }
        }
        /* here d2 == traced_to (the old value) */
        dpd_free(B[traced_to + 1]); B[traced_to + 1] = NULL;
        if (F[traced_to]) { live -= dg_width(F[traced_to]->d); dpd_free(F[traced_to]); F[traced_to] = NULL; }
        traced_to = traced_from;
        if (at_end) break;
    }
    for (int64_t i = 0; i <= nd; i++) { dpd_free(F[i]); dpd_free(B[i]); }
    free(F); free(B); free(band);
}

/* ------------------------------------------------------------------ */
/* Anchors and split points (cactus_realign main, A.5, A.7)            */
/* ------------------------------------------------------------------ */

typedef struct { int64_t n, cap; int64_t *v; } vec64;
static void v_push(vec64 *v, int64_t a) {
    if (v->n == v->cap) { v->cap = v->cap ? v->cap * 2 : 256; v->v = (int64_t *)realloc(v->v, sizeof(int64_t) * (size_t)v->cap); }
    v->v[v->n++] = a;
}

/* ops: (len<<2)|code with SAM codes.  Anchor = every position of an M run
 * except `trim` at each end. */
static void anchors_from_ops(const uint32_t *ops, int64_t nops, int64_t trim, vec64 *ax, vec64 *ay) {
    int64_t x = 0, y = 0;
    for (int64_t i = 0; i < nops; i++) {
        int64_t len = ops[i] >> 2; int code = ops[i] & 3;
        if (code == 0) {
            for (int64_t l = trim; l < len - trim; l++) { v_push(ax, x + l); v_push(ay, y + l); }
            x += len; y += len;
        } else if (code == 1) y += len;
        else if (code == 2) x += len;
    }
}

static void split_p(int64_t *x1, int64_t *y1, int64_t x2, int64_t y2, int64_t x3, int64_t y3, vec64 *sp, int64_t area, int64_t side) {
    int64_t lx2 = x3 - x2, ly2 = y3 - y2;
    if (lx2 * ly2 > area) {
        int64_t hx = lx2 / 2 > side ? side : lx2 / 2;
        int64_t hy = ly2 / 2 > side ? side : ly2 / 2;
        v_push(sp, *x1); v_push(sp, *y1); v_push(sp, x2 + hx); v_push(sp, y2 + hy);
        *x1 = x3 - hx; *y1 = y3 - hy;
    }
}

/* returns quadruples (x1,y1,x2,y2) */
static void split_points(const vec64 *ax, const vec64 *ay, int64_t lX, int64_t lY, int64_t side, vec64 *sp) {
    int64_t x1 = 0, y1 = 0, x2 = 0, y2 = 0;
    int64_t area = side * side;
    for (int64_t i = 0; i < ax->n; i++) {
        int64_t x3 = ax->v[i], y3 = ay->v[i];
        split_p(&x1, &y1, x2, y2, x3, y3, sp, area, side);
        x2 = x3 + 1; y2 = y3 + 1;
    }
    split_p(&x1, &y1, x2, y2, lX, lY, sp, area, side);
    v_push(sp, x1); v_push(sp, y1); v_push(sp, lX); v_push(sp, lY);
}

typedef struct { int64_t x1, y1, x2, y2, a0, a1; int rl, rr; } region_t;

static int64_t make_regions(const vec64 *ax, const vec64 *ay, int64_t lX, int64_t lY, int64_t side, region_t **out) {
    vec64 sp = {0, 0, NULL};
    split_points(ax, ay, lX, lY, side, &sp);
    int64_t nr = sp.n / 4, j = 0;
    region_t *r = (region_t *)malloc(sizeof(region_t) * (size_t)nr);
    for (int64_t i = 0; i < nr; i++) {
        r[i].x1 = sp.v[4 * i]; r[i].y1 = sp.v[4 * i + 1]; r[i].x2 = sp.v[4 * i + 2]; r[i].y2 = sp.v[4 * i + 3];
        r[i].a0 = j;
        while (j < ax->n && ax->v[j] + ay->v[j] < r[i].x2 + r[i].y2) j++;
        r[i].a1 = j;
        r[i].rl = i > 0; r[i].rr = i < nr - 1;
    }
    free(sp.v);
    *out = r;
    return nr;
}

/* ------------------------------------------------------------------ */
/* Decode: reweight + maximum-expected-accuracy chain + ops (A.9)       */
/* ------------------------------------------------------------------ */

typedef struct { int64_t s; int64_t last; } mea_cell;

static int cmp_pair_idx(const void *a, const void *b, void *ctx) {
    const pairs_t *p = (const pairs_t *)ctx;
    int64_t i = *(const int64_t *)a, j = *(const int64_t *)b;
    int64_t di = p->x[i] + p->y[i], dj = p->x[j] + p->y[j];
    if (di != dj) return di < dj ? -1 : 1;
    if (p->x[i] != p->x[j]) return p->x[i] < p->x[j] ? -1 : 1;
    return 0;
}

/* banded wavefront chain DP over one region; pairs are in region-local
 * sequence coordinates with reweighted weights wr (only wr>0 usable).
 * Appends chosen pair indices (ascending) to chain. */
static int64_t mea_region(const diag_t *band, int64_t lX, int64_t lY, const pairs_t *p, const int64_t *wr,
                          const int64_t *idx, int64_t nidx, vec64 *chain) {
    int64_t nd = lX + lY;
    int64_t *pred = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nidx + 1));
    mea_cell *ring[3];
    diag_t rd[3];
    memset(rd, 0, sizeof(rd));
    int64_t maxw = 1;
    for (int64_t d = 0; d <= nd; d++) if (dg_width(band[d]) > maxw) maxw = dg_width(band[d]);
    for (int k = 0; k < 3; k++) ring[k] = (mea_cell *)malloc(sizeof(mea_cell) * (size_t)maxw);
    ring[0][0].s = 0; ring[0][0].last = -1; rd[0] = band[0];
    int64_t pi = 0;
    for (int64_t d = 1; d <= nd; d++) {
        diag_t cd = band[d];
        mea_cell *cur = ring[d % 3], *m1 = ring[(d - 1) % 3], *m2 = ring[(d - 2 + 3) % 3];
        diag_t d1 = rd[(d - 1) % 3], d2 = rd[(d - 2 + 3) % 3];
        rd[d % 3] = cd;
        while (pi < nidx && p->x[idx[pi]] + p->y[idx[pi]] + 2 < d) pi++;
        for (int64_t xmy = cd.xmyL; xmy <= cd.xmyR; xmy += 2) {
            mea_cell best = { -1, -1 };
            if (xmy - 1 >= d1.xmyL && xmy - 1 <= d1.xmyR) best = m1[(xmy - 1 - d1.xmyL) / 2];
            if (xmy + 1 >= d1.xmyL && xmy + 1 <= d1.xmyR) {
                mea_cell u = m1[(xmy + 1 - d1.xmyL) / 2];
                if (u.s > best.s) best = u;
            }
            int64_t x = dg_x(d, xmy), y = dg_y(d, xmy);
            /* pair at matrix cell (x,y) == sequence (x-1,y-1) */
            while (pi < nidx && p->x[idx[pi]] + p->y[idx[pi]] + 2 == d && p->x[idx[pi]] + 1 < x) pi++;
            if (pi < nidx && p->x[idx[pi]] + p->y[idx[pi]] + 2 == d && p->x[idx[pi]] + 1 == x && p->y[idx[pi]] + 1 == y) {
                int64_t k = pi;
                if (wr[idx[k]] > 0 && d >= 2 && xmy >= d2.xmyL && xmy <= d2.xmyR) {
                    mea_cell mm = m2[(xmy - d2.xmyL) / 2];
                    if (mm.s >= 0) {
                        pred[k] = mm.last;
                        int64_t cand = mm.s + wr[idx[k]];
                        if (cand > best.s) { best.s = cand; best.last = k; }
                    }
                }
            }
            cur[(xmy - cd.xmyL) / 2] = best;
        }
    }
    mea_cell fin = ring[nd % 3][0];
    int64_t score = fin.s;
    /* traceback */
    vec64 rev = {0, 0, NULL};
    for (int64_t k = fin.last; k >= 0; k = pred[k]) v_push(&rev, idx[k]);
    for (int64_t i = rev.n - 1; i >= 0; i--) v_push(chain, rev.v[i]);
    free(rev.v); free(pred);
    for (int k = 0; k < 3; k++) free(ring[k]);
    return score;
}

/* Recalled upstream scheme for "make the pairs ordered" (filterPairwiseAlignmentToMakePairsOrdered over a poset
 * alignment): take the usable pairs (wr > 0) of the whole read by descending reweighted weight (ties: smaller x, then
 * smaller y) and keep a pair when it is consistent with every pair kept so far, i.e. its neighbours in x among the kept
 * pairs lie strictly below / above it in both coordinates.  Appends the kept pair indices in ascending x to chain.
 * PO_UP_GREEDY_ORDER only. */
typedef struct { const int64_t *wr; const pairs_t *p; } greedy_ctx;
static int cmp_greedy(const void *a, const void *b, void *ctx) {
    const greedy_ctx *g = (const greedy_ctx *)ctx;
    int64_t i = *(const int64_t *)a, j = *(const int64_t *)b;
    if (g->wr[i] != g->wr[j]) return g->wr[i] > g->wr[j] ? -1 : 1;
    if (g->p->x[i] != g->p->x[j]) return g->p->x[i] < g->p->x[j] ? -1 : 1;
    if (g->p->y[i] != g->p->y[j]) return g->p->y[i] < g->p->y[j] ? -1 : 1;
    return 0;
}
static int64_t greedy_order(const pairs_t *p, const int64_t *wr, int64_t lX, vec64 *chain) {
    int64_t n = 0, score = 0;
    int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(p->n + 1));
    for (int64_t i = 0; i < p->n; i++) if (wr[i] > 0) idx[n++] = i;
    greedy_ctx g = { wr, p };
    qsort_r(idx, (size_t)n, sizeof(int64_t), cmp_greedy, &g);
    /* kept pair (index + 1) at each x, and a bit tree over x for neighbour queries */
    int64_t size = 1; while (size < lX + 1) size <<= 1;
    int64_t *at = (int64_t *)calloc((size_t)(lX + 1), sizeof(int64_t));
    unsigned char *tree = (unsigned char *)calloc((size_t)(2 * size), 1);
    for (int64_t k = 0; k < n; k++) {
        int64_t i = idx[k], x = p->x[i], y = p->y[i];
        if (at[x]) continue;
        /* predecessor: rightmost kept x' < x */
        int64_t pre = -1, suc = -1, node;
        for (node = x + size; node > 1; node >>= 1)
            if ((node & 1) && tree[node - 1]) { node = node - 1; while (node < size) node = tree[2 * node + 1] ? 2 * node + 1 : 2 * node; pre = node - size; break; }
        for (node = x + size; node > 1; node >>= 1)
            if (!(node & 1) && tree[node + 1]) { node = node + 1; while (node < size) node = tree[2 * node] ? 2 * node : 2 * node + 1; suc = node - size; break; }
        if (pre >= 0 && p->y[at[pre] - 1] >= y) continue;
        if (suc >= 0 && p->y[at[suc] - 1] <= y) continue;
        at[x] = i + 1;
        for (node = x + size; node >= 1; node >>= 1) tree[node] = 1;
        score += wr[i];
    }
    for (int64_t x = 0; x <= lX; x++) if (at[x]) v_push(chain, at[x] - 1);
    free(idx); free(at); free(tree);
    return score;
}

/* convertAlignedPairsToPairwiseAlignment: D (reference-only) before I between
 * matched pairs; ops as (len<<2)|code.  Returns number of ops. */
static int64_t pairs_to_ops(const int64_t *cx, const int64_t *cy, int64_t n, int64_t lX, int64_t lY, vec64 *ops) {
    int64_t px = -1, py = -1, ml = 0;
    for (int64_t i = 0; i <= n; i++) {
        int64_t x = i < n ? cx[i] : lX, y = i < n ? cy[i] : lY;
        if (x - px > 1) {
            if (ml > 0) { v_push(ops, (ml << 2) | 0); ml = 0; }
            v_push(ops, ((x - px - 1) << 2) | 2);
        }
        if (y - py > 1) {
            if (ml > 0) { v_push(ops, (ml << 2) | 0); ml = 0; }
            v_push(ops, ((y - py - 1) << 2) | 1);
        }
        ml++; px = x; py = y;
    }
    if (ml > 1) v_push(ops, ((ml - 1) << 2) | 0);
    return ops->n;
}

/* ------------------------------------------------------------------ */
/* Public entry points                                                  */
/* ------------------------------------------------------------------ */

typedef struct po_result {
    int64_t n_ops; uint32_t *ops;                    /* realigned cigar, (len<<2)|code */
    int64_t n_pairs; int64_t *px, *py, *pw;          /* posterior pairs (upstream emission order) */
    int64_t n_chain; int64_t *cx, *cy;               /* MEA-selected pairs */
    int64_t mea_score;
    po_stats stats;
} po_result;

void po_result_free(po_result *r) {
    if (!r) return;
    free(r->ops); free(r->px); free(r->py); free(r->pw); free(r->cx); free(r->cy); free(r);
}

/* One read: X = reference[ref_start, ref_end), Y = read, in_ops = guide cigar
 * spanning both exactly (utils.py:381-382).  Mirrors one cactus_realign call
 * (utils.py:587). */
po_result *po_realign(const po_model *m, const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY,
                      const uint32_t *in_ops, int64_t n_in_ops, const po_params *p) {
    po_result *res = (po_result *)calloc(1, sizeof(po_result));
    vec64 ax = {0, 0, NULL}, ay = {0, 0, NULL};
    anchors_from_ops(in_ops, n_in_ops, p->trim, &ax, &ay);
    region_t *reg; int64_t nr = make_regions(&ax, &ay, lX, lY, p->split_side, &reg);
    pairs_t all = {0, 0, NULL, NULL, NULL};
    int64_t *rstart = (int64_t *)malloc(sizeof(int64_t) * (size_t)(nr + 1));
    for (int64_t i = 0; i < nr; i++) {
        rstart[i] = all.n;
        int64_t na = reg[i].a1 - reg[i].a0;
        int64_t *lax = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1)), *lay = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1));
        for (int64_t k = 0; k < na; k++) { lax[k] = ax.v[reg[i].a0 + k] - reg[i].x1; lay[k] = ay.v[reg[i].a0 + k] - reg[i].y1; }
        posteriors_banded(m, X + reg[i].x1, reg[i].x2 - reg[i].x1, Y + reg[i].y1, reg[i].y2 - reg[i].y1, lax, lay, na, p,
                          reg[i].rl, reg[i].rr, 0, &all, reg[i].x1, reg[i].y1, NULL, NULL, &res->stats);
        free(lax); free(lay);
    }
    rstart[nr] = all.n;
    /* indel probabilities and reweighting (getIndelProbabilities / reweightAlignedPairs) */
    int64_t *ipx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(lX + 1)), *ipy = (int64_t *)malloc(sizeof(int64_t) * (size_t)(lY + 1));
    for (int64_t i = 0; i < lX; i++) ipx[i] = PROB_1;
    for (int64_t i = 0; i < lY; i++) ipy[i] = PROB_1;
    for (int64_t i = 0; i < all.n; i++) { ipx[all.x[i]] -= all.w[i]; ipy[all.y[i]] -= all.w[i]; }
    for (int64_t i = 0; i < lX; i++) if (ipx[i] < 0) ipx[i] = 0;
    for (int64_t i = 0; i < lY; i++) if (ipy[i] < 0) ipy[i] = 0;
    int64_t *wr = (int64_t *)malloc(sizeof(int64_t) * (size_t)(all.n + 1));
    for (int64_t i = 0; i < all.n; i++) {
        wr[i] = all.w[i] - (int64_t)(p->gap_gamma * (double)(ipx[all.x[i]] + ipy[all.y[i]]));
        if ((double)all.w[i] < p->match_gamma * (double)PROB_1) wr[i] = 0;
    }
    /* MEA per region */
    vec64 chain = {0, 0, NULL};
    if (g_upstream & PO_UP_GREEDY_ORDER) res->mea_score = greedy_order(&all, wr, lX, &chain);
    for (int64_t i = 0; i < nr && !(g_upstream & PO_UP_GREEDY_ORDER); i++) {
        int64_t np = rstart[i + 1] - rstart[i];
        int64_t lx = reg[i].x2 - reg[i].x1, ly = reg[i].y2 - reg[i].y1;
        if (lx + ly == 0) continue;
        pairs_t loc = {np, np, NULL, NULL, NULL};
        loc.x = (int64_t *)malloc(sizeof(int64_t) * (size_t)(np + 1));
        loc.y = (int64_t *)malloc(sizeof(int64_t) * (size_t)(np + 1));
        loc.w = (int64_t *)malloc(sizeof(int64_t) * (size_t)(np + 1));
        int64_t *idx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(np + 1));
        for (int64_t k = 0; k < np; k++) {
            loc.x[k] = all.x[rstart[i] + k] - reg[i].x1; loc.y[k] = all.y[rstart[i] + k] - reg[i].y1; loc.w[k] = all.w[rstart[i] + k];
            idx[k] = k;
        }
        qsort_r(idx, (size_t)np, sizeof(int64_t), cmp_pair_idx, &loc);
        int64_t na = reg[i].a1 - reg[i].a0;
        int64_t *lax = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1)), *lay = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1));
        for (int64_t k = 0; k < na; k++) { lax[k] = ax.v[reg[i].a0 + k] - reg[i].x1; lay[k] = ay.v[reg[i].a0 + k] - reg[i].y1; }
        diag_t *band = band_build(lax, lay, na, lx, ly, p->expansion);
        vec64 lc = {0, 0, NULL};
        res->mea_score += mea_region(band, lx, ly, &loc, wr + rstart[i], idx, np, &lc);
        for (int64_t k = 0; k < lc.n; k++) v_push(&chain, rstart[i] + lc.v[k]);
        free(lc.v); free(band); free(lax); free(lay); free(idx); pairs_free(&loc);
    }
    res->n_chain = chain.n;
    res->cx = (int64_t *)malloc(sizeof(int64_t) * (size_t)(chain.n + 1));
    res->cy = (int64_t *)malloc(sizeof(int64_t) * (size_t)(chain.n + 1));
    for (int64_t i = 0; i < chain.n; i++) { res->cx[i] = all.x[chain.v[i]]; res->cy[i] = all.y[chain.v[i]]; }
    vec64 ops = {0, 0, NULL};
    pairs_to_ops(res->cx, res->cy, chain.n, lX, lY, &ops);
    res->n_ops = ops.n;
    res->ops = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(ops.n + 1));
    for (int64_t i = 0; i < ops.n; i++) res->ops[i] = (uint32_t)ops.v[i];
    res->n_pairs = all.n; res->px = all.x; res->py = all.y; res->pw = all.w;
    free(ops.v); free(chain.v); free(wr); free(ipx); free(ipy); free(rstart); free(reg); free(ax.v); free(ay.v);
    return res;
}

/* E-step for one read (cactus_realign --outputExpectations, utils.py:528 via
 * cactus_expectationMaximisation).  Adds into T[25], E[80], *loglik. */
void po_expectations(const po_model *m, const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY,
                     const uint32_t *in_ops, int64_t n_in_ops, const po_params *p,
                     double *T, double *E, double *loglik, po_stats *st) {
    vec64 ax = {0, 0, NULL}, ay = {0, 0, NULL};
    anchors_from_ops(in_ops, n_in_ops, p->trim, &ax, &ay);
    region_t *reg; int64_t nr = make_regions(&ax, &ay, lX, lY, p->split_side, &reg);
    for (int64_t i = 0; i < nr; i++) {
        int64_t na = reg[i].a1 - reg[i].a0;
        int64_t *lax = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1)), *lay = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1));
        for (int64_t k = 0; k < na; k++) { lax[k] = ax.v[reg[i].a0 + k] - reg[i].x1; lay[k] = ay.v[reg[i].a0 + k] - reg[i].y1; }
        expect_acc acc; memset(&acc, 0, sizeof(acc));
        double ll = 0.0;
        posteriors_banded(m, X + reg[i].x1, reg[i].x2 - reg[i].x1, Y + reg[i].y1, reg[i].y2 - reg[i].y1, lax, lay, na, p,
                          reg[i].rl, reg[i].rr, 1, NULL, 0, 0, &acc, &ll, st);
        for (int k = 0; k < 25; k++) T[k] += (double)acc.T[k] / EXPECT_SCALE;
        for (int k = 0; k < 80; k++) E[k] += (double)acc.E[k] / EXPECT_SCALE;
        *loglik += ll;
        free(lax); free(lay);
    }
    free(reg); free(ax.v); free(ay.v);
}

/* Same E-step with exact integer accumulation (include/phmm.h phmm_expectations_batch_fixed): value k is
 * hi[k] + lo[k] / 2^32 for the 105 expectations, hi + lo / 2^20 for the log-likelihood (per-region values
 * rounded to 2^-20).  ADDS into hi[106], lo[106] and renormalises lo into [0, 2^bits). */
void po_expectations_fixed(const po_model *m, const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY,
                           const uint32_t *in_ops, int64_t n_in_ops, const po_params *p,
                           int64_t *hi, int64_t *lo, po_stats *st) {
    vec64 ax = {0, 0, NULL}, ay = {0, 0, NULL};
    anchors_from_ops(in_ops, n_in_ops, p->trim, &ax, &ay);
    region_t *reg; int64_t nr = make_regions(&ax, &ay, lX, lY, p->split_side, &reg);
    for (int64_t i = 0; i < nr; i++) {
        int64_t na = reg[i].a1 - reg[i].a0;
        int64_t *lax = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1)), *lay = (int64_t *)malloc(sizeof(int64_t) * (size_t)(na + 1));
        for (int64_t k = 0; k < na; k++) { lax[k] = ax.v[reg[i].a0 + k] - reg[i].x1; lay[k] = ay.v[reg[i].a0 + k] - reg[i].y1; }
        expect_acc acc; memset(&acc, 0, sizeof(acc));
        double ll = 0.0;
        posteriors_banded(m, X + reg[i].x1, reg[i].x2 - reg[i].x1, Y + reg[i].y1, reg[i].y2 - reg[i].y1, lax, lay, na, p,
                          reg[i].rl, reg[i].rr, 1, NULL, 0, 0, &acc, &ll, st);
        for (int k = 0; k < 106; k++) {
            const int bits = k < 105 ? 32 : 20;
            int64_t q = k < 25 ? acc.T[k] : (k < 105 ? acc.E[k - 25] : (int64_t)llrint(ll * 1048576.0));
            int64_t qh = q >> bits;                       /* arithmetic shift: floor */
            lo[k] += q - (qh << bits);
            hi[k] += qh + (lo[k] >> bits);
            lo[k] &= (((int64_t)1) << bits) - 1;
        }
        free(lax); free(lay);
    }
    free(reg); free(ax.v); free(ay.v);
}

/* ---- introspection helpers for tests ---- */

/* band of one region: writes xmyL,xmyR for xay = 0..lX+lY */
void po_band(const int64_t *ax, const int64_t *ay, int64_t na, int64_t lX, int64_t lY, int64_t expansion, int64_t *xmyL, int64_t *xmyR) {
    diag_t *b = band_build(ax, ay, na, lX, lY, expansion);
    for (int64_t i = 0; i <= lX + lY; i++) { xmyL[i] = b[i].xmyL; xmyR[i] = b[i].xmyR; }
    free(b);
}

/* regions of one read: returns count, writes up to cap rows of 8 int64
 * (x1,y1,x2,y2,a0,a1,ragged_left,ragged_right) */
int64_t po_regions(const uint32_t *in_ops, int64_t n_in_ops, int64_t lX, int64_t lY, int64_t trim, int64_t split_side, int64_t *out, int64_t cap) {
    vec64 ax = {0, 0, NULL}, ay = {0, 0, NULL};
    anchors_from_ops(in_ops, n_in_ops, trim, &ax, &ay);
    region_t *reg; int64_t nr = make_regions(&ax, &ay, lX, lY, split_side, &reg);
    for (int64_t i = 0; i < nr && i < cap; i++) {
        out[8 * i] = reg[i].x1; out[8 * i + 1] = reg[i].y1; out[8 * i + 2] = reg[i].x2; out[8 * i + 3] = reg[i].y2;
        out[8 * i + 4] = reg[i].a0; out[8 * i + 5] = reg[i].a1; out[8 * i + 6] = reg[i].rl; out[8 * i + 7] = reg[i].rr;
    }
    free(reg); free(ax.v); free(ay.v);
    return nr;
}

/* posterior pairs for explicit anchors on one (sub)problem, no splitting */
po_result *po_posteriors(const po_model *m, const uint8_t *X, int64_t lX, const uint8_t *Y, int64_t lY,
                         const int64_t *ax, const int64_t *ay, int64_t na, const po_params *p, int ragged_left, int ragged_right) {
    po_result *res = (po_result *)calloc(1, sizeof(po_result));
    pairs_t all = {0, 0, NULL, NULL, NULL};
    posteriors_banded(m, X, lX, Y, lY, ax, ay, na, p, ragged_left, ragged_right, 0, &all, 0, 0, NULL, NULL, &res->stats);
    res->n_pairs = all.n; res->px = all.x; res->py = all.y; res->pw = all.w;
    return res;
}

/* batch driver used by the CPU-baseline timing legs (one read per task like
 * utils.py:565-570); single-threaded, callers fan out over processes/threads.
 * Offsets as in include/phmm.h.  out_n_ops[i] receives the op count; ops are
 * appended to out_ops (cap out_cap) -- returns total ops or -1 on overflow. */
int64_t po_realign_batch(const po_model *m, const uint8_t *ref, int64_t n_reads, const uint8_t *reads, const int64_t *read_off,
                         const int64_t *ref_start, const int64_t *ref_end, const uint32_t *in_ops, const int64_t *in_off,
                         const po_params *p, uint32_t *out_ops, int64_t out_cap, int64_t *out_off, int64_t *cells) {
    int64_t n = 0;
    out_off[0] = 0;
    for (int64_t i = 0; i < n_reads; i++) {
        po_result *r = po_realign(m, ref + ref_start[i], ref_end[i] - ref_start[i], reads + read_off[i], read_off[i + 1] - read_off[i],
                                  in_ops + in_off[i], in_off[i + 1] - in_off[i], p);
        if (n + r->n_ops > out_cap) { po_result_free(r); return -1; }
        memcpy(out_ops + n, r->ops, sizeof(uint32_t) * (size_t)r->n_ops);
        n += r->n_ops;
        out_off[i + 1] = n;
        if (cells) *cells += r->stats.cells;
        po_result_free(r);
    }
    return n;
}
