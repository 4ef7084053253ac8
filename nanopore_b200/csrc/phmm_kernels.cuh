// phmm_kernels.cuh -- sm_100a kernels of the realignment path.
//
//   k_geometry   one thread per region: band sweep -> cells, widths, window sizes
//   k_fwdbwd     one thread block (NW warps) per region: banded 5-state forward,
//                windowed backward, total probability, posterior pairs >= threshold
//                (or Baum-Welch expectations when EXPECT)
//   k_decode     one thread block per region: indel reweighting, banded
//                maximum-expected-accuracy chain DP, match runs out
//   k_compact    gathers the per-region match runs into one dense array
//
// Replaces the body of `cactus_realign` (reference nanopore/analyses/utils.py:587);
// algorithm per SURVEY.md Appendix A.4-A.9.  All recurrences are fp64 and keep
// the association order of the scalar definition, so results are bit-exact
// against the CPU checker.
#pragma once
#include "phmm_device.cuh"

namespace phmm {

struct FbArgs {
    const uint8_t *ref;
    const uint8_t *reads;
    const Region *regions;
    const Run *runs;
    const int32_t *order;        // regions sorted by descending cells
    int32_t n_regions;
    int32_t *counter;            // work queue head
    DevModel m;
    DevParams p;
    // per-slot scratch (slot = blockIdx.x)
    double *fring;   int64_t ring_cells;     // ring_cells * 5 doubles per slot
    DiagRec *dtab;   int32_t dcap;           // live-diagonal table per slot
    double *bring;   int32_t bw;             // 3 * bw * 5 doubles per slot
    double *dots;                            // 2 * bw doubles per slot
    // outputs
    int32_t *px, *py, *pw;                   // posterior pairs (region-local sequence coords)
    int32_t *npairs;                         // per region: pairs produced (may exceed pair_cap = overflow)
    unsigned long long *expT;                // per region 25 (EXPECT)
    unsigned long long *expE;                // per region 80
    double *expLL;                           // per region
};

template <int NW>
__device__ __forceinline__ void block_sync() {
    if (NW == 1) __syncwarp(); else __syncthreads();
}

// ---------------------------------------------------------------------------
// geometry: data-independent sweep of the band and of the traceback schedule
// ---------------------------------------------------------------------------
__global__ void k_geometry(const Region *regions, const Run *runs, int n_regions, DevParams p, RegionGeom *out,
                           const int64_t *tb_off, int32_t *tbp) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_regions) return;
    const Region reg = regions[r];
    RegionGeom g;
    const int nd = reg.lx + reg.ly;
    g.diagonals = nd > 0 ? nd + 1 : 0;              // an empty pair has no DP (the scalar path returns at once)
    g.cells = nd > 0 ? 1 : 0; g.max_width = 1; g.tracebacks = 0;
    int64_t live = 1, max_live = 1;
    int live_d = 1, max_live_d = 1;
    int wbuf[256];                      // widths of the most recent diagonals (tb_diags + 2 <= 256)
    wbuf[0] = 1;
    BandIter it;
    it.init(runs + reg.run0, reg.nrun, reg.lx, reg.ly, p.expansion);
    int traced_to = 0;
    int32_t *tb = tbp + tb_off[r];
    const int tb_cap = (int)(tb_off[r + 1] - tb_off[r]);
    int plo = 0, phi = 0, regular = 1;                  // band edges of the previous diagonal (diagonal 0: x = 0)
    for (int d = 1; d <= nd; d++) {
        int xlo, w;
        it.diag(d, xlo, w);
        const int xhi = xlo + w - 1;
        if (xlo < plo || xlo > plo + 1 || xhi < phi || xhi > phi + 1 || w < 1) regular = 0;
        plo = xlo; phi = xhi;
        wbuf[d & 255] = w;
        g.cells += w;
        live += w; live_d++;
        if (w > g.max_width) g.max_width = w;
        if (live > max_live) max_live = live;
        if (live_d > max_live_d) max_live_d = live_d;
        const bool at_end = d == nd;
        const bool tb_here = d >= traced_to + p.min_diags && w <= 2 * p.expansion + 1;
        if (at_end || tb_here) {
            if (g.tracebacks < tb_cap) tb[g.tracebacks] = d;
            g.tracebacks++;
            const int traced_from = d - (at_end ? 0 : p.tb_diags + 1);
            // diagonals traced_from..d stay live for the next window
            live = 0;
            for (int k = traced_from; k <= d; k++) live += wbuf[k & 255];
            live_d = d - traced_from + 1;
            traced_to = traced_from;
        }
    }
    g.max_live_cells = max_live;
    g.max_live_diags = max_live_d;
    g.regular = regular; g.pad = 0;
    // second sweep: ring usage of the windowed kernel (1 double per cell, 6 on total-probability diagonals)
    g.max_live_doubles = 0;
    if (nd > 0 && g.tracebacks <= tb_cap) {
        it.init(runs + reg.run0, reg.nrun, reg.lx, reg.ly, p.expansion);
        int k = 0;
        int P = tb[0];
        int TF = P - (P == nd ? 0 : p.tb_diags + 1);
        int Pn = g.tracebacks > 1 ? tb[1] : nd;
        int TFn = Pn - (Pn == nd ? 0 : p.tb_diags + 1);
        int64_t lv = 0, mx = 0;
        for (int d = 1; d <= nd; d++) {
            int xlo, w;
            it.diag(d, xlo, w);
            const int tf = d <= TF ? TF : TFn;
            const bool tot = (tf - d) % TOTAL_EVERY == 0;
            const int es = w * (tot ? 6 : 1);
            wbuf[d & 255] = es;
            lv += es;
            if (lv > mx) mx = lv;
            if (d == P) {
                lv = 0;
                for (int q = TF + 1; q <= d; q++) lv += wbuf[q & 255];
                k++;
                P = Pn; TF = TFn;
                Pn = k + 1 < g.tracebacks ? tb[k + 1] : nd;
                TFn = Pn - (Pn == nd ? 0 : p.tb_diags + 1);
            }
        }
        g.max_live_doubles = mx;
    }
    out[r] = g;
}

// ---------------------------------------------------------------------------
// forward / backward / posterior
// ---------------------------------------------------------------------------
struct EmisTables {                 // staged in shared memory
    double eM[25], eX[5], eY[5];
};

__device__ __forceinline__ double ldv(const double *p, int idx, bool ok) {
    return ok ? p[idx] : PHMM_NEG_INF;
}

// Forward cell (x,y) on diagonal d from diagonals d-1 (F1) and d-2 (F2); SoA [state][cell].
template <bool SWITCH>
__device__ __forceinline__ void fwd_cell(const DevModel &m, const EmisTables &t,
                                         const double *F1, int xlo1, int w1,
                                         const double *F2, int xlo2, int w2,
                                         int x, int cX, int cY, double out[NS]) {
    const int jl = x - 1 - xlo1;
    const int ju = jl + 1;
    const int jm = x - 1 - xlo2;
    const bool okl = (unsigned)jl < (unsigned)w1;
    const bool oku = (unsigned)ju < (unsigned)w1;
    const bool okm = (unsigned)jm < (unsigned)w2;
    const double eXc = t.eX[cX], eYc = t.eY[cY], eMc = t.eM[cX * 5 + cY];
    // lower = (x-1,y): gap-X transitions
    {
        const double Ml = ldv(F1, jl, okl), sXl = ldv(F1 + w1, jl, okl), lXl = ldv(F1 + 3 * w1, jl, okl);
        double a = Ml + (eXc + m.tr[S_M * 5 + S_SX]);
        a = logadd(a, sXl + (eXc + m.tr[S_SX * 5 + S_SX]));
        if (SWITCH) { const double sYl = ldv(F1 + 2 * w1, jl, okl); a = logadd(a, sYl + (eXc + m.tr[S_SY * 5 + S_SX])); }
        out[S_SX] = a;
        double b = Ml + (eXc + m.tr[S_M * 5 + S_LX]);
        b = logadd(b, lXl + (eXc + m.tr[S_LX * 5 + S_LX]));
        out[S_LX] = b;
    }
    // middle = (x-1,y-1): match transitions
    {
        double a = ldv(F2, jm, okm) + (eMc + m.tr[S_M * 5 + S_M]);
        a = logadd(a, ldv(F2 + w2, jm, okm) + (eMc + m.tr[S_SX * 5 + S_M]));
        a = logadd(a, ldv(F2 + 2 * w2, jm, okm) + (eMc + m.tr[S_SY * 5 + S_M]));
        a = logadd(a, ldv(F2 + 3 * w2, jm, okm) + (eMc + m.tr[S_LX * 5 + S_M]));
        a = logadd(a, ldv(F2 + 4 * w2, jm, okm) + (eMc + m.tr[S_LY * 5 + S_M]));
        out[S_M] = a;
    }
    // upper = (x,y-1): gap-Y transitions
    {
        const double Mu = ldv(F1, ju, oku), sYu = ldv(F1 + 2 * w1, ju, oku), lYu = ldv(F1 + 4 * w1, ju, oku);
        double a = Mu + (eYc + m.tr[S_M * 5 + S_SY]);
        a = logadd(a, sYu + (eYc + m.tr[S_SY * 5 + S_SY]));
        if (SWITCH) { const double sXu = ldv(F1 + w1, ju, oku); a = logadd(a, sXu + (eYc + m.tr[S_SX * 5 + S_SY])); }
        out[S_SY] = a;
        double b = Mu + (eYc + m.tr[S_M * 5 + S_LY]);
        b = logadd(b, lYu + (eYc + m.tr[S_LY * 5 + S_LY]));
        out[S_LY] = b;
    }
}

// Backward cell (x,y) on diagonal d from diagonals d+1 (B1) and d+2 (B2), pull form.
// cXn = X[x] (symbol consumed when x advances), cYn = Y[y].
template <bool SWITCH>
__device__ __forceinline__ void bwd_cell(const DevModel &m, const EmisTables &t,
                                         const double *B1, int xlo1, int w1,
                                         const double *B2, int xlo2, int w2,
                                         int x, int cXn, int cYn, double out[NS]) {
    const int ju = x - xlo1;           // successor (x, y+1): this cell is its `upper`
    const int jl = ju + 1;             // successor (x+1, y): this cell is its `lower`
    const int jm = x + 1 - xlo2;       // successor (x+1, y+1)
    const bool oku = (unsigned)ju < (unsigned)w1;
    const bool okl = (unsigned)jl < (unsigned)w1;
    const bool okm = (unsigned)jm < (unsigned)w2;
    const double eXn = t.eX[cXn], eYn = t.eY[cYn], eMn = t.eM[cXn * 5 + cYn];
    const double Bm = ldv(B2, jm, okm);
    const double BsY = ldv(B1 + 2 * w1, ju, oku), BlY = ldv(B1 + 4 * w1, ju, oku);
    const double BsX = ldv(B1 + w1, jl, okl), BlX = ldv(B1 + 3 * w1, jl, okl);
    {
        double a = Bm + (eMn + m.tr[S_M * 5 + S_M]);
        a = logadd(a, BsY + (eYn + m.tr[S_M * 5 + S_SY]));
        a = logadd(a, BlY + (eYn + m.tr[S_M * 5 + S_LY]));
        a = logadd(a, BsX + (eXn + m.tr[S_M * 5 + S_SX]));
        a = logadd(a, BlX + (eXn + m.tr[S_M * 5 + S_LX]));
        out[S_M] = a;
    }
    {
        double a = Bm + (eMn + m.tr[S_SX * 5 + S_M]);
        if (SWITCH) a = logadd(a, BsY + (eYn + m.tr[S_SX * 5 + S_SY]));
        a = logadd(a, BsX + (eXn + m.tr[S_SX * 5 + S_SX]));
        out[S_SX] = a;
    }
    {
        double a = Bm + (eMn + m.tr[S_SY * 5 + S_M]);
        a = logadd(a, BsY + (eYn + m.tr[S_SY * 5 + S_SY]));
        if (SWITCH) a = logadd(a, BsX + (eXn + m.tr[S_SY * 5 + S_SX]));
        out[S_SY] = a;
    }
    {
        double a = Bm + (eMn + m.tr[S_LX * 5 + S_M]);
        a = logadd(a, BlX + (eXn + m.tr[S_LX * 5 + S_LX]));
        out[S_LX] = a;
    }
    {
        double a = Bm + (eMn + m.tr[S_LY * 5 + S_M]);
        a = logadd(a, BlY + (eYn + m.tr[S_LY * 5 + S_LY]));
        out[S_LY] = a;
    }
}

// Sequential left-to-right logAdd fold (dpDiagonal_dotProduct order).
__device__ __forceinline__ double fold_logadd(const double *v, int n) {
    double t = PHMM_NEG_INF;
    for (int i = 0; i < n; i++) {
        const double y = v[i];
        if (t < y) {
            t = (t == PHMM_NEG_INF || y - t >= 7.5) ? y : lookup_cubic(y - t) + t;
        } else if (!(y == PHMM_NEG_INF || t - y >= 7.5)) {
            t = lookup_cubic(t - y) + y;
        }
    }
    return t;
}

// Expectation updates of one cell: same transition enumeration as the forward
// cell, p = exp(from + to + (eP + tP) - total), accumulated in 2^-32 fixed point.
// Integer addition is associative, so the sums may be regrouped freely: the 15
// transition counts stay in per-thread registers (accT, flushed once per region)
// and the emission counts are first summed per target state (the symbols are
// fixed for a cell), which leaves 5 shared-memory atomics per cell instead of 30.
constexpr int EXP_NT = 15;   // accT slots: lower M>sX sX>sX sY>sX M>lX lX>lX | middle M sX sY lX lY >M | upper M>sY sY>sY sX>sY M>lY lY>lY
template <bool SWITCH>
__device__ __forceinline__ void expect_cell(const DevModel &m, const EmisTables &t,
                                            const double *F1, int xlo1, int w1,
                                            const double *F2, int xlo2, int w2,
                                            int x, int cX, int cY, const double B[NS], double total,
                                            unsigned long long accT[EXP_NT], unsigned long long *sE) {
    const int jl = x - 1 - xlo1;
    const int ju = jl + 1;
    const int jm = x - 1 - xlo2;
    const bool okl = (unsigned)jl < (unsigned)w1;
    const bool oku = (unsigned)ju < (unsigned)w1;
    const bool okm = (unsigned)jm < (unsigned)w2;
    const bool emit = cX < 4 && cY < 4;
    const int ecol = cX * 4 + cY;
    unsigned long long qE[NS] = {0ull, 0ull, 0ull, 0ull, 0ull};
#define PHMM_EXPECT(SLOT, FROMV, F_, T_, EP)                                                    \
    do {                                                                                        \
        const double pr = exp_det((FROMV) + B[T_] + ((EP) + m.tr[(F_) * 5 + (T_)]) - total);   \
        const unsigned long long q = (unsigned long long)__double2ll_rd(pr * 4294967296.0);    \
        accT[SLOT] += q;                                                                        \
        qE[T_] += q;                                                                            \
    } while (0)
    if (okl) {
        const double eP = t.eX[cX];
        PHMM_EXPECT(0, F1[jl], S_M, S_SX, eP);
        PHMM_EXPECT(1, F1[w1 + jl], S_SX, S_SX, eP);
        if (SWITCH) PHMM_EXPECT(2, F1[2 * w1 + jl], S_SY, S_SX, eP);
        PHMM_EXPECT(3, F1[jl], S_M, S_LX, eP);
        PHMM_EXPECT(4, F1[3 * w1 + jl], S_LX, S_LX, eP);
    }
    if (okm) {
        const double eP = t.eM[cX * 5 + cY];
        PHMM_EXPECT(5, F2[jm], S_M, S_M, eP);
        PHMM_EXPECT(6, F2[w2 + jm], S_SX, S_M, eP);
        PHMM_EXPECT(7, F2[2 * w2 + jm], S_SY, S_M, eP);
        PHMM_EXPECT(8, F2[3 * w2 + jm], S_LX, S_M, eP);
        PHMM_EXPECT(9, F2[4 * w2 + jm], S_LY, S_M, eP);
    }
    if (oku) {
        const double eP = t.eY[cY];
        PHMM_EXPECT(10, F1[ju], S_M, S_SY, eP);
        PHMM_EXPECT(11, F1[2 * w1 + ju], S_SY, S_SY, eP);
        if (SWITCH) PHMM_EXPECT(12, F1[w1 + ju], S_SX, S_SY, eP);
        PHMM_EXPECT(13, F1[ju], S_M, S_LY, eP);
        PHMM_EXPECT(14, F1[4 * w1 + ju], S_LY, S_LY, eP);
    }
#undef PHMM_EXPECT
    if (emit) {
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (qE[s]) atomicAdd(&sE[s * 16 + ecol], qE[s]);
    }
}

// accT slot -> transition index from*5+to
__device__ __constant__ int EXP_SLOT_TR[EXP_NT] = {S_M * 5 + S_SX, S_SX * 5 + S_SX, S_SY * 5 + S_SX, S_M * 5 + S_LX, S_LX * 5 + S_LX,
                                                   S_M * 5 + S_M, S_SX * 5 + S_M, S_SY * 5 + S_M, S_LX * 5 + S_M, S_LY * 5 + S_M,
                                                   S_M * 5 + S_SY, S_SY * 5 + S_SY, S_SX * 5 + S_SY, S_M * 5 + S_LY, S_LY * 5 + S_LY};

template <int NW, bool SWITCH, bool EXPECT>
__global__ void __launch_bounds__(NW * 32) k_fwdbwd(const __grid_constant__ FbArgs a) {
    constexpr int NT = NW * 32;
    const int tid = threadIdx.x;
    __shared__ EmisTables tab;
    __shared__ int s_region;
    __shared__ int s_npairs;
    __shared__ double s_total[2];
    __shared__ unsigned long long sT[25];
    __shared__ unsigned long long sE[80];

    for (int i = tid; i < 25; i += NT) tab.eM[i] = a.m.eM[i];
    if (tid < 5) { tab.eX[tid] = a.m.eX[tid]; tab.eY[tid] = a.m.eY[tid]; }

    const int slot = blockIdx.x;
    double *const fr = a.fring + (int64_t)slot * a.ring_cells * NS;
    DiagRec *const dt = a.dtab + (int64_t)slot * a.dcap;
    double *const br = a.bring + (int64_t)slot * 3 * a.bw * NS;
    double *const dots = a.dots + (int64_t)slot * 2 * a.bw;
    const int e = a.p.expansion;

    for (;;) {
        block_sync<NW>();
        if (tid == 0) s_region = atomicAdd(a.counter, 1);
        block_sync<NW>();
        const int qi = s_region;
        if (qi >= a.n_regions) break;
        const int ridx = a.order[qi];
        const Region reg = a.regions[ridx];
        const uint8_t *X = a.ref + reg.xoff;
        const uint8_t *Y = a.reads + reg.yoff;
        const int lx = reg.lx, ly = reg.ly, nd = lx + ly;
        if (tid == 0) s_npairs = 0;
        if (EXPECT) {
            for (int i = tid; i < 25; i += NT) sT[i] = 0ull;
            for (int i = tid; i < 80; i += NT) sE[i] = 0ull;
        }
        double ll = 0.0;                              // thread 0 only
        unsigned long long accT[EXPECT ? EXP_NT : 1];  // this thread's transition counts of the region
        if (EXPECT) {
#pragma unroll
            for (int k = 0; k < EXP_NT; k++) accT[k] = 0ull;
        }
        if (nd > 0) {
            BandIter it;
            it.init(a.runs + reg.run0, reg.nrun, lx, ly, e);
            // diagonal 0
            if (tid < NS) {
                double v;
                if (reg.ragged_left) v = (tid == S_LX || tid == S_LY) ? 0.0 : PHMM_NEG_INF;
                else v = (tid == S_M) ? 0.0 : PHMM_NEG_INF;
                fr[tid] = v;
            }
            if (tid == 0) { DiagRec r0; r0.off = 0; r0.xlo = 0; r0.w = 1; r0.pad = 0; dt[0] = r0; }
            int off1 = 0, xlo1 = 0, w1 = 1;           // diagonal d-1
            int off2 = 0, xlo2 = 0, w2 = 0;           // diagonal d-2 (w2 = 0: absent)
            int traced_to = 0;
            block_sync<NW>();
            for (int d = 1; d <= nd; d++) {
                int xlo, w;
                it.diag(d, xlo, w);
                int off = off1 + w1;
                if ((int64_t)off + w > a.ring_cells) off = 0;
                if (tid == 0) { DiagRec rc; rc.off = off; rc.xlo = xlo; rc.w = w; rc.pad = 0; dt[d % a.dcap] = rc; }
                {
                    const double *F1 = fr + (int64_t)off1 * NS;
                    const double *F2 = fr + (int64_t)off2 * NS;
                    double *F0 = fr + (int64_t)off * NS;
                    for (int i = tid; i < w; i += NT) {
                        const int x = xlo + i, y = d - x;
                        const int cX = x >= 1 ? X[x - 1] : 4;
                        const int cY = y >= 1 ? Y[y - 1] : 4;
                        double o[NS];
                        fwd_cell<SWITCH>(a.m, tab, F1, xlo1, w1, F2, xlo2, w2, x, cX, cY, o);
#pragma unroll
                        for (int s = 0; s < NS; s++) F0[s * w + i] = o[s];
                    }
                }
                block_sync<NW>();
                const bool at_end = d == nd;
                const bool tbp = d >= traced_to + a.p.min_diags && w <= 2 * e + 1;
                if (at_end || tbp) {
                    // ---------------- traceback from diagonal d ----------------
                    const int traced_from = d - (at_end ? 0 : a.p.tb_diags + 1);
                    const double *endv = (at_end && !reg.ragged_right) ? a.m.endp : a.m.rendp;
                    {
                        double *B0 = br + (int64_t)(d % 3) * a.bw * NS;
                        for (int i = tid; i < w; i += NT) {
#pragma unroll
                            for (int s = 0; s < NS; s++) B0[s * w + i] = endv[s];
                        }
                    }
                    block_sync<NW>();
                    double total = PHMM_NEG_INF;
                    int ncalc = 0;
                    int bxlo1 = 0, bw1 = 0, bxlo2 = 0, bw2 = 0;      // diagonals dd+1, dd+2 of the backward sweep
                    for (int dd = d; dd > traced_to; dd--) {
                        const DiagRec rc = dt[dd % a.dcap];
                        const double *Fd = fr + (int64_t)rc.off * NS;
                        double *Bd = br + (int64_t)(dd % 3) * a.bw * NS;
                        const double *B1 = br + (int64_t)((dd + 1) % 3) * a.bw * NS;
                        const double *B2 = br + (int64_t)((dd + 2) % 3) * a.bw * NS;
                        const bool post = dd <= traced_from;
                        const bool need_total = post && (ncalc % 10 == 0);
                        const bool fuse = post && !need_total && !EXPECT;
                        // forward window neighbours for EXPECT
                        DiagRec f1 = rc, f2 = rc;
                        int fw2 = 0;
                        if (EXPECT && post) {
                            f1 = dt[(dd - 1) % a.dcap];
                            if (dd - 2 >= traced_to) { f2 = dt[(dd - 2) % a.dcap]; fw2 = f2.w; }
                        }
                        if (dd < d || fuse || (EXPECT && post && !need_total)) {
                            for (int i = tid; i < rc.w; i += NT) {
                                const int x = rc.xlo + i, y = dd - x;
                                double o[NS];
                                if (dd < d) {
                                    const int cXn = x < lx ? X[x] : 4;
                                    const int cYn = y < ly ? Y[y] : 4;
                                    bwd_cell<SWITCH>(a.m, tab, B1, bxlo1, bw1, B2, bxlo2, bw2, x, cXn, cYn, o);
#pragma unroll
                                    for (int s = 0; s < NS; s++) Bd[s * rc.w + i] = o[s];
                                } else {
#pragma unroll
                                    for (int s = 0; s < NS; s++) o[s] = endv[s];
                                }
                                if (fuse && x > 0 && y > 0) {
                                    const double lp = (Fd[i] + o[S_M]) - total;
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
                                if (EXPECT && post && !need_total) {
                                    const int cX = x >= 1 ? X[x - 1] : 4;
                                    const int cY = y >= 1 ? Y[y - 1] : 4;
                                    expect_cell<SWITCH>(a.m, tab, fr + (int64_t)f1.off * NS, f1.xlo, f1.w,
                                                        fr + (int64_t)f2.off * NS, f2.xlo, fw2, x, cX, cY, o, total, accT, sE);
                                }
                            }
                        }
                        if (need_total) {
                            block_sync<NW>();
                            // paths through diagonal dd
                            for (int i = tid; i < rc.w; i += NT) {
                                double t = Fd[i] + Bd[i];
#pragma unroll
                                for (int s = 1; s < NS; s++) t = logadd(t, Fd[s * rc.w + i] + Bd[s * rc.w + i]);
                                dots[i] = t;
                            }
                            // paths that step over dd with a match into dd+1
                            const bool have2 = dd < d;
                            if (have2) {
                                const DiagRec r1 = dt[(dd + 1) % a.dcap];
                                const double *Fn = fr + (int64_t)r1.off * NS;
                                for (int i = tid; i < r1.w; i += NT) dots[a.bw + i] = Fn[i] + B1[i];
                            }
                            block_sync<NW>();
                            if (tid < 2) {
                                if (tid == 0) s_total[0] = fold_logadd(dots, rc.w);
                                else s_total[1] = have2 ? fold_logadd(dots + a.bw, bw1) : PHMM_NEG_INF;
                            }
                            block_sync<NW>();
                            total = have2 ? logadd(s_total[0], s_total[1]) : s_total[0];
                            // posterior / expectation pass for this diagonal
                            for (int i = tid; i < rc.w; i += NT) {
                                const int x = rc.xlo + i, y = dd - x;
                                if (!EXPECT) {
                                    if (x > 0 && y > 0) {
                                        const double lp = (Fd[i] + Bd[i]) - total;
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
                                } else {
                                    double o[NS];
#pragma unroll
                                    for (int s = 0; s < NS; s++) o[s] = Bd[s * rc.w + i];
                                    const int cX = x >= 1 ? X[x - 1] : 4;
                                    const int cY = y >= 1 ? Y[y - 1] : 4;
                                    expect_cell<SWITCH>(a.m, tab, fr + (int64_t)f1.off * NS, f1.xlo, f1.w,
                                                        fr + (int64_t)f2.off * NS, f2.xlo, fw2, x, cX, cY, o, total, accT, sE);
                                }
                            }
                        }
                        if (post) { ncalc++; if (EXPECT && tid == 0) ll += total; }
                        bxlo2 = bxlo1; bw2 = bw1;
                        bxlo1 = rc.xlo; bw1 = rc.w;
                        block_sync<NW>();
                    }
                    traced_to = traced_from;
                }
                off2 = off1; xlo2 = xlo1; w2 = w1;
                off1 = off; xlo1 = xlo; w1 = w;
            }
        }
        if (EXPECT) {
#pragma unroll
            for (int k = 0; k < EXP_NT; k++)
                if (accT[k]) atomicAdd(&sT[EXP_SLOT_TR[k]], accT[k]);
        }
        block_sync<NW>();
        if (tid == 0) {
            a.npairs[ridx] = s_npairs;
            if (EXPECT) a.expLL[ridx] = ll;
        }
        if (EXPECT) {
            for (int i = tid; i < 25; i += NT) a.expT[(int64_t)ridx * 25 + i] = sT[i];
            for (int i = tid; i < 80; i += NT) a.expE[(int64_t)ridx * 80 + i] = sE[i];
        }
    }
}

// ---------------------------------------------------------------------------
// decode: reweight by indel probability, banded MEA chain, match runs
// ---------------------------------------------------------------------------
struct DecArgs {
    const Region *regions;
    const Run *runs;
    const int32_t *order;
    int32_t n_regions;
    const int32_t *n_dev;                                     // not NULL: the number of regions in `order` lives on the device
                                                              // (the list k_decode_w leaves behind, phmm_decode_w.cuh)
    int32_t *counter;
    DevParams p;
    const int32_t *px, *py, *pw;
    const int32_t *npairs;
    // per-slot scratch
    int32_t *sumx; int32_t *sumy; int32_t max_lx, max_ly;   // posterior mass per reference / read position
    int32_t *dstart; int32_t *dfill; int32_t max_nd;          // pairs bucketed by anti-diagonal
    int32_t *sidx; int64_t *wre; int32_t *pred; int32_t max_pairs;
    int64_t *colmap;                                          // 2 * max_lx: (diagonal << 32 | sorted position) by column
    int64_t *sring; int32_t *lring; int32_t bw;               // 4 * bw each: three diagonals + a snapshot for skips
    const int32_t *regular; int32_t regular_stride;           // RegionGeom::regular, strided (int32 units)
    // outputs
    int32_t *mrx, *mry, *mrn;                                 // match runs in reverse order (region-local sequence coords)
    int32_t *nmruns;                                          // per region
    int64_t *score;                                           // per region
};

template <int NW>
__global__ void __launch_bounds__(NW * 32) k_decode(const __grid_constant__ DecArgs a) {
    constexpr int NT = NW * 32;
    const int tid = threadIdx.x;
    __shared__ int s_region;
    __shared__ int s_part[NT];
    __shared__ int s_dn;                              // next anti-diagonal that holds a pair (skip search)
    const int slot = blockIdx.x;
    int32_t *const sumx = a.sumx + (int64_t)slot * (a.max_lx + 1);
    int32_t *const sumy = a.sumy + (int64_t)slot * (a.max_ly + 1);
    int32_t *const dstart = a.dstart + (int64_t)slot * (a.max_nd + 4);
    int32_t *const dfill = a.dfill + (int64_t)slot * (a.max_nd + 4);
    int32_t *const sidx = a.sidx + (int64_t)slot * (a.max_pairs + 1);
    int64_t *const wre = a.wre + (int64_t)slot * (a.max_pairs + 1);
    int32_t *const pred = a.pred + (int64_t)slot * (a.max_pairs + 1);
    int64_t *const colmap = a.colmap + (int64_t)slot * 2 * (a.max_lx + 2);
    int64_t *const sr = a.sring + (int64_t)slot * 4 * a.bw;
    int32_t *const lr = a.lring + (int64_t)slot * 4 * a.bw;
    const int n_regions = a.n_dev ? *a.n_dev : a.n_regions;

    for (;;) {
        block_sync<NW>();
        if (tid == 0) s_region = atomicAdd(a.counter, 1);
        block_sync<NW>();
        const int qi = s_region;
        if (qi >= n_regions) break;
        const int ridx = a.order[qi];
        const Region reg = a.regions[ridx];
        const int lx = reg.lx, ly = reg.ly, nd = lx + ly;
        const int np = min(a.npairs[ridx], reg.pair_cap);
        const int32_t *px = a.px + reg.pair_off, *py = a.py + reg.pair_off, *pw = a.pw + reg.pair_off;
        if (nd == 0) {
            if (tid == 0) { a.nmruns[ridx] = 0; a.score[ridx] = 0; }
            continue;
        }
        // 1. clear
        for (int i = tid; i < lx; i += NT) sumx[i] = 0;
        for (int i = tid; i < ly; i += NT) sumy[i] = 0;
        for (int i = tid; i < nd + 3; i += NT) { dstart[i] = 0; dfill[i] = 0; }
        for (int i = tid; i < 2 * (lx + 2); i += NT) colmap[i] = -1;
        block_sync<NW>();
        // 2. posterior mass per position, pair count per anti-diagonal (matrix diagonal = x+y+2)
        for (int i = tid; i < np; i += NT) {
            atomicAdd(&sumx[px[i]], pw[i]);
            atomicAdd(&sumy[py[i]], pw[i]);
            atomicAdd(&dstart[px[i] + py[i] + 2], 1);
        }
        block_sync<NW>();
        // 3. exclusive scan of dstart[0..nd+1]
        {
            const int n = nd + 2;
            const int chunk = (n + NT - 1) / NT;
            const int b = tid * chunk, e_ = min(b + chunk, n);
            int s = 0;
            for (int i = b; i < e_; i++) s += dstart[i];
            s_part[tid] = s;
            block_sync<NW>();
            if (tid == 0) {
                int acc = 0;
                for (int i = 0; i < NT; i++) { const int v = s_part[i]; s_part[i] = acc; acc += v; }
            }
            block_sync<NW>();
            int acc = s_part[tid];
            for (int i = b; i < e_; i++) { const int v = dstart[i]; dstart[i] = acc; acc += v; }
        }
        block_sync<NW>();
        // 4. reweight (getIndelProbabilities / reweightAlignedPairs) and bucket
        for (int i = tid; i < np; i += NT) {
            const int x = px[i], y = py[i], w = pw[i];
            int ipx = PROB_1 - sumx[x]; if (ipx < 0) ipx = 0;
            int ipy = PROB_1 - sumy[y]; if (ipy < 0) ipy = 0;
            int64_t wr = (int64_t)w - __double2ll_rz(a.p.gap_gamma * (double)((int64_t)ipx + (int64_t)ipy));
            if ((double)w < a.p.match_gamma * (double)PROB_1) wr = 0;
            wre[i] = wr;
            const int dg = x + y + 2;
            const int pos = dstart[dg] + atomicAdd(&dfill[dg], 1);
            sidx[pos] = i;
        }
        block_sync<NW>();
        // 5. wavefront over the band.
        //    Skipping pairless stretches.  In a regular band (RegionGeom::regular: both edges move right by 0 or 1 per
        //    diagonal) cell (d1, x') reaches cell (d2, x) through band cells iff x' <= x <= x' + (d2 - d1), and without
        //    pairs in between scores only propagate, `lower` (the smaller x) preferred on ties: cell (d2, x) ends up
        //    with the LEFTMOST maximum of diagonal d1 over x' in [x - (d2 - d1), x].  So after a run of pairless
        //    diagonals the block looks for the next diagonal dn that holds a pair, fills dn-2 and dn-1 from the current
        //    diagonal with that window maximum, walks the band iterator there, and resumes at dn; if no pair is left
        //    the final cell is evaluated the same way.  The chained-global alignments of this path spend most of their
        //    diagonals in pairless leading / trailing deletions.
        const bool regular = a.regular[(int64_t)ridx * a.regular_stride] != 0;
        constexpr int SKIP_AFTER = 8, SKIP_MIN = 64;
        int empty_run = 0;
        bool finished = false;
        int64_t fin_s = 0; int fin_k = -1;                // thread 0: result when the sweep ends by a skip
        BandIter it;
        it.init(a.runs + reg.run0, reg.nrun, lx, ly, a.p.expansion);
        if (tid == 0) { sr[0] = 0; lr[0] = -1; }
        // pairs of diagonal 2 into the column map (parity 0)
        for (int k = dstart[2] + tid; k < dstart[3]; k += NT) {
            const int pi = sidx[k];
            colmap[(2 & 1) * (lx + 2) + px[pi] + 1] = ((int64_t)2 << 32) | (uint32_t)k;
        }
        int xlo1 = 0, w1 = 1, xlo2 = 0, w2 = 0;
        block_sync<NW>();
        for (int d = 1; d <= nd; d++) {
            int xlo, w;
            it.diag(d, xlo, w);
            // stage pairs of diagonal d+1
            int pairs_next = 0;
            if (d + 1 <= nd) {
                const int k0 = dstart[d + 1], k1 = dstart[d + 2];
                pairs_next = k1 - k0;
                for (int k = k0 + tid; k < k1; k += NT) {
                    const int pi = sidx[k];
                    colmap[((d + 1) & 1) * (lx + 2) + px[pi] + 1] = ((int64_t)(d + 1) << 32) | (uint32_t)k;
                }
            }
            const int64_t *S1 = sr + (int64_t)((d - 1) % 3) * a.bw;
            const int32_t *L1 = lr + (int64_t)((d - 1) % 3) * a.bw;
            const int64_t *S2 = sr + (int64_t)((d + 1) % 3) * a.bw;   // (d-2) mod 3
            const int32_t *L2 = lr + (int64_t)((d + 1) % 3) * a.bw;
            int64_t *S0 = sr + (int64_t)(d % 3) * a.bw;
            int32_t *L0 = lr + (int64_t)(d % 3) * a.bw;
            for (int i = tid; i < w; i += NT) {
                const int x = xlo + i;
                int64_t bs = -1; int32_t bl = -1;
                const int jl = x - 1 - xlo1, ju = jl + 1;
                if ((unsigned)jl < (unsigned)w1) { bs = S1[jl]; bl = L1[jl]; }
                if ((unsigned)ju < (unsigned)w1) { const int64_t us = S1[ju]; if (us > bs) { bs = us; bl = L1[ju]; } }
                const int64_t cm = colmap[(d & 1) * (lx + 2) + x];
                if ((int)(cm >> 32) == d) {
                    const int k = (int)(uint32_t)cm;
                    const int64_t wr = wre[sidx[k]];
                    const int jm = x - 1 - xlo2;
                    if (wr > 0 && (unsigned)jm < (unsigned)w2) {
                        const int64_t ms = S2[jm];
                        if (ms >= 0) {
                            pred[k] = L2[jm];
                            const int64_t cand = ms + wr;
                            if (cand > bs) { bs = cand; bl = k; }
                        }
                    }
                }
                S0[i] = bs; L0[i] = bl;
            }
            xlo2 = xlo1; w2 = w1; xlo1 = xlo; w1 = w;
            block_sync<NW>();
            // ---- skip a pairless stretch (all conditions are uniform over the block) ----
            empty_run = pairs_next == 0 ? empty_run + 1 : 0;
            if (regular && empty_run >= SKIP_AFTER && d + SKIP_MIN < nd) {
                // next diagonal > d+1 that holds a pair
                int dn = nd + 1;
                for (int base = d + 2; base <= nd; base += NT) {
                    if (tid == 0) s_dn = 0x7fffffff;
                    block_sync<NW>();
                    const int dd = base + tid;
                    if (dd <= nd && dstart[dd + 1] > dstart[dd]) atomicMin(&s_dn, dd);
                    block_sync<NW>();
                    const int f = s_dn;
                    block_sync<NW>();
                    if (f != 0x7fffffff) { dn = f; break; }
                }
                if (dn > nd || dn - d >= SKIP_MIN) {
                    // snapshot of the current diagonal d (band xlo1, w1 after the shift above)
                    const int xs = xlo1, ws = w1, xse = xlo1 + w1 - 1;
                    int64_t *const sS = sr + (int64_t)3 * a.bw;
                    int32_t *const sL = lr + (int64_t)3 * a.bw;
                    {
                        const int64_t *Sd = sr + (int64_t)(d % 3) * a.bw;
                        const int32_t *Ld = lr + (int64_t)(d % 3) * a.bw;
                        for (int i = tid; i < ws; i += NT) { sS[i] = Sd[i]; sL[i] = Ld[i]; }
                    }
                    block_sync<NW>();
                    if (dn > nd) {
                        // no pair left: the final cell (nd, lx) takes the leftmost maximum of its window on diagonal d
                        if (tid == 0) {
                            const int delta = nd - d;
                            const int ja = max(xs, lx - delta), jb = min(xse, lx);
                            int64_t best = -1; int bl = -1;
                            for (int j = ja; j <= jb; j++) { const int64_t v = sS[j - xs]; if (v > best) { best = v; bl = sL[j - xs]; } }
                            fin_s = best; fin_k = bl;
                        }
                        finished = true;
                        break;
                    }
                    // walk the band to dn-1, keeping the extents of dn-2 and dn-1
                    for (int dd = d + 1; dd <= dn - 1; dd++) {
                        int xa, wa;
                        it.diag(dd, xa, wa);
                        xlo2 = xlo1; w2 = w1; xlo1 = xa; w1 = wa;
                    }
                    for (int t = dn - 2; t <= dn - 1; t++) {
                        const int xt = t == dn - 2 ? xlo2 : xlo1, wt = t == dn - 2 ? w2 : w1, delta = t - d;
                        int64_t *St = sr + (int64_t)(t % 3) * a.bw;
                        int32_t *Lt = lr + (int64_t)(t % 3) * a.bw;
                        for (int i = tid; i < wt; i += NT) {
                            const int x = xt + i;
                            const int ja = max(xs, x - delta), jb = min(xse, x);
                            int64_t best = -1; int bl = -1;
                            for (int j = ja; j <= jb; j++) { const int64_t v = sS[j - xs]; if (v > best) { best = v; bl = sL[j - xs]; } }
                            St[i] = best; Lt[i] = bl;
                        }
                    }
                    // pairs of diagonal dn into the column map, as the sweep would have done at dn-1
                    for (int k = dstart[dn] + tid; k < dstart[dn + 1]; k += NT) {
                        const int pi = sidx[k];
                        colmap[(dn & 1) * (lx + 2) + px[pi] + 1] = ((int64_t)dn << 32) | (uint32_t)k;
                    }
                    block_sync<NW>();
                    d = dn - 1;                         // the loop continues with diagonal dn
                    empty_run = 0;
                }
                else empty_run = 0;                     // a pair is near: sweep on, search again after the next pairless run
            }
        }
        // 6. traceback into match runs (reverse order), thread 0
        if (tid == 0) {
            const int64_t fs = finished ? fin_s : sr[(int64_t)(nd % 3) * a.bw];
            int k = finished ? fin_k : lr[(int64_t)(nd % 3) * a.bw];
            int nr = 0;
            int rx = -2, ry = -2, rn = 0;             // current run: starts at (rx,ry), length rn
            while (k >= 0) {
                const int pi = sidx[k];
                const int x = px[pi], y = py[pi];
                if (rn > 0 && x == rx - 1 && y == ry - 1) { rx = x; ry = y; rn++; }
                else {
                    if (rn > 0 && nr < reg.mrun_cap) {
                        a.mrx[reg.mrun_off + nr] = rx; a.mry[reg.mrun_off + nr] = ry; a.mrn[reg.mrun_off + nr] = rn;
                    }
                    if (rn > 0) nr++;
                    rx = x; ry = y; rn = 1;
                }
                k = pred[k];
            }
            if (rn > 0) {
                if (nr < reg.mrun_cap) {
                    a.mrx[reg.mrun_off + nr] = rx; a.mry[reg.mrun_off + nr] = ry; a.mrn[reg.mrun_off + nr] = rn;
                }
                nr++;
            }
            a.nmruns[ridx] = nr;
            a.score[ridx] = fs;
        }
    }
}

// gathers region match runs into a dense array; one block per region
__global__ void k_compact(const Region *regions, const int32_t *nmruns, const int64_t *dst_off, int n_regions,
                          const int32_t *mrx, const int32_t *mry, const int32_t *mrn,
                          int32_t *ox, int32_t *oy, int32_t *on) {
    const int r = blockIdx.x;
    if (r >= n_regions) return;
    const Region reg = regions[r];
    const int n = min(nmruns[r], reg.mrun_cap);
    const int64_t dst = dst_off[r];
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        ox[dst + i] = mrx[reg.mrun_off + i];
        oy[dst + i] = mry[reg.mrun_off + i];
        on[dst + i] = mrn[reg.mrun_off + i];
    }
}

// Per-reference-position base expectations (replaces the parse-and-sum loop over every read's
// --outputAllPosteriorProbs file, reference nanopore/analyses/marginAlignSnpCaller.py:149-155): every posterior pair
// (x, y, prob) of a region adds prob (1e-7 units) to acc[5 * (reference index of x) + read base at y], columns A C G T
// and "other": the reference creates the position's entry for any pair and adds only ACGT read bases to it, so the fifth
// column keeps what it needs to tell "no pair here" from "only non-ACGT pairs here".  One block per region, a pair per thread; integer atomics
// commute, so the table does not depend on the order of pairs, regions, batches or ranks.  read_mask (one byte per
// read of the batch, may be NULL) selects the reads that contribute: the caller's coverage samples
// (marginAlignSnpCaller.py:91-97) reuse one resident set of posteriors.
__global__ void k_base_expect(const Region *regions, const int32_t *npairs, int n_regions, const uint8_t *reads,
                              const uint8_t *read_mask, const int32_t *px, const int32_t *py, const int32_t *pw,
                              unsigned long long *acc) {
    for (int r = blockIdx.x; r < n_regions; r += gridDim.x) {
        const Region reg = regions[r];
        if (read_mask && !read_mask[reg.read]) continue;
        const int n = min(npairs[r], reg.pair_cap);
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            const int64_t q = reg.pair_off + k;
            const int b = reads[reg.yoff + py[q]];
            atomicAdd(&acc[(reg.xoff + px[q]) * 5 + min(b, 4)], (unsigned long long)pw[q]);
        }
    }
}

}  // namespace phmm
