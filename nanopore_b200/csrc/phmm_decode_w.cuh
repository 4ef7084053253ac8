// phmm_decode_w.cuh -- k_decode_w: the maximum-expected-accuracy chain on the ENVELOPE of the posterior pairs, one warp
// per region, wavefront state in shared memory.
//
// k_decode (phmm_kernels.cuh) sweeps every cell of the band the forward / backward kernel used: 64-1200 cells per
// diagonal on this path.  Posterior pairs >= threshold occupy a strip a few cells wide.  In a regular band (both
// edges move right by 0 or 1 cell per diagonal -- RegionGeom::regular) the sweep's value at a band cell (x, y) is the
// best chain score over the pairs inside [0..x] x [0..y], and its tie rule (lower, then upper if strictly greater,
// then middle + weight if strictly greater) yields the lexicographically smallest (x, y) among the pairs that end a
// best chain: neither depends on the band.  So the sweep may run on ANY regular band that holds every pair, the cell
// (x-1, y-1) each pair's match step comes from (inside the original band, or the forward kernel would have given the
// pair no mass), (0,0) and (lx,ly) -- same score, same chain, same CIGAR.  The tightest such band follows from the
// extreme x of those cells per diagonal by four scans (below); on the bench workload it holds 2-3 % of the band's
// cells and is at most a few dozen cells wide.  tests/test_decode_narrow_theory.py checks both the claim and the
// procedure of this kernel (column ring, sentinels, two-phase pairs, skips) in plain Python on random bands with
// many ties; on the GPU the CIGARs are compared with the checker's and with k_decode's (tests/test_gpu_parity.py).
//
// Layout.  Cell (d, x) lives in column (x - (d >> 1)) mod DW_WC of the shared-memory buffer of parity d & 1 and is
// updated in place: its `middle` predecessor (d-2, x-1) is the same column, `lower` / `upper` sit in the same and the
// next column of the other buffer (one column lower on odd diagonals).  Two columns either side of the band of the
// diagonal a buffer holds are kept at -1 (unreachable), which is as far as the next two diagonals can reach, so no
// read needs an in-band test.  Pairs are bucketed by diagonal into sorted arrays (x, y, reweighted mass); a diagonal's
// pairs read their `middle` before the cells are updated and are compared with the updated cell afterwards.  Band
// edges, bucket offsets and the next pair-holding diagonal are staged 31 diagonals at a time in registers (one lane
// per diagonal, prefetched a chunk ahead), the pairs 32 at a time.  Stretches without pairs are skipped by the
// leftmost-maximum rule of k_decode.  The traceback walks the predecessor links through 32-entry windows held in
// registers (links point to lower sorted positions, mostly a few entries away).
//
// Regions whose band is not regular, whose envelope is wider than DW_WC - 4 or that hold more than 32 pairs on one
// diagonal are appended to a list that k_decode processes afterwards on the original band.
//
// Replaces, with k_decode, the alignment-extraction half of `cactus_realign` (reference nanopore/analyses/utils.py:587);
// algorithm per SURVEY.md Appendix A.9.  Integer arithmetic only: bit-exact.
#pragma once
#include "phmm_device.cuh"

namespace phmm {

constexpr int DW_WC = 128;          // columns per parity buffer (power of two)
constexpr int DW_BLOCKS = 32;       // one-warp blocks per SM the launch bound allows for
constexpr int DW_SKIP_MIN = 24;     // pairless diagonals worth a skip (>= 3)
constexpr int DW_INF = 0x3fffffff;

struct DecWArgs {
    const Region *regions;
    const int32_t *order;
    int32_t n_regions;
    int32_t *counter;
    DevParams p;
    const int32_t *px, *py, *pw;
    const int32_t *npairs;
    const int32_t *regular; int32_t regular_stride;           // RegionGeom::regular, strided (int32 units)
    // per-slot scratch (slot = blockIdx.x)
    int32_t *sumx; int32_t *sumy; int32_t max_lx, max_ly;     // posterior mass per reference / read position
    int32_t *dstart; int32_t *nxt; int32_t *blo; int32_t *bhi; int32_t nd_stride;   // per diagonal: bucket offsets, next
                                                              // diagonal holding a pair (bucket fill counts before), envelope
    int32_t *bx; int32_t *by; int64_t *bwr; int32_t *pred; int32_t max_pairs;       // pairs in bucket order (matrix coordinates)
    int32_t *fb_list; int32_t *fb_count;                      // regions left to k_decode
    // outputs
    int32_t *mrx, *mry, *mrn;                                 // match runs in reverse order (region-local sequence coords)
    int32_t *nmruns;
    int64_t *score;
};

__device__ __forceinline__ int dw_scan_add(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    return v;
}
__device__ __forceinline__ int dw_prefix_min(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = min(v, t); }
    return v;
}
__device__ __forceinline__ int dw_prefix_max(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = max(v, t); }
    return v;
}
__device__ __forceinline__ int dw_suffix_min(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v = min(v, t); }
    return v;
}
__device__ __forceinline__ int dw_suffix_max(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_down_sync(0xffffffffu, v, o); if (lane + o < 32) v = max(v, t); }
    return v;
}

__global__ void __launch_bounds__(32, DW_BLOCKS) k_decode_w(const __grid_constant__ DecWArgs a) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int M = DW_WC - 1;
    const int lane = threadIdx.x;
    __shared__ long long sS[2][DW_WC];                 // chain score of the cell in this column, -1: unreachable / outside
    __shared__ int sL[2][DW_WC];                       // sorted position of the last pair of that chain, -1: none
    __shared__ long long snS[DW_WC];                   // snapshot of a diagonal (skips)
    __shared__ int snL[DW_WC];
    const int slot = blockIdx.x;
    int32_t *const sumx = a.sumx + (int64_t)slot * (a.max_lx + 1);
    int32_t *const sumy = a.sumy + (int64_t)slot * (a.max_ly + 1);
    int32_t *const dstart = a.dstart + (int64_t)slot * a.nd_stride;
    int32_t *const nxt = a.nxt + (int64_t)slot * a.nd_stride;
    int32_t *const blo = a.blo + (int64_t)slot * a.nd_stride;
    int32_t *const bhi = a.bhi + (int64_t)slot * a.nd_stride;
    int32_t *const bx = a.bx + (int64_t)slot * (a.max_pairs + 1);
    int32_t *const by = a.by + (int64_t)slot * (a.max_pairs + 1);
    int64_t *const bwr = a.bwr + (int64_t)slot * (a.max_pairs + 1);
    int32_t *const pred = a.pred + (int64_t)slot * (a.max_pairs + 1);

    for (;;) {
        __syncwarp();
        int qi = 0;
        if (lane == 0) qi = atomicAdd(a.counter, 1);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= a.n_regions) break;
        const int ridx = a.order[qi];
        const Region reg = a.regions[ridx];
        const int lx = reg.lx, ly = reg.ly, nd = lx + ly;
        if (nd == 0) {
            if (lane == 0) { a.nmruns[ridx] = 0; a.score[ridx] = 0; }
            continue;
        }
        if (a.regular[(int64_t)ridx * a.regular_stride] == 0) {
            if (lane == 0) a.fb_list[atomicAdd(a.fb_count, 1)] = ridx;
            continue;
        }
        const int np = min(a.npairs[ridx], reg.pair_cap);
        if (np == 0) {
            // no pair: every cell of a regular band is reachable with score 0 and an empty chain (the split-off leading /
            // trailing deletions of a long contig are such regions, a million diagonals each)
            if (lane == 0) { a.nmruns[ridx] = 0; a.score[ridx] = 0; }
            continue;
        }
        const int32_t *const px = a.px + reg.pair_off, *const py = a.py + reg.pair_off, *const pw = a.pw + reg.pair_off;

        // 1. clear
        for (int i = lane; i < lx; i += 32) sumx[i] = 0;
        for (int i = lane; i < ly; i += 32) sumy[i] = 0;
        for (int i = lane; i < nd + 3; i += 32) { dstart[i] = 0; nxt[i] = 0; blo[i] = DW_INF; bhi[i] = -DW_INF; }
        __syncwarp();
        // 2. posterior mass per position, pairs per anti-diagonal (matrix diagonal = x + y + 2), and the extreme x of the
        //    cells each diagonal must hold: the pairs, the cell each pair's match step comes from, (0,0) and (lx,ly)
        if (lane == 0) {
            atomicMin(&blo[0], 0); atomicMax(&bhi[0], 0);
            atomicMin(&blo[nd], lx); atomicMax(&bhi[nd], lx);
        }
        for (int i = lane; i < np; i += 32) {
            const int x = px[i], y = py[i], w = pw[i];
            atomicAdd(&sumx[x], w);
            atomicAdd(&sumy[y], w);
            const int dg = x + y + 2;
            atomicAdd(&dstart[dg], 1);
            atomicMin(&blo[dg], x + 1); atomicMax(&bhi[dg], x + 1);
            atomicMin(&blo[dg - 2], x); atomicMax(&bhi[dg - 2], x);
        }
        __syncwarp();
        // 3. exclusive scan of dstart[0 .. nd+1]; most pairs on one diagonal
        int maxc = 0;
        {
            const int n = nd + 2;
            int carry = 0;
            for (int base = 0; base < n; base += 128) {
                int v[4];
#pragma unroll
                for (int u = 0; u < 4; u++) { const int i = base + 32 * u + lane; v[u] = i < n ? dstart[i] : 0; }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = base + 32 * u + lane;
                    maxc = max(maxc, v[u]);
                    const int inc = dw_scan_add(v[u], lane);
                    if (i < n) dstart[i] = carry + inc - v[u];
                    carry += __shfl_sync(FULL, inc, 31);
                }
            }
            maxc = __reduce_max_sync(FULL, maxc);
        }
        __syncwarp();
        // 4. reweight (getIndelProbabilities / reweightAlignedPairs) and bucket by diagonal; nxt counts the fill
        for (int i = lane; i < np; i += 32) {
            const int x = px[i], y = py[i], w = pw[i];
            int ipx = PROB_1 - sumx[x]; if (ipx < 0) ipx = 0;
            int ipy = PROB_1 - sumy[y]; if (ipy < 0) ipy = 0;
            int64_t wr = (int64_t)w - __double2ll_rz(a.p.gap_gamma * (double)((int64_t)ipx + (int64_t)ipy));
            if ((double)w < a.p.match_gamma * (double)PROB_1) wr = 0;
            const int dg = x + y + 2;
            const int pos = dstart[dg] + atomicAdd(&nxt[dg], 1);
            bx[pos] = x + 1; by[pos] = y + 1; bwr[pos] = wr;
        }
        __syncwarp();
        // 5. downward scans over the diagonals nd+1 .. 0:
        //      blo <- T[d] - d,  T[d] = min of the required x over the diagonals >= d
        //      bhi <- Q[d] + d,  Q[d] = max of (required x - diagonal) over the diagonals >= d
        //      nxt <- first diagonal >= d that holds a pair (nd + 1: none)
        {
            const int n = nd + 2;
            int cT = DW_INF, cQ = -DW_INF, cE = DW_INF;
            for (int base = ((n - 1) >> 7) << 7; base >= 0; base -= 128) {
                int vt[4], vq[4], ve[4];
#pragma unroll
                for (int u = 3; u >= 0; u--) {
                    const int i = base + 32 * u + lane;
                    const int mn = i <= nd ? blo[i] : DW_INF;
                    const int mx = i <= nd ? bhi[i] : -DW_INF;
                    const int c0 = i <= nd ? dstart[i] : 0, c1 = i <= nd ? dstart[i + 1] : 0;
                    vt[u] = mn;
                    vq[u] = mx > -DW_INF ? mx - i : -DW_INF;
                    ve[u] = c1 > c0 ? i : DW_INF;
                }
#pragma unroll
                for (int u = 3; u >= 0; u--) {
                    const int i = base + 32 * u + lane;
                    const int t = min(dw_suffix_min(vt[u], lane), cT);
                    const int q = max(dw_suffix_max(vq[u], lane), cQ);
                    const int e = min(dw_suffix_min(ve[u], lane), cE);
                    cT = __shfl_sync(FULL, t, 0); cQ = __shfl_sync(FULL, q, 0); cE = __shfl_sync(FULL, e, 0);
                    if (i <= nd) { blo[i] = t - i; bhi[i] = q + i; }
                    if (i <= nd + 1) nxt[i] = min(e, nd + 1);
                }
            }
        }
        __syncwarp();
        // 6. upward scans: lo[d] = d + min_{d' <= d} (T[d'] - d'),  hi[d] = max(max_{d' <= d} (Q[d'] + d'), lo[d]); widest diagonal
        int wmax = 0;
        {
            int cA = DW_INF, cB = -DW_INF;
            for (int base = 0; base <= nd; base += 128) {
                int va[4], vb[4];
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = base + 32 * u + lane;
                    va[u] = i <= nd ? blo[i] : DW_INF;
                    vb[u] = i <= nd ? bhi[i] : -DW_INF;
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    const int i = base + 32 * u + lane;
                    const int pa = min(dw_prefix_min(va[u], lane), cA);
                    const int pb = max(dw_prefix_max(vb[u], lane), cB);
                    cA = __shfl_sync(FULL, pa, 31); cB = __shfl_sync(FULL, pb, 31);
                    if (i <= nd) {
                        const int lo = i + pa, hi = max(pb, lo);
                        blo[i] = lo; bhi[i] = hi;
                        wmax = max(wmax, hi - lo + 1);
                    }
                }
            }
            wmax = __reduce_max_sync(FULL, wmax);
        }
        __syncwarp();
        if (maxc > 32 || wmax + 4 > DW_WC) {
            if (lane == 0) a.fb_list[atomicAdd(a.fb_count, 1)] = ridx;
            continue;
        }

        // 7. wavefront over the envelope
        for (int i = lane; i < 2 * DW_WC; i += 32) { (&sS[0][0])[i] = -1; (&sL[0][0])[i] = -1; }
        __syncwarp();
        if (lane == 0) sS[0][0] = 0;                                  // diagonal 0: the cell (0,0), column 0 of the even buffer
        __syncwarp();
        // diagonals D0 .. D0+30 from the registers of lanes 0 .. 31 (a diagonal also needs the entries of the next one)
        int D0 = 1, c_lo, c_hi, c_ds, c_nx;
        int nD, n_lo, n_hi, n_ds, n_nx;
#define DW_CHUNK(D, vlo, vhi, vds, vnx)                                                                      \
        do {                                                                                                 \
            const int i_ = (D) + lane;                                                                       \
            vlo = i_ <= nd ? blo[i_] : 0; vhi = i_ <= nd ? bhi[i_] : 0;                                      \
            vds = i_ <= nd + 1 ? dstart[i_] : 0; vnx = i_ <= nd + 1 ? nxt[i_] : nd + 1;                      \
        } while (0)
        DW_CHUNK(D0, c_lo, c_hi, c_ds, c_nx);
        nD = D0 + 31;
        DW_CHUNK(nD, n_lo, n_hi, n_ds, n_nx);
        // pairs kW .. kW+31 (bucket order) in the registers of lanes 0 .. 31
        int kW = 0;
        int w_x = lane < np ? bx[lane] : 0;
        long long w_wr = lane < np ? bwr[lane] : 0;
        bool finished = false;
        long long fin_s = 0; int fin_k = -1;
        int d = 1;
        while (d <= nd) {
            if (d - D0 > 30) {
                if (d == nD) { c_lo = n_lo; c_hi = n_hi; c_ds = n_ds; c_nx = n_nx; D0 = nD; }
                else { D0 = d; DW_CHUNK(D0, c_lo, c_hi, c_ds, c_nx); }
                nD = D0 + 31;
                DW_CHUNK(nD, n_lo, n_hi, n_ds, n_nx);
            }
            const int j = d - D0;
            const int lo = __shfl_sync(FULL, c_lo, j), hi = __shfl_sync(FULL, c_hi, j);
            const int k0 = __shfl_sync(FULL, c_ds, j), k1 = __shfl_sync(FULL, c_ds, j + 1);
            const int dn = __shfl_sync(FULL, c_nx, j + 1);            // next diagonal > d that holds a pair
            const int npd = k1 - k0;                                  // <= 32
            const int par = d & 1, h = d >> 1, dl = par ? -1 : 0;
            long long *const ownS = sS[par]; int *const ownL = sL[par];
            const long long *const othS = sS[par ^ 1]; const int *const othL = sL[par ^ 1];
            // pairs of this diagonal, one per lane: read the middle predecessor (own column) before the cells overwrite it
            bool pv = false; int ppos = 0, pk = 0; long long pcand = 0;
            if (npd > 0) {
                if (k0 < kW || k1 > kW + 32) {
                    kW = k0;
                    w_x = kW + lane < np ? bx[kW + lane] : 0;
                    w_wr = kW + lane < np ? bwr[kW + lane] : 0;
                }
                const int src = (k0 - kW + lane) & 31;
                const int xk = __shfl_sync(FULL, w_x, src);
                const long long wr = __shfl_sync(FULL, w_wr, src);
                if (lane < npd) {
                    pk = k0 + lane;
                    ppos = (xk - h) & M;
                    const long long ms = ownS[ppos];
                    if (wr > 0 && ms >= 0) { pred[pk] = ownL[ppos]; pcand = ms + wr; pv = true; }
                }
                __syncwarp();
            }
            // cells lo-2 .. hi+2: the band and two sentinel columns either side
            for (int x = lo - 2 + lane; x <= hi + 2; x += 32) {
                const int c = x - h;
                long long bs = -1; int bl = -1;
                if (x >= lo && x <= hi) {
                    const int p1 = (c + dl) & M, p2 = (c + dl + 1) & M;
                    bs = othS[p1]; bl = othL[p1];
                    const long long us = othS[p2];
                    if (us > bs) { bs = us; bl = othL[p2]; }
                }
                ownS[c & M] = bs; ownL[c & M] = bl;
            }
            __syncwarp();
            if (npd > 0) {
                if (pv && pcand > ownS[ppos]) { ownS[ppos] = pcand; ownL[ppos] = pk; }
                __syncwarp();
            }
            // a pairless stretch ahead: cell (t, x) of a later diagonal holds the leftmost maximum of this one over
            // x' in [x - (t - d), x] (regular band, scores only propagate); fill dn-2 and dn-1 that way and resume at dn
            if (dn - d >= DW_SKIP_MIN) {
                for (int i = lane; i <= hi - lo; i += 32) { const int pos = (lo + i - h) & M; snS[i] = ownS[pos]; snL[i] = ownL[pos]; }
                __syncwarp();
                if (dn > nd) {
                    const int ja = max(lo, lx - (nd - d)), jb = min(hi, lx);
                    long long best = -1; int bl = -1;
                    for (int q = ja; q <= jb; q++) { const long long v = snS[q - lo]; if (v > best) { best = v; bl = snL[q - lo]; } }
                    fin_s = best; fin_k = bl;
                    finished = true;
                    break;
                }
                for (int t = dn - 2; t <= dn - 1; t++) {
                    const int lot = blo[t], hit = bhi[t], delta = t - d, ht = t >> 1;
                    long long *const tS = sS[t & 1]; int *const tL = sL[t & 1];
                    for (int i = lane; i < DW_WC; i += 32) { tS[i] = -1; tL[i] = -1; }
                    __syncwarp();
                    for (int x = lot + lane; x <= hit; x += 32) {
                        const int ja = max(lo, x - delta), jb = min(hi, x);
                        long long best = -1; int bl = -1;
                        for (int q = ja; q <= jb; q++) { const long long v = snS[q - lo]; if (v > best) { best = v; bl = snL[q - lo]; } }
                        tS[(x - ht) & M] = best; tL[(x - ht) & M] = bl;
                    }
                    __syncwarp();
                }
                d = dn;
                continue;
            }
            d++;
        }
#undef DW_CHUNK
        if (!finished) {
            const int pos = (lx - (nd >> 1)) & M;
            fin_s = sS[nd & 1][pos]; fin_k = sL[nd & 1][pos];
        }
        // 8. traceback into match runs (reverse order).  Every lane follows the chain (warp-uniform); the links, x and y of
        //    32 consecutive sorted positions sit in registers and are read by shuffle.
        {
            int k = fin_k, nr = 0, rx = -2, ry = -2, rn = 0;
            int wb = -64, t_x = 0, t_y = 0, t_p = -1;
            int guard = np + 1;
            while (k >= 0 && guard-- > 0) {
                if (k < wb || k >= wb + 32) {
                    wb = max(k - 31, 0);
                    const int i = wb + lane;
                    t_x = i < np ? bx[i] : 0; t_y = i < np ? by[i] : 0; t_p = i < np ? pred[i] : -1;
                }
                const int src = k - wb;
                const int x = __shfl_sync(FULL, t_x, src) - 1, y = __shfl_sync(FULL, t_y, src) - 1;
                const int kn = __shfl_sync(FULL, t_p, src);
                if (rn > 0 && x == rx - 1 && y == ry - 1) { rx = x; ry = y; rn++; }
                else {
                    if (rn > 0) {
                        if (lane == 0 && nr < reg.mrun_cap) {
                            a.mrx[reg.mrun_off + nr] = rx; a.mry[reg.mrun_off + nr] = ry; a.mrn[reg.mrun_off + nr] = rn;
                        }
                        nr++;
                    }
                    rx = x; ry = y; rn = 1;
                }
                k = kn;
            }
            if (rn > 0) {
                if (lane == 0 && nr < reg.mrun_cap) {
                    a.mrx[reg.mrun_off + nr] = rx; a.mry[reg.mrun_off + nr] = ry; a.mrn[reg.mrun_off + nr] = rn;
                }
                nr++;
            }
            if (lane == 0) { a.nmruns[ridx] = nr; a.score[ridx] = fin_s; }
        }
    }
}

}  // namespace phmm
