"""The two claims behind k_decode_w ("decode on the envelope of the pairs"), checked in plain Python with
MANY ties, against the full sweep of tests/test_decode_skip_theory.py (the oracle's mea_region semantics).

Claim 1 (band independence).  In a regular band (both edges move right by 0 or 1 per diagonal; every cell reachable,
rows and columns of the band contiguous) the sweep computes, at every band cell (x, y), the best chain score over
the pairs inside the rectangle [0..x] x [0..y] and -- through its tie rule `lower, then upper if strictly greater,
then middle + weight if strictly greater` -- the lexicographically smallest (x, y) among the pairs that end a best
chain.  Neither depends on the band.  So the sweep may run on ANY regular band that holds every pair, (0,0) and
(lx,ly): same final score, same chain.  The tightest such band is
    lo[d] = d + min_{d' <= d} (T[d'] - d'),  T[d] = min_{d' >= d} minx[d']
    hi[d] = max(P[d], Q[d] + d, lo[d]),      P[d] = max_{d' <= d} maxx[d'],  Q[d] = max_{d' >= d} (maxx[d'] - d')
with minx / maxx the extreme x of the cells diagonal d must hold -- the pairs, the cell (x-1, y-1) every pair's match
step comes from (in the band by construction: the forward kernel gives a pair no mass otherwise), (0,0) and (lx,ly) --
four scans.

Claim 2 (the kernel's data layout).  Cells live in column (x - (d >> 1)) mod W of the buffer of parity d & 1, updated
in place; two columns either side of the band of the diagonal a buffer holds are kept at -1; pairs are applied in two
phases around the cell update (read `middle` before, compare after); pairless stretches are skipped by the
leftmost-maximum rule.  model_sweep() below is that procedure, statement for statement what the CUDA kernel does,
with W small enough that columns wrap."""
import random

from test_decode_skip_theory import full_sweep, random_regular_band, regular

INF = 1 << 30


def narrow_band(lx, ly, pairs):
    nd = lx + ly
    minx = [INF] * (nd + 1)
    maxx = [-INF] * (nd + 1)
    minx[0] = maxx[0] = 0
    minx[nd] = min(minx[nd], lx)
    maxx[nd] = max(maxx[nd], lx)
    for (x, y) in pairs:
        minx[x + y] = min(minx[x + y], x)
        maxx[x + y] = max(maxx[x + y], x)
        minx[x + y - 2] = min(minx[x + y - 2], x - 1)          # the cell a match step comes from
        maxx[x + y - 2] = max(maxx[x + y - 2], x - 1)
    T = minx[:]
    Q = [maxx[d] - d if maxx[d] > -INF else -INF for d in range(nd + 1)]
    for d in range(nd - 1, -1, -1):
        T[d] = min(T[d], T[d + 1])
        Q[d] = max(Q[d], Q[d + 1])
    lo, hi = [0] * (nd + 1), [0] * (nd + 1)
    m, P = INF, -INF
    for d in range(nd + 1):
        m = min(m, T[d] - d)
        P = max(P, maxx[d])
        lo[d] = d + m
        hi[d] = max(P, Q[d] + d, lo[d])
    return lo, hi


def model_sweep(lo, hi, lx, ly, pairs, W, skip_min=4):
    """The kernel's procedure on the band lo/hi (widths + 4 <= W)."""
    nd = lx + ly
    by_diag = {}
    for (x, y), (w, k) in pairs.items():
        by_diag.setdefault(x + y, []).append((x, w, k))
    nxt = [nd + 1] * (nd + 3)
    for d in range(nd, -1, -1):
        nxt[d] = d if d in by_diag else nxt[d + 1]
    S = [[-1] * W, [-1] * W]
    L = [[-1] * W, [-1] * W]
    S[0][0] = 0
    pred = {}
    fin = None
    d = 1
    while d <= nd:
        par, h = d & 1, d >> 1
        own_s, own_l, oth_s, oth_l = S[par], L[par], S[par ^ 1], L[par ^ 1]
        dl = -1 if par else 0
        # phase P1: pairs read their middle predecessor (own column, before it is overwritten)
        pend = []
        for (x, w, k) in by_diag.get(d, []):
            pos = (x - h) % W
            ms, ml = own_s[pos], own_l[pos]
            if w > 0 and ms >= 0:
                pred[k] = ml
                pend.append((pos, ms + w, k))
        # phase C: cells lo-2 .. hi+2; out-of-band ones become -1 (the sentinels)
        new = []
        for x in range(lo[d] - 2, hi[d] + 3):
            c = x - h
            if lo[d] <= x <= hi[d]:
                bs, bl = oth_s[(c + dl) % W], oth_l[(c + dl) % W]
                us = oth_s[(c + dl + 1) % W]
                if us > bs:
                    bs, bl = us, oth_l[(c + dl + 1) % W]
            else:
                bs, bl = -1, -1
            new.append((c % W, bs, bl))
        for pos, bs, bl in new:
            own_s[pos], own_l[pos] = bs, bl
        # phase P2
        for pos, cand, k in pend:
            if cand > own_s[pos]:
                own_s[pos], own_l[pos] = cand, k
        # skip a pairless stretch
        dn = nxt[d + 1]
        if dn - d >= skip_min:
            snap = [(own_s[(x - h) % W], own_l[(x - h) % W]) for x in range(lo[d], hi[d] + 1)]

            def window_max(x, delta):
                best = (-1, -1)
                for xp in range(max(lo[d], x - delta), min(hi[d], x) + 1):
                    if snap[xp - lo[d]][0] > best[0]:
                        best = snap[xp - lo[d]]
                return best

            if dn > nd:
                fin = window_max(lx, nd - d)
                break
            for t in (dn - 2, dn - 1):
                bs_, bl_ = S[t & 1], L[t & 1]
                for i in range(W):
                    bs_[i], bl_[i] = -1, -1
                if t == d:                      # dn - 2 == d cannot happen with skip_min >= 3; kept for clarity
                    continue
                for x in range(lo[t], hi[t] + 1):
                    v = window_max(x, t - d)
                    bs_[(x - (t >> 1)) % W], bl_[(x - (t >> 1)) % W] = v
            d = dn
            continue
        d += 1
    if fin is None:
        pos = (lx - (nd >> 1)) % W
        fin = (S[nd & 1][pos], L[nd & 1][pos])
    s, k = fin
    chain = []
    while k >= 0:
        chain.append(k)
        k = pred[k]
    return s, chain[::-1]


def random_case(rng):
    lx, ly = rng.randint(3, 40), rng.randint(3, 40)
    lo, hi = random_regular_band(rng, lx, ly)
    if not regular(lo, hi, lx, ly):
        return None
    # pairs as the forward / backward kernel produces them: a match cell whose diagonal predecessor is in the band
    cells = [(x, d - x) for d in range(2, lx + ly + 1) for x in range(lo[d], hi[d] + 1)
             if x >= 1 and d - x >= 1 and lo[d - 2] <= x - 1 <= hi[d - 2]]
    if not cells:
        return None
    centre = rng.choice(cells)
    near = [c for c in cells if abs(c[0] + c[1] - centre[0] - centre[1]) <= rng.randint(1, 8)] + rng.sample(cells, min(2, len(cells)))
    chosen = rng.sample(near, min(len(near), rng.randint(1, 12)))
    pairs = {c: (rng.choice([0, 1, 1, 2]), k) for k, c in enumerate(dict.fromkeys(chosen))}
    return lx, ly, lo, hi, pairs


def test_envelope_is_a_regular_band_holding_every_pair():
    rng = random.Random(11)
    n = 0
    for _ in range(400):
        case = random_case(rng)
        if case is None:
            continue
        lx, ly, _, _, pairs = case
        lo, hi = narrow_band(lx, ly, pairs)
        assert regular(lo, hi, lx, ly), (lx, ly, pairs, lo, hi)
        assert lo[0] == hi[0] == 0
        for (x, y) in pairs:
            assert lo[x + y] <= x <= hi[x + y] and lo[x + y - 2] <= x - 1 <= hi[x + y - 2]
        n += 1
    assert n > 150


def test_sweep_on_the_envelope_equals_sweep_on_the_band():
    rng = random.Random(13)
    n = 0
    for _ in range(600):
        case = random_case(rng)
        if case is None:
            continue
        lx, ly, lo, hi, pairs = case
        nlo, nhi = narrow_band(lx, ly, pairs)
        assert full_sweep(nlo, nhi, lx, ly, pairs) == full_sweep(lo, hi, lx, ly, pairs), (lx, ly, lo, hi, pairs)
        n += 1
    assert n > 250


def test_kernel_procedure_equals_full_sweep():
    rng = random.Random(17)
    n = 0
    for _ in range(600):
        case = random_case(rng)
        if case is None:
            continue
        lx, ly, lo, hi, pairs = case
        nlo, nhi = narrow_band(lx, ly, pairs)
        wmax = max(h - l + 1 for l, h in zip(nlo, nhi))
        W = 8
        while W < wmax + 4:
            W *= 2
        for skip_min in (3, 4, 1000):
            assert model_sweep(nlo, nhi, lx, ly, pairs, W, skip_min) == full_sweep(lo, hi, lx, ly, pairs), (lx, ly, pairs, W, skip_min)
        n += 1
    assert n > 250
