"""The claim behind k_decode's skipping of pairless stretches, checked in plain Python with MANY ties.

Full sweep (the oracle's mea_region semantics, oracle/phmm_oracle.c): every band cell takes `lower` (x-1), then `upper`
if strictly greater, then `middle + weight` if strictly greater.  Claim: in a regular band (both edges move right by 0
or 1 per diagonal) a cell separated from diagonal d1 by pairless diagonals holds the LEFTMOST maximum of diagonal d1
over x' in [x - (d2 - d1), x].  The test builds random regular bands and random pairs with tiny integer weights (so
that equal scores are the rule, not the exception) and compares a sweep that skips pairless stretches that way with
the full sweep: same final score, same chain."""
import random


def random_regular_band(rng, lx, ly):
    """lo[d], hi[d] for d = 0..lx+ly: monotone edges with steps 0/1, containing (0,0) and (lx,ly), inside the matrix."""
    nd = lx + ly
    lo, hi = [0], [0]
    for d in range(1, nd + 1):
        xmin, xmax = max(0, d - ly), min(lx, d)
        l = lo[-1] + (1 if rng.random() < 0.5 else 0)
        h = hi[-1] + (1 if rng.random() < 0.6 else 0)
        l = min(max(l, xmin), xmax)
        h = min(max(h, xmin, l), xmax)
        # keep the steps legal after clipping
        l = min(max(l, lo[-1]), lo[-1] + 1)
        h = min(max(h, hi[-1]), hi[-1] + 1)
        if l > h:
            l = h
        lo.append(l); hi.append(h)
    return lo, hi


def regular(lo, hi, lx, ly):
    nd = lx + ly
    ok = lo[nd] == lx == hi[nd]
    for d in range(1, nd + 1):
        ok &= lo[d - 1] <= lo[d] <= lo[d - 1] + 1 and hi[d - 1] <= hi[d] <= hi[d - 1] + 1 and lo[d] <= hi[d]
        ok &= max(0, d - ly) <= lo[d] and hi[d] <= min(lx, d)
    return ok


def full_sweep(lo, hi, lx, ly, pairs):
    """pairs: {(cx, cy): (weight, id)} at matrix cells. Returns (score, chain ids)."""
    nd = lx + ly
    S = {(0, 0): (0, -1)}
    pred = {}
    for d in range(1, nd + 1):
        for x in range(lo[d], hi[d] + 1):
            best = S.get((d - 1, x - 1), (-1, -1))
            up = S.get((d - 1, x), (-1, -1))
            if up[0] > best[0]:
                best = up
            y = d - x
            if (x, y) in pairs and d >= 2:
                w, k = pairs[(x, y)]
                mm = S.get((d - 2, x - 1))
                if w > 0 and mm is not None and mm[0] >= 0:
                    pred[k] = mm[1]
                    if mm[0] + w > best[0]:
                        best = (mm[0] + w, k)
            S[(d, x)] = best
    s, k = S[(nd, lx)]
    chain = []
    while k >= 0:
        chain.append(k)
        k = pred[k]
    return s, chain[::-1]


def skipping_sweep(lo, hi, lx, ly, pairs, skip_after=2, skip_min=4):
    nd = lx + ly
    has = [False] * (nd + 2)
    for (x, y) in pairs:
        has[x + y] = True
    S = {(0, 0): (0, -1)}
    pred = {}
    empty_run, d = 0, 1

    def window_max(d1, x, delta):
        best = (-1, -1)
        for xp in range(max(lo[d1], x - delta), min(hi[d1], x) + 1):      # left to right, strict >: leftmost maximum
            if S[(d1, xp)][0] > best[0]:
                best = S[(d1, xp)]
        return best

    while d <= nd:
        for x in range(lo[d], hi[d] + 1):
            best = S.get((d - 1, x - 1), (-1, -1))
            up = S.get((d - 1, x), (-1, -1))
            if up[0] > best[0]:
                best = up
            y = d - x
            if (x, y) in pairs and d >= 2:
                w, k = pairs[(x, y)]
                mm = S.get((d - 2, x - 1))
                if w > 0 and mm is not None and mm[0] >= 0:
                    pred[k] = mm[1]
                    if mm[0] + w > best[0]:
                        best = (mm[0] + w, k)
            S[(d, x)] = best
        empty_run = empty_run + 1 if not has[d + 1] else 0
        if empty_run >= skip_after and d + skip_min < nd:
            dn = next((t for t in range(d + 2, nd + 1) if has[t]), nd + 1)
            if dn > nd:
                s, k = window_max(d, lx, nd - d)
                break
            if dn - d >= skip_min:
                for t in (dn - 2, dn - 1):
                    for x in range(lo[t], hi[t] + 1):
                        S[(t, x)] = window_max(d, x, t - d)
                d = dn
                empty_run = 0
                continue
            empty_run = 0
        d += 1
    else:
        s, k = S[(nd, lx)]
    chain = []
    while k >= 0:
        chain.append(k)
        k = pred[k]
    return s, chain[::-1]


def test_skipping_equals_full_sweep_with_many_ties():
    rng = random.Random(7)
    cases = 0
    for trial in range(400):
        lx, ly = rng.randint(3, 40), rng.randint(3, 40)
        lo, hi = random_regular_band(rng, lx, ly)
        if not regular(lo, hi, lx, ly):
            continue
        cells = [(x, d - x) for d in range(2, lx + ly + 1) for x in range(lo[d], hi[d] + 1) if x >= 1 and d - x >= 1]
        if not cells:
            continue
        # clustered pairs (so that long pairless stretches exist), weights in {0, 1, 2}: ties everywhere
        centre = rng.choice(cells)
        near = [c for c in cells if abs(c[0] + c[1] - centre[0] - centre[1]) <= rng.randint(1, 8)] + rng.sample(cells, min(2, len(cells)))
        chosen = rng.sample(near, min(len(near), rng.randint(1, 12)))
        pairs = {c: (rng.choice([0, 1, 1, 2]), k) for k, c in enumerate(dict.fromkeys(chosen))}
        assert skipping_sweep(lo, hi, lx, ly, pairs) == full_sweep(lo, hi, lx, ly, pairs), (lx, ly, lo, hi, pairs)
        cases += 1
    assert cases > 150
