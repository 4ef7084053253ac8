"""First-light check on a GPU box: GPU library vs CPU oracle on small seeded batches."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import oracle
from nanopore_b200 import synth, capi
from nanopore_b200.hmm import Hmm


def compare(ctx, model, b, params, op, label):
    t = time.time()
    ops, off, post = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, params, want_posteriors=True)
    tg = time.time() - t
    st = ctx.stats()
    bad_ops = bad_pairs = 0
    tc = 0
    for i in range(b.n):
        X = b.ref[b.ref_start[i]:b.ref_end[i]]
        t = time.time()
        r = oracle.realign(model, X, b.read(i), b.ops(i), op)
        tc += time.time() - t
        g = ops[off[i]:off[i + 1]]
        if len(g) != len(r["ops"]) or (g != r["ops"]).any():
            bad_ops += 1
            if bad_ops <= 2:
                print("  ops differ read", i, synth.unpack_ops(g)[:12], "vs", synth.unpack_ops(r["ops"])[:12])
        o = np.lexsort((r["py"], r["px"]))
        ox, oy, ow = r["px"][o], r["py"][o], r["pw"][o]
        gx = post["ref_pos"][post["off"][i]:post["off"][i + 1]]
        gy = post["read_pos"][post["off"][i]:post["off"][i + 1]]
        gw = post["prob_1e7"][post["off"][i]:post["off"][i + 1]]
        if len(gx) != len(ox) or (gx != ox).any() or (gy != oy).any() or (gw != ow).any():
            bad_pairs += 1
            if bad_pairs <= 2:
                print("  pairs differ read", i, len(gx), len(ox))
                n = min(len(gx), len(ox))
                d = np.flatnonzero((gx[:n] != ox[:n]) | (gy[:n] != oy[:n]) | (gw[:n] != ow[:n]))
                print("   first diffs", d[:5], [(gx[k], gy[k], gw[k], ox[k], oy[k], ow[k]) for k in d[:5]])
    print("%s: reads %d regions %d cells %d  bad_ops %d bad_pairs %d  gpu %.3fs (fb %.1f ms dec %.1f ms slots %d slotMB %.1f nw?) cpu %.2fs" % (
        label, b.n, st["n_regions"], st["cells"], bad_ops, bad_pairs, tg, st["ms_fwdbwd"], st["ms_decode"], st["n_slots"], st["slot_bytes"] / 1e6, tc))
    return bad_ops + bad_pairs


def main():
    print("version", capi.load_library().phmm_version())
    bad = 0
    # stock model
    ctx = capi.PhmmContext(0)
    model = oracle.Model()
    for (n, L, R, band, seed, gf) in [(4, 300, 900, 10, 1, True), (8, 1000, 3000, 20, 2, True), (6, 2000, 2000, 50, 3, False), (3, 3000, 12000, 50, 4, True)]:
        b = synth.make_batch(n, L, R, seed=seed, global_form=gf)
        ctx.set_reference(b.ref)
        bad += compare(ctx, model, b, capi.default_params(band=band), oracle.make_params(expansion=band), "stock band=%d" % band)
    # trained asymmetric model
    h = Hmm.loadHmm(os.path.join(os.path.dirname(__file__), "..", "golden", "blasr_hmm_0.txt"))
    t, e = h.arrays()
    ctx2 = capi.PhmmContext(0, t, e, 1)
    model2 = oracle.Model(t, e)
    b = synth.make_batch(6, 1500, 4000, seed=7)
    ctx2.set_reference(b.ref)
    bad += compare(ctx2, model2, b, capi.default_params(band=10), oracle.make_params(expansion=10), "trained band=10")
    # splitting
    b = synth.make_batch(3, 400, 20000, seed=9)
    ctx2.set_reference(b.ref)
    bad += compare(ctx2, model2, b, capi.default_params(band=10, split_side=100), oracle.make_params(expansion=10, split_side=100), "split_side=100")
    # expectations
    b = synth.make_batch(4, 800, 2400, seed=11)
    ctx2.set_reference(b.ref)
    out = ctx2.expectations_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, capi.default_params(band=10, split_side=300))
    T = np.zeros(25); E = np.zeros(80); ll = 0.0
    op = oracle.make_params(expansion=10, split_side=300)
    for i in range(b.n):
        T, E, ll, _ = oracle.expectations(model2, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), op, T, E, ll)
    dT = np.abs(out[:25] - T).max(); dE = np.abs(out[25:105] - E).max(); dl = abs(out[105] - ll)
    print("expectations: max|dT| %.3g max|dE| %.3g |dLL| %.3g  exact=%s" % (dT, dE, dl, (out[:25] == T).all() and (out[25:105] == E).all() and out[105] == ll))
    bad += int(dT > 1e-6 or dE > 1e-6)
    print("TOTAL BAD", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
