"""Both decode kernels on one bench-shaped batch (GPU box): k_decode_w on the envelope of the posterior pairs (default)
against k_decode on the band (option decode_block) -- identical CIGAR ops required, decode time of each printed.
usage: python scripts/decode_compare.py [READS] [READ_LEN] [REF_LEN] [BAND]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanopore_b200 import capi, synth

n, L, R, band = (int(sys.argv[i]) if len(sys.argv) > i else v for i, v in ((1, 2368), (2, 10000), (3, 50000), (4, 50)))
lengths = synth.pareto_lengths(n, seed=5) if os.environ.get("PARETO") == "1" else None
b = synth.make_batch(n, L, R, seed=1001, lengths=lengths)
out = {"reads": n, "read_len": L, "ref_len": R, "band": band}
res = {}
for name, opts in (("k_decode_w", {}), ("k_decode", {"decode_block": 1})):
    ctx = capi.PhmmContext(0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.set_reference(b.ref)
    ctx.prepare(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, capi.default_params(band=band))
    ctx.run()
    ctx.run()
    st = ctx.stats()
    ops, off, _ = ctx.fetch()
    res[name] = (ops, off)
    out[name] = {"ms_decode": st["ms_decode"], "ms_fwdbwd": st["ms_fwdbwd"], "regions": st["n_regions"], "launches": st["launches"]}
    ctx.close()
out["identical_ops"] = bool(np.array_equal(res["k_decode_w"][0], res["k_decode"][0]) and np.array_equal(res["k_decode_w"][1], res["k_decode"][1]))
print(json.dumps(out))
sys.exit(0 if out["identical_ops"] else 1)
