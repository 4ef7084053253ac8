"""Kernel-only timing of the realignment path under library options (GPU box).
usage: python scripts/tune.py READS "opt=val,opt=val" ["opt=val" ...]

Loads build/libphmm_tune.so (built HERE with -DPHMM_TUNE by `python -c "from nanopore_b200 import build; build.build_tune()"`,
it travels with the snapshot) when it exists, so that the timing_experiment switches are available; the shipped
library rejects them."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nanopore_b200 import capi, synth
_tune = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "build", "libphmm_tune.so")
if os.environ.get("TUNE_LIB"):                       # a variant built by scripts/build_variant.sh
    capi.LIB_PATH = os.path.abspath(os.environ["TUNE_LIB"])
elif os.path.exists(_tune) and os.environ.get("SHIPPED_LIB") != "1":
    capi.LIB_PATH = _tune

n = int(sys.argv[1])
band = int(os.environ.get("BAND", "50"))
L, R = int(os.environ.get("READ_LEN", "10000")), int(os.environ.get("REF_LEN", "50000"))
gf = os.environ.get("GLOBAL_FORM", "1") == "1"
ch = dict(sub=float(os.environ.get("SUB", "0.05")), ins=float(os.environ.get("INS", "0.04")), dele=float(os.environ.get("DEL", "0.06")))
lengths = synth.pareto_lengths(n, seed=5) if os.environ.get("PARETO") == "1" else None      # BASELINE.json configs[4] shape
b = synth.make_batch(n, L, R, seed=1001, global_form=gf, lengths=lengths, **ch)
for spec in sys.argv[2:] or [""]:
    ctx = capi.PhmmContext(0)
    for kv in [s for s in spec.split(",") if s]:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    if os.environ.get("MEM_BUDGET_GB"):                  # fewer resident regions per SM: how throughput scales with occupancy
        ctx.set_memory_budget(int(float(os.environ["MEM_BUDGET_GB"]) * 1e9))
    ctx.set_reference(b.ref)
    ctx.prepare(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, capi.default_params(band=band))
    ctx.run()
    ts = []
    for _ in range(int(os.environ.get("REPS", "2"))):
        ctx.run()
        ts.append(ctx.stats()["ms_fwdbwd"])
    st = ctx.stats()
    print("%-40s fwdbwd %9.2f ms  decode %7.2f ms  slots %4d  slotMB %7.1f  cells %.3e  -> %.1f Gcell/s  %.0f GB/s(80B/cell)  regions %d  %.0f reads/s" % (
        spec or "(default)", min(ts), st["ms_decode"], st["n_slots"], st["slot_bytes"] / 1e6, st["cells"],
        st["cells"] / min(ts) * 1e-6, 80.0 * st["cells"] / min(ts) * 1e-6, st["n_regions"], n / ((min(ts) + st["ms_decode"]) * 1e-3)), flush=True)
    ctx.close()
