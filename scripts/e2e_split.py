import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np
from nanopore_b200 import capi, synth
b = synth.make_batch(10000, 10000, 50000, seed=1001)
ctx = capi.PhmmContext(0); ctx.set_reference(b.ref); p = capi.default_params(band=50)
ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, p)
for _ in range(2):
    t0=time.perf_counter(); ctx.prepare(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, p); t1=time.perf_counter()
    ctx.run(); t2=time.perf_counter(); ops,off,_=ctx.fetch(); t3=time.perf_counter()
    st=ctx.stats()
    print("prepare %.1f ms (geometry kernels %.1f)  run %.1f ms  fetch %.1f ms  ops %d" % ((t1-t0)*1e3, st["ms_geometry"], (t2-t1)*1e3, (t3-t2)*1e3, len(ops)))
