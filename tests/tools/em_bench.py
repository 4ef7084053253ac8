"""E-step throughput (BASELINE.json configs[3] shape at reduced read count): Baum-Welch iterations over N reads x 8 kb
vs a 50 kb reference, realign options of the reference's EM (--diagonalExpansion=10 --splitMatrixBiggerThanThis=300,
reference nanopore/analyses/utils.py:511).  Prints one JSON line; the CPU oracle is timed on a sample beside it.
usage: python tests/tools/em_bench.py [reads] [iterations]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import oracle                                               # noqa: E402  (CPU baseline leg only)
from nanopore_b200 import capi, em, synth                   # noqa: E402
from nanopore_b200.engine import Realigner                  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
b = synth.make_batch(n, 8000, 50000, seed=4)
band = int(os.environ.get("EM_BAND", "10"))                 # the reference's EM runs at 10 (utils.py:511); 50 = the realignment metric's band
p = em.parseRealignOptions("--diagonalExpansion=%d --splitMatrixBiggerThanThis=300" % band)
hmm = em._stock_start("fiveStateAsymmetric")
r = Realigner(0)
if os.environ.get("EM_WARPS"):
    r.ctx.set_option("warps", int(os.environ["EM_WARPS"]))
r.set_reference(b.ref)
r.set_hmm(hmm)
r.expectations(b, p)                                        # warm-up (allocations)
t0 = time.perf_counter()
rl = em.expectationMaximisation(r, b, hmm, p, iters, trainEmissions=True)
dt = time.perf_counter() - t0
st = r.ctx.stats()
cells = r.cells
# CPU oracle: one E-step over a sample, one read per task on all cores
from concurrent.futures import ThreadPoolExecutor
cores = os.cpu_count() or 1
m = oracle.Model()
op = oracle.make_params(expansion=band, split_side=300)
idx = list(range(min(n, int(os.environ.get("EM_CPU_SAMPLE", str(4 * cores))))))
t0 = time.perf_counter()
with ThreadPoolExecutor(cores) as ex:
    list(ex.map(lambda i: oracle.expectations_fixed(m, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), op), idx))
cdt = time.perf_counter() - t0
print(json.dumps({"workload": "EM: %d reads x 8 kb vs 50 kb, band %d, split 300, %d iterations" % (n, band, iters),
                  "reads_iter_per_s": n * iters / dt, "s_per_iteration": dt / iters, "cells_per_iteration": cells,
                  "gcells_per_s": cells * iters / dt / 1e9, "regions": st["n_regions"], "ms_fwdbwd_last": st["ms_fwdbwd"],
                  "running_likelihoods": rl, "monotone": all(b2 >= a - 1e-9 * abs(a) for a, b2 in zip(rl[1:], rl[2:])),
                  "cpu_reads_iter_per_s": len(idx) / cdt, "cpu_cores": cores}), flush=True)
