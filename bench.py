#!/usr/bin/env python
"""bench.py -- aligned reads/sec of the pair-HMM realignment hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--reads R] ...
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path (banded 5-state forward/backward, posterior extraction, MEA decode
to CIGARs) over one batch of synthetic reads: BASELINE.json configs[1], 10,000 reads x 10 kb against a
50 kb reference, band (diagonalExpansion) 50, chained-global guide alignments.  With
N > 1 every rank realigns its own 10,000-read shard (weak scaling, reads are independent units; the
only collectives are the start-up broadcast of reference + HMM and the max-reduction of the timings).

value      reads/s, inputs and outputs resident in HBM (phmm_batch_run: kernels only)
e2e        reads/s through the host-buffer entry point phmm_realign_batch (host planning, H2D, kernels,
           D2H, CIGAR assembly all inside the timed region)
roofline   dominant kernel k_fb2 (forward / backward / posterior): 80 algorithmic bytes per DP cell (5 fp64
           forward values written once and read back once) / its CUDA-event time on
           the library's stream, against MEASURED_PEAKS.json hbm_gbs; `traffic` is the ncu DRAM traffic per launch
           scaled from the committed capture (profiles/), bytes per cell x cells of this launch
cpu_baseline  the CPU oracle (oracle/phmm_oracle.c, a port: the reference's cactus_realign sources are
           absent) on a bounded sample of the same reads, one read per task on all host cores

--impl reference times that CPU implementation as the reference arm (the reference's own
realigner cannot be built or installed here).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "aligned reads/sec (10 kb reads, band=50)"
BYTES_PER_CELL = 80.0


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=10000, help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=10000)
    ap.add_argument("--ref-len", type=int, default=50000)
    ap.add_argument("--band", type=int, default=50)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-sample", type=int, default=0, help="reads in the CPU baseline sample (0 = 4 x cores, about 15 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="library option (phmm_set_option) for kernel experiments, e.g. register_path=0; recorded in config")
    return ap.parse_args()


TRAFFIC_SRC = "profiles/ncu_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one k_fb2 launch / its cells"


def traffic_per_cell():
    """DRAM bytes per DP cell of k_fb2 from the committed `ncu --set full` capture (None if absent)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        t = json.load(open(p))
        return (float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])) / float(t["cells"])
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def make_workload(args, rank):
    from nanopore_b200 import synth
    # same reference on every rank (seed), different reads per rank
    rng_ref = np.random.default_rng(args.seed)
    ref = synth.random_reference(args.ref_len, rng_ref)
    return synth.make_batch(args.reads, args.read_len, args.ref_len, seed=args.seed * 1000 + rank, ref=ref)


def cpu_realign_sample(b, idx, band, threads):
    """Oracle on reads idx, one read per task over `threads` host threads (ctypes drops the GIL).
    Returns (seconds, cells)."""
    import oracle
    from concurrent.futures import ThreadPoolExecutor
    model = oracle.Model()
    op = oracle.make_params(expansion=band)
    oracle.lib()

    def one(i):
        r = oracle.realign(model, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), op)
        return r["cells"], r["ops"]

    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(one, idx))
    dt = time.perf_counter() - t0
    return dt, int(sum(c for c, _ in res)), [o for _, o in res]


def run_reference(args, rank, world):
    """Reference arm: the CPU implementation of the path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or 4 * cores          # 4 reads per thread: keeps the threads busy to the end of a step
    a2 = argparse.Namespace(**vars(args))
    a2.reads = sample * (args.steps + args.warmup)
    b = make_workload(a2, 0)
    times, cells = [], 0
    for s in range(args.warmup + args.steps):
        idx = list(range(s * sample, (s + 1) * sample))
        dt, c, _ = cpu_realign_sample(b, idx, args.band, cores)
        if s >= args.warmup:
            times.append(dt); cells += c
    tot = sum(times)
    v = sample * args.steps / tot
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "reads/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, sample_reads=sample),
            "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port",
                             "sample": "%d reads per step (one read per task, %d threads), same generator and seed as the GPU arm"
                                       % (sample, cores)},
            "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "cells_per_read": cells / max(1, sample * args.steps)}
    print(json.dumps(line), flush=True)


def workload_config(args, sample_reads=None):
    c = {"workload": "BASELINE.json configs[1]: %d synthetic 2D-like reads x %d bp vs %d bp random reference, band=%d, "
                     "chained-global guide alignments, stock 5-state HMM" % (args.reads, args.read_len, args.ref_len, args.band),
         "reads_per_gpu_per_step": args.reads, "read_len": args.read_len, "ref_len": args.ref_len, "band": args.band,
         "split_side": 3000, "gap_gamma": 0.5, "match_gamma": 0.0, "channel": "sub 5% ins 4% del 6%, geometric p=0.6",
         "seed": args.seed}
    if sample_reads is not None:
        c["reads_per_step_in_this_arm"] = sample_reads
    return c


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return 0

    import torch
    import torch.distributed as dist
    from nanopore_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    b = make_workload(args, rank)
    # reference + HMM: rank 0 owns them, NCCL broadcast to the other ranks (the only data-path collective)
    ref_t = torch.from_numpy(b.ref.copy()).to(dev)
    if world > 1:
        dist.broadcast(ref_t, src=0)
        assert np.array_equal(ref_t.cpu().numpy(), b.ref)
    ctx = capi.PhmmContext(local)
    for kv in args.opt:
        k, v = kv.split("=")
        ctx.set_option(k, int(v))
    stream = torch.cuda.Stream(device=dev)           # the library launches on this stream; the events below time it
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    ctx.set_reference(ref_t.cpu().numpy())
    params = capi.default_params(band=args.band)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- kernel-only: everything resident in HBM ----------------
    ctx.prepare(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, params)
    for _ in range(args.warmup):
        ctx.run()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fb_ms, dec_ms, launches = 0.0, 0.0, 0
    for _ in range(args.steps):
        ctx.run()
        st = ctx.stats()
        fb_ms += st["ms_fwdbwd"]; dec_ms += st["ms_decode"]; launches += st["run_launches"]
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    st = ctx.stats()
    ops, off, _ = ctx.fetch()
    assert off[-1] == len(ops) and len(off) == b.n + 1

    # ---------------- end to end: host buffers in, CIGARs out ----------------
    e2e_ms = None
    h2d = int(b.reads.nbytes + b.read_off.nbytes + b.ref_start.nbytes + b.ref_end.nbytes + b.in_ops.nbytes + b.in_off.nbytes)
    d2h = 0
    if not args.no_e2e:
        pin = [torch.from_numpy(a).pin_memory().numpy() for a in (b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off)]
        ctx.realign_batch(*pin, params)                           # warm-up (allocations)
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(max(1, args.steps)):
            ops2, off2, _ = ctx.realign_batch(*pin, params)
            launches_e2e = ctx.stats()["launches"]
        e1.record()
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / max(1, args.steps)
        est = ctx.stats()
        h2d += int(est["h2d_bytes"]) - int(b.reads.nbytes)        # + planned regions / anchor runs the library uploads
        d2h = int(est["d2h_bytes"])
        assert np.array_equal(ops2, ops) and np.array_equal(off2, off)

    # ---------------- reduce over ranks: max time, sums of work ----------------
    vec = torch.tensor([ms, fb_ms, dec_ms, e2e_ms or 0.0], dtype=torch.float64, device=dev)
    tot = torch.tensor([float(b.n), float(st["cells"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms, fb_ms, dec_ms, e2e_max = [float(x) for x in vec.tolist()]
    reads_total, cells_total = [float(x) for x in tot.tolist()]

    if rank == 0:
        peak, peak_src = peaks()
        cells_rank = float(st["cells"])
        achieved = BYTES_PER_CELL * cells_rank / (fb_ms / args.steps * 1e-3) / 1e9      # algorithmic GB/s of the k_fb2 launch (per GPU)
        line = {"metric": METRIC, "value": reads_total * args.steps / (ms * 1e-3), "unit": "reads/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": dict(workload_config(args), l2="working set (forward windows, %.1f GB of scratch) exceeds the 126 MB L2"
                               % (st["slot_bytes"] * st["n_slots"] / 1e9), parallelism="reads sharded by rank, %d per GPU" % args.reads),
                "roofline": {"bound": "hbm", "kernel": "k_fb2", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic_per_cell() * cells_rank if traffic_per_cell() else None,
                             "traffic_source": TRAFFIC_SRC,
                             "algorithmic_bytes_per_launch": BYTES_PER_CELL * cells_rank, "cells_per_launch": cells_rank,
                             "kernel_ms": fb_ms / args.steps, "kernel_share_of_step": fb_ms / ms,
                             "frac_of_datasheet_8TBs": achieved / 8000.0},
                "clocks": clocks, "gpu_launches": launches,
                "cells_per_read": cells_total / reads_total, "pairs": st["pairs"], "regions": st["n_regions"],
                "resident_regions": st["n_slots"], "decode_ms": dec_ms / args.steps}
        if args.opt:
            line["config"]["library_options"] = list(args.opt)
        if e2e_ms is not None:
            line["e2e"] = {"value": reads_total / (e2e_max * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": h2d,
                           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_max}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n = args.cpu_sample or 4 * cores
            idx = list(range(min(n, b.n)))
            dt, ccells, cops = cpu_realign_sample(b, idx, args.band, cores)
            for k, i in enumerate(idx):                          # the checker role: same CIGARs as the GPU path
                assert np.array_equal(cops[k], ops[off[i]:off[i + 1]]), "GPU CIGAR of read %d differs from the CPU oracle" % i
            line["cpu_baseline"] = {"value": len(idx) / dt, "unit": "reads/s", "cores": cores, "kind": "port",
                                    "sample": "first %d reads of the step's batch, one read per task on %d threads (%.1f s); "
                                              "CIGARs compared with the GPU output" % (len(idx), cores, dt),
                                    "cells_per_s": ccells / dt}
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
