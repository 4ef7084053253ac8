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
    ap.add_argument("--no-files", action="store_true", help="skip the SAM-in / SAM-out leg (e2e_files)")
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


def thread_inst_per_cell():
    """Thread-instructions per DP cell of k_fb2 from the same committed capture (None if absent)."""
    try:
        return float(json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["thread_inst_per_cell"])
    except Exception:
        return None


ISSUE_PEAK = 148 * 4 * 32 * 1.965e9        # thread-instructions/s: 148 SMs x 4 schedulers x 32 lanes x 1965 MHz


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


def write_experiment(b, dirname):
    """The batch as the files a mapper leaves behind (reference FASTA, read FASTQ, SAM of one local hit per read): the
    input of AbstractMapper.realignSamFile.  Reverse-strand reads sit reverse-complemented in the FASTQ, as reads do."""
    from nanopore_b200 import synth
    from nanopore_b200.bioio import reverseComplement
    os.makedirs(dirname, exist_ok=True)
    fa, fq, sam = (os.path.join(dirname, n) for n in ("reference.fa", "reads.fq", "mapping.sam"))
    ref = synth.decode(b.ref)
    with open(fa, "w") as fh:
        fh.write(">ref synthetic\n")
        fh.write("\n".join(ref[i:i + 100] for i in range(0, len(ref), 100)) + "\n")
    letters = np.array(list("MID"))
    with open(fq, "w") as fqh, open(sam, "w") as sh:
        sh.write("@HD\tVN:1.0\tSO:unsorted\n@SQ\tSN:ref\tLN:%d\n" % len(ref))
        for i in range(b.n):
            seq = synth.decode(b.read(i))
            o = b.ops(i)
            code, ln = (o & 3).astype(np.int64), (o >> 2).astype(np.int64)
            lead = int(ln[0]) if code[0] == 2 else 0                     # chained-global guide: D(window start) local D(tail)
            a = 1 if code[0] == 2 else 0
            z = len(o) - 1 if code[-1] == 2 and len(o) > a + 1 else len(o)
            cig = "".join(np.char.add(ln[a:z].astype("U12"), letters[code[a:z]]).tolist())
            rev = bool(b.reverse[i]) if b.reverse is not None else False
            name = b.names[i] if b.names else "read_%d" % i
            fqh.write("@%s\n%s\n+\n%s\n" % (name, reverseComplement(seq) if rev else seq, "2" * len(seq)))
            sh.write("\t".join([name, "16" if rev else "0", "ref", str(int(b.ref_start[i]) + lead + 1), "30", cig, "*", "0", "0", seq, "*"]) + "\n")
    return fa, fq, sam


def realign_files(sam, fq, fa, out_sam, band):
    """What a `*Realign*` mapper does after mapping (reference nanopore/mappers/abstractMapper.py:25-39): chain, realign,
    rewrite the SAM -- through the plugin class and the in-process Target runner.  The reference hard-codes
    --diagonalExpansion=10 on its command line (utils.py:587); the metric is quoted at band 50, so the constant is set."""
    import shutil
    from nanopore_b200 import realign
    realign.REALIGN_DIAGONAL_EXPANSION = band
    from nanopore_b200.mappers.abstractMapper import AbstractMapper
    from nanopore_b200.target import Stack

    class Realign(AbstractMapper):
        def run(self):
            self.realignSamFile(gapGamma=0.5, matchGamma=0.0)

    shutil.copyfile(sam, out_sam)
    failed = Stack(Realign(fq, "2D", fa, out_sam)).startJobTree(None)
    if failed:
        raise RuntimeError("%d job(s) failed" % failed)


def per_read_process_overhead(b, n=16):
    """What the reference pays per read besides the DP (utils.py:582-587): the whole reference and the read written as
    FASTA, one `sh -c "echo cigar | ..."` process, the cigar file read back.  `cat` stands in for cactus_realign."""
    import subprocess
    import tempfile
    from nanopore_b200 import synth
    ref = synth.decode(b.ref)
    t0 = time.perf_counter()
    with tempfile.TemporaryDirectory() as d:
        for i in range(min(n, b.n)):
            with open(os.path.join(d, "ref.fa"), "w") as fh:
                fh.write(">ref\n%s\n" % ref)
            with open(os.path.join(d, "read.fa"), "w") as fh:
                fh.write(">read\n%s\n" % synth.decode(b.read(i)))
            cig = "cigar: read 0 %d + ref 0 %d + 1 M %d" % (len(b.read(i)), len(ref), len(b.read(i)))
            subprocess.check_call("echo '%s' | cat > %s" % (cig, os.path.join(d, "out.cig")), shell=True)
            open(os.path.join(d, "out.cig")).read()
    return 1e3 * (time.perf_counter() - t0) / max(1, min(n, b.n))


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
        launches_e2e = 0
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

    # ---------------- end to end on N GPUs: ONE batch on rank 0 through the rank-sharded realigner ----------------
    sharded = None
    if not args.no_e2e and world > 1:
        from nanopore_b200 import parallel
        from nanopore_b200.batch import Batch
        ctx.close()                                                  # the realigners of nanopore_b200.parallel own the GPUs now
        fields = ("reads", "read_off", "ref_start", "ref_end", "in_ops", "in_off")
        if rank == 0:
            parts = [[getattr(b, f) for f in fields]] + [parallel._unpack(parallel._recv_blob(r)) for r in range(1, world)]
            cat = lambda k: np.concatenate([p[k] for p in parts])
            offs = lambda k, d: np.concatenate([[0]] + [p[k][1:] + sum(int(q[k][-1]) for q in parts[:j]) for j, p in enumerate(parts)]).astype(np.int64)
            big = Batch(b.ref, cat(0), offs(1, 0), cat(2), cat(3), cat(4), offs(5, 0))
            sr = parallel.ShardedRealigner(None)
            sr.set_reference(b.ref)
            sr.realign(big, params)                                  # warm-up (contexts, allocations)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(max(1, args.steps)):
                sr._sent = None                                      # a new batch every step: the scatter is inside the timed region
                ops_s, off_s, _ = sr.realign(big, params)
            e1.record()
            torch.cuda.synchronize()
            sh_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0)) / max(1, args.steps)
            assert np.array_equal(ops_s[:len(ops)], ops) and np.array_equal(off_s[:b.n + 1], off)     # rank 0's own reads: same CIGARs
            sharded = {"value": big.n / (sh_ms * 1e-3), "unit": "reads/s", "ms_per_step": sh_ms, "reads_per_step": int(big.n),
                       "h2d_bytes_per_step": int(sum(getattr(big, f).nbytes for f in fields)), "d2h_bytes_per_step": int(ops_s.nbytes + off_s.nbytes),
                       "cells": int(sr.cells),
                       "path": "nanopore_b200.parallel.ShardedRealigner: one batch in rank 0's host memory, shards balanced on "
                               "estimated DP cells and sent point to point, CIGAR ops returned to rank 0 in input order"}
            sr.close()
            parallel.shutdown()
        else:
            parallel._send_blob(parallel._pack([getattr(b, f) for f in fields]), 0)
            parallel.worker_loop()

    # ---------------- end to end from files (1 GPU): SAM + FASTQ + FASTA in, realigned SAM out ----------------
    files = None
    if not args.no_e2e and world == 1 and not args.no_files:
        import shutil
        import tempfile
        ctx.close()
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        d = tempfile.mkdtemp(prefix="phmm_bench_", dir=base)
        try:
            fa, fq, sam = write_experiment(b, d)
            out_sam = os.path.join(d, "realigned.sam")
            realign_files(sam, fq, fa, out_sam, args.band)           # warm-up
            nrun = max(1, min(args.steps, 2))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(nrun):
                realign_files(sam, fq, fa, out_sam, args.band)
            torch.cuda.synchronize()
            f_ms = 1e3 * (time.perf_counter() - t0) / nrun
            # same CIGARs as the packed path (the chained file is sorted by name: compare by name)
            got = {}
            for ln in open(out_sam):
                if not ln.startswith("@"):
                    f = ln.split("\t", 6)
                    got[f[0]] = f[5]
            assert len(got) == b.n
            letters = np.array(list("MID"))
            for i in range(0, b.n, max(1, b.n // 64)):
                o = ops[off[i]:off[i + 1]]
                want = "".join(np.char.add((o >> 2).astype("U12"), letters[(o & 3).astype(np.int64)]).tolist())
                assert got[b.names[i]] == want, "CIGAR of %s from files differs from the packed path" % b.names[i]
            files = {"value": b.n / (f_ms * 1e-3), "unit": "reads/s", "ms_per_step": f_ms,
                     "bytes_in": int(sum(os.path.getsize(x) for x in (fa, fq, sam))), "bytes_out": int(os.path.getsize(out_sam)),
                     "path": "AbstractMapper.realignSamFile: native chain (libphmm_io.so) -> temp.sam -> native load + pack -> "
                             "phmm_realign_batch -> native SAM emit"}
        finally:
            shutil.rmtree(d, ignore_errors=True)

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
                             "frac_of_datasheet_8TBs": achieved / 8000.0,
                             # the bound the counters show (DRAM is ~5 % busy): instruction issue.  cells/s x thread-instructions
                             # per cell (committed ncu capture) / what 148 SMs can issue
                             "issue_frac": (cells_rank / (fb_ms / args.steps * 1e-3)) * thread_inst_per_cell() / ISSUE_PEAK
                             if thread_inst_per_cell() else None,
                             "thread_inst_per_cell": thread_inst_per_cell()},
                "clocks": clocks, "gpu_launches": launches,
                "cells_per_read": cells_total / reads_total, "pairs": st["pairs"], "regions": st["n_regions"],
                "resident_regions": st["n_slots"], "decode_ms": dec_ms / args.steps}
        if args.opt:
            line["config"]["library_options"] = list(args.opt)
        if e2e_ms is not None:
            line["e2e"] = {"value": reads_total / (e2e_max * 1e-3), "unit": "reads/s", "h2d_bytes_per_step": h2d,
                           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_max,
                           "path": "phmm_realign_batch from pinned host buffers on every rank (its own shard)"}
        if sharded is not None:
            # N GPUs: the product path.  e2e = one batch on rank 0 through the sharded realigner; the per-rank call is kept beside it
            line["e2e_per_rank_call"] = line.get("e2e")
            line["e2e"] = sharded
        if files is not None:
            line["e2e_files"] = files
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
            # the reference's own default: 4 workers (Makefile:1 maxThreads=4), and what it pays per read besides the DP
            i4 = list(range(min(8, b.n)))
            dt4, _, _ = cpu_realign_sample(b, i4, args.band, 4)
            line["cpu_baseline"]["four_workers"] = {"value": len(i4) / dt4, "unit": "reads/s", "cores": 4,
                                                    "sample": "%d reads on 4 threads (reference Makefile:1 maxThreads=4)" % len(i4)}
            line["cpu_baseline"]["per_read_process_overhead_ms"] = per_read_process_overhead(b)
        print(json.dumps(line), flush=True)
    ctx.close()                                                      # idempotent
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
