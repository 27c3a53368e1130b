#!/usr/bin/env python
"""bench.py -- batched SSW GCUPS (score + coords + CIGAR) on B200 vs host libssw.

    python bench.py [--config C2|C3|C4|S1|S2|C5] [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the whole hot path (forward score pass, deciding byte-flavour pass, reverse pass,
CIGAR pass) over one batch of synthetic pairs.  GCUPS = sum(len(query) * len(ref)) / seconds / 1e9 --
forward-matrix cells only; the reverse pass, the 16-bit re-run and the banded CIGAR DP are not counted as extra
cells, for the GPU and the CPU alike.

Configurations (BASELINE.json `configs`; the default, C2, is the headline the metric is quoted on):
  C2  1,048,576 BSJ-refinement pairs, 300-800 nt consensus segment vs 2 kb flank, find_bsj scoring 1/1/1/1
  C3  ~800k segment-vs-segment pairs of 200k NanoSim-style rolling-circle reads, collapse scoring 10/4/8/2
  C4  length sweep 64..4096 (m ~ n), score-only (flag=0) and full (flag=1), 1/1/1/1 and 10/4/8/2: one JSON line
      with a `sweep` table
  S1  find_bsj.py:191-215 as it really is: 20-600 nt clipped ends vs 400-600 kb genomic windows, 1/1/1/1
  S2  collapse.py:165-172: 4 M junction pairs (40-60 nt vs 20 nt), 10/4/8/2
  C5  ONE 10 M-pair mixed batch (60 % S2, 25 % C2, 10 % C3, 5 % S4/S6 shapes) in 8 slabs; under torchrun the slabs
      are dealt to the ranks (strong scaling: the same 10 M pairs for every N, no collective on the data path);
      in one process with --gpus N the product's own ssw_align_batch_multi spreads them over N devices

  value     inputs resident in HBM, K x ssw_batch_run timed with CUDA events on the launching stream
  e2e       the public one-shot call (ssw_align_batch) from pinned HOST buffers, H2D and D2H inside the timing
  roofline  forward score-pass kernels: achieved GCUPS vs the DPX peak (VIADDMNMX.S16x2 issue rate measured
            live by ssw_cuda_dpx_peak / 3 lane-instructions per cell); whole-step and per-stage
            fractions beside it
  cpu_baseline / --impl reference: the unmodified reference libssw.so (oracle/_ref, built from /root/reference by
            oracle/Makefile) in a multiprocessing pool over all host cores, chunks of 250 pairs like
            find_bsj.py:338-345, on a bounded sample of the SAME pairs; every result of that sample is compared
            with the GPU's (the parity check).  Two figures: the C library on pre-encoded arrays, and the
            per-call Python wrapper path of ssw_wrap.py:174-252 (oracle/ref_wrap.py) the way CIRI-long drives it.
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
sys.path.insert(0, ROOT)

METRIC = "batched SSW GCUPS (score+coords+CIGAR)"

CONFIGS = {
    "C2": dict(pairs=1 << 20, params=(1, 1, 1, 1), cpu_sample=32768, wrap_sample=8192,
               workload="C2 BSJ-refinement pairs: 300-800 nt consensus segment (5/4/4 % sub/ins/del, 1 % N) vs 2 kb genomic "
                        "flank, find_bsj params 1/1/1/1, flag=1 (score+coords+CIGAR)",
               l2="inputs (2.7 GB per GPU) are larger than L2"),
    "C3": dict(pairs=200000, params=(10, 4, 8, 2), cpu_sample=16384, wrap_sample=4096,
               workload="C3 rolling-circle reads (200k reads of 2-6 kb, 2-8 tandem copies, 5/4/4 % noise): segment k vs "
                        "segment 0 of every read (~800k pairs of 250-3000 nt), collapse params 10/4/8/2, flag=1",
               l2="inputs (0.8 GB per GPU) are larger than L2"),
    "S1": dict(pairs=4096, params=(1, 1, 1, 1), cpu_sample=512, wrap_sample=64,
               workload="S1 find_bsj clip refinement: 20-600 nt clipped ends vs 400-600 kb genomic windows (views into one "
                        "64 Mb synthetic genome), params 1/1/1/1, flag=1",
               l2="every pair streams a 0.4-0.6 MB window; the windows of a batch (64 MB genome) fit L2, the direction and "
                  "column-record scratch does not"),
    "S2": dict(pairs=1 << 22, params=(10, 4, 8, 2), cpu_sample=262144, wrap_sample=65536,
               workload="S2 junction pairs (collapse.curate_junction): 40-60 nt junction consensus vs 20 nt genomic "
                        "junction, params 10/4/8/2, flag=1",
               l2="inputs (0.3 GB) are larger than L2"),
    "C5": dict(pairs=10000000, params=(10, 4, 8, 2), cpu_sample=32768, wrap_sample=8192,
               workload="C5 one 10,000,000-pair mixed batch in 8 slabs: 60 % S2-like (40-60 vs 20 nt), 25 % C2-like (300-800 nt "
                        "vs 2 kb), 10 % C3-like (250-3000 nt segments), 5 % S4/S6-like (0.2-5 kb vs 50 nt), shuffled; params "
                        "10/4/8/2, flag=1",
               l2="inputs (about 9 GB) are larger than L2"),
    "C4": dict(pairs=0, params=None, cpu_sample=0, wrap_sample=0,
               workload="C4 length sweep: m ~ n in {64,128,256,512,1024,2048,4096} (5/4/4 % noise), params 1/1/1/1 and "
                        "10/4/8/2, score-only (flag=0) and full (flag=1); value = the full-mode rows together",
               l2="each row's inputs are larger than L2 or the row is flushed by the next row's inputs"),
}
C4_LENGTHS = (64, 128, 256, 512, 1024, 2048, 4096)
C5_SLABS = 8


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference libssw.so (or the oracle port) over all host cores.  Workers are spawned, not
# forked (the parent may hold a CUDA context); the sample travels to them once, through the initializer.
_W = {}


def _cpu_worker_init(kind, arrays, params, strings):
    from oracle import oracle as O
    _W["lib"] = O.RefLib() if kind == "reference" else O.Oracle()
    _W["mat"] = O.make_mat(params[0], params[1])
    _W["arrays"], _W["params"], _W["strings"] = arrays, params, strings


def _cpu_worker(job):
    lo, hi, mode, flag = job
    seqs, q_off, q_len, r_off, r_len = _W["arrays"]
    p = _W["params"]
    out = []
    if mode == "c":
        lib, mat = _W["lib"], _W["mat"]
        for i in range(lo, hi):
            r = lib.align(seqs[q_off[i]:q_off[i] + q_len[i]], seqs[r_off[i]:r_off[i] + r_len[i]], mat, p[2], p[3], flag=flag)
            out.append(None if r is None else (r["score"], r["score2"], r["ref_begin"], r["ref_end"], r["read_begin"],
                                               r["read_end"], r["ref_end2"], tuple(r["cigar"])))
    else:
        # the way CIRI-long drives it: a new Aligner per pair, strings in, a result object out (find_bsj.py:204-205)
        from oracle.ref_wrap import RefAligner
        qs, rs = _W["strings"]
        for i in range(lo, hi):
            a = RefAligner(rs[i], p[0], p[1], p[2], p[3]).align(qs[i])
            out.append(None if a is None else (a.score, a.ref_begin, a.ref_end, a.query_begin, a.query_end))
    return out


class CpuArm(object):
    """A pool over all host cores holding one sample of pairs."""

    def __init__(self, sample, need_strings=0):
        import multiprocessing as mp
        from oracle import oracle as O
        O.build(ref=True)
        self.kind = "reference" if O.RefLib.available() else "port"
        self.cores = os.cpu_count() or 1
        self.sample = sample
        self.params = (sample.match, sample.mismatch, sample.gap_open, sample.gap_extend)
        strings = None
        if need_strings and self.kind == "reference":
            lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
            k = min(need_strings, len(sample))
            strings = ([lut[sample.query(i)].tobytes().decode() for i in range(k)], [lut[sample.ref(i)].tobytes().decode() for i in range(k)])
        self.n_strings = 0 if strings is None else len(strings[0])
        arrays = (sample.seqs, sample.q_off, sample.q_len, sample.r_off, sample.r_len)
        self.pool = mp.get_context("spawn").Pool(self.cores, initializer=_cpu_worker_init, initargs=(self.kind, arrays, self.params, strings))
        self.pool.map(_cpu_worker, [(0, min(4, len(sample)), "c", 1)] * self.cores)       # start and warm every worker

    def run(self, mode="c", flag=1, n=None):
        """-> (seconds, results) for the first n pairs of the sample, chunks of 250 like find_bsj.py:338-345"""
        n = len(self.sample) if n is None else min(n, len(self.sample))
        jobs = [(i, min(i + 250, n), mode, flag) for i in range(0, n, 250)]
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, jobs)
        dt = time.perf_counter() - t0
        return dt, [x for c in res for x in c]

    def cells(self, n=None):
        n = len(self.sample) if n is None else min(n, len(self.sample))
        return int((self.sample.q_len[:n].astype(np.int64) * self.sample.r_len[:n].astype(np.int64)).sum())

    def close(self):
        self.pool.close()
        self.pool.join()


def head_sample(batch, n):
    """the first n pairs of a batch as their own compact batch (the CPU sample; same pairs as the GPU's)"""
    from ciri_long_b200 import workloads as W
    n = min(n, len(batch))
    sub = W.PairBatch(batch.seqs, batch.q_off[:n], batch.q_len[:n], batch.r_off[:n], batch.r_len[:n], batch.match,
                      batch.mismatch, batch.gap_open, batch.gap_extend, batch.name)
    lo = int(min(sub.q_off.min(), sub.r_off.min()))
    hi = int(max((sub.q_off + sub.q_len).max(), (sub.r_off + sub.r_len).max()))
    return W.PairBatch(np.ascontiguousarray(batch.seqs[lo:hi]), sub.q_off - lo, sub.q_len.copy(), sub.r_off - lo, sub.r_len.copy(),
                       batch.match, batch.mismatch, batch.gap_open, batch.gap_extend, batch.name)


def parity_check(rec, cig, cpu_results, what):
    """every pair of the CPU sample against the GPU records (same pairs, same order): all seven fields and the CIGAR"""
    bad, escapes = [], 0
    for i, e in enumerate(cpu_results):
        r = rec[i]
        st = int(r["status"]) & 0xff
        got = (int(r["score1"]), int(r["score2"]), int(r["ref_begin1"]), int(r["ref_end1"]), int(r["read_begin1"]),
               int(r["read_end1"]), int(r["ref_end2"]))
        if st == 1 and (e is None or got == e[:7]):
            escapes += 1            # traceback left the band: the reference's CIGAR is undefined there, coordinates compared
            continue
        if e is None or st != 0 or got != e[:7] or tuple(cig[r["cigar_off"]:r["cigar_off"] + r["cigar_len"]].tolist()) != e[7]:
            bad.append((i, st, got, None if e is None else e[:7]))
    if bad:
        raise SystemExit("bench.py: parity check failed (%s): %d of %d pairs differ, first %r" % (what, len(bad), len(cpu_results), bad[:3]))
    return len(cpu_results), escapes


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) >= 7 and r[3 + k] == "Active" for r in self.rows)]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


# ------------------------------------------------------------------------------------------------
def c5_slabs_of(rank, world):
    """which of the C5_SLABS slabs of the one C5 batch a rank owns: round robin, so every slab has exactly one owner
    for any world size, and the statistically identical slabs make that the cost-balanced assignment"""
    return [s for s in range(C5_SLABS) if s % world == rank]


def make_batches(cfg_name, args, dev, rank, world):
    """the batches this rank owns (host numpy PairBatches): one for most configurations, its slabs for C5"""
    from ciri_long_b200 import workloads as W
    cfg = CONFIGS[cfg_name]
    n = args.pairs if args.pairs else cfg["pairs"]
    seed = W.SEED_BASE + 1000 * rank
    p = cfg["params"]
    if cfg_name == "C2":
        return [W.bsj_refinement_pairs_torch(n, dev, seed=seed + 2, params=p)]
    if cfg_name == "C3":
        return [W.repack(W.rolling_circle_pairs_torch(n, dev, seed=seed + 3, params=p))]
    if cfg_name == "S1":
        return [W.clip_window_pairs_torch(n, dev, seed=seed + 8, params=p)]
    if cfg_name == "S2":
        return [W.junction_pairs_torch(n, dev, seed=seed + 5, params=p)]
    if cfg_name == "C5":
        per = n // C5_SLABS
        # the same 8 slabs whatever the number of ranks; statistically identical, so dealing them round-robin is
        # the cost-balanced (LPT) assignment
        slabs = [W.mixed_slab_torch(per, dev, seed=W.SEED_BASE + 50 + 10 * s, params=p) for s in c5_slabs_of(rank, world)]
        return [W.concat_batches(slabs, name="C5-mixed")]          # this rank's share as ONE batch (one set of device scratch)
    raise SystemExit("unknown config " + cfg_name)


def cpu_sample_numpy(cfg_name, n):
    """the reference arm's sample: the configuration's recipe at sample size (numpy stream; no GPU needed)"""
    from ciri_long_b200 import workloads as W
    p = CONFIGS[cfg_name]["params"]
    if cfg_name == "C2":
        return W.bsj_refinement_pairs(n, seed=W.SEED_BASE + 2, params=p)
    if cfg_name == "C3":
        return W.repack(W.rolling_circle_pairs(max(8, n // 4), seed=W.SEED_BASE + 3, params=p))
    if cfg_name == "S2":
        return W.junction_pairs(n, seed=W.SEED_BASE + 5, params=p)
    import torch
    if cfg_name == "S1":
        return W.clip_window_pairs_torch(n, torch.device("cpu"), seed=W.SEED_BASE + 8, params=p)
    if cfg_name == "C5":
        return W.mixed_slab_torch(n, torch.device("cpu"), seed=W.SEED_BASE + 50, params=p)
    raise SystemExit("unknown config " + cfg_name)


def base_line(cfg_name, args, world, impl):
    cfg = CONFIGS[cfg_name]
    n = args.pairs if args.pairs else cfg["pairs"]
    strong = cfg_name == "C5"
    return {
        "metric": METRIC, "value": None, "unit": "GCUPS", "n_gpus": max(world, args.gpus if impl == "ours" and world == 1 else world),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": None, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "s16x2 (DPX) score passes and CIGAR fill, int32 for pairs near the 16-bit range", "data": "synthetic",
        "config": {"workload": cfg["workload"], "id": cfg_name,
                   ("pairs_total" if strong else "pairs_per_gpu"): n, "l2": cfg["l2"]},
    }


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    cfg_name = args.config
    if cfg_name == "C4":
        return run_reference_c4(args)
    cfg = CONFIGS[cfg_name]
    sample = cpu_sample_numpy(cfg_name, args.cpu_sample or cfg["cpu_sample"])
    arm = CpuArm(sample)
    times = []
    for s in range(args.warmup + args.steps):
        dt, _ = arm.run("c", 1)
        if s >= args.warmup:
            times.append(dt)
    arm.close()
    ms = 1e3 * float(np.mean(times))
    val = arm.cells() / (ms * 1e-3) / 1e9
    out = base_line(cfg_name, args, world, "reference")
    out.update({"impl": "reference", "value": val, "ms_per_step": ms, "dtype": "u8/s16 (SSE2)",
                "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": arm.cores, "kind": arm.kind,
                                 "sample": "%d pairs of the workload per step (same recipe, numpy stream), unmodified reference libssw.so "
                                           "via ctypes on pre-encoded int8 arrays, Pool(%d, spawn) x chunks of 250, flag=1" % (len(sample), arm.cores)},
                "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(out)


def c4_rows(args):
    rows = []
    for params in ((1, 1, 1, 1), (10, 4, 8, 2)):
        for L in C4_LENGTHS:
            n = int(min(1 << 20, max(2048, 3e10 / (L * L))))
            if args.pairs:
                n = min(n, args.pairs)
            rows.append((params, L, n, int(min(n, 32768, max(256, 1.5e10 / (L * L))))))
    return rows


def run_reference_c4(args):
    from ciri_long_b200 import workloads as W
    cells = secs = 0.0
    sweep = []
    cores = kind = None
    for params, L, n, ns in c4_rows(args):
        arm = CpuArm(W.square_pairs(ns, L, params=params))
        cores, kind = arm.cores, arm.kind
        row = {"length": L, "params": "%d/%d/%d/%d" % params, "pairs": ns}
        for flag in (0, 1):
            best = min(arm.run("c", flag)[0] for _ in range(max(1, args.steps)))
            row["score_only" if flag == 0 else "full"] = arm.cells() / best / 1e9
            if flag == 1:
                cells += arm.cells(); secs += best
        arm.close()
        sweep.append(row)
    val = cells / secs / 1e9
    out = base_line("C4", args, 1, "reference")
    out.update({"impl": "reference", "value": val, "ms_per_step": secs * 1e3, "dtype": "u8/s16 (SSE2)", "sweep": sweep,
                "cpu_baseline": {"value": val, "unit": "GCUPS", "cores": cores, "kind": kind,
                                 "sample": "per row 256-32768 pairs of the row's recipe, reference libssw.so via ctypes, Pool(%d) x chunks of 250" % cores},
                "e2e": {"value": val, "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    emit(out)


# ------------------------------------------------------------------------------------------------
def pin_batch(batch):
    import torch
    keep = []
    for k in ("seqs", "q_off", "q_len", "r_off", "r_len"):
        t = torch.from_numpy(np.ascontiguousarray(getattr(batch, k))).pin_memory()
        keep.append(t)
        setattr(batch, k, t.numpy())
    batch._pinned = keep
    return batch


def run_ours(args, rank, world, local_rank):
    import torch
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import ssw_wrap as sw

    if not torch.cuda.is_available() or sw.Aligner.libssw.ssw_cuda_device_count() <= 0:
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU implementation")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.config == "C4":
        if rank == 0:
            run_ours_c4(args, sw, torch, dev, local_rank)
        if world > 1:
            dist.destroy_process_group()
        return

    cfg_name = args.config
    cfg = CONFIGS[cfg_name]
    params = cfg["params"]
    batches = [pin_batch(b) for b in make_batches(cfg_name, args, dev, rank, world)]
    cells = sum(b.cells for b in batches)
    n_pairs = sum(len(b) for b in batches)
    in_bytes = sum(int(b.seqs.nbytes) for b in batches)
    multi = list(range(args.gpus)) if (world == 1 and args.gpus > 1) else None      # one process, several devices

    # ---- CPU arm on the first pairs of this rank's first batch (rank 0, single-GPU run only)
    cpu = cpu_res = wrap = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample = head_sample(batches[0], args.cpu_sample or cfg["cpu_sample"])
        arm = CpuArm(sample, need_strings=cfg["wrap_sample"])
        dt, cpu_res = arm.run("c", 1)
        cpu = dict(value=arm.cells() / dt / 1e9, unit="GCUPS", cores=arm.cores, kind=arm.kind,
                   sample="the first %d pairs of the GPU's own batch, %s via ctypes on pre-encoded int8 arrays, Pool(%d, spawn) x chunks of "
                          "250, flag=1 (score+coords+CIGAR), %.1f s; every result compared with the GPU's"
                          % (len(sample), "unmodified reference libssw.so (oracle/_ref)" if arm.kind == "reference" else "oracle port", arm.cores, dt))
        if arm.n_strings:
            dtw, _ = arm.run("wrapper", 1, n=arm.n_strings)
            wrap = dict(value=arm.cells(arm.n_strings) / dtw / 1e9, unit="GCUPS", cores=arm.cores, pairs=arm.n_strings, seconds=dtw,
                        how="a new Aligner(ref).align(query) per pair on Python strings inside the pool workers: the per-call path of "
                            "ssw_wrap.py:102-252 (per-base encode loop included) restated in oracle/ref_wrap.py over the unmodified libssw.so")
            cpu["wrapper_driven"] = wrap
        arm.close()

    stream = torch.cuda.Stream(device=dev)       # the library enqueues on this stream; events are recorded on it
    torch.cuda.set_stream(stream)

    # ---- device-resident throughput
    ds = [sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, *params, flag=1, device=local_rank, stream=stream.cuda_stream)
          for b in batches]
    for _ in range(args.warmup):
        for d in ds:
            d.run()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        for d in ds:
            d.run()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop()
    stage = np.sum([d.stage_ms() for d in ds], axis=0)
    launches = sum(d.launch_count() for d in ds)
    rec0, cig0 = ds[0].fetch()
    n_bad_status = int(((rec0["status"] & 0xff) > 1).sum())
    out_bytes = int(rec0.nbytes + 4 * len(cig0)) * len(ds)
    for d in ds:
        d.close()

    parity = None
    if cpu_res is not None:
        n_chk, esc = parity_check(rec0, cig0, cpu_res, cfg_name + " resident batch")
        parity = "%d pairs (the whole CPU sample) bit-exact vs %s: all seven fields and every CIGAR op; %d band-escape pairs " \
                 "(coordinates compared); %d pairs with a refused status in the batch" % (n_chk, cpu["kind"], esc, n_bad_status)

    # ---- end to end through the public one-shot C call: pinned host buffers in, pinned host results out; upload,
    # all kernels and download of every chunk inside the timed region
    outs = []
    for b in batches:
        o = torch.empty(len(b) * sw.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
        c = torch.empty(int(len(cig0) * 1.1 * len(b) / max(1, len(batches[0]))) + 65536, dtype=torch.int32).pin_memory()
        outs.append((o, c, o.numpy().view(sw.RESULT_DTYPE), c.numpy().view(np.uint32)))

    def e2e_step():
        res = []
        for b, (_, _, o, c) in zip(batches, outs):
            res.append(sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, *params, flag=1, device=local_rank, out=o, cig=c,
                                       devices=multi))
        return res
    r2 = e2e_step()
    for k in ("score1", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "score2", "ref_end2", "cigar_len"):
        if not (r2[0][0][k] == rec0[k]).all():
            raise SystemExit("bench.py: e2e call disagrees with the resident batch on %s" % k)
    if cpu_res is not None:
        parity_check(r2[0][0], r2[0][1], cpu_res, cfg_name + " one-shot call")
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        r2 = e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    h2d = int(in_bytes + 28 * n_pairs)
    d2h = int(sum(r.nbytes + 4 * len(c) for r, c in r2))

    # ---- the same call on 4-bit packed input (two bases per byte): half the upload
    e2e_packed = None
    if not multi:
        packs = []
        for b in batches:
            t = torch.from_numpy(sw.pack_dna4(b.seqs)).pin_memory()
            packs.append((t, t.numpy(), len(b.seqs)))

        def e2e_packed_step():
            return [sw.align_arrays(pk, b.q_off, b.q_len, b.r_off, b.r_len, *params, flag=1, device=local_rank, out=o, cig=c, packed_bases=nb)
                    for b, (_, _, o, c), (_, pk, nb) in zip(batches, outs, packs)]
        r3 = e2e_packed_step()
        for k in ("score1", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "score2", "ref_end2", "cigar_len"):
            if not (r3[0][0][k] == rec0[k]).all():
                raise SystemExit("bench.py: packed-input call disagrees with the resident batch on %s" % k)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_packed_step()
        barrier()
        e2e_packed = (time.perf_counter() - t0) / args.e2e_steps
        del packs

    peak_lane, _ = sw.dpx_peak(local_rank)

    # ---- max over ranks, whole-job aggregate
    ms_step = ms_total / args.steps
    vals = torch.tensor([ms_step, e2e_s, float(cells), float(stage[0]), float(launches), float(h2d), float(d2h), float(n_pairs), float(e2e_packed or 0.0)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = vals.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_step, e2e_s = float(mx[0]), float(mx[1])
        e2e_packed = float(mx[8]) or None
        cells_all, launches_all, h2d, d2h, pairs_all = float(sm[2]), float(sm[4]), int(sm[5]), int(sm[6]), float(sm[7])
    else:
        cells_all, launches_all, pairs_all = float(cells), float(launches), float(n_pairs)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    traffic = tnote = None
    try:
        with open(os.path.join(ROOT, "profiles", "r2_traffic.json")) as f:
            tj = json.load(f)
        if cfg_name == "C2":
            traffic = tj["forward_dram_bytes_per_step"] * n_pairs / tj["pairs"]
            tnote = tj.get("note")
    except Exception:
        pass
    fwd_ms = float(stage[0])
    peak_gcups = peak_lane / 3.0 / 1e9
    gc = lambda ms: cells / (ms * 1e-3) / 1e9 if ms > 0 else None
    fwd_gcups = gc(fwd_ms)
    out = base_line(cfg_name, args, world, "ours")
    out.update({
        "value": cells_all / (ms_step * 1e-3) / 1e9, "ms_per_step": ms_step,
        "stage_ms": {"forward": fwd_ms, "deciding": float(stage[1]), "reverse": float(stage[2]), "cigar": float(stage[3])},
        "roofline": {"bound": "dpx", "achieved": fwd_gcups, "peak": peak_gcups, "unit": "GCUPS", "frac": fwd_gcups / peak_gcups,
                     "traffic": traffic, "traffic_note": tnote,
                     "kernel": "score_kernel<K,TRUNC|GOTOH,fwd> (forward score pass, all strip heights)",
                     "peak_source": "ssw_cuda_dpx_peak: %.3e VIADDMNMX.S16x2 lane-instr/s measured in this run / 3 per cell (SURVEY 8d)" % peak_lane,
                     "whole_step_frac": cells / (ms_total / args.steps * 1e-3) / 1e9 / peak_gcups,
                     "stage_frac": {"note": "this rank's forward cells / stage time / peak: what the whole step would reach if it ran at that stage's pace",
                                    "forward+deciding": gc(fwd_ms + float(stage[1])) / peak_gcups,
                                    "reverse_share_of_step": float(stage[2]) / float(stage.sum()), "cigar_share_of_step": float(stage[3]) / float(stage.sum())},
                     "hbm": {"achieved": float(in_bytes + 24 * n_pairs + out_bytes) / (ms_total / args.steps * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                             "note": "algorithmic bytes per step / step time; the path is DPX-bound, not HBM-bound"}},
        "e2e": {"value": cells_all / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
                "devices": "ssw_align_batch_multi over %d devices in one process" % len(multi) if multi else "one device per rank"},
        "gpu_launches": int(launches_all) * args.steps,
        "pairs_per_second": pairs_all / (ms_step * 1e-3),
        "clocks": clocks,
    })
    if e2e_packed:
        out["e2e_packed"] = {"value": cells_all / e2e_packed / 1e9, "unit": "GCUPS", "ms_per_step": e2e_packed * 1e3,
                             "h2d_bytes_per_step": int(h2d - (in_bytes - in_bytes // 2) * (world if cfg_name != "C5" else 1)) if world == 1 else None,
                             "note": "ssw_align_batch_multi_packed: two bases per byte on the host side (codes 0..4), expanded on the device"}
    if multi:
        # one process driving several devices: only the one-shot call spreads over them, so it is the figure
        out["value"], out["ms_per_step"] = out["e2e"]["value"], out["e2e"]["ms_per_step"]
        out["value_note"] = "single-process multi-device run: value = the one-shot call over all devices (host buffers); stage_ms / roofline are device 0 alone"
    if parity:
        out["parity"] = parity
    if cpu is not None:
        out["cpu_baseline"] = cpu
    emit(out)
    if world > 1:
        dist.destroy_process_group()


def run_ours_c4(args, sw, torch, dev, local_rank):
    from ciri_long_b200 import workloads as W
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    peak_lane, _ = sw.dpx_peak(local_rank)
    peak_gcups = peak_lane / 3.0 / 1e9
    sweep = []
    tot_cells = tot_ms = tot_fwd = 0.0
    launches = 0
    e2e_cells = e2e_s = 0.0
    h2d = d2h = 0
    cpu_cells = cpu_s = 0.0
    cores = kind = None
    checked = 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    for params, L, n, ns in c4_rows(args):
        b = pin_batch(W.square_pairs_torch(n, L, dev, params=params))
        row = {"length": L, "params": "%d/%d/%d/%d" % params, "pairs": n}
        cpu_res = {}
        if not args.no_cpu:
            arm = CpuArm(head_sample(b, ns))
            cores, kind = arm.cores, arm.kind
            for flag in (0, 1):
                dt, cpu_res[flag] = arm.run("c", flag)
                row["cpu_score_only" if flag == 0 else "cpu_full"] = arm.cells() / dt / 1e9
                if flag == 1:
                    cpu_cells += arm.cells(); cpu_s += dt
            arm.close()
        for flag in (0, 1):
            with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, *params, flag=flag, device=local_rank, stream=stream.cuda_stream) as d:
                for _ in range(max(1, args.warmup)):
                    d.run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(args.steps):
                    d.run()
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                st = d.stage_ms()
                rec, cig = d.fetch()
                row["score_only" if flag == 0 else "full"] = b.cells / (ms * 1e-3) / 1e9
                if flag == 1:
                    tot_cells += b.cells; tot_ms += ms; tot_fwd += float(st[0]); launches += d.launch_count() * args.steps
                    row["stage_ms"] = [float(x) for x in st]
                if flag in cpu_res:
                    if flag == 1:
                        checked += parity_check(rec, cig, cpu_res[1], "C4 L=%d %s full" % (L, row["params"]))[0]
                    else:
                        for i, e in enumerate(cpu_res[0]):
                            g = (int(rec[i]["score1"]), int(rec[i]["score2"]), int(rec[i]["ref_end1"]), int(rec[i]["read_end1"]), int(rec[i]["ref_end2"]))
                            if e is None or g != (e[0], e[1], e[3], e[5], e[6]):
                                raise SystemExit("bench.py: parity check failed (C4 L=%d score-only) on pair %d: %r vs %r" % (L, i, g, e))
                        checked += len(cpu_res[0])
        # end to end, full mode
        o = np.zeros(len(b), dtype=sw.RESULT_DTYPE)
        c = np.empty(int(len(cig) * 1.1) + 65536, dtype=np.uint32)
        sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, *params, flag=1, device=local_rank, out=o, cig=c)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(max(2, args.e2e_steps // 2)):
            r2, c2 = sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, *params, flag=1, device=local_rank, out=o, cig=c)
        dt = (time.perf_counter() - t0) / max(2, args.e2e_steps // 2)
        row["e2e_full"] = b.cells / dt / 1e9
        e2e_cells += b.cells; e2e_s += dt
        h2d += int(b.seqs.nbytes + 28 * len(b)); d2h += int(r2.nbytes + 4 * len(c2))
        sweep.append(row)
        del b
    clocks = sampler.stop()
    out = base_line("C4", args, 1, "ours")
    val = tot_cells / (tot_ms * 1e-3) / 1e9
    fwd = tot_cells / (tot_fwd * 1e-3) / 1e9
    out.update({"value": val, "ms_per_step": tot_ms, "sweep": sweep,
                "roofline": {"bound": "dpx", "achieved": fwd, "peak": peak_gcups, "unit": "GCUPS", "frac": fwd / peak_gcups, "traffic": None,
                             "kernel": "score_kernel / score32_kernel forward launches of the full-mode rows", "whole_step_frac": val / peak_gcups,
                             "peak_source": "ssw_cuda_dpx_peak: %.3e VIADDMNMX.S16x2 lane-instr/s measured in this run / 3 per cell" % peak_lane},
                "e2e": {"value": e2e_cells / e2e_s / 1e9, "unit": "GCUPS", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_s * 1e3, "steps": max(2, args.e2e_steps // 2)},
                "gpu_launches": int(launches), "clocks": clocks,
                "parity": "%d sample pairs over the rows bit-exact vs the CPU arm (full: seven fields + CIGAR; score-only: score, ends, second best)" % checked})
    if cpu_s > 0:
        out["cpu_baseline"] = {"value": cpu_cells / cpu_s / 1e9, "unit": "GCUPS", "cores": cores, "kind": kind,
                               "sample": "per row the first 256-32768 pairs of the GPU's own batch, full mode; per-row figures in `sweep`"}
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """the ONE JSON line of the contract, on the real stdout (everything else -- NCCL's version banner, library
    chatter -- goes to stderr, see main)"""
    line = json.dumps(obj) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(line); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line.encode())


def main():
    global _REAL_STDOUT
    # keep stdout for the JSON line alone: C libraries (NCCL prints its version there) write to fd 1 directly
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=sorted(CONFIGS))
    ap.add_argument("--pairs", type=int, default=0, help="override the configuration's pair count (per GPU; C5: total)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (and with it the parity check)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
