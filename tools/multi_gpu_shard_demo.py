"""One batch split over all visible GPUs with sharding.py (cost-balanced shards, one thread per device,
host-side gather by original index) and checked against the single-GPU result.  No NCCL involved."""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ciri_long_b200
from ciri_long_b200 import sharding, ssw_wrap as sw, workloads as W
n_dev = sw.Aligner.libssw.ssw_cuda_device_count()
b = W.bsj_refinement_pairs(int(sys.argv[1]) if len(sys.argv) > 1 else 65536, seed=8)
ref_rec, ref_cig = sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1, device=0)
parts = [None] * n_dev
def work(rank):
    idx, sub = sharding.shard_batch(b, rank, n_dev)
    rec, cig = sw.align_arrays(sub["seqs"], sub["q_off"], sub["q_len"], sub["r_off"], sub["r_len"], 1, 1, 1, 1, device=rank)
    parts[rank] = (idx, rec, cig)
for rep in range(2):                                    # first pass warms the per-device memory pools
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(r,)) for r in range(n_dev)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
rec, cig = sharding.gather_results(len(b), parts)
ok = all((rec[k] == ref_rec[k]).all() for k in ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len"))
for i in range(0, len(b), 97):
    ok = ok and (cig[rec["cigar_off"][i]:rec["cigar_off"][i] + rec["cigar_len"][i]] == ref_cig[ref_rec["cigar_off"][i]:ref_rec["cigar_off"][i] + ref_rec["cigar_len"][i]]).all()
print("devices", n_dev, "pairs", len(b), "sharded e2e %.1f ms" % (dt * 1e3), "%.0f GCUPS" % (b.cells / dt / 1e9), "identical to single GPU:", bool(ok))
