"""Throughput of the GPU-owner service (ciri-long_b200/server.py) fed by forked pool workers, next to the
reference's way of doing the same work (a new Aligner per pair inside Pool workers, oracle/ref_wrap.py over the
unmodified libssw.so).  argv: workers, pairs per request, requests per worker.
Shape: find_bsj-like pairs (300-800 nt vs 2 kb, 1/1/1/1) as Python strings, which is what the pipeline holds."""
import json, multiprocessing as mp, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

_SVC = None
_DATA = None


def _init(svc, data):
    global _SVC, _DATA
    _SVC, _DATA = svc, data
    if svc is not None:
        svc.attach()


def _gpu_job(k):
    from ciri_long_b200 import ssw_wrap as sw
    qs, rs = _DATA
    rec, _ = sw.align_pairs(rs, qs, 1, 1, 1, 1, need_cigar=False, as_records=True)
    return int(rec["score1"].sum())


def _gpu_job_single(k):
    from ciri_long_b200 import ssw_wrap as sw
    qs, rs = _DATA
    tot = 0
    for q, r in zip(qs[:64], rs[:64]):
        tot += sw.Aligner(r, 1, 1, 1, 1).align(q).score          # the unmodified per-call site, blocking on the service
    return tot


def _cpu_job(k):
    from oracle.ref_wrap import RefAligner
    qs, rs = _DATA
    tot = 0
    for q, r in zip(qs, rs):
        tot += RefAligner(r, 1, 1, 1, 1).align(q).score
    return tot


def main():
    workers = int(sys.argv[1]) if len(sys.argv) > 1 else (os.cpu_count() or 8)
    per_req = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
    reqs = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import server, workloads as W
    b = W.bsj_refinement_pairs(per_req, seed=77)
    lut = np.frombuffer(b"ACGTN", dtype=np.uint8)
    data = ([lut[b.query(i)].tobytes().decode() for i in range(per_req)], [lut[b.ref(i)].tobytes().decode() for i in range(per_req)])
    cells = b.cells
    out = dict(workers=workers, pairs_per_request=per_req, requests_per_worker=reqs)
    with server.AlignService(devices=(0,), n_clients=workers, arena_mb=64, flush_ms=2.0) as svc:
        with mp.get_context("fork").Pool(workers, initializer=_init, initargs=(svc, data)) as pool:
            pool.map(_gpu_job, range(workers))                                     # warm
            t0 = time.perf_counter()
            sums = pool.map(_gpu_job, range(workers * reqs), chunksize=1)
            dt = time.perf_counter() - t0
            out["service_batched"] = dict(pairs_per_s=workers * reqs * per_req / dt, gcups=workers * reqs * cells / dt / 1e9, seconds=dt)
            t0 = time.perf_counter()
            s1 = pool.map(_gpu_job_single, range(workers * 2), chunksize=1)
            dt1 = time.perf_counter() - t0
            out["service_per_call"] = dict(pairs_per_s=workers * 2 * 64 / dt1, seconds=dt1,
                                           note="unmodified Aligner(ref).align(query) sites blocking on the service: concurrency = number of workers")
    from oracle import oracle as O
    if O.RefLib.available():
        with mp.get_context("fork").Pool(workers, initializer=_init, initargs=(None, data)) as pool:
            t0 = time.perf_counter()
            cs = pool.map(_cpu_job, range(workers), chunksize=1)
            dtc = time.perf_counter() - t0
        out["reference_pool"] = dict(pairs_per_s=workers * per_req / dtc, gcups=workers * cells / dtc / 1e9, seconds=dtc,
                                     how="new Aligner per pair on strings inside Pool(%d) workers (oracle/ref_wrap.py + unmodified libssw.so)" % workers)
        out["same_scores"] = bool(cs[0] == sums[0])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
