"""GPU-owner service (ciri-long_b200/server.py): transport and routing with forked pool workers.  The CPU tests use
the service's "echo" backend (no device, no alignment): requests of many forked workers share owner batches and
every answer comes back to the worker that asked, in order.  The GPU test runs the product backend."""
import multiprocessing as mp
import os

import numpy as np
import pytest

_SVC = None


def _init(svc):
    global _SVC
    _SVC = svc
    svc.client()                                  # claim a slot in this forked worker


def _work(seed):
    rng = np.random.default_rng(seed)
    c = _SVC.client()
    out = []
    for rep in range(4):
        n = int(rng.integers(1, 300))
        qs = ["ACGT"[0] * int(rng.integers(1, 400)) for _ in range(n)]
        rs = ["C" * int(rng.integers(1, 90)) for _ in range(n)]
        params = (10, 4, 8, 2) if rep % 2 else (1, 1, 1, 1)
        rec, cig = c.align_pairs(rs, qs, *params, need_cigar=bool(rep % 2))
        ok = len(rec) == n and (rec["score1"] == np.array([len(q) * 1000 + len(r) for q, r in zip(qs, rs)])).all()
        if rep % 2:
            ok = ok and len(cig) == n and (cig[rec["cigar_off"]] >> 4 == np.array([len(q) for q in qs])).all()
        out.append(bool(ok))
    return os.getpid(), all(out)


def test_service_routes_answers_to_forked_workers():
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import server
    with server.AlignService(devices=(0, 1), n_clients=6, arena_mb=1, flush_ms=5.0, backend="echo") as svc:
        with mp.get_context("fork").Pool(6, initializer=_init, initargs=(svc,)) as pool:
            res = pool.map(_work, range(24))
        assert all(ok for _, ok in res)
        assert len({pid for pid, _ in res}) > 1

def test_more_clients_than_slots_is_an_error():
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import server
    with server.AlignService(devices=(0,), n_clients=1, arena_mb=1, backend="echo") as svc:
        svc.client()
        svc._client = None
        with pytest.raises(RuntimeError):
            svc.client()


def test_large_request_is_split_across_arena_sized_round_trips():
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import server
    with server.AlignService(devices=(0,), n_clients=1, arena_mb=1, backend="echo") as svc:
        qs = ["A" * 60] * 30000                       # 30000 * (8 + 60 + 50) bytes > 1 MiB
        rs = ["C" * 50] * 30000
        rec, cig = svc.client().align_pairs(rs, qs, 10, 4, 8, 2, need_cigar=True)
        assert len(rec) == 30000 and (rec["score1"] == 60050).all()
        assert len(cig) == 30000 and (np.diff(rec["cigar_off"]) == 1).all()


def _gpu_work(seed):
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import ssw_wrap as sw
    rng = np.random.default_rng(seed)
    out = []
    L = "ACGT"
    for rep in range(3):
        params = [(1, 1, 1, 1), (10, 4, 8, 2), (2, 2, 3, 1)][rep]
        refs, qs = [], []
        for k in range(int(rng.integers(200, 700))):
            n = int(rng.integers(20, 1500))
            r = "".join(L[i] for i in rng.integers(0, 4, n))
            a = int(rng.integers(0, max(1, n - 20)))
            q = list(r[a:a + int(rng.integers(15, 400))])
            for j in range(len(q)):
                if rng.random() < 0.08:
                    q[j] = L[int(rng.integers(0, 4))]
            refs.append(r); qs.append("".join(q))
        res = sw.align_pairs(refs, qs, *params, report_cigar=True)                       # through the service (attach())
        one = sw.Aligner(refs[0], *params, report_cigar=True).align(qs[0])                # the per-call site, too
        out.append((params, refs, qs, [(x.score, x.ref_begin, x.ref_end, x.query_begin, x.query_end, x.cigar_string) for x in res],
                    (one.score, one.ref_begin, one.cigar_string)))
    return out


@pytest.mark.gpu
def test_forked_workers_through_the_service_match_the_local_library():
    """8 forked workers, a few thousand mixed pairs each, three scoring schemes interleaved: identical to the
    in-process batched path, pair by pair (coordinates and CIGAR strings)."""
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import server, ssw_wrap as sw
    if sw.Aligner.libssw.ssw_cuda_device_count() <= 0:
        pytest.skip("no CUDA device")
    with server.AlignService(devices=(0,), n_clients=8, arena_mb=32, flush_ms=3.0) as svc:
        with mp.get_context("fork").Pool(8, initializer=svc.attach) as pool:
            res = pool.map(_gpu_work, range(8))
        stats = None
    n = 0
    for worker in res:
        for params, refs, qs, got, one in worker:
            exp = sw.align_pairs(refs, qs, *params, report_cigar=True)
            assert [(x.score, x.ref_begin, x.ref_end, x.query_begin, x.query_end, x.cigar_string) for x in exp] == got
            assert one == (exp[0].score, exp[0].ref_begin, exp[0].cigar_string)
            n += len(refs)
    assert n > 8000
