"""GPU-owner service for CIRI-long's ``-t`` worker pools (second half).

The reference runs its SSW calls inside forked ``multiprocessing.Pool(threads, env.initializer, ...)`` workers
(find_bsj.py:338-345, collapse.py:842-851, env.py:9-21; ``-t`` is main.py:231).  A CUDA context does not survive
``fork``, and a GPU is only efficient on batches, so the device is owned by ONE process per GPU -- spawned, never
forked, and the only place where libssw_cuda.so touches the device -- and the pool workers become clients:

    service = AlignService(devices=[0], n_clients=threads)        # in the parent, before the Pool is created
    Pool(threads, initializer=lambda: (env.initializer(...), service.attach()))
    ...
    # inside a worker, either the batched call sites (callsites.py) ...
    rec, cig = service.client().align_pairs(refs, queries, 10, 4, 8, 2)
    # ... or the unmodified per-call sites: Aligner(ref, ...).align(query) blocks on the service
    ssw_wrap.use_service(service.client())

Transport: every client slot owns two shared-memory arenas (request, response).  A request is written into the
arena as struct-of-arrays (lengths, then the sequence bytes -- ASCII, the owner encodes on the device) and
announced with a few-byte message on the shared request queue; the owner maps the arena, so sequence bytes are
never pickled.  The owner collects announcements until ``max_pairs`` pairs are waiting or ``flush_ms`` has passed
since the first one, groups them by scoring scheme, runs ONE device batch per group through the C ABI
(ssw_align_batch_multi), writes each client's records and CIGAR ops into its response arena and answers on the
client's own queue.  Requests of many workers therefore share device batches; with several devices every
device has its own owner process and all of them pull from the same request queue.

A client has one outstanding request at a time (its calls block, like ``Aligner.align``), so an arena is a
single slot, not a ring; a request larger than the arena is split by the client.
"""
import multiprocessing as mp
import os
import time
from multiprocessing import shared_memory

import numpy as np

_REC_DTYPE = np.dtype([('score1', '<i4'), ('score2', '<i4'), ('ref_begin1', '<i4'), ('ref_end1', '<i4'),
                       ('read_begin1', '<i4'), ('read_end1', '<i4'), ('ref_end2', '<i4'), ('cigar_len', '<i4'),
                       ('cigar_off', '<i8'), ('status', '<i4'), ('word', '<i4')])
_STOP = "stop"


def _owner_main(device, req_q, resp_qs, req_names, resp_names, arena_bytes, max_pairs, flush_ms, ready, backend):
    """The owner process of one GPU.  `backend` = "cuda" (the product) or "echo" (transport tests without a device:
    score1 = len(query) * 1000 + len(ref), one CIGAR op per pair)."""
    reqs = [shared_memory.SharedMemory(name=n) for n in req_names]
    resps = [shared_memory.SharedMemory(name=n) for n in resp_names]
    sw = None
    if backend == "cuda":
        import ciri_long_b200  # noqa: F401
        from ciri_long_b200 import ssw_wrap as sw
        if sw.Aligner.libssw.ssw_cuda_device_count() <= device:
            ready.put("no CUDA device %d (libssw_cuda has no CPU path)" % device)
            return
    ready.put("ok")
    stats = dict(batches=0, pairs=0, requests=0)
    pending, stop = [], False
    parts = q_len = r_len = seqs = None
    while not stop or pending:
        # ---- collect announcements: block for the first, then until the batch is full or flush_ms has passed
        if not pending:
            m = req_q.get()
            if m == _STOP:
                break
            pending.append(m)
        t_first = time.perf_counter()
        npairs = sum(x[1] for x in pending)
        while npairs < max_pairs:
            left = flush_ms * 1e-3 - (time.perf_counter() - t_first)
            if left <= 0:
                break
            try:
                m = req_q.get(timeout=left)
            except Exception:
                break
            if m == _STOP:
                stop = True
                break
            pending.append(m)
            npairs += m[1]
        # ---- one device batch per scoring scheme
        groups = {}
        for m in pending:
            groups.setdefault(m[2:], []).append(m)
        pending = []
        for key, msgs in groups.items():
            match, mismatch, gap_open, gap_extend, need_cigar = key
            parts = []
            for slot, n in [(m[0], m[1]) for m in msgs]:
                buf = reqs[slot].buf
                q_len = np.frombuffer(buf, dtype=np.int32, count=n, offset=0)
                r_len = np.frombuffer(buf, dtype=np.int32, count=n, offset=4 * n)
                total = int(q_len.sum(dtype=np.int64) + r_len.sum(dtype=np.int64))
                seqs = np.frombuffer(buf, dtype=np.int8, count=total, offset=8 * n)
                parts.append((slot, n, q_len, r_len, seqs))
            q_len = np.concatenate([p[2] for p in parts])
            r_len = np.concatenate([p[3] for p in parts])
            seqs = np.concatenate([p[4] for p in parts])
            # layout inside a request: [q0 r0 q1 r1 ...]
            lens = np.stack([q_len.astype(np.int64), r_len.astype(np.int64)], 1).reshape(-1)
            offs = np.cumsum(lens) - lens
            q_off, r_off = offs[0::2].copy(), offs[1::2].copy()
            n_all = len(q_len)
            if backend == "cuda":
                flag = 1 if need_cigar else 4
                with sw.DeviceBatch(seqs, q_off, q_len, r_off, r_len, match, mismatch, gap_open, gap_extend, flag=flag,
                                    device=device, filterd=0 if need_cigar else -1, ascii=True) as b:
                    b.run()
                    rec, cig = b.fetch(cigar_cap=None if need_cigar else 1)
            else:
                rec = np.zeros(n_all, dtype=_REC_DTYPE)
                rec["score1"] = q_len * 1000 + r_len
                rec["cigar_len"] = 1 if need_cigar else 0
                rec["cigar_off"] = np.arange(n_all) if need_cigar else 0
                cig = (q_len.astype(np.uint32) << 4) if need_cigar else np.zeros(0, np.uint32)
            stats["batches"] += 1; stats["pairs"] += n_all; stats["requests"] += len(msgs)
            # ---- route the results back: records + this request's CIGAR ops, offsets rebased to the request
            p0 = 0
            for slot, n, _, _, _ in parts:
                r = rec[p0:p0 + n].copy()
                ops = [cig[o:o + l] for o, l in zip(r["cigar_off"], r["cigar_len"])] if need_cigar else []
                clen = r["cigar_len"].astype(np.int64)
                r["cigar_off"] = np.cumsum(clen) - clen
                cflat = np.concatenate(ops) if ops else np.zeros(0, np.uint32)
                out = resps[slot].buf
                need = r.nbytes + cflat.nbytes
                if need > arena_bytes:
                    resp_qs[slot].put(("error", "response of %d bytes does not fit the %d-byte arena" % (need, arena_bytes)))
                else:
                    np.frombuffer(out, dtype=np.uint8, count=r.nbytes)[:] = r.view(np.uint8)
                    if cflat.nbytes:
                        np.frombuffer(out, dtype=np.uint32, count=len(cflat), offset=r.nbytes)[:] = cflat
                    resp_qs[slot].put(("ok", n, len(cflat)))
                p0 += n
    ready.put(stats)
    del parts, q_len, r_len, seqs
    for s in reqs + resps:
        try:
            s.close()
        except BufferError:
            pass                                     # (views into the arena still referenced by the last batch)


class AlignClient(object):
    """One client slot of an AlignService (one per pool worker)."""

    def __init__(self, service, slot):
        self.service, self.slot = service, slot
        self.req = shared_memory.SharedMemory(name=service.req_names[slot])
        self.resp = shared_memory.SharedMemory(name=service.resp_names[slot])
        self.resp_q = service.resp_qs[slot]

    def _round_trip(self, refs, queries, params, need_cigar):
        n = len(queries)
        buf = self.req.buf
        q_len = np.fromiter(map(len, queries), dtype=np.int32, count=n)
        r_len = np.fromiter(map(len, refs), dtype=np.int32, count=n)
        np.frombuffer(buf, dtype=np.int32, count=n, offset=0)[:] = q_len
        np.frombuffer(buf, dtype=np.int32, count=n, offset=4 * n)[:] = r_len
        blob = "".join([x for pair in zip(queries, refs) for x in pair]).encode("latin-1", "replace")
        buf[8 * n:8 * n + len(blob)] = blob
        self.service.req_q.put((self.slot, n) + tuple(params) + (bool(need_cigar),))
        try:
            ans = self.resp_q.get(timeout=self.service.timeout_s)
        except Exception:
            raise RuntimeError("align service: no answer within %.0f s (owner process gone?)" % self.service.timeout_s)
        if ans[0] != "ok":
            raise RuntimeError("align service: " + str(ans[1]))
        _, n_back, n_ops = ans
        rec = np.frombuffer(self.resp.buf, dtype=_REC_DTYPE, count=n_back).copy()
        cig = np.frombuffer(self.resp.buf, dtype=np.uint32, count=n_ops, offset=rec.nbytes).copy()
        return rec, cig

    def align_pairs(self, refs, queries, match=2, mismatch=2, gap_open=3, gap_extend=1, need_cigar=False):
        """Records (RESULT_DTYPE of ssw_wrap) and CIGAR ops of ``Aligner(r, ...).align(q)`` for every pair; blocks
        until the owner has run the device batch this request became part of."""
        refs, queries = list(refs), list(queries)
        if len(refs) != len(queries):
            raise ValueError("refs and queries differ in length")
        params = (match, mismatch, gap_open, gap_extend)
        arena = self.service.arena_bytes
        recs, cigs, base, i = [], [], 0, 0
        while i < len(queries):
            # as many pairs as fit the request arena (lengths + bytes) and, conservatively, the response arena
            j, used, out = i, 0, 0
            while j < len(queries):
                add = 8 + len(queries[j]) + len(refs[j])
                add_out = _REC_DTYPE.itemsize + (4 * (2 * len(queries[j]) + 3) if need_cigar else 0)
                if j > i and (used + add > arena or out + add_out > arena):
                    break
                used += add; out += add_out; j += 1
            if used > arena:
                raise ValueError("one pair of %d bytes exceeds the %d-byte arena (AlignService(arena_mb=...))" % (used, arena))
            r, c = self._round_trip(refs[i:j], queries[i:j], params, need_cigar)
            r["cigar_off"] += base
            base += len(c)
            recs.append(r); cigs.append(c)
            i = j
        if not recs:
            return np.zeros(0, dtype=_REC_DTYPE), np.zeros(0, np.uint32)
        return np.concatenate(recs), np.concatenate(cigs)

    def align(self, ref, query, match=2, mismatch=2, gap_open=3, gap_extend=1, need_cigar=True):
        rec, cig = self.align_pairs([ref], [query], match, mismatch, gap_open, gap_extend, need_cigar)
        return rec[0], cig


class AlignService(object):
    """Owner processes (one per device) + ``n_clients`` client slots.  Create it in the parent BEFORE the worker
    pool; forked workers call ``attach()`` (or ``client()``) once to claim a slot."""

    def __init__(self, devices=(0,), n_clients=None, arena_mb=64, max_pairs=262144, flush_ms=2.0, backend="cuda", timeout_s=600.0):
        ctx = mp.get_context("spawn")                      # the owners are spawned: no CUDA state is ever forked
        self.timeout_s = float(timeout_s)                  # a client gives up (with an error) if no answer comes back
        self.n_clients = n_clients or (os.cpu_count() or 1)
        self.arena_bytes = int(arena_mb) << 20
        self.req_q = ctx.Queue()
        self.resp_qs = [ctx.Queue() for _ in range(self.n_clients)]
        self._shm = []
        self.req_names, self.resp_names = [], []
        for _ in range(self.n_clients):
            a = shared_memory.SharedMemory(create=True, size=self.arena_bytes)
            b = shared_memory.SharedMemory(create=True, size=self.arena_bytes)
            self._shm += [a, b]
            self.req_names.append(a.name); self.resp_names.append(b.name)
        self._next_slot = ctx.Value("i", 0)
        self._ready = ctx.Queue()
        self._client = None
        self.owners = []
        for d in devices:
            p = ctx.Process(target=_owner_main, args=(d, self.req_q, self.resp_qs, self.req_names, self.resp_names, self.arena_bytes,
                                                       max_pairs, flush_ms, self._ready, backend), daemon=True)
            p.start()
            self.owners.append(p)
        for _ in self.owners:
            msg = self._ready.get(timeout=300)
            if msg != "ok":
                self.close()
                raise RuntimeError("align service: " + str(msg))

    def client(self):
        """This process's client (claims a slot on first use; call it in the worker, after the fork)."""
        if self._client is None or self._client[0] != os.getpid():
            with self._next_slot.get_lock():
                slot = self._next_slot.value
                self._next_slot.value += 1
            if slot >= self.n_clients:
                raise RuntimeError("align service: more clients than slots (n_clients=%d)" % self.n_clients)
            self._client = (os.getpid(), AlignClient(self, slot))
        return self._client[1]

    def attach(self):
        """Pool initializer helper: claim a slot and make ``ssw_wrap.Aligner.align`` / ``align_pairs`` of this worker
        go through the service (the worker itself never touches CUDA)."""
        from . import ssw_wrap
        ssw_wrap.use_service(self.client())

    def close(self):
        stats = []
        for _ in self.owners:
            self.req_q.put(_STOP)
        for p in self.owners:
            p.join(timeout=30)
        try:
            while True:
                s = self._ready.get_nowait()
                if isinstance(s, dict):
                    stats.append(s)
        except Exception:
            pass
        for s in self._shm:
            try:
                s.close(); s.unlink()
            except Exception:
                pass
        self._shm = []
        self.owners = []
        return stats

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
