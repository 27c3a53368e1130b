"""Multi-GPU host logic on CPU: shards are disjoint, cover every pair, are cost-balanced, and a
world_size-2 gloo job gathers per-rank results back into original pair order."""
import os
import socket

import numpy as np
import pytest


def test_lpt_shards_cover_and_balance():
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import sharding, workloads as W
    b = W.bsj_refinement_pairs(4000, seed=4)
    j = W.junction_pairs(6000, seed=5)
    q_len = np.concatenate([b.q_len, j.q_len]); r_len = np.concatenate([b.r_len, j.r_len])
    for world in (1, 2, 4, 8):
        sh = sharding.lpt_shards(q_len, r_len, world)
        allidx = np.concatenate(sh)
        assert len(allidx) == len(q_len) and len(np.unique(allidx)) == len(q_len)
        cost = np.array([sharding.pair_cost(q_len[s], r_len[s]).sum() for s in sh], dtype=np.float64)
        assert cost.max() / cost.mean() < 1.01


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import sharding, ssw_wrap as sw, workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = W.junction_pairs(1000, seed=9)
    idx, sub = sharding.shard_batch(b, rank, world)
    # stand-in for the device: a deterministic function of the pair, so the gather can be verified
    rec = np.zeros(len(idx), dtype=sw.RESULT_DTYPE)
    rec["score1"] = sub["q_len"] * 3 + sub["r_len"]
    rec["cigar_len"] = 1 + (idx % 3)
    rec["cigar_off"] = np.cumsum(rec["cigar_len"]) - rec["cigar_len"]
    cig = np.repeat(idx.astype(np.uint32), rec["cigar_len"])
    parts = [None] * world
    dist.all_gather_object(parts, (idx, rec, cig))
    if rank == 0:
        full, cigs = sharding.gather_results(len(b), parts)
        ok = bool((full["score1"] == b.q_len * 3 + b.r_len).all())
        for i in range(len(b)):
            seg = cigs[full["cigar_off"][i]:full["cigar_off"][i] + full["cigar_len"][i]]
            ok = ok and len(seg) == 1 + i % 3 and bool((seg == i).all())
        out.put(ok)
    dist.destroy_process_group()


def test_gloo_world2_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    ok = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
    assert ok


def test_shard_batch_compacts_its_sequences():
    """a shard carries only its own bytes (pair-major, rebased offsets) and the same sequences as the original pairs"""
    import ciri_long_b200  # noqa: F401
    from ciri_long_b200 import sharding, workloads as W
    b = W.concat_batches([W.junction_pairs(600, seed=3), W.square_pairs(40, 200, params=(10, 4, 8, 2))], shuffle_seed=1)
    tot = 0
    for rank in range(4):
        idx, sub = sharding.shard_batch(b, rank, 4)
        assert len(sub["seqs"]) == int(b.q_len[idx].sum() + b.r_len[idx].sum())
        tot += len(sub["seqs"])
        for k in (0, len(idx) // 2, len(idx) - 1):
            i = idx[k]
            assert (sub["seqs"][sub["q_off"][k]:sub["q_off"][k] + sub["q_len"][k]] == b.query(i)).all()
            assert (sub["seqs"][sub["r_off"][k]:sub["r_off"][k] + sub["r_len"][k]] == b.ref(i)).all()
    assert tot == int(b.q_len.sum() + b.r_len.sum())


def test_bench_c5_slabs_have_one_owner_each():
    """bench.py --config C5: the same 8 slabs for every world size, each owned by exactly one rank (strong scaling)"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for world in (1, 2, 3, 4, 8):
        owned = [s for r in range(world) for s in bench.c5_slabs_of(r, world)]
        assert sorted(owned) == list(range(bench.C5_SLABS))
    assert {len(bench.c5_slabs_of(r, 8)) for r in range(8)} == {1}
