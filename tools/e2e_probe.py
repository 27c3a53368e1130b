"""e2e timing of the one-shot C call for several chunk sizes (pinned host buffers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ciri_long_b200
from ciri_long_b200 import ssw_wrap as sw, workloads as W
n = int(sys.argv[1]) if len(sys.argv) > 1 else 524288
dev = torch.device("cuda", 0)
b = W.bsj_refinement_pairs_torch(n, dev)
for k in ("seqs", "q_off", "q_len", "r_off", "r_len"):
    t = torch.from_numpy(getattr(b, k)).pin_memory(); setattr(b, k, t.numpy()); globals()["_keep_" + k] = t
out_pin = torch.empty(n * sw.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
cig_pin = torch.empty(int(b.q_len.sum() * 0.3) + 4096, dtype=torch.int32).pin_memory()
out_np = out_pin.numpy().view(sw.RESULT_DTYPE); cig_np = cig_pin.numpy().view(np.uint32)
with sw.DeviceBatch(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1) as d:
    d.run(); d.run(); ms = d.stage_ms(); print("resident ms", float(ms.sum()), ms.tolist())
for chunk in (65536, 131072, 262144, 524288, 1 << 22):
    os.environ["SSW_CUDA_CHUNK"] = str(chunk)
    sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1, out=out_np, cig=cig_np)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(2):
        sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1, out=out_np, cig=cig_np)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 2
    print("chunk %8d  e2e %.1f ms  %.0f GCUPS" % (chunk, dt * 1e3, b.cells / dt / 1e9))
