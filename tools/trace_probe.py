import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import ciri_long_b200
from ciri_long_b200 import ssw_wrap as sw, workloads as W
n = 262144
dev = torch.device("cuda", 0)
b = W.bsj_refinement_pairs_torch(n, dev)
keep = []
for k in ("seqs", "q_off", "q_len", "r_off", "r_len"):
    t = torch.from_numpy(getattr(b, k)).pin_memory(); setattr(b, k, t.numpy()); keep.append(t)
out_pin = torch.empty(n * sw.RESULT_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
cig_pin = torch.empty(int(b.q_len.sum() * 0.3) + 4096, dtype=torch.int32).pin_memory()
out_np = out_pin.numpy().view(sw.RESULT_DTYPE); cig_np = cig_pin.numpy().view(np.uint32)
os.environ["SSW_CUDA_CHUNK"] = str(1 << 22)
for it in range(3):
    print("---- call", it, file=sys.stderr)
    t0 = time.perf_counter()
    sw.align_arrays(b.seqs, b.q_off, b.q_len, b.r_off, b.r_len, 1, 1, 1, 1, out=out_np, cig=cig_np)
    print("python total %.1f ms" % ((time.perf_counter() - t0) * 1e3), file=sys.stderr)
