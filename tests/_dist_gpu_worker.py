"""Worker of tests/test_gpu_parity.py::test_sharded_search_under_nccl (one process per GPU, launched by torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batchelor_b200 import device as dev, synth  # noqa: E402

local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
b1, b2 = synth.pc_batches(2, [30000, 26000], d=50, ncomp=8)
h1, h2 = torch.from_numpy(b1).pin_memory(), torch.from_numpy(b2).pin_memory()
d1, d2 = dev.upload_sharded(h1, device), dev.upload_sharded(h2, device)        # each rank uploads its rows, NCCL all-gather
assert torch.equal(d1.cpu(), torch.from_numpy(b1)) and torch.equal(d2.cpu(), torch.from_numpy(b2))
f, s, w21, w12 = dev.find_mutual_nn(d1, d2, 20, 20, sharded=True)              # query rows sharded over the ranks
f1, s1, a1, c1 = dev.find_mutual_nn(d1, d2, 20, 20, sharded=False)             # this rank alone
ok = torch.equal(f, f1) and torch.equal(s, s1) and torch.equal(w21, a1) and torch.equal(w12, c1)
idx, dd = dev.query_knn_sharded(d2, d1, 20, want_dist=True)
idx1, dd1 = dev.query_knn(d2, d1, 20, want_dist=True)
ok = ok and torch.equal(idx, idx1) and torch.equal(dd, dd1)
flag = torch.tensor([int(ok)], device=device)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if dist.get_rank() == 0:
    print("DIST_GPU_OK" if flag.item() == 1 else "DIST_GPU_MISMATCH", int(f.shape[0]))
dist.destroy_process_group()
