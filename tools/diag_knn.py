"""First-contact diagnostics for the tcgen05 kNN kernel (run on the GPU box): compares the tensor-core scores of tiny
problems with exact fp64 squared distances so that a wrong descriptor/layout shows up as a pattern, not just a failed test."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from batchelor_b200 import device as dev, synth

torch.cuda.set_device(0)
print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))


def one(n, nq, d, k=5, verbose=False):
    X, Q = synth.pc_batches(2, [n, nq], d=d, ncomp=4)
    Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
    cidx, cd2, thr = dev.debug_candidates(Xd, Qd, k)
    torch.cuda.synchronize()
    cidx, cd2 = cidx.cpu().numpy(), cd2.cpu().numpy()
    valid = cidx >= 0
    exact = ((Q[:, None, :] - X[np.where(valid, cidx, 0)]) ** 2).sum(-1)
    err = np.where(valid, np.abs(cd2 - exact), 0)
    nvalid = valid.sum(1)
    print(f"n={n} nq={nq} d={d}: candidates/query min {nvalid.min()} max {nvalid.max()}  max|err| {err.max():.3e}  "
          f"median d2 {np.median(exact[valid]):.3f}  rows with err>1e-2: {(err.max(1) > 1e-2).sum()}")
    if verbose or err.max() > 1e-2:
        r = int(np.argmax(err.max(1)))
        print("  worst row", r, "idx", cidx[r][:8], "approx", np.round(cd2[r][:8], 3), "exact", np.round(exact[r][:8], 3))
        print("  row 0 approx", np.round(cd2[0][:8], 3), "exact", np.round(exact[0][:8], 3))
    idx, dist = dev.query_knn(Xd, Qd, k)
    torch.cuda.synchronize()
    d2 = ((Q[:, None, :] - X[None, :, :]) ** 2).sum(-1) if n * nq <= 4_000_000 else None
    if d2 is not None:
        order = np.lexsort((np.broadcast_to(np.arange(n), d2.shape), d2), axis=1)[:, :k]
        print("  final kNN mismatching slots:", int((idx.cpu().numpy() != order).sum()))


for d in (16, 2, 50, 64, 100):
    one(32, 128, d, verbose=(d == 16))
one(256, 128, 50)
one(1000, 300, 50, k=20)
one(5000, 1000, 50, k=20)
t = time.time()
X, Q = synth.pc_batches(2, [200_000, 200_000], d=50)
Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
stats = torch.zeros(4, dtype=torch.int64, device="cuda")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.time()
    idx, dist = dev.query_knn(Xd, Qd, 20, stats=stats)
    torch.cuda.synchronize(); print(f"200k x 200k: {time.time() - t0:.4f} s, stats {stats.cpu().numpy()}")
