"""Times one device-resident query_knn call (used under ncu for the per-kernel launch list)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batchelor_b200 import device as dev, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
nq = int(sys.argv[4]) if len(sys.argv) > 4 else n
X, Q = synth.pc_batches(2, [n, nq], d=50)
Xd, Qd = torch.from_numpy(X).cuda(), torch.from_numpy(Q).cuda()
for rep in range(reps):
    torch.cuda.synchronize(); t0 = time.time()
    idx, dist = dev.query_knn(Xd, Qd, k)
    torch.cuda.synchronize(); print(f"{n} refs x {nq} queries k={k}: {time.time() - t0:.4f} s")
