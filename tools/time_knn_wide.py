"""Times the wide-d exact kNN (device-resident): usage time_knn_wide.py n nq d [k]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batchelor_b200 import device as dev, synth
n, nq, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]); k = int(sys.argv[4]) if len(sys.argv) > 4 else 20
A, B = synth.gene_batches(2, [n, nq], G=d)
cuda = torch.device("cuda")
X = dev.cosine_norm(torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda))[0]; Q = dev.cosine_norm(torch.from_numpy(np.ascontiguousarray(B.T)).to(cuda))[0]
stats = torch.zeros(8, dtype=torch.int64, device=cuda)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    idx, dist = dev.query_knn(X, Q, k, stats=stats)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"wide kNN {nq} queries vs {n} x {d}, k={k}: {dt:.3f} s  {2.0*n*nq*d/dt/1e12:.1f} algorithmic TFLOP/s  {nq/dt/1e6:.3f} M queries/s  stats {stats.tolist()}")
