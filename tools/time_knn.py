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
import ctypes as C
from batchelor_b200 import _lib
for rep in range(reps):
    _lib.call("b200mnn_profile_enable", 1)
    torch.cuda.synchronize(); t0 = time.time()
    stats = torch.zeros(8, dtype=torch.int64, device='cuda')
    idx, dist = dev.query_knn(Xd, Qd, k, stats=stats)
    t_enq = time.time() - t0
    torch.cuda.synchronize(); dt = time.time() - t0
    kms, kl, kf = C.c_double(0), C.c_int64(0), C.c_double(0)
    _lib.call("b200mnn_profile_collect", C.byref(kms), C.byref(kl), C.byref(kf))
    _lib.call("b200mnn_profile_enable", 0)
    print(f"{n} refs x {nq} queries k={k}: {dt:.4f} s (host enqueue {t_enq * 1e3:.2f} ms; candidates kernels {kms.value / 1e3:.4f} s, {kl.value} launches; rescued {int(stats[0])}, tier-2 queries {int(stats[3])}, path {int(stats[2])}, tiles scored {int(stats[4])}+{int(stats[5])} of {int(stats[6])} dense)")
