"""Times b200mnn_dev_adjust_shift_variance (device-resident) for both tile modes; usage: time_shiftvar.py n1 n2 G [sigma]."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batchelor_b200 import device as dev, synth

n1, n2, G = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
A, B = synth.gene_batches(2, [n1, n2], G=G)
cuda = torch.device("cuda")
d1 = torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda); d2 = torch.from_numpy(np.ascontiguousarray(B.T)).to(cuda)
d1 = dev.cosine_norm(d1)[0]; d2 = dev.cosine_norm(d2)[0]
g = torch.Generator(device="cuda").manual_seed(1)
vect = torch.randn((n2, G), dtype=torch.float64, device=cuda, generator=g) * 0.01
r1 = torch.arange(n1, device=cuda, dtype=torch.int32); r2 = torch.arange(n2, device=cuda, dtype=torch.int32)
res = {}
for mode in ("fast", "fast_simt", "exact"):
    os.environ["B200MNN_SHIFTVAR"] = mode
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = dev.adjust_shift_variance(d1, d2, vect, sigma, r1, r2)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    res[mode] = out
    pairs = n2 * (n1 + n2)
    print(f"{mode}: {dt:.3f} s  ({pairs * G * (10 if mode == 'exact' else 2) / dt / 1e12:.2f} T fp64 op/s, {pairs / dt / 1e9:.2f} G pairs/s)")
print("modes identical:", bool(torch.equal(res["fast"], res["exact"]) and torch.equal(res["fast_simt"], res["exact"])), " differing cells:", int((res["fast"] != res["exact"]).sum()))
