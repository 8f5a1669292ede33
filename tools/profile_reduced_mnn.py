"""cProfile of api.reducedMNN on two synthetic 1M-cell batches (host in/out): where the non-kernel time goes."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from batchelor_b200 import api, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
b1, b2 = synth.pc_batches(2, n, d=50)
api.reducedMNN(b1[:20000], b2[:20000], k=20)   # warm-up (library load, pools)
for rep in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    pr = cProfile.Profile(); pr.enable()
    res = api.reducedMNN(b1, b2, k=20)
    torch.cuda.synchronize()
    pr.disable(); dt = time.perf_counter() - t0
    print(f"reducedMNN 2 x {n} cells: {dt:.3f} s -> {2 * n / dt / 1e6:.2f} M cells/s")
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(22); print(s.getvalue()[:3500])
