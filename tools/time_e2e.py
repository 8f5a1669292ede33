"""Times the host-buffer C-ABI findMutualNN (the e2e figure of bench.py) call by call."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import batchelor_b200 as bb
from batchelor_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
b1, b2 = synth.pc_batches(2, [n, n], d=50)
h1 = torch.from_numpy(b1).pin_memory(); h2 = torch.from_numpy(b2).pin_memory()
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    res = bb.findMutualNN(h1.numpy(), h2.numpy(), k1=20, k2=20)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"findMutualNN {n} x {n} (host buffers): {dt * 1e3:.1f} ms, {len(res['first'])} pairs")
