"""Times b200mnn_dev_smooth_gaussian_kernel (device-resident): usage time_smooth.py ncells nmnn G [sigma]"""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from batchelor_b200 import device as dev, synth, _lib

ncells, nmnn, G = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
sigma = float(sys.argv[4]) if len(sys.argv) > 4 else 0.1
(B,) = synth.gene_batches(1, [ncells], G=G)
cuda = torch.device("cuda")
mat = dev.cosine_norm(torch.from_numpy(np.ascontiguousarray(B.T)).to(cuda))[0]
g = torch.Generator(device="cuda").manual_seed(1)
avg = torch.randn((nmnn, G), dtype=torch.float64, device=cuda, generator=g) * 0.01
idx = torch.sort(torch.randperm(ncells, device=cuda, generator=g)[:nmnn])[0].to(torch.int32)
outs = {}
for mode in ("tensor", "fp64"):
    if mode == "fp64" and ncells * nmnn * G > 3e13:
        continue
    os.environ["B200MNN_SMOOTH"] = mode
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = dev.smooth_gaussian_kernel(avg, idx, mat, sigma)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
    path, e0, e1 = C.c_int(0), C.c_double(0), C.c_double(0)
    _lib.call("b200mnn_smooth_last_check", C.byref(path), C.byref(e0), C.byref(e1))
    flops = 2.0 * nmnn * ncells * (2 * G) + 2.0 * nmnn * nmnn * G
    print(f"{mode}: {dt:.3f} s  {flops / dt / 1e12:.1f} algorithmic TFLOP/s  path {path.value} check rows {e0.value:.2e} dens {e1.value:.2e}")
    outs[mode] = out
if len(outs) == 2:
    a, b = outs["tensor"], outs["fp64"]
    print("tensor vs fp64: max rel err", float((a - b).abs().max() / b.abs().max()))
