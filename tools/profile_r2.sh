#!/bin/bash
# Round-2 profiling pass (run under gpurun on ONE GPU): launch lists and one `ncu --set full` capture per new kernel.
# Everything lands in gpurun_out/; tools/summarise_ncu.py turns the captures into the CSV summaries kept under profiles/.
set -x
O=gpurun_out
# 1. launch list of the headline bench step (shares of the step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-legs --no-parity --no-cpu-baseline --e2e-steps 1 > $O/r2_bench_under_ncu.log 2>&1
# 2. launch list of config 3 at 2 x 20k x 2000
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r2_launches_config3_20k.csv \
    python tools/run_config3.py --cells 20000 --no-check > $O/r2_config3_under_ncu.log 2>&1
# 3. full captures: the split-fp16 GEMM (smoothing logits, three-term), the DMMA pairs kernel, the selection kernels
ncu --set full --clock-control none --import-source on -k regex:gemm_split_kernel -s 2 -c 2 -o $O/r2_gemm_split -f \
    python tools/time_smooth.py 20000 8000 2000 > $O/r2_ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"sv_pairs_dmma_kernel|sv_select_kernel" -c 3 -o $O/r2_shiftvar -f \
    env B200MNN_SHIFTVAR=fast python -c "
import os, sys, numpy as np, torch
sys.path.insert(0, '.')
from batchelor_b200 import device as dev, synth
A, B = synth.gene_batches(2, [8000, 8000], G=2000)
cuda = torch.device('cuda')
d1 = dev.cosine_norm(torch.from_numpy(np.ascontiguousarray(A.T)).to(cuda))[0]; d2 = dev.cosine_norm(torch.from_numpy(np.ascontiguousarray(B.T)).to(cuda))[0]
v = torch.randn((8000, 2000), dtype=torch.float64, device=cuda) * 0.01
r = torch.arange(8000, device=cuda, dtype=torch.int32)
dev.adjust_shift_variance(d1, d2, v, 0.1, r, r); torch.cuda.synchronize()
" > $O/r2_ncu_shiftvar.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"wide_select_kernel|wide_rerank_kernel" -c 2 -o $O/r2_knn_wide -f \
    python tools/time_knn_wide.py 30000 30000 2000 > $O/r2_ncu_knn_wide.log 2>&1
# 4. race check of a small parity subset (compute-sanitizer serialises heavily: keep it small and bounded)
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
    -k "split_gemm and 300-700-200-3 or shift_variance_matches_golden and fast or tensor_path_matches_golden and vanilla or query_knn_matches_oracle_bit_exact and 1000-700" \
    > $O/r2_racecheck.log 2>&1
tail -5 $O/r2_racecheck.log
ls -la $O/*.ncu-rep
