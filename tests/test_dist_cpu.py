"""World-size-2 `gloo` test (CPU) of the multi-GPU host logic: query rows are sharded into contiguous blocks, each rank
computes its block, padded blocks are all-gathered and every rank ends with the full result in row order.  On GPUs the
same function runs over NCCL with the CUDA kNN as `compute_local`; here `compute_local` is the CPU oracle (tests may use it)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nq, k, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from batchelor_b200 import device as dev, synth
    from oracle import capi

    X, Q = synth.pc_batches(2, [700, nq], d=12, ncomp=4)

    def compute_local(lo, hi):
        if hi <= lo:
            return torch.zeros((0, k), dtype=torch.int32), torch.zeros((0, k), dtype=torch.float64)
        idx, d = capi.query_knn(X, Q[lo:hi], k, nthreads=1)
        return torch.from_numpy(np.ascontiguousarray(idx - 1)), torch.from_numpy(np.ascontiguousarray(d))

    idx, d = dev.shard_rows_and_gather(nq, k, compute_local, torch.device("cpu"), want_dist=True)
    np.save(os.path.join(out_dir, f"idx_{rank}.npy"), idx.numpy())
    np.save(os.path.join(out_dir, f"dist_{rank}.npy"), d.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("nq", [501, 2, 1])   # ragged split, fewer rows than ranks
def test_query_sharding_world_size_2_gloo(tmp_path, nq):
    from batchelor_b200 import device as dev, synth
    from oracle import capi

    k, world = 5, 2
    mp.spawn(_worker, args=(world, _free_port(), nq, k, str(tmp_path)), nprocs=world, join=True)
    X, Q = synth.pc_batches(2, [700, nq], d=12, ncomp=4)
    want_idx, want_d = capi.query_knn(X, Q, k)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"idx_{r}.npy"), want_idx - 1)
        assert np.array_equal(np.load(tmp_path / f"dist_{r}.npy"), want_d)
    # shard bounds: contiguous, ordered, cover everything
    for n in (0, 1, 2, 7, 1000):
        for w in (1, 2, 4, 8):
            blocks = [dev.shard_bounds(n, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))


def _worker_split(rank, world, port, n1, n2, k1, k2, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from batchelor_b200 import device as dev, synth
    from oracle import capi

    A, B = synth.pc_batches(2, [n1, n2], d=12, ncomp=4)
    direction, lo, hi, _, _ = dev.direction_split(n1, n2, world, rank)
    X, Q, k = (B, A, k2) if direction == 0 else (A, B, k1)
    if hi > lo:
        idx = capi.query_knn(X, Q[lo:hi], k, nthreads=1)[0] - 1
    else:
        idx = np.zeros((0, k), dtype=np.int32)
    w21, w12 = dev.direction_split_gather(torch.from_numpy(np.ascontiguousarray(idx.astype(np.int32))), n1, k2, n2, k1, world)
    np.save(os.path.join(out_dir, f"w21_{rank}.npy"), w21.numpy())
    np.save(os.path.join(out_dir, f"w12_{rank}.npy"), w12.numpy())
    # row-sharded pair extraction: this rank's block of batch-1 rows through the CPU oracle, then the exchange
    lo, hi, _ = dev.shard_bounds(n1, world, rank)
    if hi > lo:
        f, s = capi.find_mutual_nns(np.asfortranarray(w21.numpy()[lo:hi] + 1), np.asfortranarray(w12.numpy() + 1 - lo))
        f = f - 1 + lo; s = s - 1
    else:
        f = np.zeros(0, np.int32); s = np.zeros(0, np.int32)
    first, second = dev.gather_pair_blocks(torch.from_numpy(f.astype(np.int32)), torch.from_numpy(s.astype(np.int32)), world)
    np.save(os.path.join(out_dir, f"first_{rank}.npy"), first.numpy())
    np.save(os.path.join(out_dir, f"second_{rank}.npy"), second.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n1,n2,k1,k2", [(2, 301, 257, 5, 5), (3, 400, 90, 4, 7), (2, 3, 2, 1, 2)])
def test_direction_split_gloo(tmp_path, world, n1, n2, k1, k2):
    """findMutualNN over ranks: directions first, rows second; one all-gather; every rank ends with both index matrices."""
    from batchelor_b200 import device as dev, synth
    from oracle import capi

    mp.spawn(_worker_split, args=(world, _free_port(), n1, n2, k1, k2, str(tmp_path)), nprocs=world, join=True)
    A, B = synth.pc_batches(2, [n1, n2], d=12, ncomp=4)
    want21 = capi.query_knn(B, A, k2)[0] - 1
    want12 = capi.query_knn(A, B, k1)[0] - 1
    wf, ws_ = capi.find_mutual_nns(np.asfortranarray(want21 + 1), np.asfortranarray(want12 + 1))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"w21_{r}.npy"), want21)
        assert np.array_equal(np.load(tmp_path / f"w12_{r}.npy"), want12)
        assert np.array_equal(np.load(tmp_path / f"first_{r}.npy"), wf - 1)     # the reference's order survives the row sharding
        assert np.array_equal(np.load(tmp_path / f"second_{r}.npy"), ws_ - 1)
    # the split covers both directions with contiguous blocks for every world size
    for w in (2, 3, 4, 8):
        for (a, b) in ((1000, 1000), (10, 100000), (100000, 10)):
            parts = [dev.direction_split(a, b, w, r) for r in range(w)]
            h = parts[0][4]
            assert 1 <= h <= w - 1
            d0 = [p for p in parts if p[0] == 0]; d1 = [p for p in parts if p[0] == 1]
            assert len(d0) == h and d0[0][1] == 0 and d0[-1][2] == a and d1[0][1] == 0 and d1[-1][2] == b
