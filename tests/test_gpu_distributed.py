"""Multi-GPU (NCCL, one rank per GPU) sharded sketch == single-GPU sketch.  Skipped with fewer than
two visible GPUs (the world_size-2 host logic is covered on CPU by test_distributed_cpu.py)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tt-sketch_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device(f"cuda:{rank}"))
    try:
        from tt_sketch.distributed import distributed_stream_sketch
        from tt_sketch.drm import SparseGaussianDRM, TensorTrainDRM
        from tt_sketch.sketch import stream_sketch
        from tt_sketch.tensor import CPTensor, SparseTensor, TensorTrain

        shape = (10000, 10000, 10000, 500)
        rng = np.random.default_rng(0)
        nnz = 400_001
        idx = np.stack([rng.integers(0, n, nnz) for n in shape]).astype(np.int64)
        sp = SparseTensor(shape, idx, rng.standard_normal(nnz))
        lr, rr = (20,) * 3, (40,) * 3
        L = SparseGaussianDRM(lr, shape=shape, transpose=False, seed=1)
        R = SparseGaussianDRM(rr, shape=shape, transpose=True, seed=2)
        got = distributed_stream_sketch(sp, L, R)
        want = stream_sketch(sp, lr, rr, left_drm=L, right_drm=R)
        a, b = got.pack(), want.sketch_.pack()
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
        shape2 = (30, 40, 50, 20)
        idx2 = np.stack([rng.integers(0, n, 5000) for n in shape2]).astype(np.int64)
        tsum = (TensorTrain.random(shape2, 3, seed=5) + CPTensor.random(shape2, 4, seed=6)
                + SparseTensor(shape2, idx2, rng.standard_normal(5000)))
        Lt = TensorTrainDRM((6, 7, 8), shape=shape2, transpose=False, seed=3)
        Rt = TensorTrainDRM((9, 10, 11), shape=shape2, transpose=True, seed=4)
        got = distributed_stream_sketch(tsum, Lt, Rt)
        want = stream_sketch(tsum, (6, 7, 8), (9, 10, 11), left_drm=Lt, right_drm=Rt)
        a, b = got.pack(), want.sketch_.pack()
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
        # dense tensor sharded in slabs of the first mode (BASELINE configs[0] shape)
        from tt_sketch.distributed import distributed_orthogonal_sketch
        from tt_sketch.sketch import orthogonal_sketch
        from tt_sketch.tensor import DenseTensor

        shape3 = (20,) * 5
        dn = DenseTensor(np.random.default_rng(0).standard_normal(shape3))
        Ld = TensorTrainDRM((10,) * 4, shape=shape3, transpose=False, seed=1)
        Rd = TensorTrainDRM((15,) * 4, shape=shape3, transpose=True, seed=2)
        got = distributed_stream_sketch(dn, Ld, Rd)
        want = stream_sketch(dn, (10,) * 4, (15,) * 4, left_drm=Ld, right_drm=Rd)
        a, b = got.pack(), want.sketch_.pack()
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
        # orthogonal sketch of a TensorSum: Psi_mu all-reduced before every QR
        cores = distributed_orthogonal_sketch(tsum, Lt, Rt)
        want_tt = orthogonal_sketch(tsum, (6, 7, 8), (9, 10, 11), left_drm=Lt, right_drm=Rt)
        for c, w in zip(cores, want_tt.cores):
            assert c.shape == w.shape and np.max(np.abs(c - w)) <= 1e-9 * np.max(np.abs(w))
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_sketch_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
