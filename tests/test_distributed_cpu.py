"""world_size-2 gloo test of the multi-GPU host logic on CPU: sharding, packing, the single
all-reduce and unpacking.  The per-rank partial sketch is produced by the CPU oracle here (the
test harness is allowed to call it; the product's own local sketch needs a GPU), so what is
verified is that  shard -> local sketch -> all-reduce(sum)  reproduces the unsharded sketch."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_local_sketch(part, left_drm, right_drm, total):
    """Packed partial sketch of a shard through the oracle (test infrastructure only)."""
    from oracle import sketch_oracle as orc
    from tt_sketch.sketch_container import SketchContainer
    from tt_sketch.tensor import CPTensor, SparseTensor, TensorSum, TensorTrain

    if part is None:
        return np.zeros(total)

    def desc(t):
        if isinstance(t, SparseTensor):
            return ("sparse", t.shape, np.asarray(t.indices), t.entries)
        if isinstance(t, TensorTrain):
            return ("tt", t.cores)
        if isinstance(t, CPTensor):
            return ("cp", t.cores)
        if isinstance(t, TensorSum):
            return ("sum", [desc(x) for x in t.tensors])
        raise TypeError(type(t))

    def odrm(d):
        kind = "gauss" if type(d).__name__ == "SparseGaussianDRM" else "tt"
        return orc.Drm(kind, d.transpose, d.shape, d.bond_rank_min, d.bond_rank_max, d.seed, list(getattr(d, "cores", [])))

    Psi, Om = orc.general_sketch(desc(part), odrm(left_drm), odrm(right_drm), "streaming", fast_sparse=True)
    return SketchContainer(Psi, Om).pack()


def _worker(rank, world, port, out_dir):
    for p in (ROOT, os.path.join(ROOT, "tt-sketch_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist

    from tt_sketch.distributed import (distributed_blocked_stream_sketch, distributed_stream_sketch, shard_bounds,
                                       shard_tensor)
    from tt_sketch.drm import SparseGaussianDRM, TensorTrainDRM
    from tt_sketch.tensor import CPTensor, SparseTensor, TensorTrain

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        shape = (7, 8, 9, 6)
        rng = np.random.default_rng(0)  # same data on every rank
        idx = np.stack([rng.integers(0, n, 501) for n in shape]).astype(np.int64)
        sp = SparseTensor(shape, idx, rng.standard_normal(501))
        # 1. sparse + lazy Gaussian DRMs
        L = SparseGaussianDRM((3, 4, 5), shape=shape, transpose=False, seed=1)
        R = SparseGaussianDRM((5, 6, 7), shape=shape, transpose=True, seed=2)
        got = distributed_stream_sketch(sp, L, R, local_sketch=_oracle_local_sketch)
        want = _oracle_local_sketch(sp, L, R, got.pack().size)
        assert np.allclose(got.pack(), want, rtol=1e-12, atol=1e-12)
        lo, hi = shard_bounds(sp.nnz, world, rank)
        assert shard_tensor(sp, world, rank).nnz == hi - lo
        # 2. TensorSum of TT + CP + sparse with TT-DRMs: summands dealt round-robin, sparse range-split
        Lt = TensorTrainDRM((3, 4, 5), shape=shape, transpose=False, seed=3)
        Rt = TensorTrainDRM((5, 6, 7), shape=shape, transpose=True, seed=4)
        tsum = TensorTrain.random(shape, 2, seed=5) + CPTensor.random(shape, 3, seed=6) + sp + TensorTrain.random(shape, 3, seed=7)
        got = distributed_stream_sketch(tsum, Lt, Rt, local_sketch=_oracle_local_sketch)
        want = _oracle_local_sketch(tsum, Lt, Rt, got.pack().size)
        assert np.allclose(got.pack(), want, rtol=1e-11, atol=1e-11)
        mine = shard_tensor(tsum, world, rank)
        assert sum(1 for x in mine.tensors if not isinstance(x, SparseTensor)) in (1, 2)
        # 3. blocked sketch == unblocked
        blk = distributed_blocked_stream_sketch(sp, L, R, [(0, 0, 0), (2, 2, 2), (3, 4, 5)], [(0, 0, 0), (3, 3, 3), (5, 6, 7)],
                                                local_sketch=_oracle_local_sketch)
        full = distributed_stream_sketch(sp, L, R, local_sketch=_oracle_local_sketch)
        assert np.allclose(blk.pack(), full.pack(), rtol=1e-12, atol=1e-12)
        # 4. a lone TT does not shard: rank 0 sketches it, the others add zeros
        tt = TensorTrain.random(shape, 2, seed=8)
        assert (shard_tensor(tt, world, rank) is None) == (rank != 0)
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_sharded_sketch_world2_gloo(tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_shard_bounds_cover_everything():
    sys.path.insert(0, os.path.join(ROOT, "tt-sketch_b200"))
    from tt_sketch.distributed import shard_bounds

    for n in (0, 1, 7, 100, 1001):
        for world in (1, 2, 3, 8):
            cuts = [shard_bounds(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cuts, cuts[1:]))
            assert max(h - l for l, h in cuts) - min(h - l for l, h in cuts) <= 1
