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
        if type(t).__name__ == "DenseTensor":
            return ("dense", t.data)
        raise TypeError(type(t))

    def odrm(d):
        kind = "gauss" if type(d).__name__ == "SparseGaussianDRM" else "tt"
        return orc.Drm(kind, d.transpose, d.shape, d.bond_rank_min, d.bond_rank_max, d.seed, list(getattr(d, "cores", [])))

    Psi, Om = orc.general_sketch(desc(part), odrm(left_drm), odrm(right_drm), "streaming", fast_sparse=True)
    return SketchContainer(Psi, Om).pack()


def _desc(t):
    from tt_sketch.tensor import CPTensor, SparseTensor, TensorSum, TensorTrain

    if isinstance(t, SparseTensor):
        return ("sparse", t.shape, np.asarray(t.indices), t.entries)
    if isinstance(t, TensorTrain):
        return ("tt", t.cores)
    if isinstance(t, CPTensor):
        return ("cp", t.cores)
    if isinstance(t, TensorSum):
        return ("sum", [_desc(x) for x in t.tensors])
    raise TypeError(type(t))


def _odrm(d):
    from oracle import sketch_oracle as orc

    kind = "gauss" if type(d).__name__ == "SparseGaussianDRM" else "tt"
    return orc.Drm(kind, d.transpose, d.shape, d.bond_rank_min, d.bond_rank_max, d.seed, list(getattr(d, "cores", [])))


class _OracleSequentialOps:
    """Per-rank contributions to the orthogonal sketch through the oracle (test infrastructure only): the
    same interface as tt_sketch.distributed._GpuSequentialOps."""

    def __init__(self, part, left_drm, right_drm, shape):
        from oracle import sketch_oracle as orc

        self.orc, self.shape, self.d = orc, tuple(shape), len(shape)
        self.t = _desc(part) if part is not None else None
        self.L, self.R = _odrm(left_drm), _odrm(right_drm)
        self.rL, self.rR = self.L.rank, self.R.rank
        self.Rc = orc.drm_contractions(self.R, self.t) if self.t is not None else None
        self.cores = []

    def omegas(self):
        import torch

        if self.t is None:
            return [torch.zeros(a, b, dtype=torch.float64) for a, b in zip(self.rL, self.rR)]
        Lc = self.orc.drm_contractions(self.L, self.t)
        return [torch.from_numpy(self.orc.omega(self.t, Lc[mu], self.Rc[mu], mu)) for mu in range(self.d - 1)]

    def psi(self, mu, prev_core):
        import torch

        if prev_core is not None:
            self.cores.append(np.asarray(prev_core))
        r1 = self.rL[mu - 1] if mu > 0 else 1
        r2 = self.rR[mu] if mu < self.d - 1 else 1
        if self.t is None:
            return torch.zeros(r1, self.shape[mu], r2, dtype=torch.float64)
        Lm = None
        if mu > 0:
            od = self.orc.Drm("tt", False, self.shape, (0,) * (self.d - 1), tuple(self.rL), cores=list(self.cores))
            Lm = self.orc.drm_contractions_prefix(od, self.t, mu - 1)
        Rm = self.Rc[mu] if mu < self.d - 1 else None
        return torch.from_numpy(self.orc.psi(self.t, Lm, Rm, mu, (r1, self.shape[mu], r2), True))

    def orth(self, P, Omega):
        import torch

        return torch.from_numpy(self.orc.orth_step(np.asarray(P), np.asarray(Omega)))

    def to_host(self, t):
        return np.asarray(t)


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
        # 5. dense tensor: slabs along the first mode, DRMs restricted to the slab
        from tt_sketch.distributed import DenseSlab, distributed_orthogonal_sketch
        from tt_sketch.tensor import DenseTensor

        dn = DenseTensor(np.random.default_rng(3).standard_normal(shape))
        slab = shard_tensor(dn, world, rank)
        assert isinstance(slab, DenseSlab) and (slab.lo, slab.hi) == shard_bounds(shape[0], world, rank)
        got = distributed_stream_sketch(dn, Lt, Rt, local_sketch=_oracle_local_sketch)
        want = _oracle_local_sketch(dn, Lt, Rt, got.pack().size)
        assert np.allclose(got.pack(), want, rtol=1e-11, atol=1e-11)
        # 6. ranks holding different DRMs are detected before anything is reduced
        bad = SparseGaussianDRM((5, 6, 7), shape=shape, transpose=True, seed=100 + rank)
        try:
            distributed_stream_sketch(sp, L, bad, local_sketch=_oracle_local_sketch)
            raise AssertionError("DRM mismatch not detected")
        except ValueError as e:
            assert "different DRMs" in str(e)
        # 7. orthogonal sketch of a TensorSum: Omega reduced once, every Psi_mu before its QR
        cores = distributed_orthogonal_sketch(tsum, Lt, Rt, ops_factory=_OracleSequentialOps)
        from oracle import sketch_oracle as orc
        want_cores, _ = orc.general_sketch(_desc(tsum), _odrm(Lt), _odrm(Rt), "orthogonal")
        for a, b in zip(cores, want_cores):
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-9, atol=1e-10)
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
