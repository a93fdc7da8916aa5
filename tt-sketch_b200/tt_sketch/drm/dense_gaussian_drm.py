"""Dense Gaussian DRM: one stored (rank, prod(shape[:mu+1])) Gaussian matrix per bond (small tensors only).

Mirror of tt_sketch/drm/dense_gaussian_drm.py:17-80 (reference): same constructor, same matrices -- they are drawn
on the HOST with NumPy's legacy MT19937 stream seeded exactly like the reference does (a seed offset hashed from one
uniform draw per bond), so entries are bit-identical; a private RandomState replaces the reference's reseeding of the
global NumPy generator (same numbers, no side effect) -- and uploaded once.  Contractions run on the device:
`sketch_sparse` is a column gather at the nonzeros' flat indices (the level-0 row gather of `ttsk_ttdrm_sparse_step`),
`sketch_tt` the partial left-to-right contraction of the TT multiplied by the matrix (`ttsk_gemm`), `sketch_dense` the
matrix itself.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.drm_base import CanIncreaseRank, handle_transpose
from tt_sketch.sketching_methods.abstract_methods import CansketchDense, CansketchSparse, CansketchTT
from tt_sketch.tensor import DenseTensor, SparseTensor, TensorTrain


class DenseGaussianDRM(CansketchTT, CansketchSparse, CansketchDense, CanIncreaseRank):
    def __init__(self, rank: Union[Tuple[int, ...], int], shape: Tuple[int, ...], transpose: bool,
                 seed: Optional[int] = None, **kwargs) -> None:
        super().__init__(rank, shape, transpose, seed=seed, **kwargs)
        dims = self.shape[::-1] if transpose else self.shape
        self.sketching_mats = []
        cols = 1
        for mu, (r, n) in enumerate(zip(self.true_rank, dims[:-1])):
            cols *= int(n)
            # reference :47-49: the offset comes from a generator seeded with the constructor's `seed` argument
            # (not self.seed), the matrix from one seeded with (self.seed + offset) mod (2^32 - 1)
            offset = hash(np.random.RandomState(seed).uniform(0, cols))
            gen = np.random.RandomState(int(np.mod(self.seed + offset, 2**32 - 1)))
            self.sketching_mats.append(gen.normal(size=(int(r), cols))[self.rank_min[mu]:self.rank_max[mu]])
        self._dev_mats = {}

    def device_mat(self, mu: int):
        """sketching_mats[mu] on the current device (uploaded once)."""
        m = self.sketching_mats[mu]
        key = (be.device_index(), mu, id(m))
        hit = self._dev_mats.get(key)
        if hit is None:
            hit = self._dev_mats[key] = be.to_device(m, np.float64)
        return hit

    @handle_transpose
    def sketch_sparse_device(self, tensor: SparseTensor):
        """(rank[mu], nnz): columns of the stored matrix at the flat index of the leading mu+1 modes."""
        d = len(tensor.shape)
        idx = np.asarray(tensor.indices)
        for mu in range(d - 1):
            flat = np.ravel_multi_index(tuple(idx[: mu + 1]), tensor.shape[: mu + 1]).astype(np.int64)
            mt = self.device_mat(mu).T.contiguous()          # (cols, r): one row per flat index
            rows = be.ttdrm_sparse_step(be.to_device(flat, np.int64), None, mt.reshape(1, mt.shape[0], mt.shape[1]))
            yield rows.T

    @handle_transpose
    def sketch_tt_device(self, tensor: TensorTrain):
        """(tensor.rank[mu], rank[mu]): (matrix @ X[0]...X[mu] as a dense (prod n, r_T) unfolding)^T."""
        cores = tensor.device()["cores"]
        pc = None
        for mu in range(len(self.sketching_mats)):
            c = cores[mu]
            r0, n, r1 = c.shape
            pc = c.reshape(n, r1) if mu == 0 else be.gemm(pc, c.reshape(r0, n * r1)).reshape(-1, r1)
            yield be.gemm(self.device_mat(mu), pc).T

    @handle_transpose
    def sketch_dense_device(self, tensor: DenseTensor):
        for mu in range(len(self.sketching_mats)):
            yield self.device_mat(mu)
