"""Sparse sign DRM: every row of the (nnz, rank) sketching matrix holds a fixed number of non-zero "sign" entries
at hash-seeded positions.

Mirror of tt_sketch/drm/sparse_sign_drm.py:12-51 (reference): same constructor (`num_non_zero_per_row`, default the
true rank), same `sketch_sparse` generator.  The rows are produced by `ttsk_lazy_sparse_sign`, an integer-exact
restatement of the reference's Cython routine (fast_lazy_gaussian.pyx:121-180) (signs from the exponent parity of a hashed uniform, positions from a
partial Fisher-Yates shuffle driven by its mantissa) and feed the operator-level sketching kernels like any other per-nonzero rows.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

from tt_sketch import _backend as be
from tt_sketch.drm_base import CanSlice, handle_transpose
from tt_sketch.sketching_methods.abstract_methods import CansketchSparse
from tt_sketch.tensor import SparseTensor


class SparseSignDRM(CansketchSparse, CanSlice):
    def __init__(self, rank: Union[Tuple[int, ...], int], shape: Tuple[int, ...], transpose: bool,
                 seed: Optional[int] = None, num_non_zero_per_row: Optional[Tuple[int, ...]] = None, **kwargs) -> None:
        super().__init__(rank, shape, transpose, seed=seed, **kwargs)
        if num_non_zero_per_row is None:
            num_non_zero_per_row = self.true_rank
        self.nnz = num_non_zero_per_row
        for k, r in zip(self.nnz, self.true_rank):
            if not 0 < int(k) <= int(r):
                raise ValueError(f"num_non_zero_per_row {tuple(self.nnz)} must lie in [1, rank] = {tuple(self.true_rank)} "
                                 "(the reference writes past the row otherwise)")

    # like the reference, a slice does not inherit a custom `num_non_zero_per_row` (drm_base.py:92-109 passes only
    # rank / seed / slice bounds): it falls back to the true rank

    @handle_transpose
    def sketch_sparse_device(self, tensor: SparseTensor):
        """Yields (rank[mu], nnz) device views; bond mu uses the first mu+1 index rows of the (possibly transposed)
        tensor and seed mu + self.seed (sparse_sign_drm.py:34-51)."""
        dev = tensor.device()
        d = len(tensor.shape)
        for mu in range(d - 1):
            rows = be.lazy_sparse_sign(dev["indices"], mu + 1, tensor.nnz, tensor.shape, self.true_rank[mu],
                                       self.rank_min[mu], self.rank_max[mu], self.nnz[mu], (mu + self.seed) % 2**63)
            yield rows.T
