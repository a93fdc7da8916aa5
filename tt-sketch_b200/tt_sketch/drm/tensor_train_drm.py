"""Random tensor-train DRM: sketches by partial contractions with a fixed Gaussian TT.

Mirror of tt_sketch/drm/tensor_train_drm.py:23-122 (reference): same constructor, same core
values (cores come from TensorTrain.random(..., norm_goal="norm-preserve") on the HOST, because
the reference's generator depends on the host's cpu_count(); they are uploaded once), same
`sketch_sparse / sketch_tt / sketch_cp / sketch_dense` generators.  Every contraction is a
libttsk kernel: strided FP64 GEMMs for TT / CP / dense input, the per-nonzero chain kernel for
sparse input, `sketch_tucker` (:124-145) included.
"""
from __future__ import annotations

from typing import Optional, Tuple, Union

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.drm_base import CanSlice, handle_transpose
from tt_sketch.sketching_methods.abstract_methods import (CansketchCP, CansketchDense, CansketchSparse,
                                                          CanSketchTucker, CansketchTT)
from tt_sketch.tensor import CPTensor, DenseTensor, SparseTensor, TensorTrain, TuckerTensor


class TensorTrainDRM(CansketchSparse, CansketchTT, CansketchCP, CanSlice, CansketchDense, CanSketchTucker):
    kind = be.DRM_TT

    def __init__(self, rank: Union[Tuple[int, ...], int], shape: Tuple[int, ...], transpose: bool,
                 seed: Optional[int] = None, **kwargs) -> None:
        super().__init__(rank, shape, transpose, seed=seed, **kwargs)
        if "cores" in kwargs:
            self.cores = kwargs["cores"]
        else:
            tt_shape = self.shape[::-1] if transpose else self.shape
            tt = TensorTrain.random(tt_shape, self.true_rank, self.seed, norm_goal="norm-preserve")
            self.cores = tt.cores[:-1]
        self._dev_cores = kwargs.get("_dev_cores", {})

    def _slice_kwargs(self):
        # a slice shares the parent's cores (the reference regenerates identical ones)
        return {"cores": self.cores, "_dev_cores": self._dev_cores}

    def restrict_first_mode(self, lo: int, hi: int) -> "TensorTrainDRM":
        """The same DRM for the slab X[lo:hi] of a tensor (multi-GPU slab sharding): a left DRM keeps rows
        lo:hi of its first core; a right DRM's cores never touch the first mode (its TT is built for the
        reversed shape and drops that mode's core, reference tensor_train_drm.py:46-56), so only `shape` changes."""
        cores = list(self.cores)
        if not self.transpose:
            c0 = cores[0]
            cores[0] = c0[:, lo:hi, :].contiguous() if be.is_device(c0) else np.ascontiguousarray(c0[:, lo:hi, :])
        return type(self)(rank=self.bond_rank_max, shape=(hi - lo,) + tuple(self.shape[1:]), transpose=self.transpose,
                          seed=self.seed, rank_min=self.bond_rank_min, rank_max=self.bond_rank_max,
                          true_rank=self.bond_true_rank, cores=cores)

    def device_core(self, k: int):
        """Core k (DRM orientation) as a contiguous device tensor, uploaded once."""
        c = self.cores[k]
        if be.is_device(c):
            return c if c.is_contiguous() else c.contiguous()
        key = (be.device_index(), id(c))
        hit = self._dev_cores.get(key)
        if hit is None or hit[0] is not c:  # the host array is kept alive next to its upload, so ids cannot be recycled
            hit = (c, be.to_device(c, np.float64))
            self._dev_cores[key] = hit
        return hit[1]

    # ------------------------------------------------------------------ sparse
    @handle_transpose
    def sketch_sparse_device(self, tensor: SparseTensor):
        """Chained per-nonzero core products v_mu = v_{mu-1} @ core_mu[:, i_mu, :]; yields the
        column slice [rank_min, rank_max) as a (rank, nnz) view."""
        idx = tensor.device()["indices"]
        v, mu = None, 0
        while mu < len(self.cores):  # re-read every step: OrthogTTDRM appends cores lazily
            v = be.ttdrm_sparse_step(idx[mu], v, self.device_core(mu))
            yield v[:, self.rank_min[mu]:self.rank_max[mu]].T
            mu += 1

    # ------------------------------------------------------------------ tensor train
    @handle_transpose
    def sketch_tt_device(self, tensor: TensorTrain):
        """lr_mu (r_T, r_D): DRM contracted with the first mu+1 cores of the TT."""
        cores = tensor.device()["cores"]
        lr, mu = None, 0
        while mu < len(self.cores):
            c, g = cores[mu], self.device_core(mu)
            rT0, n, rT1 = c.shape
            rD0, _, rD1 = g.shape
            if mu == 0:
                lr = be.gemm(c.reshape(n, rT1).T, g.reshape(n, rD1))
            else:
                w = be.gemm(lr.T, c.reshape(rT0, n * rT1))            # (rD0, n*rT1)
                lr = be.gemm(w.reshape(rD0 * n, rT1).T, g.reshape(rD0 * n, rD1))
            yield lr[:, self.rank_min[mu]:self.rank_max[mu]]
            mu += 1

    # ------------------------------------------------------------------ CP
    @handle_transpose
    def sketch_cp_device(self, tensor: CPTensor):
        """lr_mu (R_cp, r_D):  lr_mu[i, l] = sum_{j,k} lr_{mu-1}[i, j] A_mu[k, i] G_mu[j, k, l]."""
        cores = tensor.device()["cores"]
        lr, mu = None, 0
        while mu < len(self.cores):
            a, g = cores[mu], self.device_core(mu)
            n, R = a.shape
            rD0, _, rD1 = g.shape
            if mu == 0:
                lr = be.gemm(a.T, g.reshape(n, rD1))
            else:
                w = be.gemm(lr, g.reshape(rD0, n * rD1)).reshape(R, n, rD1)   # w[i, k, l]
                out = be.empty((R, 1, rD1))
                # batched over the CP index i: (1 x n) row A[:, i] times (n x rD1) slab w[i]
                be.gemm_batched(a.T.unsqueeze(1), w, out)
                lr = out.reshape(R, rD1)
            yield lr[:, self.rank_min[mu]:self.rank_max[mu]]
            mu += 1

    # ------------------------------------------------------------------ dense
    @handle_transpose
    def sketch_dense_device(self, tensor: DenseTensor):
        """Dense unfoldings of the DRM itself, (true_rank[mu], prod(shape[:mu+1])); like the
        reference this ignores rank_min/rank_max (no blocked dense sketch)."""
        g0 = self.device_core(0)
        pc = g0.reshape(-1, g0.shape[-1])
        yield pc.T
        mu = 1
        while mu < len(self.cores):
            g = self.device_core(mu)
            r0, n, r1 = g.shape
            pc = be.gemm(pc, g.reshape(r0, n * r1)).reshape(-1, r1)
            yield pc.T
            mu += 1

    # ------------------------------------------------------------------ Tucker
    @handle_transpose
    def sketch_tucker_device(self, tensor: TuckerTensor):
        """pc_mu (prod(s_0..s_mu), r_D): the DRM cores contracted with the Tucker factors mode by mode.  Like the
        reference (tensor_train_drm.py:124-145) this takes the whole DRM: a rank slice is refused."""
        if tuple(self.rank) != tuple(self.true_rank):
            raise ValueError("a sliced TensorTrainDRM cannot sketch a TuckerTensor (the reference reshapes to the "
                             "unsliced rank)")
        factors = tensor.device()["factors"]
        pc, mu = None, 0
        while mu < len(self.cores):
            g, U = self.device_core(mu), factors[mu]      # (r0, n, r1), (s, n)
            r0, n, r1 = g.shape
            s = U.shape[0]
            if mu == 0:
                pc = be.gemm(U, g.reshape(n, r1))
            else:
                red = be.empty((r0, s, r1))               # red[j] = U @ g[j]
                be.gemm_batched(U.unsqueeze(0).expand(r0, s, n), g, red)
                pc = be.gemm(pc, red.reshape(r0, s * r1)).reshape(-1, r1)
            yield pc
            mu += 1
