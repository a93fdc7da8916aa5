"""Capability mix-ins of a DRM: which tensor formats it can contract with.

Interface mirror of tt_sketch/sketching_methods/abstract_methods.py:15-63 (reference).  Each
capability is an abstract generator method yielding one array per bond mu = 0..d-2:
    sketch_sparse -> (rank[mu], nnz)             rows of the DRM sampled at every nonzero
    sketch_tt     -> (tensor.rank[mu], rank[mu])  DRM contracted with the first mu+1 TT cores
    sketch_cp     -> (cp_rank, rank[mu])          same for CP factors
    sketch_dense  -> (rank[mu], prod(shape[:mu+1])) dense unfolding of the DRM itself
Every capability has a `<name>_device` twin yielding device tensors.
"""
from abc import ABC, abstractmethod

from tt_sketch.drm_base import DRM
from tt_sketch.utils import ArrayGenerator


class _DeviceTwin:
    """sketch_x(tensor) = host copies of sketch_x_device(tensor)."""

    @staticmethod
    def host(gen):
        from tt_sketch import _backend as be

        for m in gen:
            yield be.to_host(m)


class CansketchTT(DRM, ABC):
    @abstractmethod
    def sketch_tt_device(self, tensor):
        ...

    def sketch_tt(self, tensor) -> ArrayGenerator:
        return _DeviceTwin.host(self.sketch_tt_device(tensor))


class CansketchSparse(DRM, ABC):
    @abstractmethod
    def sketch_sparse_device(self, tensor):
        ...

    def sketch_sparse(self, tensor) -> ArrayGenerator:
        return _DeviceTwin.host(self.sketch_sparse_device(tensor))


class CansketchDense(DRM, ABC):
    @abstractmethod
    def sketch_dense_device(self, tensor):
        ...

    def sketch_dense(self, tensor) -> ArrayGenerator:
        return _DeviceTwin.host(self.sketch_dense_device(tensor))


class CansketchCP(DRM, ABC):
    @abstractmethod
    def sketch_cp_device(self, tensor):
        ...

    def sketch_cp(self, tensor) -> ArrayGenerator:
        return _DeviceTwin.host(self.sketch_cp_device(tensor))


class CanSketchTucker(DRM, ABC):
    """sketch_tucker -> (prod(tensor.rank[:mu+1]), rank[mu]): the DRM contracted with the first mu+1 Tucker
    factors (reference abstract_methods.py:54-63)."""

    @abstractmethod
    def sketch_tucker_device(self, tensor):
        ...

    def sketch_tucker(self, tensor) -> ArrayGenerator:
        return _DeviceTwin.host(self.sketch_tucker_device(tensor))
