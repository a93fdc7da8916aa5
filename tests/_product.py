"""Build product-package tensors / DRMs from the golden fixtures (see tests/_golden.py)."""
import numpy as np


def make_tensor(desc):
    from tt_sketch.tensor import CPTensor, DenseTensor, SparseTensor, TensorSum, TensorTrain

    kind = desc[0]
    if kind == "sparse":
        return SparseTensor(desc[1], np.ascontiguousarray(desc[2]), np.ascontiguousarray(desc[3]))
    if kind == "dense":
        return DenseTensor(np.ascontiguousarray(desc[1]))
    if kind == "tt":
        return TensorTrain([np.ascontiguousarray(c) for c in desc[1]])
    if kind == "cp":
        return CPTensor([np.ascontiguousarray(c) for c in desc[1]])
    if kind == "sum":
        return TensorSum([make_tensor(s) for s in desc[1]])
    raise ValueError(kind)


def make_drm(odrm):
    """oracle `Drm` record -> product DRM object with identical seed / slices / cores."""
    from tt_sketch.drm import SparseGaussianDRM, TensorTrainDRM

    full = tuple(odrm.rank_max)
    if odrm.kind == "gauss":
        drm = SparseGaussianDRM(full, shape=odrm.shape, transpose=odrm.right, seed=odrm.seed)
    else:
        true_rank = tuple(c.shape[2] for c in odrm.cores)
        if odrm.right:
            true_rank = true_rank[::-1]
        drm = TensorTrainDRM(full, shape=odrm.shape, transpose=odrm.right, seed=odrm.seed,
                             cores=[np.ascontiguousarray(c) for c in odrm.cores], true_rank=true_rank)
    if any(odrm.rank_min):
        drm = drm.slice(tuple(odrm.rank_min), tuple(odrm.rank_max))
    return drm
