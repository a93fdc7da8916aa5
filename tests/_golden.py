"""Helpers that turn the committed golden fixtures (tests/golden/*.npz, written by the
unmodified reference through tests/golden/make_golden.py) into oracle descriptors."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def tensor_desc(z, prefix):
    kind = str(z[prefix + "_kind"])
    if kind == "sparse":
        return ("sparse", tuple(int(x) for x in z[prefix + "_shape"]), z[prefix + "_indices"], z[prefix + "_entries"])
    if kind == "dense":
        return ("dense", z[prefix + "_data"])
    if kind in ("tt", "cp"):
        cores, i = [], 0
        while f"{prefix}_c{i}" in z:
            cores.append(z[f"{prefix}_c{i}"])
            i += 1
        return (kind, cores)
    if kind == "sum":
        return ("sum", [tensor_desc(z, f"{prefix}_s{i}") for i in range(int(z[prefix + "_n"]))])
    raise ValueError(kind)


def drm_desc(z, name, side, shape):
    """Oracle `Drm` record for side 'L' or 'R' of golden case `name`."""
    from oracle.sketch_oracle import Drm

    p = f"{name}_{side}"
    kind = {"SparseGaussianDRM": "gauss", "TensorTrainDRM": "tt"}[str(z[f"{name}_{side}kind"])]
    right = side == "R"
    rmin = tuple(int(x) for x in z[p + "_rank_min"])
    rmax = tuple(int(x) for x in z[p + "_rank_max"])
    if right:  # stored in the reference's internal (reversed) orientation
        rmin, rmax = rmin[::-1], rmax[::-1]
    cores, i = [], 0
    while f"{p}_core{i}" in z:
        cores.append(z[f"{p}_core{i}"])
        i += 1
    return Drm(kind, right, tuple(shape), rmin, rmax, int(z[p + "_seed"]), cores)


def stored_list(z, prefix):
    out, i = [], 0
    while f"{prefix}{i}" in z:
        out.append(z[f"{prefix}{i}"])
        i += 1
    return out


def rel_err(a, b):
    """max-norm relative error (SURVEY.md App. C accuracy note)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a)))
