"""FROSTT sparse tensors (".tns" / ".tns.gz" text) -> SparseTensor.

Mirror of the ingestion half of scripts/frostt.py of the reference (`process_frostt_tensor` :51-66, `get_frostt_tensor`
:69-89): same arguments, same `.npz` cache next to the data file (keys `indices`, `entries`, `shape`).  The text is
parsed by the multi-threaded native parser `ttsk_tns_parse` (csrc/ttsk_tns.cu) instead of a per-line Python loop; the
result is the (d, nnz) int64 / (nnz,) float64 COO pair the sketching entry points stream to the GPU.  There is no
network here: `get_frostt_tensor` takes the data file where it lies and raises FileNotFoundError otherwise.
"""
from __future__ import annotations

import gzip
import os
from ctypes import byref, c_int
from typing import Optional, Tuple

import numpy as np

from tt_sketch import _backend as be
from tt_sketch.tensor import SparseTensor


def parse_tns(text: bytes, nnz: Optional[int] = None):
    """(indices (d, nnz) int64 0-based, entries (nnz,) float64, largest index per mode) of a .tns byte string."""
    lib = be.lib()
    d = c_int(0)
    n = int(lib.ttsk_tns_count(text, len(text), byref(d)))
    if n < 0:
        raise ValueError(lib.ttsk_last_error().decode())
    if nnz is not None and int(nnz) != n:
        raise ValueError(f"the file holds {n} nonzeros, {nnz} were announced")
    if n == 0:
        raise ValueError("no nonzeros in the .tns text")
    idx = np.empty((d.value, n), dtype=np.int64)
    val = np.empty(n, dtype=np.float64)
    mx = np.empty(d.value, dtype=np.int64)
    be.check(lib.ttsk_tns_parse(text, len(text), d.value, n, idx.ctypes.data, val.ctypes.data, mx.ctypes.data))
    return idx, val, mx


def process_frostt_tensor(filepath: str, nnz: Optional[int] = None, shape: Optional[Tuple[int, ...]] = None) -> SparseTensor:
    """Read a FROSTT file (gzip or plain text).  `shape` defaults to the largest coordinate per mode."""
    opener = gzip.open if filepath.endswith(".gz") else open
    with opener(filepath, "rb") as f:
        text = f.read()
    idx, val, mx = parse_tns(text, nnz)
    if shape is None:
        shape = tuple(int(m) + 1 for m in mx)
    shape = tuple(int(s) for s in shape)
    if len(shape) != idx.shape[0] or any(int(m) >= s for m, s in zip(mx, shape)):
        raise ValueError(f"coordinates up to {tuple(int(m) + 1 for m in mx)} do not fit the shape {shape}")
    return SparseTensor(shape, idx, val)


def get_frostt_tensor(file_url: str, nnz: Optional[int] = None, shape: Optional[Tuple[int, ...]] = None,
                      data_dir: str = "data") -> SparseTensor:
    """The tensor of `file_url` (only its file name is used) from `data_dir`: the `.npz` cache if present, else the
    `.tns(.gz)` file, which is parsed and cached like the reference does."""
    filename = file_url.split("/")[-1]
    filepath = os.path.join(data_dir, filename)
    npzpath = filepath.split(".gz")[0] + ".npz"
    if os.path.exists(npzpath):
        z = np.load(npzpath)
        return SparseTensor(tuple(int(s) for s in z["shape"]), z["indices"], z["entries"])
    if not os.path.exists(filepath):
        raise FileNotFoundError(f"{filepath} (no network here: place the FROSTT file there)")
    tensor = process_frostt_tensor(filepath, nnz, shape)
    np.savez_compressed(npzpath, indices=tensor.indices, entries=tensor.entries, shape=np.array(tensor.shape))
    return tensor
