"""TT-DRM on sparse input at 2e7 nonzeros (two calls; profile the second under ncu)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tt-sketch_b200")): sys.path.insert(0, p)
import torch
from tt_sketch.drm import TensorTrainDRM
from tt_sketch.sketch import stream_sketch
from tt_sketch.tensor import SparseTensor
shape = (10000, 10000, 10000, 500)
lr, rr = (20,) * 3, (40,) * 3
L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
nnz = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
idx = np.stack([np.random.default_rng(200 + k).integers(0, n, nnz) for k, n in enumerate(shape)]).astype(np.int64)
sp = SparseTensor(shape, idx, np.random.default_rng(99).standard_normal(nnz))
for _ in range(2):
    t0 = time.perf_counter(); stream_sketch(sp, lr, rr, left_drm=L, right_drm=R); torch.cuda.synchronize()
    print(f"nnz={nnz}: {(time.perf_counter() - t0) * 1e3:.2f} ms")
