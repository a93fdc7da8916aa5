"""Soak check of the sparse path: odd sizes, repeated runs; the device-resident call and the chunked host call must
agree (two different tilings of the same sum), and repeated runs must agree to summation-order noise."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tt-sketch_b200")): sys.path.insert(0, p)
import torch
from ctypes import byref
from tt_sketch import _backend as be
from tt_sketch.drm import SparseGaussianDRM
from tt_sketch.sketch_container import SketchContainer
from tt_sketch.sketch_dispatch import drm_descriptor

shape = (10000, 10000, 10000, 500)
rl, rr = (20,) * 3, (40,) * 3
left = SparseGaussianDRM(rl, shape=shape, transpose=False, seed=1)
right = SparseGaussianDRM(rr, shape=shape, transpose=True, seed=2)
ld, _ = drm_descriptor(left); rd, _ = drm_descriptor(right)
_, total = SketchContainer.layout(shape, rl, rr)
lib, ctx = be.lib(), be.ctx()
worst = 0.0
for nnz in (7_654_321, 12_345_678, 33_333_333):
    rng = np.random.default_rng(nnz)
    idx = np.stack([rng.integers(0, n, nnz) for n in shape]).astype(np.int64)
    val = rng.standard_normal(nnz)
    d_idx, d_val = torch.from_numpy(idx).cuda(), torch.from_numpy(val).cuda()
    outs = []
    for rep in range(3):
        out = torch.empty(total, dtype=torch.float64, device="cuda")
        be.check(lib.ttsk_sparse_sketch(ctx, 4, be.as_i64(shape), nnz, be.ptr(d_idx), d_idx.stride(0), be.ptr(d_val),
                                        byref(ld), byref(rd), be.ptr(out), 0, be.stream()))
        torch.cuda.synchronize()
        outs.append(out.cpu().numpy())
    h_idx, h_val = torch.from_numpy(idx).pin_memory(), torch.from_numpy(val).pin_memory()
    h_out = torch.empty(total, dtype=torch.float64).pin_memory()
    be.check(lib.ttsk_sparse_sketch_host(ctx, 4, be.as_i64(shape), nnz, h_idx.data_ptr(), h_idx.stride(0), h_val.data_ptr(),
                                         byref(ld), byref(rd), h_out.data_ptr(), 0))
    outs.append(h_out.numpy().copy())
    scale = np.max(np.abs(outs[0]))
    errs = [float(np.max(np.abs(o - outs[0])) / scale) for o in outs[1:]]
    worst = max(worst, max(errs))
    print(f"nnz={nnz}: max rel deviation between runs / tilings {errs}", flush=True)
assert worst < 1e-11, worst
print("soak ok", worst)
