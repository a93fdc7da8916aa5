#!/usr/bin/env python
"""Per-chunk, per-pass kernel times of the host-buffer sparse sketch (ttsk_sparse_sketch_host) next to the
device-resident call: where the chunking tax of the end-to-end path goes.
    python tools/e2e_passes.py [nnz]"""
import os
import sys
from ctypes import byref, c_double, c_int

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tt-sketch_b200")]
import bench  # noqa: E402
from tt_sketch import _backend as be  # noqa: E402
from tt_sketch.drm import SparseGaussianDRM  # noqa: E402
from tt_sketch.sketch_container import SketchContainer  # noqa: E402
from tt_sketch.sketch_dispatch import drm_descriptor  # noqa: E402

nnz = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
S = bench.SHAPE
idx, val = bench.make_coo(nnz, 0, nnz)
left = SparseGaussianDRM(bench.RL, shape=S, transpose=False, seed=1)
right = SparseGaussianDRM(bench.RR, shape=S, transpose=True, seed=2)
ld, _ = drm_descriptor(left)
rd, _ = drm_descriptor(right)
_, total = SketchContainer.layout(S, bench.RL, bench.RR)
lib, ctx = be.lib(), be.ctx()
h_idx = torch.from_numpy(idx).pin_memory()
h_val = torch.from_numpy(val).pin_memory()
h_out = torch.empty(total, dtype=torch.float64).pin_memory()


def passes():
    pm, pn = (c_double * 256)(), c_int()
    be.check(lib.ttsk_last_pass_ms(ctx, pm, 256, byref(pn)))
    a, b = c_double(), c_double()
    be.check(lib.ttsk_last_kernel_ms(ctx, byref(a), byref(b)))
    return [round(pm[i], 2) for i in range(min(pn.value, 256))], a.value, b.value


for rep in range(2):
    be.check(lib.ttsk_sparse_sketch_host(ctx, 4, be.as_i64(S), nnz, h_idx.data_ptr(), h_idx.stride(0), h_val.data_ptr(),
                                         byref(ld), byref(rd), h_out.data_ptr(), 0))
p, span, tot = passes()
print("host call: span %.1f ms, pass kernels %.1f ms, %d pass launches" % (span, tot, len(p)))
for c in range(0, len(p), 4):
    print("  chunk", c // 4, p[c:c + 4])
d_idx = torch.from_numpy(idx).cuda()
d_val = torch.from_numpy(val).cuda()
packed = torch.empty(total, dtype=torch.float64, device="cuda")
for rep in range(2):
    be.check(lib.ttsk_sparse_sketch(ctx, 4, be.as_i64(S), nnz, be.ptr(d_idx), d_idx.stride(0), be.ptr(d_val), byref(ld),
                                    byref(rd), be.ptr(packed), 0, be.stream()))
torch.cuda.synchronize()
p, span, tot = passes()
print("device call: span %.1f ms, pass kernels %.1f ms" % (span, tot), p)
