#!/usr/bin/env python
"""Wall time + parity of BASELINE.json configs 1-3 (and a scaled config 5) on one B200 through
the tt_sketch API, next to the CPU oracle port of the reference on the same box.  These are the
parity-test cases of bench.py's contract, not bench lines; results go to profiles/.

    python tools/bench_configs.py > profiles/r01_configs.json
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tt-sketch_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

from oracle import sketch_oracle as orc  # noqa: E402
from tt_sketch import _backend as be  # noqa: E402
from tt_sketch.drm import TensorTrainDRM  # noqa: E402
from tt_sketch.sketch import blocked_stream_sketch, orthogonal_sketch, stream_sketch  # noqa: E402
from tt_sketch.tensor import CPTensor, DenseTensor, SparseTensor, TensorSum, TensorTrain  # noqa: E402


def odrm(d):
    return orc.Drm("tt", d.transpose, d.shape, d.bond_rank_min, d.bond_rank_max, d.seed, list(d.cores))


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best, out


def rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def main():
    res = {"gpu": torch.cuda.get_device_name(0), "host_cpus": os.cpu_count()}
    # ---- C1: dense 20^5, TT-DRM rL=10 / rR=15, stream_sketch
    shape = (20,) * 5
    X = DenseTensor(np.random.default_rng(0).standard_normal(shape))
    lr, rr = (10,) * 4, (15,) * 4
    L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    l0 = be.launch_count()
    t_gpu, stt = timed(lambda: stream_sketch(X, lr, rr, left_drm=L, right_drm=R))
    t0 = time.perf_counter(); Psi, Om = orc.general_sketch(("dense", X.data), odrm(L), odrm(R), "streaming"); t_cpu = time.perf_counter() - t0
    err = max(rel(a, b) for a, b in zip(stt.Psi_cores + stt.Omega_mats, Psi + Om))
    res["C1_dense_20^5_stream"] = {"gpu_s": t_gpu, "cpu_oracle_s": t_cpu, "entries_per_s_gpu": 20**5 / t_gpu,
                                   "max_rel_err_vs_oracle": err, "launches_per_sketch": (be.launch_count() - l0) // 6}
    # ---- C2: TT d=10 n=50 r=100 -> orthogonal_sketch rL=20 rR=40
    shape = (50,) * 10
    T = TensorTrain.random(shape, 100, seed=2)
    lr, rr = (20,) * 9, (40,) * 9
    L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    t_gpu, tt = timed(lambda: orthogonal_sketch(T, lr, rr, left_drm=L, right_drm=R))
    t0 = time.perf_counter(); Psi, _ = orc.general_sketch(("tt", T.cores), odrm(L), odrm(R), "orthogonal"); t_cpu = time.perf_counter() - t0
    # Householder QR with LAPACK's sign convention makes the orthogonal cores comparable element-wise
    err = max(rel(a, b) for a, b in zip(tt.cores, Psi))
    res["C2_tt_orthogonal"] = {"gpu_s": t_gpu, "cpu_oracle_s": t_cpu, "max_rel_err_cores_vs_oracle": err}
    t_gpu, _ = timed(lambda: stream_sketch(T, lr, rr, left_drm=L, right_drm=R))
    res["C2_tt_stream"] = {"gpu_s": t_gpu}
    # ---- C3: CP d=8 n=100 R=200 -> stream_sketch rL=30 rR=60
    shape = (100,) * 8
    C = CPTensor.random(shape, 200, seed=3)
    lr, rr = (30,) * 7, (60,) * 7
    L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    t_gpu, stt = timed(lambda: stream_sketch(C, lr, rr, left_drm=L, right_drm=R))
    t0 = time.perf_counter(); Psi, Om = orc.general_sketch(("cp", C.cores), odrm(L), odrm(R), "streaming"); t_cpu = time.perf_counter() - t0
    err = max(rel(a, b) for a, b in zip(stt.Psi_cores + stt.Omega_mats, Psi + Om))
    res["C3_cp_stream"] = {"gpu_s": t_gpu, "cpu_oracle_s": t_cpu, "max_rel_err_vs_oracle": err}
    # ---- C5 scaled: TensorSum(20 TT rank 10 + sparse 2e6 nnz), TT-DRMs 20/40, blocked 2x2
    shape = (10000, 10000, 10000, 500)
    nnz = 2_000_000
    idx = np.stack([np.random.default_rng(200 + k).integers(0, n, nnz) for k, n in enumerate(shape)]).astype(np.int64)
    sp = SparseTensor(shape, idx, np.random.default_rng(99).standard_normal(nnz))
    tts = [TensorTrain.random(shape, 10, seed=1000 + k) for k in range(20)]
    S = TensorSum(tts + [sp])
    lr, rr = (20,) * 3, (40,) * 3
    L = TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    ls, rs = [(0,) * 3, (10,) * 3, (20,) * 3], [(0,) * 3, (20,) * 3, (40,) * 3]
    t_blk, blk = timed(lambda: blocked_stream_sketch(S, L, R, ls, rs), reps=2)
    t_full, full = timed(lambda: stream_sketch(S, lr, rr, left_drm=L, right_drm=R), reps=2)
    err = max(rel(a, b) for a, b in zip(blk.Psi_cores + blk.Omega_mats, full.Psi_cores + full.Omega_mats))
    res["C5_scaled_sum_20tt_plus_2e6nnz_ttdrm"] = {"blocked_2x2_gpu_s": t_blk, "unblocked_gpu_s": t_full,
                                                   "blocked_vs_unblocked_max_rel": err,
                                                   "sparse_nnz_per_s_unblocked": nnz / t_full}
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
