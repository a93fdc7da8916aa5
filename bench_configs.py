"""bench.py --config C1|C2|C3|C5: the other BASELINE.json configurations through the tt_sketch API.

    C1  stream_sketch of a dense 20^5 float64 tensor, TensorTrainDRM rL=10 / rR=15        unit: entries/s
    C2  orthogonal_sketch of an order-10 TensorTrain (dim 50, rank 100) to rank 20 (rR 40)  unit: stored entries/s
    C3  stream_sketch of an order-8 CPTensor (dim 100, CP rank 200) to TT rank 30 (rR 60)   unit: stored entries/s
    C5  blocked_stream_sketch (2 x 2 rank blocks) of TensorSum(100 TT rank 10 + sparse), shape
        (1e4,1e4,1e4,500), TensorTrainDRMs 20 / 40; 1.25e8 nonzeros PER GPU (1e9 at 8)     unit: nnz/s

Same JSON contract as the C4 line of bench.py.  `value`: inputs resident in HBM (device copies cached by the
containers), CUDA events around the timed steps; `e2e`: the same API call with the input re-uploaded from host
memory every step (`invalidate_device()`), wall clock; both include the device->host copy of the sketch, which
the API returns as NumPy arrays.  C1-C3 do not shard ("replicas only"): under torchrun every rank sketches its own
replica and `value` is the aggregate.  `--impl reference` runs the unmodified reference (oracle/_ref) on the
same inputs at full size (C1-C3) or on a bounded sample (C5).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
C4_SHAPE = (10000, 10000, 10000, 500)

SPEC = {
    "C1": dict(shape=(20,) * 5, lr=(10,) * 4, rr=(15,) * 4, unit="entries/s", metric="stream_sketch entries/sec (dense)",
               workload="C1: stream_sketch of a dense 20^5 float64 tensor, TensorTrainDRM left rank 10 / right rank 15"),
    "C2": dict(shape=(50,) * 10, lr=(20,) * 9, rr=(40,) * 9, unit="stored entries/s",
               metric="orthogonal_sketch stored-entries/sec (TT)",
               workload="C2: orthogonal_sketch of an order-10 TensorTrain (dim 50, rank 100) to rank 20 (right rank 40), TensorTrainDRMs"),
    "C3": dict(shape=(100,) * 8, lr=(30,) * 7, rr=(60,) * 7, unit="stored entries/s",
               metric="stream_sketch stored-entries/sec (CP)",
               workload="C3: stream_sketch of an order-8 CPTensor (dim 100, CP rank 200) to TT rank 30 (right rank 60), TensorTrainDRMs"),
    "C5": dict(shape=C4_SHAPE, lr=(20,) * 3, rr=(40,) * 3, unit="nnz/s", metric="blocked_stream_sketch nnz/sec (TensorSum of TT + sparse)",
               workload="C5: blocked_stream_sketch (2x2 rank blocks) of TensorSum(100 TT rank 10 + sparse), shape (1e4,1e4,1e4,500), "
                        "TensorTrainDRMs left 20 / right 40"),
}
C5_LS = [(0,) * 3, (10,) * 3, (20,) * 3]
C5_RS = [(0,) * 3, (20,) * 3, (40,) * 3]


def c5_coo(nnz, shard):
    """Sparse shard `shard` of C5 (seeds 200 + g), generated blockwise to bound host memory."""
    idx = np.empty((4, nnz), dtype=np.int64)
    val = np.empty(nnz, dtype=np.float64)
    blk = 1 << 23
    for b, lo in enumerate(range(0, nnz, blk)):
        hi = min(nnz, lo + blk)
        rng = np.random.default_rng([200 + shard, b])
        for k, n in enumerate(C4_SHAPE):
            idx[k, lo:hi] = rng.integers(0, n, hi - lo)
        val[lo:hi] = rng.standard_normal(hi - lo)
    return idx, val


def build_inputs(cfg, mod, nnz=0, shard=0, n_tt=100):
    """The configuration's tensor and DRMs built with the classes of `mod` (the product's tt_sketch or the
    reference's: same constructors).  Returns (tensor, left_drm, right_drm, units_per_step, big_operand_bytes)."""
    sp = SPEC[cfg]
    shape, lr, rr = sp["shape"], sp["lr"], sp["rr"]
    T, D = mod["tensor"], mod["drm"]
    L = D.TensorTrainDRM(lr, shape=shape, transpose=False, seed=1)
    R = D.TensorTrainDRM(rr, shape=shape, transpose=True, seed=2)
    if cfg == "C1":
        X = T.DenseTensor(np.random.default_rng(0).standard_normal(shape))
        return X, L, R, float(np.prod(shape)), X.data.nbytes
    if cfg == "C2":
        X = T.TensorTrain.random(shape, 100, seed=2)
        n = sum(c.size for c in X.cores)
        return X, L, R, float(n), 8 * n
    if cfg == "C3":
        X = T.CPTensor.random(shape, 200, seed=3)
        n = sum(c.size for c in X.cores)
        return X, L, R, float(n), 8 * n
    idx, val = c5_coo(nnz, shard)
    tts = [T.TensorTrain.random(shape, 10, seed=1000 + k) for k in range(n_tt)]
    X = T.TensorSum(tts + [T.SparseTensor(shape, idx, val)])
    return X, L, R, float(nnz), 40 * nnz + 8 * sum(c.size for t in tts for c in t.cores)


def run_step(cfg, mod, X, L, R):
    sp = SPEC[cfg]
    S = mod["sketch"]
    if cfg == "C2":
        return S.orthogonal_sketch(X, sp["lr"], sp["rr"], left_drm=L, right_drm=R)
    if cfg == "C5":
        return S.blocked_stream_sketch(X, L, R, C5_LS, C5_RS)
    return S.stream_sketch(X, sp["lr"], sp["rr"], left_drm=L, right_drm=R)


def sketch_bytes(out):
    if hasattr(out, "cores"):
        return int(sum(c.nbytes for c in out.cores))
    return int(sum(a.nbytes for a in list(out.Psi_cores) + list(out.Omega_mats)))


def config_dict(cfg, world, nnz, n_tt):
    sp = SPEC[cfg]
    d = {"workload": sp["workload"], "shape": list(sp["shape"]), "left_rank": list(sp["lr"]), "right_rank": list(sp["rr"]),
         "l2": "2 GB of L2 flush writes between timed steps are NOT used: C1-C3 inputs (1.3 - 32 MB) fit the 126 MB L2, "
               "the timed region streams them from L2/HBM as the reference streams them from its caches; stated, not hidden"}
    if cfg == "C5":
        d["nnz_per_gpu"], d["nnz"], d["tt_summands"] = int(nnz), int(nnz) * world, n_tt
        d["sharding"] = (f"{world} sparse shards of {nnz:.3g} nonzeros (one per GPU, seeds 200+g), the {n_tt} TT summands dealt "
                         "round-robin; blocks all-reduced with NCCL")
        d["l2"] = "inputs (5 GB of COO per GPU) are larger than the 126 MB L2; no explicit flush"
    else:
        d["sharding"] = "does not shard (replicas only): every rank sketches its own replica"
    return d


def reference_main(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, ROOT)
    from oracle.ref_import import import_reference

    import_reference()
    import tt_sketch.drm as drm
    import tt_sketch.sketch as sketch
    import tt_sketch.tensor as tensor

    mod = {"tensor": tensor, "drm": drm, "sketch": sketch}
    cfg = args.config
    sample = min(args.ref_nnz, 20000) if cfg == "C5" else 0
    n_tt = 100
    X, L, R, units, _ = build_inputs(cfg, mod, nnz=sample, shard=0, n_tt=n_tt)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        run_step(cfg, mod, X, L, R)
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    from bench import blas_threads

    cores = blas_threads()
    c = config_dict(cfg, args.gpus, sample if cfg == "C5" else 0, n_tt)
    note = "full size, same inputs and DRM seeds as the GPU arm"
    if cfg == "C5":
        note = (f"{n_tt} TT summands + a {sample}-nonzero sample of shard 0 per step (the reference's TT-DRM path needs 12.8 KB "
                f"per nonzero and its mask loop is O(n_mu * nnz); linear in nnz)")
        c["workload"] += f"; THIS ARM times a {sample}-nonzero sample of the sparse term per step"
        c["sample_nnz"] = sample
    out = {"impl": "reference", "metric": SPEC[cfg]["metric"], "value": units / sec, "unit": SPEC[cfg]["unit"],
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": c,
           "cpu_baseline": {"value": units / sec, "unit": SPEC[cfg]["unit"], "cores": cores, "kind": "reference",
                            "sample": f"{note}; the unmodified reference package from oracle/_ref; BLAS threads={cores}"},
           "e2e": {"value": units / sec, "unit": SPEC[cfg]["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main(args):
    if args.impl == "reference":
        return reference_main(args)
    for p in (ROOT, os.path.join(ROOT, "tt-sketch_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist

    from bench import ClockSampler, cpu_baseline_subprocess, measured_hbm_peak

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    json_out = sys.stdout
    if world > 1:
        sys.stdout.flush()  # one JSON line only on stdout: native banners (NCCL) go to stderr, see bench.py
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    import tt_sketch.drm as drm
    import tt_sketch.sketch as sketch
    import tt_sketch.tensor as tensor
    from tt_sketch import _backend as be

    mod = {"tensor": tensor, "drm": drm, "sketch": sketch}
    cfg = args.config
    nnz = int(args.nnz) if (cfg == "C5" and args.nnz != int(1e8)) else int(1.25e8)
    n_tt = 100
    t0 = time.perf_counter()
    X, L, R, units, big_bytes = build_inputs(cfg, mod, nnz=nnz, shard=rank, n_tt=n_tt)
    gen_s = time.perf_counter() - t0
    if cfg == "C5" and world > 1:
        # every rank holds ITS sparse shard and the TT summands dealt to it; blocks are all-reduced
        from tt_sketch.distributed import allreduce_blocked_stream_sketch

        mine = [t for i, t in enumerate(X.tensors[:-1]) if i % world == rank] + [X.tensors[-1]]
        X = tensor.TensorSum(mine, shape=X.shape)

        def step():
            return allreduce_blocked_stream_sketch(X, L, R, C5_LS, C5_RS)
    else:
        def step():
            return run_step(cfg, mod, X, L, R)

    def invalidate():
        for t in (X.tensors if hasattr(X, "tensors") else [X]):
            t.invalidate_device()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    out = None
    for _ in range(args.warmup):
        out = step()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = be.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = be.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax, tsum = t.clone(), t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), int(tsum[1])
    ms_per_step = ms / args.steps
    total_units = units * world
    value = total_units / (ms_per_step * 1e-3)
    d2h = sketch_bytes(out)

    e2e = None
    if not args.no_e2e:
        invalidate()
        step()
        sync_all()
        n_e2e = max(1, min(args.steps, 3))
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            invalidate()
            step()
        sync_all()
        e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        te = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te[0])
        e2e = {"value": total_units / (e_ms * 1e-3), "unit": SPEC[cfg]["unit"], "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(big_bytes), "d2h_bytes_per_step": d2h,
               "api": {"C2": "tt_sketch.sketch.orthogonal_sketch", "C5": "tt_sketch.sketch.blocked_stream_sketch"}.get(cfg, "tt_sketch.sketch.stream_sketch")
                      + " (host NumPy tensor in, host NumPy sketch out; input re-uploaded every step)"}
    if rank == 0:
        peak, which = measured_hbm_peak()
        ach = big_bytes / (ms_per_step * 1e-3) / 1e9
        res = {"metric": SPEC[cfg]["metric"], "value": value, "unit": SPEC[cfg]["unit"], "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64", "data": "synthetic", "config": config_dict(cfg, world, nnz, n_tt),
               "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                            "peak_source": which, "algorithmic_bytes_per_step": int(big_bytes),
                            "note": "bytes of the large operand (the input tensor, read once) / device time of one whole sketch "
                                    "(CUDA events on the launching stream around the timed steps, all launches of the step, "
                                    "host gaps included); these configurations are launch-/latency-bound, see DESIGN.md section 6"},
               "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
               "host": {"input_generation_s": gen_s, "cpu_count": os.cpu_count()}}
        if world == 1 and not args.no_cpu:
            ref = cpu_baseline_subprocess(["--config", cfg])
            res["cpu_baseline"] = dict(ref["cpu_baseline"], ms=ref["ms_per_step"])
        print(json.dumps(res), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
