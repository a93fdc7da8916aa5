#!/usr/bin/env python
"""bench.py -- stream_sketch throughput of the B200 sketching path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--nnz NNZ] [--impl reference] [--config C1|C2|C3|C4|C5]

Workload (BASELINE.json configs[3], SURVEY.md section 8d "C4"): one step = one streaming sketch
(all d Psi cores + d-1 Omega matrices) of a synthetic order-4 COO tensor, shape
(10^4, 10^4, 10^4, 500), nnz = 1e8 (uniform int64 indices, N(0,1) float64 values), with lazy
SparseGaussianDRMs, left rank 20 (seed 1) / right rank 40 (seed 2).  `value` = nonzeros per
second with the COO arrays resident in HBM; `e2e` = the same through the host-buffer C-ABI call
(pinned host COO -> chunked H2D overlapped with the kernels -> packed sketch D2H).
Multi-GPU (torchrun, one rank per GPU): the SAME 1e8 nonzeros are sharded into equal contiguous
ranges (strong scaling), every rank sketches its shard, one NCCL all-reduce of the packed
sketch (16.4 M doubles) combines them inside the timed region.

`--impl reference` times THE REFERENCE ITSELF (the unmodified package that oracle/build_ref.sh installs
under oracle/_ref/, git-ignored) on a bounded sample of the same workload on the host cores
(`cpu_baseline.kind` = "reference"); the same run, in a subprocess, is the `cpu_baseline` of the GPU arm.

`--config` selects another BASELINE.json configuration (default C4, the one the metric is quoted on):
C1 dense 20^5 stream_sketch, C2 TT orthogonal_sketch, C3 CP stream_sketch (all three at full size, one GPU
each: they do not shard -- N ranks run N replicas), C5 TensorSum(100 TT + sparse) blocked_stream_sketch with
TensorTrainDRMs (1.25e8 nonzeros PER GPU: weak scaling, 1e9 at 8 GPUs).  See bench_configs.py.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tt-sketch_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

SHAPE = (10000, 10000, 10000, 500)
RL, RR = (20, 20, 20), (40, 40, 40)
SEED_L, SEED_R = 1, 2
METRIC = "stream_sketch nnz/sec (sparse)"
UNIT = "nnz/s"
ALGO_BYTES_PER_NNZ = 8 * (len(SHAPE) + 1)  # int64 COO indices + fp64 value, read once


def make_coo(nnz, lo, hi):
    """Nonzeros [lo, hi) of the synthetic tensor: generated blockwise from per-block seeds so
    every rank can build exactly its shard of the same global tensor."""
    blk = 1 << 22
    idx = np.empty((len(SHAPE), hi - lo), dtype=np.int64)
    val = np.empty(hi - lo, dtype=np.float64)
    b0 = lo // blk
    pos = 0
    b = b0
    while pos < hi - lo:
        start = max(lo, b * blk)
        stop = min(hi, (b + 1) * blk)
        n_blk = min(blk, nnz - b * blk)
        rng = np.random.default_rng([1234, b])
        full_idx = [rng.integers(0, n, n_blk) for n in SHAPE]
        full_val = rng.standard_normal(n_blk)
        s0, s1 = start - b * blk, stop - b * blk
        for k in range(len(SHAPE)):
            idx[k, pos:pos + (s1 - s0)] = full_idx[k][s0:s1]
        val[pos:pos + (s1 - s0)] = full_val[s0:s1]
        pos += s1 - s0
        b += 1
    return idx, val


def _pass_profile(nnz):
    """Summary of the committed ncu --set full capture of this workload's mode-pass kernels
    (profiles/r02_pass_traffic_1e8nnz.json, written by tools/ncu_summary.py); None for other sizes."""
    path = os.path.join(ROOT, "profiles", "r02_pass_traffic_1e8nnz.json")
    try:
        d = json.load(open(path))
        if int(d["nnz"]) == int(nnz):
            return d
    except (OSError, ValueError, KeyError):
        pass
    return None


def measured_traffic(nnz):
    """dram__bytes_read.sum + dram__bytes_write.sum of the pass kernels, per launch."""
    d = _pass_profile(nnz)
    return float(d["dram_bytes_per_launch_avg"]) if d else None


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def reference_sketch_seconds(nnz_sample):
    """One stream_sketch of the first `nnz_sample` nonzeros of the workload by THE REFERENCE (oracle/_ref/pkg:
    unmodified tt_sketch package + its Cython generator), same shape / DRM kinds / ranks / seeds."""
    from oracle.ref_import import import_reference

    import_reference()
    from tt_sketch.drm import SparseGaussianDRM as RefGauss
    from tt_sketch.sketch import stream_sketch as ref_stream_sketch
    from tt_sketch.tensor import SparseTensor as RefSparse

    idx, val = make_coo(nnz_sample, 0, nnz_sample)
    X = RefSparse(SHAPE, idx, val)
    left = RefGauss(RL, shape=SHAPE, transpose=False, seed=SEED_L)
    right = RefGauss(RR, shape=SHAPE, transpose=True, seed=SEED_R)
    t0 = time.perf_counter()
    ref_stream_sketch(X, RL, RR, left_drm=left, right_drm=right)
    return time.perf_counter() - t0


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p.get("num_threads", 1) for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


def reference_sample_note(sample, cores):
    return (f"first {sample} nonzeros of the workload per step (the reference cannot run nnz=1e8: >=144 GB of "
            f"(r x nnz) intermediates, ~5.7 h, int overflow at fast_lazy_gaussian.pyx:95; it is linear in nnz); "
            f"the unmodified reference package from oracle/_ref (NumPy + its Cython generator); BLAS threads={cores}, "
            f"generator and mask loop single-threaded by construction")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_nnz
    times = []
    for i in range(args.warmup + args.steps):
        sec = reference_sketch_seconds(sample)
        if i >= args.warmup:
            times.append(sec)
    sec = float(np.mean(times))
    value = sample / sec
    cores = blas_threads()
    cfg = config_dict(args.nnz, args.gpus)
    cfg["workload"] += f"; THIS ARM times a {sample}-nonzero sample of it per step"
    cfg["sample_nnz"] = int(sample)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": reference_sample_note(sample, cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def cpu_baseline_subprocess(extra_args, timeout=900):
    """The reference arm in a subprocess (the reference's package is also called tt_sketch, so it cannot share
    a process with the product); returns its parsed JSON line."""
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"] + extra_args
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        env.pop(k, None)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
    if p.returncode != 0 or not lines:
        raise RuntimeError("reference arm failed: " + p.stderr[-2000:])
    return json.loads(lines[-1])


def config_dict(nnz, gpus):
    return {"workload": "C4: stream_sketch of sparse order-4 COO tensor, shape (1e4,1e4,1e4,500), "
                        f"nnz={nnz:.3g}, SparseGaussianDRM lazy DRMs, left rank 20 / right rank 40, float64",
            "nnz": int(nnz), "shape": list(SHAPE), "left_rank": list(RL), "right_rank": list(RR),
            "sharding": f"nnz split into {gpus} equal contiguous ranges, one NCCL all-reduce of the packed sketch",
            "l2": "inputs (4.0 GB COO) are larger than the 126 MB L2; no explicit flush",
            "drm_tables": "prefix tables of the Gaussian DRM (L_0, R_1, R_2; 1.6 GB, 2.5 ms to build) are DRM state: "
                          "built in the first warm-up step and reused, like the reference builds TT-DRM cores once"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--nnz", type=float, default=1e8)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--ref-nnz", type=int, default=100000, help="nonzeros of one --impl reference step (BASELINE.md section 3)")
    ap.add_argument("--cpu-nnz", type=int, default=50000, help="nonzeros of the cpu_baseline leg of the GPU arm")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--config", default="C4", choices=["C1", "C2", "C3", "C4", "C5"])
    args = ap.parse_args()
    args.nnz = int(args.nnz)
    if args.config != "C4":
        import bench_configs

        return bench_configs.main(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    json_out = sys.stdout
    if world > 1:
        # one JSON line only on stdout: NCCL / torch print a version banner on fd 1 from native code, so fd 1 is
        # pointed at stderr for the whole run and the line is written to a duplicate of the original stdout
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    from ctypes import byref, c_double

    from tt_sketch import _backend as be
    from tt_sketch.drm import SparseGaussianDRM
    from tt_sketch.sketch_container import SketchContainer
    from tt_sketch.sketch_dispatch import drm_descriptor

    nnz = args.nnz
    lo, hi = rank * nnz // world, (rank + 1) * nnz // world
    n_loc = hi - lo
    t0 = time.perf_counter()
    idx, val = make_coo(nnz, lo, hi)
    gen_s = time.perf_counter() - t0
    left = SparseGaussianDRM(RL, shape=SHAPE, transpose=False, seed=SEED_L)
    right = SparseGaussianDRM(RR, shape=SHAPE, transpose=True, seed=SEED_R)
    ld, _ = drm_descriptor(left)
    rd, _ = drm_descriptor(right)
    _, total = SketchContainer.layout(SHAPE, RL, RR)
    lib, ctx = be.lib(), be.ctx()
    shape_c = be.as_i64(SHAPE)

    # ---------------- device-resident arm (`value`)
    d_idx = torch.from_numpy(idx).cuda()
    d_val = torch.from_numpy(val).cuda()
    packed = torch.empty(total, dtype=torch.float64, device="cuda")

    def step_device():
        be.check(lib.ttsk_sparse_sketch(ctx, len(SHAPE), shape_c, n_loc, be.ptr(d_idx), d_idx.stride(0), be.ptr(d_val),
                                        byref(ld), byref(rd), be.ptr(packed), 0, be.stream()))
        if world > 1:
            dist.all_reduce(packed)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    sync_all()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = be.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pass_ms, total_ms = [], []
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    launches = be.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel timing of the dominant kernel (events recorded by the library on the same stream
    # around every pass launch of the LAST timed step)
    a, b = c_double(), c_double()
    be.check(lib.ttsk_last_kernel_ms(ctx, byref(a), byref(b)))
    step_kernels_ms, pass_kernels_ms = a.value, b.value
    from ctypes import c_int
    pm, pn = (c_double * 64)(), c_int()
    be.check(lib.ttsk_last_pass_ms(ctx, pm, 64, byref(pn)))
    per_pass_ms = [pm[i] for i in range(min(pn.value, 64))]
    t = torch.tensor([ms, float(launches)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, launches = float(tmax[0]), int(tsum[1])
    ms_per_step = ms / args.steps
    value = nnz / (ms_per_step * 1e-3)
    checksum = float(packed.sum().item())

    # ---------------- end-to-end arm (host COO buffers through the C ABI)
    e2e = None
    if not args.no_e2e:
        h_idx = torch.from_numpy(idx).pin_memory()
        h_val = torch.from_numpy(val).pin_memory()
        h_out = torch.empty(total, dtype=torch.float64).pin_memory()

        def step_e2e():
            if world == 1:
                be.check(lib.ttsk_sparse_sketch_host(ctx, len(SHAPE), shape_c, n_loc, h_idx.data_ptr(), h_idx.stride(0),
                                                     h_val.data_ptr(), byref(ld), byref(rd), h_out.data_ptr(), 0))
            else:
                be.check(lib.ttsk_sparse_sketch_stream(ctx, len(SHAPE), shape_c, n_loc, h_idx.data_ptr(),
                                                       h_idx.stride(0), h_val.data_ptr(), byref(ld), byref(rd),
                                                       be.ptr(packed), 0))
                dist.reduce(packed, dst=0)  # the result is needed on ONE rank: one reduce, one device->host copy
                if rank == 0:
                    h_out.copy_(packed, non_blocking=True)
                torch.cuda.synchronize()

        step_e2e()
        sync_all()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            step_e2e()
        sync_all()
        e_ms = (time.perf_counter() - t0) * 1e3 / n_e2e
        te = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te[0])
        a2, b2 = c_double(), c_double()
        be.check(lib.ttsk_last_kernel_ms(ctx, byref(a2), byref(b2)))
        e2e = {"value": nnz / (e_ms * 1e-3), "unit": UNIT, "ms_per_step": e_ms,
               "device_span_ms_last_step": a2.value, "pass_kernels_ms_last_step": b2.value,
               "h2d_bytes_per_step": int(nnz * ALGO_BYTES_PER_NNZ), "d2h_bytes_per_step": int(total * 8),
               "api": "ttsk_sparse_sketch_host (C ABI, pinned host COO in, packed sketch out)"
               if world == 1 else "ttsk_sparse_sketch_stream + NCCL reduce to rank 0 + D2H on rank 0",
               "checksum_matches_device_arm": bool(rank != 0 or abs(float(h_out.sum()) - checksum) <= 1e-6 * max(1.0, abs(checksum)))}

    if rank == 0:
        peak, which = measured_hbm_peak()
        n_pass = len(SHAPE)
        achieved = ALGO_BYTES_PER_NNZ * n_loc / (pass_kernels_ms * 1e-3) / 1e9 if pass_kernels_ms > 0 else 0.0
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(nnz, world),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(nnz) if world == 1 else None, "peak_source": which,
                         "kernel": "ttsk::gw_kernel (modes 0, 2, 3: fused lazy-Gaussian generator) / ttsk::gt_kernel (mode 1: table-row gather on the payload partition, bulk-copy staged), one launch per mode", "algorithmic_bytes_per_launch": ALGO_BYTES_PER_NNZ * n_loc / n_pass,
                         "launches_per_step": n_pass,
                         "note": "algorithmic 40 B/nnz over the d=4 mode passes (10 B/nnz per launch) / summed pass "
                                 "time of the last step; the kernel is FP64-issue-bound (bit-exact ndtri), see "
                                 "fp64_pipe and DESIGN.md section 3"},
            "fp64_pipe": fp64_model(n_loc, pass_kernels_ms, nnz),
            "kernel_ms": {"pass_kernels_last_step": pass_kernels_ms, "all_kernels_last_step": step_kernels_ms,
                          "per_pass_last_step": per_pass_ms},
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "host": {"coo_generation_s": gen_s, "cpu_count": os.cpu_count()},
            "checksum": checksum,
        }
        if world == 1 and not args.no_cpu:
            ref = cpu_baseline_subprocess(["--ref-nnz", str(args.cpu_nnz)])
            out["cpu_baseline"] = dict(ref["cpu_baseline"], ms=ref["ms_per_step"])
        print(json.dumps(out), file=json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# FP64-pipe peak of this pool's B200 on the generator's instruction mix (unfused DMUL + DADD chains), thread
# instructions per second: tools/fp64_mix_microbench (profiles/r02_fp64_mix_microbench.txt); plain DFMA: 1.70e13
FP64_PIPE_PEAK = 1.85e13


def fp64_model(n_loc, pass_ms, nnz_total):
    """FP64-pipe view of the same kernels: FP64-pipe instructions per nonzero MEASURED by ncu
    (smsp__inst_executed_pipe_fp64.sum of the four pass launches, profiles/r02_pass_traffic_1e8nnz.json; the
    instruction count per nonzero does not depend on the shard size) over the measured issue rate of the pipe."""
    d = _pass_profile(nnz_total)
    if d and d.get("fp64_pipe_thread_instructions_per_nnz"):
        per_nnz, src = float(d["fp64_pipe_thread_instructions_per_nnz"]), "ncu smsp__inst_executed_pipe_fp64.sum x 32 / nnz (profiles/r02_pass_traffic_1e8nnz.json)"
    else:  # other sizes: 80 on-the-fly variates x (0.73 x 38 central + 0.27 x 140 tail incl. the discarded central evaluation) + accumulation
        per_nnz, src = 80 * (0.73 * 38 + 0.27 * 140) + 1400.0, "model (no ncu capture at this size)"
    ach = per_nnz * n_loc / (pass_ms * 1e-3) if pass_ms > 0 else 0.0
    return {"fp64_instr_per_nnz": per_nnz, "fp64_instr_per_nnz_source": src, "achieved_instr_per_s": ach,
            "peak_instr_per_s": FP64_PIPE_PEAK, "frac": ach / FP64_PIPE_PEAK,
            "peak_source": "measured DMUL+DADD issue rate, tools/fp64_mix_microbench on this pool's B200 (1.85e13 thread-instr/s)"}


if __name__ == "__main__":
    main()
