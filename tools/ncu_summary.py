#!/usr/bin/env python
"""Summarise an `ncu --set full` report of the mode-pass kernels into profiles/:
    python tools/ncu_summary.py gpurun_out/X.ncu-rep NNZ TAG
writes profiles/TAG_ncu_full_pass_kernels.csv (selected raw metrics per launch) and
profiles/r02_pass_traffic_1e8nnz.json (DRAM bytes and FP64-pipe instructions per launch, read by bench.py)."""
import csv
import io
import json
import os
import subprocess
import sys

rep, nnz, tag = sys.argv[1], int(float(sys.argv[2])), sys.argv[3]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__warps_active.avg.per_cycle_active",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
cols = [hdr.index(w) for w in want if w in hdr]
out = os.path.join(ROOT, "profiles", f"{tag}_ncu_full_pass_kernels.csv")
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow([hdr[c] for c in cols])
    w.writerow([units[c] for c in cols])
    for r in data:
        w.writerow([r[c] for c in cols])


def scale(v, u):
    v = float(v)
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)


kn, ti = hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
# FP64-pipe warp instructions: the --set full sections carry the pipe's utilisation, not the raw count; the pipe
# issues one warp instruction per two cycles and sub-partition, so count = pct / 100 * 0.5 * (active SM cycles * 4)
pi = hdr.index("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
ci = hdr.index("TPC.TriageCompute.sm__cycles_active.avg") if "TPC.TriageCompute.sm__cycles_active.avg" in hdr else hdr.index("sm__cycles_active.avg")
gi = hdr.index("launch__grid_size")
N_SM = 148
ai = hdr.index("smsp__inst_executed.sum")
kernels = []
for r in data:
    kernels.append({"kernel": r[kn], "ms": float(r[ti]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(units[ti], 1.0),
                    "dram_bytes": scale(r[ri], units[ri]) + scale(r[wi], units[wi]),
                    "warp_instructions": float(r[ai]),
                    "fp64_pipe_pct_of_peak": float(r[pi]),
                    "fp64_pipe_warp_instructions": float(r[pi]) / 100.0 * 0.5 * float(r[ci]) * 4 * min(N_SM, int(float(r[gi])))})
tot = sum(k["dram_bytes"] for k in kernels)
fp64 = sum(k["fp64_pipe_warp_instructions"] or 0.0 for k in kernels)
json.dump({"nnz": nnz, "source": f"ncu --set full, profiles/{os.path.basename(out)}", "kernels": kernels,
           "dram_bytes_all_pass_launches": tot, "dram_bytes_per_launch_avg": tot / max(1, len(kernels)),
           "fp64_pipe_thread_instructions_per_nnz": 32.0 * fp64 / nnz},
          open(os.path.join(ROOT, "profiles", "r02_pass_traffic_1e8nnz.json"), "w"), indent=1)
print(out, len(kernels), "launches", tot / 1e9, "GB")
