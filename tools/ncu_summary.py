"""Summarise one frame's `ncu --set full` capture (tools/profile_frame.sh) into the files under profiles/:
  <tag>_ncu_full_summary.csv  selected metrics per kernel launch
  traffic.json                DRAM bytes (read + write) per stage
  pipes.json                  issue / FMA-pipe utilisation of the compute-bound kernels
Usage: python tools/ncu_summary.py gpurun_out/<tag>_frame.ncu-rep <tag>"""
import csv, io, json, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
METRICS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
col = {h: i for i, h in enumerate(hdr)}
ki = col["Kernel Name"]


def val(r, m):
    v = r[col[m]].replace(",", "")
    try:
        return float(v)
    except ValueError:
        return v


def to_bytes(r, m):
    u = units[col[m]].lower()
    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return val(r, m) * scale


def to_us(r):
    u = units[col["gpu__time_duration.sum"]].lower()
    return val(r, "gpu__time_duration.sum") * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1.0)


with open(f"profiles/{tag}_ncu_full_summary.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["id", "kernel"] + [f"{m} [{units[col[m]]}]" for m in METRICS])
    for i, r in enumerate(data):
        w.writerow([i, r[ki]] + [r[col[m]] for m in METRICS])


def stage_of(name):
    if "project_cull" in name or "compact_visible" in name:
        return "project"
    if "onesweep_pass_kernel<unsigned int" in name or "bucket_rank" in name or "bucket_scatter" in name or "bucket_local_sort" in name:
        return "depthSort"
    if "tile_chunk_count" in name or "tile_chunk_place" in name:
        return "tileSort"
    if "onesweep_pass_kernel<unsigned short" in name or "onesweep_pair_kernel" in name:
        return "tileSort"
    if "create_instances" in name:
        return "expand"
    if "tile_lower_bounds" in name:
        return "ranges"
    if "blend" in name:
        return "blend"
    return None


traffic = {"note": f"dram__bytes_read.sum + dram__bytes_write.sum per stage of one warm frame, ncu --set full (cold L2 under ncu), "
                   f"profiles/{tag}_ncu_full_summary.csv", "unit": "bytes"}
for r in data:
    st = stage_of(r[ki])
    if st:
        traffic[st] = traffic.get(st, 0) + int(to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum"))
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)

pipes = {"note": f"pipe utilisation of the compute-bound kernels from the same capture (profiles/{tag}_ncu_full_summary.csv)"}
for key, pat in (("project", "project_cull"), ("expand", "create_instances"), ("blend", "blend_mono")):
    for r in data:
        if pat in r[ki]:
            pipes[key] = {"kernel": r[ki], "issue_active_pct": val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                          "fma_pipe_cycles_active_pct": val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                          "xu_pipe_inst_pct": val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                          "duration_us_under_ncu": round(to_us(r), 3)}
            break
# every kernel of the frame with its stage and duration: bench.py splits a stage interval among its kernels by these shares
pipes["kernels_us"] = [{"stage": stage_of(r[ki]), "kernel": r[ki].split("(")[0], "us": round(to_us(r), 3)} for r in data if stage_of(r[ki])]
json.dump(pipes, open("profiles/pipes.json", "w"), indent=1)
print(json.dumps({"traffic": traffic, "pipes": pipes}, indent=1))
