"""Summarise .ncu-rep captures (raw page) into a small text/JSON file for profiles/."""
import csv, json, subprocess, sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
        "sm__ops_path_tensor_src_fp64.sum.per_second", "sm__ops_path_tensor_src_fp64.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second"]

def summarise(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    idx = {n: i for i, n in enumerate(h)}
    out = []
    for r in rows[2:]:
        d = {"kernel": r[idx["Kernel Name"]], "grid": r[idx["Grid Size"]], "block": r[idx["Block Size"]]}
        for w in WANT:
            if w in idx:
                d[w] = "%s %s" % (r[idx[w]], units[idx[w]])
        for name, i in idx.items():   # every tensor-pipe / TMEM / cache-throughput metric of the capture
            if name not in d and any(t in name for t in ("pipe_tensor", "tmem", "lts__throughput", "l1tex__throughput",
                                                          "sm__inst_executed_pipe_uniform", "shared_op")):
                d[name] = "%s %s" % (r[i], units[i])
        out.append(d)
    return out

if __name__ == "__main__":
    res = {p: summarise(p) for p in sys.argv[1:]}
    print(json.dumps(res, indent=1))
