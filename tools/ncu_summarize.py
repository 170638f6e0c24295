"""Turn gpurun_out/*.ncu-rep captures into small text summaries for profiles/ (run in the authoring container).
usage: python tools/ncu_summarize.py <rep> <out.txt> [title]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    with open(out, "w") as f:
        f.write(f"# {title}\n# source: ncu --set full --clock-control none (one launch; cold-cache, serialised)\n")
        f.write(f"kernel: {d.get('Kernel Name', ('?', ''))[0]}\n")
        for k in KEYS:
            if k in d:
                f.write(f"{k} = {d[k][0]} {d[k][1]}\n")
    print(open(out).read())
    if len(sys.argv) > 4:      # <traffic.json> <key>: record DRAM bytes per launch for bench.py's roofline.traffic
        import json
        import os
        tj, key = sys.argv[4], sys.argv[5]
        def to_bytes(k):
            v, u = d[k]
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        cur = json.load(open(tj)) if os.path.exists(tj) else {}
        cur[key] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
        json.dump(cur, open(tj, "w"), indent=1)
        print("traffic", key, cur[key])


if __name__ == "__main__":
    main()
