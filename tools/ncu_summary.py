"""Summarise an .ncu-rep (read here with `ncu -i … --page raw --csv`) into the small JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep "how it was captured" > profiles/r01_ncu_full_x.json
"""
import csv, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]


def main():
    rep, how = sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    kernels = []
    for r in rows[2:]:
        d = dict(zip(head, r))
        k = {"name": d.get("Kernel Name", "?")}
        for key in KEYS:
            if key in d and d[key] != "":
                k[key] = "%s %s" % (d[key], units[head.index(key)])
        kernels.append(k)
    json.dump({"source": how, "kernels": kernels}, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
