"""Summarise an `ncu --set full` report of one bench step: per-launch DRAM bytes, throughput,
occupancy and pipe utilisation of the library's own kernels, plus the traffic table bench.py reads.

    ncu -i gpurun_out/prof.ncu-rep --page raw --csv > raw.csv
    python scripts/ncu_summary.py raw.csv profiles/r01_ncu_full_cfg2_summary.txt profiles/traffic.json cfg2

Only the LAST `npass` library kernels of the capture are kept (the final step of the run), in
launch order, named "<index>:<MODE>" like bench.py's plan description.
"""
import csv
import json
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    # where the warps wait (per issued instruction)
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
]
MODES = {"0": "FWD", "1": "MID", "2": "INV"}


def to_bytes(val, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    return float(val.replace(",", "")) * scale.get(unit, 1)


def main(raw_csv, out_txt, traffic_json, workload, header=""):
    with open(raw_csv, newline="") as f:
        rows = list(csv.reader(f))
    # the raw page has a header row, a units row, then one row per kernel launch
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, body = rows[start], rows[start + 1], rows[start + 2:]
    col = {n: i for i, n in enumerate(names)}
    mine = [r for r in body if len(r) > col["Kernel Name"] and
            ("fast_pass_kernel" in r[col["Kernel Name"]] or "pass_kernel" in r[col["Kernel Name"]]
             or "downsample" in r[col["Kernel Name"]])]
    # one step = from a first-pass kernel to the next first-pass kernel: keep the last step
    firsts = [i for i, r in enumerate(mine) if "_pass_kernel<0" in r[col["Kernel Name"]].replace("(int)", "")
              and i + 1 < len(mine)]
    # a step starts at a forward pass that does not follow another forward pass
    starts = [i for i in firsts if i == 0 or
              "_pass_kernel<0" not in mine[i - 1][col["Kernel Name"]].replace("(int)", "")]
    step_len = None
    for a, b in zip(starts, starts[1:]):
        step_len = b - a
    if step_len is None:
        step_len = len(mine)
    last = mine[-step_len:]
    traffic = {}
    lines = [header or f"ncu --set full --clock-control none, one step of {workload}",
             "units: " + ", ".join(f"{k} [{units[col[k]]}]" for k in ["Kernel Name", "Grid Size",
                                                                     "Block Size"] + KEEP if k in col),
             ""]
    for idx, r in enumerate(last):
        kn = r[col["Kernel Name"]].replace("(int)", "").replace("(bool)", "").replace("pbk::", "")
        mode = "DS"
        if "pass_kernel<" in kn:
            mode = MODES.get(kn.split("pass_kernel<")[1].split(",")[0].strip(), "?")
        key = f"{idx}:{mode}"
        lines.append(f"== pass {key}")
        for k in ["Kernel Name", "Grid Size", "Block Size"] + KEEP:
            if k in col:
                v = kn if k == "Kernel Name" else r[col[k]]
                lines.append(f"   {k:<84s}  {v}")
        rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
        wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        traffic[key] = int(rd + wr)
    with open(out_txt, "w") as f:
        f.write("\n".join(lines) + "\n")
    try:
        with open(traffic_json) as f:
            tj = json.load(f)
    except (OSError, ValueError):
        tj = {}
    tj[workload] = traffic
    tj["_source"] = ("profiles/*_ncu_full_*_summary.txt (dram__bytes_read.sum + "
                     "dram__bytes_write.sum per launch, scripts/ncu_summary.py)")
    with open(traffic_json, "w") as f:
        json.dump(tj, f, indent=1)
        f.write("\n")
    print(json.dumps(traffic))


if __name__ == "__main__":
    main(*sys.argv[1:6])
