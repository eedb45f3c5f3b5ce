"""Dev tool: condense `ncu --set full` reports into the text tables kept under profiles/.

usage: python tools/ncu_summary.py [--mean] out.txt "header line" rep1.ncu-rep [rep2.ncu-rep ...]
One column per captured launch (or the mean over each report's launches with --mean)."""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
]
STALL = "smsp__average_warps_issue_stalled_"


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit)
    return v * scale if scale else v


def load(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"_name": r[head.index("Kernel Name")]}
        for i, k in enumerate(head):
            short = k.split("TriageCompute.")[-1]
            if short in METRICS or short.startswith(STALL):
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                if units[i].endswith("byte"):
                    v, u = to_bytes(v, units[i]) / 1e6, "MB"
                elif units[i] in ("us", "ns", "ms", "s") and short == "gpu__time_duration.sum":
                    v, u = v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}[units[i]], "ms"
                else:
                    u = units[i]
                d[short] = (v, u)
        launches.append(d)
    return launches


def main():
    args = sys.argv[1:]
    mean = "--mean" in args
    args = [a for a in args if a != "--mean"]
    out_path, header, reps = args[0], args[1], args[2:]
    cols = []
    for rep in reps:
        ls = load(rep)
        if mean and ls:
            keys = set().union(*[set(l) for l in ls]) - {"_name"}
            m = {"_name": f"{ls[0]['_name'][:22]} (mean of {len(ls)})"}
            for k in keys:
                vals = [l[k][0] for l in ls if k in l]
                m[k] = (sum(vals) / len(vals), ls[0][k][1] if k in ls[0] else "")
            cols.append(m)
        else:
            cols.extend(ls)
    keys = [k for k in METRICS if any(k in c for c in cols)]
    stalls = sorted({k for c in cols for k in c if k.startswith(STALL) and k.endswith("_per_warp_active.pct") is False and k.endswith(".ratio")})
    lines = [header, "metric".ljust(64) + "unit".ljust(10) + " | ".join(c["_name"][:28].rjust(28) for c in cols)]
    for k in keys + stalls:
        unit = next((c[k][1] for c in cols if k in c), "")
        name = k.replace(STALL, "stall_").replace("_per_warp_active.pct", "").replace(".ratio", "")
        lines.append(name.ljust(64) + unit[:9].ljust(10) + " | ".join((f"{c[k][0]:.6g}" if k in c else "-").rjust(28) for c in cols))
    open(out_path, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
