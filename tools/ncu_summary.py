"""Print the metrics we quote from an .ncu-rep (ncu -i ... --page raw --csv): time, DRAM bytes, pipe utilisation, stalls."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "lts__t_sector_hit_rate.pct",
]
STALL = "smsp__average_warps_issue_stalled_"


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    head, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(head, r))
        print("==", d.get("Kernel Name", "?")[:90])
        for w in WANT:
            if w in d:
                print(f"  {w:75s} {d[w]:>16s} {units[head.index(w)]}")
        st = sorted(((float(v.replace(',', '')), k) for k, v in d.items()
                     if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and v not in ("", "n/a")), reverse=True)
        print("  stalls per issue:", ", ".join(f"{k[len(STALL):-len('_per_issue_active.ratio')]} {v:.2f}" for v, k in st[:7]))


if __name__ == "__main__":
    for p in sys.argv[1:]:
        main(p)
