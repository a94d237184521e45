"""Summarise an .ncu-rep (ncu --set full) into the handful of numbers the roofline tables use.
  python tools/ncu_summary.py report.ncu-rep [kernel-substring] > profiles/….txt
Reads the report with `ncu -i … --page raw --csv` (works without a GPU)."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->L1/SMEM bytes"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor-pipe instructions"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) active %"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait / issue"),
    ("smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio", "stall sleeping / issue"),
    ("sm__inst_executed_pipe_xu.sum", "XU (MUFU) instructions"),
    ("smsp__cycles_elapsed.avg.per_second", "SM clock"),
]


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}  (ncu --set full --clock-control none; one row per captured launch)")
    seen = {}
    for r in rows[2:]:
        seen[r[col["Kernel Name"]]] = seen.get(r[col["Kernel Name"]], 0) + 1
    done = set()
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if (pat and pat not in name) or name in done:
            continue
        done.add(name)        # first captured launch of each kernel; the others repeat it
        print(f"\n== {name[:150]}   ({seen[name]} launches captured, first shown)")
        for key, label in WANT:
            if key in col and r[col[key]] != "":
                print(f"  {label:36s} {r[col[key]]:>18s} {units[col[key]]}")


if __name__ == "__main__":
    main()
