"""Summarise an .ncu-rep (ncu --set full) into a small JSON + text table for profiles/ (run here, no GPU needed).
Usage: python scripts/ncu_summary.py gpurun_out/X.ncu-rep profiles/NAME"""
import csv
import io
import json
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_membar", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_lg_throttle", "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_sleeping",
    "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_dispatch_stall",
    "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_no_instructions",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        rec = {"kernel": r[hdr.index("Kernel Name")], "id": r[0]}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                try:
                    rec[w] = float(r[i].replace(",", ""))
                except ValueError:
                    rec[w] = r[i]
                rec[w + "__unit"] = units[i]
        res.append(rec)
    with open(out + ".json", "w") as fh:
        json.dump(res, fh, indent=1)
    with open(out + ".txt", "w") as fh:
        for rec in res:
            fh.write(f"== launch {rec['id']}: {rec['kernel']}\n")
            for w in WANT:
                if w in rec:
                    fh.write(f"   {w:75s} {rec[w]!s:>20s} {rec[w + '__unit']}\n")
    print(open(out + ".txt").read())


if __name__ == "__main__":
    main()
