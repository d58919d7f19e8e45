"""DRAM traffic of the solver launches from an `ncu --set full` capture → JSON that bench.py quotes as `roofline.traffic`.

    python scripts/ncu_traffic.py gpurun_out/X.ncu-rep out.json family d nsims

Run where `ncu` exists (the GPU box, right after the capture, or here on the copied .ncu-rep).  The record carries the
sha256 of the library's SOURCES: bench.py only quotes it while the sources are the ones that were profiled."""
import csv
import io
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from sass_census import src_sha256  # noqa: E402


def main():
    rep, out, family, d, nsims = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), int(sys.argv[5])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]

    def col(r, name, scale_unit=True):
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        if scale_unit:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1)
        return v

    launches = []
    for r in rows[2:]:
        launches.append({"kernel": r[hdr.index("Kernel Name")], "ms": col(r, "gpu__time_duration.sum"),
                         "dram_read": col(r, "dram__bytes_read.sum"), "dram_write": col(r, "dram__bytes_write.sum"),
                         "dram_pct": col(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
                         "grid": col(r, "launch__grid_size", False), "regs": col(r, "launch__registers_per_thread", False)})
    tot = sum(l["dram_read"] + l["dram_write"] for l in launches)
    rec = {"source": f"ncu --set full --clock-control none, {os.path.basename(rep)}", "src_sha256": src_sha256(),
           "family": family, "d": d, "nsims": nsims, "launches": launches,
           "dram_bytes_per_launch_avg": tot / max(1, len(launches))}
    with open(out, "w") as fh:
        json.dump(rec, fh, indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
