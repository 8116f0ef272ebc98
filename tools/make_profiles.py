#!/usr/bin/env python
"""Turn an `ncu --set full` report (read here with `ncu -i <rep> --page raw --csv`) into the committed summaries:

    python tools/make_profiles.py gpurun_out/r02_full.ncu-rep profiles/r02_ncu_summary_512cubed.md profiles/r02_traffic_512cubed.json

The markdown lists every captured launch (time, DRAM bytes, achieved DRAM GB/s and fraction of the measured copy peak, issue
utilisation, fp64 pipe, occupancy, registers, top stall reason); the JSON holds DRAM bytes per launch per kernel and the DRAM
bytes of the max-flow phase of one step (everything between the build kernel and the read-out), which bench.py divides by the
live phase time."""
import csv
import json
import os
import re
import subprocess
import sys

PEAK = 6583.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*", "", name)


def main():
    rep, md, js = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)
    hdr, units, rows = raw_rows(rep)
    ci = {h: i for i, h in enumerate(hdr)}

    def val(r, key, default=0.0):
        try:
            return float(r[ci[key]].replace(",", ""))
        except Exception:
            return default

    def scaled(r, key):
        """value in base units (ncu prints Gbyte / Mbyte / ms / us per column)."""
        v = val(r, key)
        u = units[ci[key]] if key in ci else ""
        mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
                "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
        return v * mult

    stall_keys = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    lines = []
    per_kernel = {}
    phase_bytes = 0.0
    in_phase = False
    for r in rows:
        name = short(r[ci["Kernel Name"]])
        t = scaled(r, "gpu__time_duration.sum")
        rd, wr = scaled(r, "dram__bytes_read.sum"), scaled(r, "dram__bytes_write.sum")
        gbs = (rd + wr) / t / 1e9 if t > 0 else 0.0
        stalls = sorted(((val(r, k), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stall_keys
                         if "selected" not in k), reverse=True)
        top = ", ".join("%s %.1f" % (n, v) for v, n in stalls[:2])
        lines.append("| %s | %s | %.3f | %.3f | %.3f | %.0f | %.1f | %.1f | %.1f | %.1f | %d | %s |" % (
            name, r[ci["Grid Size"]] if "Grid Size" in ci else "", t * 1e3, rd / 1e9, wr / 1e9, gbs, 100 * gbs / PEAK,
            val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), val(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
            val(r, "sm__warps_active.avg.pct_of_peak_sustained_active"), int(val(r, "launch__registers_per_thread")), top))
        per_kernel.setdefault(name.split("<")[0], []).append(rd + wr)
        if name.startswith("k_build_tile") or name.startswith("k_init_tile"):
            in_phase, phase_bytes = True, 0.0
        elif name.startswith("k_readout"):
            in_phase = False
        elif in_phase and not name.startswith("k_sum_partials"):
            phase_bytes += rd + wr
    with open(md, "w") as fh:
        fh.write("# %s\n\n" % title)
        fh.write("Source: `ncu --set full --clock-control none --import-source on` (report not committed), read with `ncu -i ... --page raw --csv` "
                 "by tools/make_profiles.py.  Per-launch numbers are cold-cache and serialised (and taken at the profiler's clocks): compare shares, not absolutes.  "
                 "Peak = %.1f GB/s measured copy (MEASURED_PEAKS.json).\n\n" % PEAK)
        fh.write("| kernel | grid | time ms | DRAM read GB | DRAM write GB | DRAM GB/s | % of measured peak | sm issue % | fp64 pipe % | occupancy % | regs | top stalls (warps per issue) |\n")
        fh.write("|---|---|---|---|---|---|---|---|---|---|---|---|\n")
        fh.write("\n".join(lines) + "\n")
    out = {"bytes_per_launch": {k: sum(v) / len(v) for k, v in per_kernel.items()}, "launches": {k: len(v) for k, v in per_kernel.items()},
           "maxflow_phase_bytes_per_step": phase_bytes, "source": os.path.basename(rep),
           "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (mean over the captured launches of each kernel); "
                   "maxflow_phase_bytes_per_step = all launches between the build kernel and k_readout of the last captured step"}
    with open(js, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", md, js)


if __name__ == "__main__":
    main()
