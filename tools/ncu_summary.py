#!/usr/bin/env python3
"""Summarise an ncu report (read HERE, no GPU needed) into the text files committed under profiles/ and, optionally, into
profiles/roofline_inputs.json (the one place bench.py takes profiler-only figures from).

  python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--kernel regex] [--out profiles/r02x_ncu_<kernel>.txt]
                              [--update loglik_kernel|var_kernel] [--note "..."]
Per matching launch: duration, DRAM bytes, registers, occupancy limits, pipe utilisations, issue statistics and the warp
stall breakdown; --update writes the LAST matching launch's figures into roofline_inputs.json.
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum$|launch__(block_size|grid_size|registers_per_thread|"
                  r"occupancy_limit_(registers|shared_mem|warps)|shared_mem_per_block_dynamic)$|sm__cycles_elapsed\.avg$|"
                  r"sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu|uniform)\.avg\.pct_of_peak_sustained_active$|"
                  r"sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_(active|elapsed)$|sm__throughput\.avg\.pct_of_peak_sustained_elapsed$|"
                  r"sm__warps_active\.avg\.per_cycle_active$|smsp__inst_executed\.sum$|smsp__issue_active\.avg\.pct_of_peak_sustained_active$|"
                  r"smsp__thread_inst_executed_per_inst_executed\.ratio$|smsp__average_warp_latency_per_inst_issued\.ratio$|"
                  r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|"
                  r"smsp__inst_executed_op_shared_(ld|st)\.sum$|launch__occupancy_per_block_size$)")


def to_bytes(value, unit):
    m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(value) * m.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default=".*")
    ap.add_argument("--out", default="")
    ap.add_argument("--update", default="")
    ap.add_argument("--note", default="")
    ap.add_argument("--walkers", type=int, default=0, help="walkers of the profiled launch (stored with --update)")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr_i], rows[hdr_i + 1]
    lines, last = [], None
    for r in rows[hdr_i + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, r))
        if not re.search(a.kernel, d["Kernel Name"]):
            continue
        lines.append("## launch id %s: %s  grid %s block %s" % (d["ID"], d["Kernel Name"], d["Grid Size"], d["Block Size"]))
        cur = {"kernel": d["Kernel Name"]}
        for n, u in zip(names, units):
            short = n.split(".", 2)[-1] if n.count(".") >= 2 and n.split(".")[1].startswith("Triage") else n
            if KEEP.match(short) and d[n] != "":
                lines.append("%s [%s] = %s" % (short, u, d[n]))
                cur[short] = (d[n].replace(",", ""), u)
        last = cur
    text = "# %s\n# source: %s  (ncu --set full --clock-control none)\n%s\n" % (a.note, os.path.basename(a.report), "\n".join(lines))
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
    else:
        sys.stdout.write(text)
    if a.update and last:
        path = os.path.join(ROOT, "profiles", "roofline_inputs.json")
        with open(path) as f:
            ri = json.load(f)
        rd = to_bytes(*last.get("dram__bytes_read.sum", ("0", "byte")))
        wr = to_bytes(*last.get("dram__bytes_write.sum", ("0", "byte")))
        ms = last.get("gpu__time_duration.sum", ("0", "ms"))
        t_ms = float(ms[0]) * {"ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(ms[1], 1.0)
        ri[a.update] = {"kernel": last["kernel"], "dram_bytes_per_launch": rd + wr,
                        "fp64_pipe_active_pct": float(last["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
                        "issue_active_pct": float(last["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
                        "registers_per_thread": int(float(last["launch__registers_per_thread"][0])),
                        "gpu_time_ms": t_ms, "source": a.out or a.report, "note": a.note}
        if a.walkers:
            ri[a.update]["walkers_per_launch"] = a.walkers
        with open(path, "w") as f:
            json.dump(ri, f, indent=2)
            f.write("\n")


if __name__ == "__main__":
    main()
