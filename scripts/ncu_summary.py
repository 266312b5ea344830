#!/usr/bin/env python
"""Summarise one `ncu --set full` capture (.ncu-rep) into the JSON / text files kept under profiles/.
usage: python scripts/ncu_summary.py gpurun_out/X.ncu-rep profiles/NAME inband_cells "note" """
import csv
import io
import json
import re
import subprocess
import sys

rep, out, cells, note = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keep = re.compile(r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum(\.per_second)?|dram__throughput\.avg\.pct_of_peak_sustained_elapsed|"
                  r"lts__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__throughput\.avg\.pct_of_peak_sustained_elapsed|l1tex__t_sector_hit_rate\.pct|"
                  r"lts__t_sector_hit_rate\.pct|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum|smsp__warps_(eligible|active)\.avg\.per_cycle_active|"
                  r"launch__(registers_per_thread|block_size|grid_size|occupancy_limit_\w+|shared_mem_per_block_dynamic|shared_mem_config_size|waves_per_multiprocessor)|"
                  r"sm__inst_executed_pipe_(alu|fma|lsu|cbu|adu|uniform|xu)\.avg\.pct_of_peak_sustained_active|"
                  r"smsp__average_warps_issue_stalled_\w+_per_issue_active\.ratio)$")
m = {k: [v[0], v[1]] for k, v in sorted(d.items()) if keep.match(k)}
def gb(k):
    v, u = d[k]
    return float(v) * {"Tbyte": 1e12, "Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
summary = {"command": note, "inband_cells": cells, "dram_bytes_per_cell": (rd + wr) / cells, "dram_read_bytes_per_cell": rd / cells,
           "dram_write_bytes_per_cell": wr / cells, "warp_instructions_per_cell": float(d["smsp__inst_executed.sum"][0]) / cells, "metrics": m}
json.dump(summary, open(out + "_ncu_full_summary.json", "w"), indent=1)
# source page: top instructions by stall samples
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next((i for i, r in enumerate(rows) if r and r[0].strip() == "Address"), None)
if hi is not None:
    h = rows[hi]
    rows = rows[hi:]
    def col(name):
        for i, x in enumerate(h):
            if x.strip() == name:
                return i
        return None
    ci, cs, cx = col("Source"), col("Warp Stall Sampling (All Samples)"), col("Instructions Executed")
    ca = col("Address")
    cl = col("stall_long_sb")
    if None not in (ci, cs):
        body = [r for r in rows[1:] if len(r) > max(ci, cs) and r[cs].replace(".", "").isdigit()]
        tot = sum(float(r[cs]) for r in body) or 1.0
        body.sort(key=lambda r: -float(r[cs]))
        with open(out + "_ncu_source_hotspots.txt", "w") as f:
            f.write(f"# top 40 SASS instructions by warp-stall samples ({note}); address, SASS, executions, % of samples, long-scoreboard share\n")
            for r in body[:40]:
                ex = r[cx] if cx is not None else ""
                lsb = f" long_sb={100 * float(r[cl]) / tot:.2f}%" if cl is not None and r[cl].replace('.', '').isdigit() else ""
                f.write(f"{(r[ca][-5:] if ca is not None else ''):>6}  {r[ci].strip()[:70]:70s} inst={ex:>12} samples={100 * float(r[cs]) / tot:5.2f}%{lsb}\n")
print(json.dumps({k: summary[k] for k in ("dram_bytes_per_cell", "dram_write_bytes_per_cell", "warp_instructions_per_cell")}))
