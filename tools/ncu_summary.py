#!/usr/bin/env python
"""Condenses an ncu report (--set full) into the per-launch numbers DESIGN.md / profiles/README.md quote:
duration, DRAM bytes, achieved DRAM GB/s, L2 hit rate, atomics, occupancy, issue utilisation, top stall reasons.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_ncu_summary.csv"""
import csv
import io
import subprocess
import sys

METRICS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
           ("lts__t_sector_hit_rate.pct", "l2_hit_pct"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_atom.sum", "global_atom_sectors"),
           ("l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "global_red_sectors"),
           ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_bank_conflicts"),
           ("launch__registers_per_thread", "registers"), ("launch__grid_size", "grid"), ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
           ("launch__occupancy_limit_registers", "occ_limit_regs"), ("launch__occupancy_limit_shared_mem", "occ_limit_smem"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"), ("smsp__inst_executed.sum", "warp_instructions")]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "ms": 1e-3, "usecond": 1e-6, "us": 1e-6, "second": 1.0, "s": 1.0,
        "nsecond": 1e-9, "ns": 1e-9}


def main():
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
    out = csv.writer(sys.stdout)
    out.writerow(["id", "kernel"] + [n for _m, n in METRICS] + ["dram_GBps", "top_stalls(cycles per issue)"])
    for d in data:
        vals = {}
        for m, n in METRICS:
            if m in idx:
                v = float(d[idx[m]].replace(",", "") or 0)
                vals[n] = v * UNIT.get(units[idx[m]], 1.0)
            else:
                vals[n] = ""
        gbps = (vals["dram_read"] + vals["dram_write"]) / vals["duration"] / 1e9 if vals.get("duration") else ""
        top = sorted(((float(d[idx[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in stalls),
                     reverse=True)[:4]
        out.writerow([d[idx["ID"]], d[idx["Kernel Name"]][:60]] + [f"{vals[n]:.6g}" if vals[n] != "" else "" for _m, n in METRICS] +
                     [f"{gbps:.1f}" if gbps != "" else "", " ".join(f"{b}={a:.1f}" for a, b in top)])


if __name__ == "__main__":
    main()
