#!/usr/bin/env python
"""Summarises a `ncu --set full` capture (one kernel) into a small JSON file for profiles/:
    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.json --command "..." --reading "..."
Reads the report with `ncu -i ... --page raw --csv` and keeps the metrics the roofline discussion uses."""
import argparse
import csv
import io
import json
import subprocess

KEEP = ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "launch__block_size",
        "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "lts__t_sector_hit_rate.pct",
        "sm__cycles_elapsed.avg.per_second", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__ops_path_tensor_src_fp64.sum", "sm__ops_path_tensor_src_tf32_dst_fp32.sum")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("out")
    ap.add_argument("--command", default="")
    ap.add_argument("--reading", default="")
    a = ap.parse_args()
    raw = subprocess.check_output(["ncu", "-i", a.report, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {}
    for h, u, v in zip(hdr, units, vals):
        if h in KEEP:
            out[h] = (v + " " + u).strip()
    def num(key):
        s = out.get(key, "").split()
        if not s:
            return None
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(s[1] if len(s) > 1 else "byte", 1)
        return float(s[0]) * mult
    r, w = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    if r is not None and w is not None:
        out["_traffic_bytes_per_launch"] = r + w
    if a.command:
        out["_command"] = a.command
    if a.reading:
        out["_reading"] = a.reading
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(a.out, out.get("gpu__time_duration.sum"), out.get("_traffic_bytes_per_launch"))


if __name__ == "__main__":
    main()
