#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the few counters DESIGN.md / bench.py quote.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
    "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none (per launch; cold-cache, serialised replays)")
    for r in rows[2:]:
        print(f"kernel: {r[hdr.index('Kernel Name')]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:66s} {r[i]:>18s} {units[i]}")
        if "dram__bytes_read.sum" in hdr:
            def val(name):
                i = hdr.index(name)
                v = float(r[i].replace(",", ""))
                u = units[i].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            print(f"  {'traffic (dram read + write), bytes':66s} {val('dram__bytes_read.sum') + val('dram__bytes_write.sum'):18.0f}")
        print()


if __name__ == "__main__":
    main()
