#!/usr/bin/env python
"""Digest of an .ncu-rep (read here, no GPU needed): the counters DESIGN.md / bench.py quote PLUS the warp-stall
breakdown (smsp__average_warps_issue_stalled_*_per_issue_active, smsp__issue_active) and the hottest source lines
(--page source), so that "register-bound" / "shared-memory bound" claims are shown, not asserted.

  python tools/ncu_digest.py gpurun_out/x.ncu-rep [n_source_lines] > profiles/rNN_x.txt
"""
import csv
import io
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
    "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread", "launch__block_size",
    "launch__grid_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__cycles_elapsed.avg.per_second",
]


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none --import-source on (one launch; cold-cache, serialised replays)")
    for r in rows[2:]:
        print(f"kernel: {r[hdr.index('Kernel Name')]}")
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:72s} {r[i]:>18s} {units[i]}")
        if "dram__bytes_read.sum" in hdr:
            def val(name):
                i = hdr.index(name)
                return num(r[i]) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(units[i].lower(), 1)
            print(f"  {'traffic (dram read + write), bytes':72s} {val('dram__bytes_read.sum') + val('dram__bytes_write.sum'):18.0f}")
        stalls = [(num(r[i]), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        print("  warp stall reasons (warps stalled per issue-active cycle, largest first):")
        for v, h in sorted(stalls, reverse=True)[:8]:
            name = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
            print(f"    {name:28s} {v:8.3f}")
        print()
    # source view "cuda,sass": one aggregate row per CUDA source line (needs -lineinfo + --import-source on), followed
    # by that line's SASS rows (empty "Line No"); several files (headers inlined into the kernel) follow each other
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                         capture_output=True, text=True).stdout
    fname, hdr2, lines = "", None, []
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr2 = r
        elif hdr2 and r[0].strip().isdigit():
            lines.append((fname, r))
    if hdr2 and lines:
        ci, cx = hdr2.index("# Samples"), hdr2.index("Instructions Executed")
        total = sum(num(r[ci]) for _, r in lines) or 1.0
        total_x = sum(num(r[cx]) for _, r in lines) or 1.0
        print(f"  hottest source lines by warp-state samples ({int(total)} samples; second column: share of the warp instructions executed):")
        for f, r in sorted(lines, key=lambda fr: -num(fr[1][ci]))[:top]:
            print(f"    {100 * num(r[ci]) / total:5.1f} %  {100 * num(r[cx]) / total_x:5.1f} %  {f}:{r[0]}  {r[1].strip()[:120]}")


if __name__ == "__main__":
    main()
