#!/usr/bin/env python
"""Summarise an .ncu-rep (captured with `ncu --set full --import-source on`) into the handful of numbers DESIGN.md and
bench.py's roofline quote: duration, DRAM bytes, L2/L1 hit rates, occupancy, stall reasons, top stalled source lines.

  python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [--kernel regex] [--source N]   > profiles/rNN_xxx.txt
"""
import argparse
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("--kernel", default=".")
    ap.add_argument("--source", type=int, default=14)
    a = ap.parse_args()
    hdr, units, rows = raw(a.rep)
    ki = hdr.index("Kernel Name")
    for r in rows:
        if not re.search(a.kernel, r[ki]):
            continue
        print("=" * 100)
        print("kernel:", r[ki][:140], " id:", r[0])
        for m in METRICS:
            if m in hdr:
                i = hdr.index(m)
                print("  %-72s %16s %s" % (m, r[i], units[i]))
        if "dram__bytes_read.sum" in hdr:
            def val(m):
                i = hdr.index(m)
                v, u = float(r[i].replace(",", "")), units[i].lower()
                return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
            tr = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
            i = hdr.index("gpu__time_duration.sum")
            t = float(r[i].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(units[i], 1e-9)
            print("  %-72s %16.0f byte  (%.1f GB/s under ncu)" % ("dram traffic (read + write)", tr, tr / t / 1e9))
        st = [(float(r[i]), hdr[i]) for i in range(len(hdr))
              if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active.ratio", hdr[i]) and r[i] not in ("", "n/a")]
        print("  stall reasons (warps stalled per issue-active cycle):")
        for v, n in sorted(st, reverse=True)[:7]:
            print("     %-28s %8.2f" % (n.split("stalled_")[1].split("_per_issue")[0], v))
    if a.source:
        out = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                              "--kernel-name", "regex:" + a.kernel], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        fpath, h, lines, first_fn, fn = "", None, [], None, None
        for r in rows:
            if not r:
                continue
            if r[0] == "File Path":
                fpath = r[1].split("/")[-1]
            elif r[0] == "Function Name":
                fn = r[1]
                if first_fn is None:
                    first_fn = fn
            elif r[0] == "Line No":
                h = r
            elif h and fn == first_fn and r[0].isdigit():
                extra = len(r) - len(h)  # unescaped quotes / commas inside the source text
                if extra > 0:
                    r = [r[0], ",".join(r[1:2 + extra])] + r[2 + extra:]
                d = dict(zip(h, r))
                try:
                    smp = float(d.get("# Samples", "0") or 0)
                except ValueError:
                    continue
                stalls = sorted(((float(v), k) for k, v in d.items() if k.startswith("stall_") and "Not Issued" not in k
                                 and v.replace(".", "").isdigit()), reverse=True)[:3]
                lines.append((smp, fpath, r[0], r[1].strip(), float(d.get("Instructions Executed", "0") or 0), stalls))
        tot = sum(x[0] for x in lines) or 1
        toti = sum(x[4] for x in lines) or 1
        print("  top source lines by warp-stall samples (%d samples, %.0f warp-instructions; first launch in the report):" % (tot, toti))
        for smp, f, ln, src, inst, stalls in sorted(lines, reverse=True)[:a.source]:
            print("   %5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100 * smp / tot, 100 * inst / toti, f, ln, src[:88]))
            print("          " + ", ".join("%s %.0f" % (k[6:], v) for v, k in stalls if v > 0))


if __name__ == "__main__":
    sys.exit(main())
