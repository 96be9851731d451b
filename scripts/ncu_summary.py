#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small text files for profiles/.

  python scripts/ncu_summary.py launches <launches.csv> <out.md>      # per-kernel launch list -> shares
  python scripts/ncu_summary.py full <report.ncu-rep> <out.md>        # --set full capture -> key metrics

Runs on the CPU box (needs the `ncu` binary only for the `full` mode, to read the report).
"""
import collections
import csv
import io
import re
import subprocess
import sys

OURS = re.compile(r"(dwt_|bpc_|bsb_|export_ref)")
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
    "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "lts__t_sector_hit_rate.pct",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("pcs::", "").lstrip("<")
    return name.strip()


def launches(path, out):
    rows = [r for r in csv.DictReader(l for l in open(path) if l.startswith('"'))]
    agg = collections.OrderedDict()
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        k = short(r["Kernel Name"])
        ns = float(r["Metric Value"].replace(",", ""))
        if r["Metric Unit"] in ("us", "usecond"):
            ns *= 1e3
        elif r["Metric Unit"] in ("ms", "msecond"):
            ns *= 1e6
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1; a[1] += ns
    ours = {k: v for k, v in agg.items() if OURS.search(k)}
    tot = sum(v[1] for v in ours.values()) or 1.0
    with open(out, "w") as f:
        f.write(f"# ncu launch list summary ({path})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`: cold-cache, serialised -- compare SHARES.\n\n")
        f.write("| kernel (library kernels only) | launches | total us | avg us | share of library time |\n|---|---|---|---|---|\n")
        for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {v[1] / 1e3 / v[0]:.1f} | {100 * v[1] / tot:.1f}% |\n")
        other = sum(v[1] for k, v in agg.items() if k not in ours)
        f.write(f"\nNon-library launches (torch fills/copies of the harness): {sum(v[0] for k, v in agg.items() if k not in ours)}, {other / 1e3:.1f} us.\n")


def full(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary ({path})\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"## `{short(d['Kernel Name'])}`  grid {d['Grid Size']} block {d['Block Size']}\n\n")
            for m in METRICS:
                if m in d:
                    f.write(f"- {m} [{units[hdr.index(m)]}] = {d[m]}\n")
            f.write("\n")


def traffic(path, out, workload="cfg2", frames=16, nlaunch=5):
    """dram bytes (read + write) of the first `nlaunch` dwt_forward launches of a --set full capture = one
    pcs_dwt_forward call over the bench batch -> profiles/dwt_traffic.json (bench.py's roofline.traffic)."""
    import json
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    ir, iw, ik, it = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name"), hdr.index("gpu__time_duration.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, t_us, n = 0.0, 0.0, 0
    for r in rows[2:]:
        if "dwt_forward" not in r[ik] or n >= nlaunch:
            continue
        tot += float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
        t_us += float(r[it].replace(",", "")) * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(units[it], 1)
        n += 1
    json.dump({"workload": workload, "frames": frames, "dram_bytes": int(tot), "launches": n, "ncu_sum_us": t_us,
               "source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum over the {n} forward-DWT launches of one call ({path})"},
              open(out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "traffic" and len(sys.argv) > 4:  # ... traffic <report> <out.json> <frames>
        traffic(sys.argv[2], sys.argv[3], frames=int(sys.argv[4]))
    else:
        {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])
