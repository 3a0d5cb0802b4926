"""Summarise an `ncu --page raw --csv` export and a launch list into profiles/ (markdown + traffic.json).

    python scratch/ncu_summary.py <raw.csv> <launches.csv> <out_prefix> "<title>" "<command>"
"""
import csv
import json
import re
import sys
from collections import OrderedDict

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
]


def short(name):
    m = re.match(r"(?:void )?(?:\w+::)*(\w+)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def to_bytes(v, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    return float(v) * scale


def main():
    raw, launches, prefix, title, command = sys.argv[1:6]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = ["# %s" % title, "", "Command: `%s`" % command, "",
           "Values per launch; a launch of `k_conv_cols_f`/`k_fit_rows_f` covers a batch of templates.", ""]
    traffic = OrderedDict()
    seen = {}
    for r in rows[2:]:
        name = short(r[ix["Kernel Name"]])
        seen[name] = seen.get(name, 0) + 1
        out += ["## %s (capture %d)" % (name, seen[name]), "", "| metric | value |", "|---|---|"]
        for m in METRICS:
            if m in ix:
                out.append("| `%s` | %s %s |" % (m, r[ix[m]], units[ix[m]]))
        out.append("")
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        key = re.sub(r"_f$", "", name.split("<")[0])
        if key not in traffic:
            traffic[key] = {"dram_bytes_per_launch": rd + wr, "grid": r[ix["launch__grid_size"]],
                            "time_ms_under_ncu": float(r[ix["gpu__time_duration.sum"]])
                            if units[ix["gpu__time_duration.sum"]] == "ms" else r[ix["gpu__time_duration.sum"]]}
    # launch list: per-kernel totals and shares
    tot = OrderedDict()
    n = 0
    with open(launches) as f:
        rd = csv.reader(l for l in f if l.startswith('"'))
        h = next(rd)
        ki, vi = h.index("Kernel Name"), h.index("Metric Value")
        for r in rd:
            if len(r) <= vi:
                continue
            k = short(r[ki])
            t = tot.setdefault(k, [0, 0.0])
            t[0] += 1
            t[1] += float(r[vi].replace(",", "")) / 1e6
            n += 1
    total_ms = sum(v[1] for v in tot.values())
    out += ["## Launch list (`--metrics gpu__time_duration.sum`, cold-cache, serialised): share of device time", "",
            "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.2f | %.1f %% |" % (k, v[0], v[1], 100 * v[1] / total_ms))
    out.append("")
    open(prefix + "_ncu_full_summary.md", "w").write("\n".join(out))
    json.dump(traffic, open(prefix + "_traffic.json", "w"), indent=1)
    print("\n".join(out[-12:]))


if __name__ == "__main__":
    main()
