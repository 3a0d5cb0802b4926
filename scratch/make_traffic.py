"""profiles/traffic.json (what bench.py scales into roofline.traffic / dram_frac and quotes as
on-chip counters) from an `ncu --page raw --csv` export of the dominant kernels.

    python scratch/make_traffic.py <raw.csv> <templates_per_launch> <pixels> "<source note>" [more raw.csv ...]
"""
import csv
import json
import re
import sys

KEY = {"k_conv_cols": "k_conv_cols", "k_fit_rows": "k_fit_rows", "k_curv_rows": "k_curv_rows", "k_curv_cols": "k_curv_cols",
       "k_tmpl_rows": "k_tmpl_rows"}


def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def main():
    raw, tpl, pixels, note = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    rows = list(csv.reader(open(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = {}
    for r in rows[2:]:
        name = re.sub(r"^void ", "", r[ix["Kernel Name"]])
        kern = re.match(r"(?:sb\w*::)?(\w+)", name).group(1)
        key = next((v for k, v in KEY.items() if kern.startswith(k)), None)
        if key is None or key in out:
            continue
        rd = to_bytes(r[ix["dram__bytes_read.sum"]], units[ix["dram__bytes_read.sum"]])
        wr = to_bytes(r[ix["dram__bytes_write.sum"]], units[ix["dram__bytes_write.sum"]])
        g = lambda m: float(r[ix[m]]) if m in ix and r[ix[m]] not in ("", "n/a") else None
        out[key] = {
            "kernel": kern, "templates_per_launch": tpl, "px_evals_per_launch": tpl * pixels,
            "dram_bytes_per_launch": rd + wr, "dram_bytes_per_px_eval": (rd + wr) / (tpl * pixels),
            "time_ms_under_ncu": g("gpu__time_duration.sum"),
            "lsu_wavefront_pct": g("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "fma_pipe_pct": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
            "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "registers_per_thread": g("launch__registers_per_thread"),
            "xbar2l1tex_read_bytes": to_bytes(r[ix["l1tex__m_xbar2l1tex_read_bytes.sum"]], units[ix["l1tex__m_xbar2l1tex_read_bytes.sum"]]),
            "source": note}
    json.dump(out, sys.stdout, indent=1)


if __name__ == "__main__":
    main()
