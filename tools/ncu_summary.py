#!/usr/bin/env python
"""Summarise an .ncu-rep or its exported raw-page CSV (read here, no GPU needed): one row per launch with the metrics DESIGN.md quotes.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.csv] [--traffic profiles/ncu_traffic.json]

--traffic merges {entry point: {"dram_bytes_per_launch": read+write}} for the kernels bench.py reports
`roofline.traffic` for (largest launch of each kernel family).
"""
import csv
import io
import json
import subprocess
import sys

COLS = [
    ("Kernel Name", "kernel"),
    ("gpu__time_duration.sum", "time_us"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_active_pct"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved_occupancy_pct"),
    ("launch__registers_per_thread", "regs"),
    ("launch__occupancy_limit_registers", "occ_limit_regs_blocks"),
    ("launch__occupancy_limit_shared_mem", "occ_limit_smem_blocks"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared_pipe_pct"),
    ("smsp__inst_executed.sum", "warp_instructions"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]
FAMILY = {"blend_bwd_kernel": "pxb_blend_backward", "blend_fwd_kernel": "pxb_blend_forward",
          "fused_bwd_kernel": "pxb_fused_backward", "fused_fwd_kernel": "pxb_fused_forward",
          "l1_ssim_fwd_kernel": "pxb_l1_ssim_loss_forward", "l1_ssim_bwd_kernel": "pxb_l1_ssim_backward"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3,
        "usecond": 1.0, "nsecond": 1e-3}


def main():
    rep = sys.argv[1]
    out = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else None
    traffic_path = sys.argv[sys.argv.index("--traffic") + 1] if "--traffic" in sys.argv else None
    if rep.endswith(".csv"):  # already exported with `ncu -i x.ncu-rep --page raw --csv` (on the GPU box)
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {n: hdr.index(n) for n, _ in COLS if n in hdr}
    res = []
    for r in data:
        d = {}
        for n, short in COLS:
            if n not in idx:
                continue
            v = r[idx[n]]
            if short == "kernel":
                d[short] = v.split("(")[0].replace("void ", "").replace("pxb::", "")
                continue
            try:
                f = float(v.replace(",", ""))
            except ValueError:
                d[short] = v
                continue
            d[short] = f * UNIT.get(units[idx[n]], 1.0) if short in ("time_us", "dram_read", "dram_write") else f
        res.append(d)
    w = csv.writer(open(out, "w", newline="") if out else sys.stdout)
    names = [s for n, s in COLS if n in idx]
    w.writerow(names)
    for d in res:
        w.writerow([round(d[k], 3) if isinstance(d.get(k), float) else d.get(k, "") for k in names])
    if traffic_path:
        try:
            tr = json.load(open(traffic_path))
        except Exception:
            tr = {}
        best = {}
        for d in res:
            for fam, entry in FAMILY.items():
                if fam in d["kernel"] and d.get("time_us", 0) > best.get(entry, (0, 0))[0]:
                    best[entry] = (d["time_us"], d.get("dram_read", 0) + d.get("dram_write", 0))
        for entry, (t, b) in best.items():
            tr[entry] = {"dram_bytes_per_launch": int(b), "time_us_under_ncu": round(t, 1), "source": rep.split("/")[-1]}
        json.dump(tr, open(traffic_path, "w"), indent=1)


if __name__ == "__main__":
    main()
