"""Compact per-launch table from `ncu -i X.ncu-rep --page raw --csv` output (tools/gpu_call*.sh write the csv on the box).

    python tools/ncu_summary.py gpurun_out/r2_small.raw.csv > profiles/r2_ncu_small_kernels.txt

Per launch: duration, DRAM bytes read + written and the achieved GB/s (with the fraction of the measured HBM peak in
MEASURED_PEAKS.json), DRAM / L2 / L1 / SM throughput as % of peak, the fp32 (fma), fp64 and tensor pipe activity, issue
slot utilisation, occupancy and registers."""
import csv
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("dram__bytes_read.sum", "rdMB", 1e-6),
    ("dram__bytes_write.sum", "wrMB", 1e-6),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%", 1),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%", 1),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%", 1),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tens%", 1),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
    ("launch__registers_per_thread", "regs", 1),
]
UNIT = {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6,
        "Gbyte": 1e9}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    peak = 6650.0
    pk = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    print(f"# {os.path.basename(path)}: ncu --set full --clock-control none, one row per captured launch; "
          f"GB/s = (DRAM read + write) / duration, hbm = fraction of the measured copy peak {peak:.0f} GB/s")
    print("%-34s %-14s %9s %9s %9s %7s %6s " % ("kernel", "grid x block", "us", "rdMB", "wrMB", "GB/s", "hbm") +
          " ".join("%6s" % c[1] for c in COLS[3:]))
    for r in data:
        name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").replace("nvs::", "")
        def val(col):
            i = ix.get(col)
            if i is None or r[i] in ("", "n/a"):
                return float("nan")
            return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        ns = val("gpu__time_duration.sum")
        rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
        gbs = (rd + wr) / ns
        grid = "%s x %s" % (r[ix["launch__grid_size"]], r[ix["launch__block_size"]])
        print("%-34s %-14s %9.1f %9.2f %9.2f %7.0f %6.3f " % (name[:34], grid, ns / 1e3, rd / 1e6, wr / 1e6, gbs, gbs / peak) +
              " ".join("%6.1f" % val(c[0]) if c[1] != "regs" else "%6d" % int(val(c[0])) for c in COLS[3:]))


if __name__ == "__main__":
    main(sys.argv[1])
