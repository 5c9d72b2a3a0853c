"""ncu raw-page CSV (ncu -i rep --page raw --csv) -> one line per libb200pt kernel launch with the counters the roofline
claims rest on. usage: ncu_summary.py raw.csv [hbm_peak_gbs]"""
import csv
import json
import sys
from pathlib import Path

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
try:
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else json.loads((Path(__file__).resolve().parents[2] / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
except Exception:
    peak = 6650.0
col = {n: i for i, n in enumerate(hdr)}
M = {
    "t_us": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
    "l2hit": "lts__t_sector_hit_rate.pct", "clk": "gpc__cycles_elapsed.max",
}


def val(r, key):
    i = col.get(M[key])
    if i is None or r[i] == "":
        return float("nan")
    v = float(r[i].replace(",", ""))
    u = units[i]
    if key == "t_us":
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
    if key in ("rd", "wr"):
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1.0)
    return v


print(f"# per-launch ncu --set full counters (cold-cache, serialised); HBM fraction = (dram read + write) / time / {peak:.0f} GB/s measured copy")
print(f"{'kernel':58s} {'time us':>9s} {'dram rd MB':>10s} {'dram wr MB':>10s} {'GB/s':>7s} {'of HBM':>6s} {'tensor%':>7s} {'sm%':>5s} {'L2hit%':>6s} {'regs':>4s} {'grid':>6s} {'blk':>4s} {'SM MHz':>6s}")
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    if "at::" in name or "cub::" in name:  # torch's own fill / random / copy kernels of the set-up code
        continue
    name = name.replace("void ", "").replace("b200::", "").split("(")[0]
    t, rd, wr = val(r, "t_us"), val(r, "rd"), val(r, "wr")
    gbs = (rd + wr) / t / 1e3
    mhz = val(r, "clk") / t
    print(f"{name[:58]:58s} {t:9.1f} {rd / 1e6:10.1f} {wr / 1e6:10.1f} {gbs:7.0f} {gbs / peak:6.2f} {val(r, 'tensor'):7.1f} {val(r, 'sm_pct'):5.1f} {val(r, 'l2hit'):6.1f} "
          f"{int(val(r, 'regs')):4d} {int(val(r, 'grid')):6d} {int(val(r, 'block')):4d} {mhz:6.0f}")
