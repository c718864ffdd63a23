"""Reads an ncu launch list (long CSV: one row per launch x metric; metrics gpu__time_duration.sum [+ dram__bytes_read.sum,
dram__bytes_write.sum]) of ONE replayed step and writes
    profiles/<tag>_ncu_launch_summary.txt   per-family / per-kernel device time shares (cold-cache, serialised)
    profiles/r02_dram_traffic.json          per-family DRAM bytes per launch (bench.py puts it into roofline.traffic)

    python tools/ncu_traffic.py gpurun_out/r02_ncu_step.csv r02
"""
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import bench
    path, tag = sys.argv[1], sys.argv[2]
    rows = {}
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        try:
            lid = int(r["ID"])
        except (KeyError, ValueError):
            continue
        d = rows.setdefault(lid, {"name": r["Kernel Name"]})
        val = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
        unit = r.get("Metric Unit", "")
        m = r["Metric Name"]
        if m == "gpu__time_duration.sum":
            d["us"] = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3 if unit in ("ms", "msecond") else val / 1e3)
        elif m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            d[m] = val * mult
    fam, ker = {}, {}
    for d in rows.values():
        name = d["name"]
        f_ = bench.kernel_family(name)
        for table, key in ((fam, f_), (ker, name)):
            t = table.setdefault(key, {"us": 0.0, "n": 0, "rd": 0.0, "wr": 0.0})
            t["us"] += d.get("us", 0.0)
            t["n"] += 1
            t["rd"] += d.get("dram__bytes_read.sum", 0.0)
            t["wr"] += d.get("dram__bytes_write.sum", 0.0)
    total = sum(t["us"] for t in fam.values()) or 1.0
    out = [f"{path}: {len(rows)} launches, {total / 1e3:.3f} ms serialised device time", ""]
    for k, t in sorted(fam.items(), key=lambda kv: -kv[1]["us"]):
        out.append(f"family {k:26s} {t['us'] / 1e3:8.3f} ms {100 * t['us'] / total:6.2f}% {t['n']:5d}x  dram rd {t['rd'] / 1e6:9.1f} MB wr {t['wr'] / 1e6:9.1f} MB")
    out.append("")
    for k, t in sorted(ker.items(), key=lambda kv: -kv[1]["us"])[:60]:
        out.append(f"{t['us'] / 1e3:8.3f} ms {100 * t['us'] / total:6.2f}% {t['n']:5d}x  rd/launch {t['rd'] / max(t['n'], 1) / 1e6:8.2f} MB wr/launch {t['wr'] / max(t['n'], 1) / 1e6:8.2f} MB  {k[:110]}")
    (ROOT / "profiles" / f"{tag}_ncu_launch_summary.txt").write_text("\n".join(out))
    traffic = {k: {"dram_bytes_per_launch": round((t["rd"] + t["wr"]) / max(t["n"], 1)), "launches": t["n"],
                   "dram_read_bytes": round(t["rd"]), "dram_write_bytes": round(t["wr"]),
                   "note": f"ncu dram__bytes_read.sum + dram__bytes_write.sum, average over the family's {t['n']} launches of one "
                           f"replayed step ({Path(path).name}); cold-cache, serialised"}
               for k, t in fam.items() if t["rd"] + t["wr"] > 0}
    if traffic:
        (ROOT / "profiles" / "r02_dram_traffic.json").write_text(json.dumps(traffic, indent=1))
    print("\n".join(out[:40]))


if __name__ == "__main__":
    main()
