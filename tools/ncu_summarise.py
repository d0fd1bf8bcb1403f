"""Summarise an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` launch list:
one row per (kernel, grid size) with the number of launches, the mean duration and the mean DRAM bytes per launch, sorted by
total time; optionally a traffic JSON {kernel<args>: mean dram bytes per launch of the largest grid} for bench.py.
    python tools/ncu_summarise.py gpurun_out/r02_ncu_L4096_launches.csv profiles/r02_launches_L4096_summary.csv [profiles/traffic_L4096.json]"""
import csv
import json
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def to_bytes(v: float, unit: str) -> float:
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def to_us(v: float, unit: str) -> float:
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)


def main(src, dst, traffic=None):
    rows = defaultdict(dict)
    with open(src, newline="") as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        key = r["ID"]
        rows[key]["k"] = (short(r["Kernel Name"]), r["Grid Size"], r["Block Size"])
        val = float(r["Metric Value"].replace(",", ""))
        m = r["Metric Name"]
        if m.startswith("dram__bytes"):
            rows[key][m] = to_bytes(val, r["Metric Unit"])
        elif m.startswith("gpu__time_duration"):
            rows[key]["us"] = to_us(val, r["Metric Unit"])
    agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for v in rows.values():
        a = agg[v["k"]]
        a[0] += 1
        a[1] += v.get("us", 0.0)
        a[2] += v.get("dram__bytes_read.sum", 0.0)
        a[3] += v.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in agg.values())
    out = sorted(agg.items(), key=lambda kv: -kv[1][1])
    with open(dst, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "grid", "block", "launches", "total_us", "share", "mean_us", "mean_dram_read_MB", "mean_dram_write_MB", "mean_dram_GBps"])
        for (k, g, b), (n, us, rd, wr) in out:
            w.writerow([k, g, b, n, f"{us:.1f}", f"{us / total:.4f}", f"{us / n:.2f}", f"{rd / n / 1e6:.2f}", f"{wr / n / 1e6:.2f}",
                        f"{(rd + wr) / us / 1e3:.0f}" if us > 0 else ""])
    if traffic:
        best = {}
        for (k, g, b), (n, us, rd, wr) in out:
            gs = int(re.sub(r"[^0-9]", " ", g).split()[0])
            if k not in best or gs > best[k][0]:
                best[k] = (gs, (rd + wr) / n)
        json.dump({"source": src, "what": "mean dram__bytes_read.sum + dram__bytes_write.sum per launch (largest grid of each kernel)",
                   **{k: v[1] for k, v in best.items()}}, open(traffic, "w"), indent=1)
    print(f"{len(rows)} launches, {total / 1e3:.2f} ms under ncu; top:")
    for (k, g, b), (n, us, rd, wr) in out[:14]:
        print(f"  {us / total:6.1%} {n:5d} x {us / n:9.1f} us  {(rd + wr) / n / 1e6:9.1f} MB  {k[:90]} grid {g}")


if __name__ == "__main__":
    main(*sys.argv[1:4])
