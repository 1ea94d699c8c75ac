"""Groups an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel and prints count, total and share.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.txt"""
import csv
import re
import sys
from collections import defaultdict


def main(path):
  rows = []
  with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
  for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
      continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    ns = v * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1.0)
    name = re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "")
    rows.append((name, ns, r["Grid Size"], r["Block Size"]))
  tot = sum(r[1] for r in rows)
  agg = defaultdict(lambda: [0, 0.0])
  for name, ns, *_ in rows:
    agg[name][0] += 1
    agg[name][1] += ns
  print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms total (ncu per-launch times are cold-cache and serialised: compare shares)")
  print(f"{'kernel':100s} {'n':>5s} {'ms':>10s} {'share':>7s} {'us/launch':>10s}")
  for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:100]:100s} {n:5d} {ns / 1e6:10.3f} {ns / tot:7.3f} {ns / n / 1e3:10.1f}")


if __name__ == "__main__":
  main(sys.argv[1])
