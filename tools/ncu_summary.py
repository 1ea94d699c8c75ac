"""Summarises an `ncu --set full` report: one line per captured launch (duration, tensor-pipe activity, DRAM bytes and
throughput, achieved occupancy, registers, grid / block) and, with --json, the per-launch DRAM traffic that bench.py
reports as `roofline.traffic`.

  ncu -i gpurun_out/prof_x.ncu-rep --page raw --csv > /tmp/x.csv   (done here)
  python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [--json profiles/rNN_ncu_x_summary.json] [--title "..."]
"""
import argparse
import csv
import io
import json
import re
import subprocess

COLS = {
    "dur": "gpu__time_duration.sum",
    "tensor": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "tensor_alt": "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "rd": "dram__bytes_read.sum",
    "wr": "dram__bytes_write.sum",
    "dram_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram_pct_alt": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "warps": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "long_sb": "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("report")
  ap.add_argument("--json", default=None)
  ap.add_argument("--title", default=None)
  ap.add_argument("--match", default=None, help="regex on the kernel name for the --json traffic mean")
  a = ap.parse_args()
  raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr, units = rows[0], rows[1]
  idx = {h: i for i, h in enumerate(hdr)}

  def get(r, key, scale=True):
    for k in (key, key + "_alt"):
      name = COLS.get(k)
      if name in idx and r[idx[name]] not in ("", "n/a"):
        v = float(r[idx[name]].replace(",", ""))
        return v * (UNIT.get(units[idx[name]], 1.0) if scale else 1.0)
    return None

  print(f"## {a.title or a.report}  (ncu --set full --clock-control none; per-launch times are cold-cache and serialised)")
  out = []
  for r in rows[2:]:
    name = re.sub(r"\(.*$", "", r[idx["Kernel Name"]]).replace("void ", "").replace("umd::", "").replace("<unnamed>::", "")
    d = dict(kernel=name, dur_us=get(r, "dur"), tensor_pipe_active_pct=get(r, "tensor", False), dram_rd=get(r, "rd"), dram_wr=get(r, "wr"),
             dram_pct=get(r, "dram_pct", False), warps_active_pct=get(r, "warps", False), regs=get(r, "regs", False),
             grid=get(r, "grid", False), block=get(r, "block", False), long_scoreboard=get(r, "long_sb", False))
    d["dram_bytes"] = (d["dram_rd"] or 0.0) + (d["dram_wr"] or 0.0)
    d["dram_gbs"] = d["dram_bytes"] / (d["dram_us"] if False else d["dur_us"]) / 1e3 if d["dur_us"] else None
    out.append(d)
    f = lambda v, fmt: (fmt % v) if v is not None else "n/a"
    print(f"{name[:58]:58s} dur={f(d['dur_us'], '%9.1f')}us tensor={f(d['tensor_pipe_active_pct'], '%5.1f')}% "
          f"dram_rd={f((d['dram_rd'] or 0) / 1e6, '%8.1f')}MB dram_wr={f((d['dram_wr'] or 0) / 1e6, '%8.1f')}MB "
          f"dram={f(d['dram_gbs'], '%7.1f')}GB/s ({f(d['dram_pct'], '%4.1f')}%) warps={f(d['warps_active_pct'], '%4.1f')}% "
          f"regs={f(d['regs'], '%3.0f')} grid={f(d['grid'], '%5.0f')} block={f(d['block'], '%4.0f')}")
  if a.json:
    sel = [d for d in out if (re.search(a.match, d["kernel"]) if a.match else True)]
    with open(a.json, "w") as fjs:
      json.dump({"launches": sel, "traffic_bytes_per_launch_mean": sum(d["dram_bytes"] for d in sel) / max(len(sel), 1),
                 "note": f"dram__bytes_read.sum + dram__bytes_write.sum per launch from one ncu --set full capture ({a.report})"},
                fjs, indent=1)


if __name__ == "__main__":
  main()
