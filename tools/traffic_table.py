"""Per-shape DRAM traffic of the GEMM launches of one training step: joins the ncu launch list of `tools/step_once.py`
(metrics dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum, sm__pipe_tensor_cycles_active...; kernel filter
regex:gemm_kernel) with the shapes the library logged in launch order (UMD_GEMM_LOG), and writes
profiles/rNN_traffic_table.json: for every (role, M, N, K, batch, epilogue) the launches per step, the algorithmic bytes
(each operand and each output once) and the measured bytes per launch.  bench.py reports the launch-weighted mean over the
forward / dgrad GEMMs as `roofline.traffic`.

  python tools/traffic_table.py gpurun_out/gemm_launches.csv gpurun_out/gemm_log.txt profiles/r02_traffic_table.json
"""
import csv
import json
import re
import subprocess
import sys
from collections import OrderedDict

EPI = {0: "bf16", 1: "f32", 2: "gelu(u,g)", 3: "gate_res", 4: "dgelu", 5: "atomic(f32 +=)", 6: "bf16+delta"}


def out_bytes(M, N, batch, epi):
  per = {0: 2, 1: 4, 2: 4, 3: 6, 4: 2 + 2, 5: 4, 6: 2 + 2}[epi]   # dgelu / delta also read a bf16 operand tile
  return M * N * batch * per


def main(csv_path, log_path, out_path):
  lines = [l for l in open(csv_path, newline="") if not l.startswith("==")]
  rows = list(csv.DictReader(lines))
  launches = OrderedDict()
  for r in rows:
    if "gemm_kernel" not in r["Kernel Name"]:
      continue
    d = launches.setdefault(r["ID"], {"kernel": re.sub(r"\(.*$", "", r["Kernel Name"]).replace("void ", "").replace("umd::", "")})
    v = float(r["Metric Value"].replace(",", ""))
    u = r.get("Metric Unit", "")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "%": 1.0}.get(u, 1.0)
    d[r["Metric Name"]] = v * scale
  launches = list(launches.values())
  shapes = [tuple(int(x) for x in l.split()) for l in open(log_path) if l.strip()]
  assert len(shapes) >= len(launches) > 0, (len(shapes), len(launches))
  shapes = shapes[:len(launches)]
  table = OrderedDict()
  for L, shp in zip(launches, shapes):
    M, N, K, batch, a_mn, b_mn, epi, split_k, bn, cta2 = shp[:10]
    a_bc, b_bc = (shp[10], shp[11]) if len(shp) >= 12 else (0, 0)
    role = "wgrad" if a_mn else ("dgrad" if not b_mn else "fwd")
    key = f"{role} M={M} N={N} K={K} batch={batch} epi={EPI[epi]}"
    e = table.setdefault(key, {"role": role, "M": M, "N": N, "K": K, "batch": batch, "epilogue": EPI[epi], "split_k": split_k,
                               "tile_n": bn, "cta_pairs": cta2, "a_broadcast": a_bc, "b_broadcast": b_bc, "kernel": L["kernel"], "launches": 0, "dram_bytes": 0.0, "dur_us": 0.0,
                               "tensor_pct": 0.0})
    e["launches"] += 1
    e["dram_bytes"] += L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
    e["dur_us"] += L.get("gpu__time_duration.sum", 0.0)
    e["tensor_pct"] += L.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
  for e in table.values():
    n = e["launches"]
    e["dram_bytes_per_launch"] = e.pop("dram_bytes") / n
    e["dur_us_per_launch"] = e.pop("dur_us") / n
    e["tensor_pipe_active_pct"] = e.pop("tensor_pct") / n
    na = 1 if e.get("a_broadcast") else e["batch"]      # a broadcast operand is read once for the whole batch
    nb = 1 if e.get("b_broadcast") else e["batch"]
    e["algorithmic_bytes_per_launch"] = 2 * (na * e["M"] * e["K"] + nb * e["K"] * e["N"]) + out_bytes(e["M"], e["N"], e["batch"], [k for k, v in EPI.items() if v == e["epilogue"]][0])
    e["traffic_over_algorithmic"] = e["dram_bytes_per_launch"] / e["algorithmic_bytes_per_launch"]
    e["flops_per_launch"] = 2.0 * e["M"] * e["N"] * e["K"] * e["batch"]
  fd = [e for e in table.values() if e["role"] != "wgrad"]
  nfd = sum(e["launches"] for e in fd)
  try:
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
  except Exception:
    commit = ""
  commit = commit or "(stamped when the table is copied into profiles/: the GPU box has no .git)"
  out = {"commit": commit, "source": f"{csv_path} + {log_path} (one step of tools/step_once.py under ncu; cold-cache, serialised launches)",
         "fwd_dgrad_launches_per_step": nfd,
         "fwd_dgrad_traffic_bytes_per_launch": sum(e["dram_bytes_per_launch"] * e["launches"] for e in fd) / nfd,
         "fwd_dgrad_algorithmic_bytes_per_launch": sum(e["algorithmic_bytes_per_launch"] * e["launches"] for e in fd) / nfd,
         "shapes": table}
  json.dump(out, open(out_path, "w"), indent=1)
  print(f"{len(launches)} launches, {len(table)} shapes -> {out_path}")
  for k, e in table.items():
    print(f"{k:64s} n={e['launches']:3d} {e['dur_us_per_launch']:8.1f} us tensor {e['tensor_pipe_active_pct']:5.1f}% "
          f"dram {e['dram_bytes_per_launch'] / 1e6:8.1f} MB alg {e['algorithmic_bytes_per_launch'] / 1e6:8.1f} MB x{e['traffic_over_algorithmic']:.2f}")


if __name__ == "__main__":
  main(*sys.argv[1:4])
